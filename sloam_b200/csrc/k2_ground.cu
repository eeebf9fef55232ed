// k2_ground.cu -- stages a3 + a4 + a5: polar ground cells, bottom-k% retention,
// per-cell plane fit and acceptance.
//
// Replaces sloam::binGroundPoints (sloam/src/core/sloam.cpp:330-386), the Plane
// constructor + computeModel (sloam/src/objects/plane.cpp:3-17,96-128) and the
// acceptance test in sloam::computeModels (sloam.cpp:394-412).
//
// Binning (stable, by polar cell, (z key, point index) records of cell 0, then cell 1, ... each
// in input order).  Fused pipeline: the split kernel (k1_project.cu) leaves one record and one
// cell tag per ground point, tile by tile in input order, plus the number of points of every
// cell in every tile; ground_offsets_kernel turns the counts into first slots and
// ground_scatter_kernel moves the records, one WARP per 1024-point tile, no barriers, no
// atomics.  Caller-supplied clouds (stage entry): ground_bin_kernel, one CTA per keyframe.
// ground_cells_kernel: one warp per (keyframe, cell) reads its members contiguously, selects
// the r lowest z with a radix select on (key - kmin) whose digits start at the top significant
// bit of the cell's key range and which finishes among the members of the target bin once that
// holds at most 32 (ties by input order), and sorts only those r; ground_fit_kernel (one warp
// per cell with enough retained points) fits the plane:
//   - centroid: float32 sequential sum in sorted order (utils.h:14-28), one lane;
//   - 3 x n JacobiSVD: column-pivoted Householder QR of the n x 3 adjoint with
//     warp-shuffle reductions, then the 3x3 two-sided Jacobi of dev_plane.h
//     (plane_finish_kernel, one thread per cell).
// Cells whose result depends on how libstdc++'s unstable std::sort orders exact z ties
// (SURVEY B-3) are redone by a second instance that replays that sort with the whole warp
// (dev_warpsort.cuh; dev_stdsort.h is the single-thread replay it must agree with).
// Algorithmic bytes: 17 G in, 8 G member records, B * (72 + 16 F_g) out per keyframe.
#include "common.cuh"
#include "dev_plane.h"
#include "dev_stdsort.h"
#include "dev_warpsort.cuh"

namespace sb {

// A cell is a few hundred points: the work per CTA is latency-bound (dependent
// passes separated by barriers), so the CTA is kept small (2 warps, ~16 KB of
// shared memory) to have many cells resident per SM.  Larger cells spill their
// member list / fit workspace to global scratch (L2-resident).
#ifndef SLOAM_K2_THREADS
#define SLOAM_K2_THREADS 32  // one warp per cell: every barrier is warp-local, 32 cells resident per SM (64 -> 32 threads: 381 -> 335 us)
#endif
constexpr int kGThreads = SLOAM_K2_THREADS;
#ifndef SLOAM_K2_CELLS_PER_CTA
#define SLOAM_K2_CELLS_PER_CTA 2
#endif
constexpr int kCellsPerCta = SLOAM_K2_CELLS_PER_CTA;
#ifndef SLOAM_K2_SELCAP
#define SLOAM_K2_SELCAP 256
#define SLOAM_K2_QRCAP 64
#define SLOAM_K2_MINCTAS 24  // measured (1000 VLP-16 kf): 1024/192/1 -> 483 us, 256/96/20 -> 387, 256/64/20 -> 377, 32/64/28 -> 359
#endif
constexpr int kSelCap = SLOAM_K2_SELCAP;   // members kept in shared memory (else global scratch)
constexpr int kQrCap = SLOAM_K2_QRCAP;     // retained points whose fit lives in shared memory

// exclusive block scan of one int per thread (256 threads); returns total in *total
__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int *total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kGThreads / 32; ++w) {
    const int c = s_warp[w];
    if (w < warp) off += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return off + inc - v;
}

// Stable multisplit of one keyframe's ground points by polar cell: members[k] receives the
// (z key, input index) records of cell 0, then cell 1, ... each in input order, so that a
// cell's CTA reads its n_c members with one contiguous load instead of scanning all G tags.
// One CTA per keyframe walks the tags in chunks of kBinThreads; inside a chunk the rank of a
// point among the points of its cell is  (points of the cell in earlier warps) + (points of
// the cell on lower lanes, from match_any).
constexpr int kBinThreads = 1024;
constexpr int kBinWarps = kBinThreads / 32;
constexpr int kBinItems = 2;  // points per thread per chunk

__global__ void __launch_bounds__(kBinThreads)
ground_bin_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ ground,
                  const int32_t *__restrict__ ground_count, int stride,
                  const uint8_t *__restrict__ ground_cell, const int32_t *__restrict__ cell_count,
                  const int32_t *__restrict__ tile_count, SelKey *__restrict__ members) {
  extern __shared__ int s_bin[];  // [B] running offsets, [tiles + 1] tile prefix, [items][warps][B] counts
  const int B = dp->B;
  const int tiles = (dp->N + kSplitTile - 1) / kSplitTile;
  int *s_run = s_bin, *s_tpre = s_bin + B, *s_wc = s_tpre + tiles + 1;
  const int k = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const sloam_point *gk = ground + (size_t)k * stride;
  const uint8_t *ck = ground_cell + (size_t)k * stride;
  SelKey *mk = members + (size_t)k * stride;
  if (warp == 0) {  // exclusive prefix of the cell counts -> first slot of every cell
    int carry = 0;
    for (int base = 0; base < B; base += 32) {
      const int c = base + lane;
      const int v = c < B ? cell_count[(size_t)k * kMaxCells + c] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
      }
      if (c < B) s_run[c] = carry + inc - v;
      carry += __shfl_sync(kFull, inc, 31);
    }
  } else if (warp == 1 && tile_count) {
    // tile-strided cloud: the v-th ground point (input order) is slot v - pre[t] of tile t
    int carry = 0;
    for (int base = 0; base < tiles; base += 32) {
      const int t = base + lane;
      const int v = t < tiles ? tile_count[(size_t)k * tiles + t] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += u;
      }
      if (t < tiles) s_tpre[t] = carry + inc - v;
      carry += __shfl_sync(kFull, inc, 31);
    }
    if (lane == 0) s_tpre[tiles] = carry;
  }
  __syncthreads();
  const int G = tile_count ? s_tpre[tiles] : ground_count[k];
  for (int base = 0; base < G; base += kBinThreads * kBinItems) {
    for (int i = threadIdx.x; i < kBinItems * kBinWarps * B; i += kBinThreads) s_wc[i] = 0;
    __syncthreads();
    int idx[kBinItems], cel[kBinItems], rank[kBinItems];
    uint32_t zk[kBinItems];
#pragma unroll
    for (int q = 0; q < kBinItems; ++q) {  // item order: q-major, then warp, then lane = input order
      const int v = base + q * kBinThreads + threadIdx.x;
      int i = v;
      if (tile_count && v < G) {
        int lo = 0, hi = tiles - 1;  // last tile with pre[t] <= v
        while (lo < hi) {
          const int mid = (lo + hi + 1) >> 1;
          if (s_tpre[mid] <= v) lo = mid; else hi = mid - 1;
        }
        i = lo * kSplitTile + (v - s_tpre[lo]);
      }
      const int c = v < G ? (int)ck[i] : 255;
      const bool valid = c < B;
      idx[q] = i; cel[q] = valid ? c : -1;
      zk[q] = valid ? float_key(gk[i].z) : 0u;
      // lanes of this warp in the same cell (invalid lanes get a private key)
      const unsigned same = __match_any_sync(kFull, valid ? c : 256 + lane);
      rank[q] = __popc(same & ((1u << lane) - 1u));
      if (valid && rank[q] == 0) s_wc[(q * kBinWarps + warp) * B + c] = __popc(same);
    }
    __syncthreads();
    // per cell: exclusive scan of the counts over (item, warp) (one warp per cell, lane = warp
    // index), turned into absolute output slots; the running offset advances by the total
    for (int cc = warp; cc < B; cc += kBinWarps) {
      int run = s_run[cc];
      __syncwarp();
#pragma unroll
      for (int q = 0; q < kBinItems; ++q) {
        const int v = s_wc[(q * kBinWarps + lane) * B + cc];
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(kFull, inc, o);
          if (lane >= o) inc += t;
        }
        s_wc[(q * kBinWarps + lane) * B + cc] = run + inc - v;
        run += __shfl_sync(kFull, inc, 31);
      }
      if (lane == 0) s_run[cc] = run;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kBinItems; ++q)
      if (cel[q] >= 0) {
        SelKey e; e.z = zk[q]; e.j = (uint32_t)idx[q];
        mk[s_wc[(q * kBinWarps + warp) * B + cel[q]] + rank[q]] = e;
      }
    __syncthreads();
  }
}

// ---- fused pipeline: binning per tile ------------------------------------------------------
// seg_tab[k][cell][tile]: in = points of the cell in the tile (split kernel); out = first slot
// of those points in the keyframe's member array = (points of lower cells) + (points of the
// cell in lower tiles).  One CTA per keyframe, one warp per cell at a time.
__global__ void __launch_bounds__(256)
ground_offsets_kernel(const DevParams *__restrict__ dp, const int32_t *__restrict__ cell_count,
                      uint32_t *__restrict__ seg_tab) {
  __shared__ int s_start[kMaxCells];
  const int B = dp->B, tiles = (dp->N + kSplitTile - 1) / kSplitTile;
  const int k = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {  // first slot of every cell: exclusive prefix of the cell counts
    int carry = 0;
    for (int base = 0; base < B; base += 32) {
      const int c = base + lane;
      const int v = c < B ? cell_count[(size_t)k * kMaxCells + c] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
      }
      if (c < B) s_start[c] = carry + inc - v;
      carry += __shfl_sync(kFull, inc, 31);
    }
  }
  __syncthreads();
  for (int c = warp; c < B; c += 8) {
    uint32_t *row = seg_tab + ((size_t)k * B + c) * tiles;
    int carry = s_start[c];
    for (int t0 = 0; t0 < tiles; t0 += 32) {
      const int t = t0 + lane;
      const int v = t < tiles ? (int)row[t] : 0;
      int inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += u;
      }
      if (t < tiles) row[t] = (uint32_t)(carry + inc - v);
      carry += __shfl_sync(kFull, inc, 31);
    }
  }
}

// One warp per (keyframe, tile): the tile's records are in input order, so walking them 32 at
// a time and handing every cell of a chunk the next slots of its run keeps the binning stable.
// The running slots of the cells live in shared memory, one row per warp.
constexpr int kScatWarps = 8;
__global__ void __launch_bounds__(kScatWarps * 32)
ground_scatter_kernel(const DevParams *__restrict__ dp, int K, const uint2 *__restrict__ recs,
                      const uint8_t *__restrict__ ground_cell, const int32_t *__restrict__ tile_count,
                      const uint32_t *__restrict__ seg_tab, SelKey *__restrict__ members) {
  __shared__ int s_run[kScatWarps][kMaxCells];
  const int N = dp->N, B = dp->B, tiles = (N + kSplitTile - 1) / kSplitTile;
  const int lane = threadIdx.x & 31;
  const int wid = blockIdx.x * kScatWarps + (threadIdx.x >> 5);
  if (wid >= K * tiles) return;
  const int k = wid / tiles, tile = wid - k * tiles;
  const int n = tile_count[(size_t)k * tiles + tile];
  if (n == 0) return;
  int *run = s_run[threadIdx.x >> 5];
  for (int c = lane; c < B; c += 32) run[c] = (int)seg_tab[((size_t)k * B + c) * tiles + tile];
  __syncwarp();
  const size_t base = (size_t)k * N + (size_t)tile * kSplitTile;
  SelKey *mk = members + (size_t)k * N;
  const unsigned lt = (1u << lane) - 1u;
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    int cell = 255;
    uint2 rec = make_uint2(0u, 0u);
    if (i < n) { cell = ground_cell[base + i]; rec = recs[base + i]; }
    unsigned todo = __ballot_sync(kFull, cell < B);
    while (todo) {
      const int l = __ffs(todo) - 1;
      const int c0 = __shfl_sync(kFull, cell, l);
      const unsigned mm = __ballot_sync(kFull, cell == c0);
      const int r = run[c0];
      if (cell == c0) { SelKey e; e.z = rec.x; e.j = rec.y; mk[r + __popc(mm & lt)] = e; }
      __syncwarp();
      if (lane == l) run[c0] = r + __popc(mm);
      __syncwarp();
      todo &= ~mm;
    }
  }
}

// REPLAY = false: grid (cells, keyframes).  Cells whose retained set involves exact z ties
// among more than 16 points are also listed in `tied`; the REPLAY = true instance (1-D grid
// over that list) redoes them with libstdc++'s std::sort replayed by one thread, because the
// reference's sort (sloam.cpp:377-380) is not stable (SURVEY B-3) and the order of tied points
// decides which of them are kept and in which order they enter the plane fit.
// a cell's CTA is one warp: its barriers are warp barriers
#define CELL_SYNC() do { if (kGThreads == 32) __syncwarp(); else __syncthreads(); } while (0)
template <bool REPLAY>
__global__ void __launch_bounds__(REPLAY ? 32 : 32 * kCellsPerCta, REPLAY ? 16 : SLOAM_K2_MINCTAS)
ground_cells_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ ground,
                    const int32_t *__restrict__ ground_count, int stride,
                    SelKey *__restrict__ members, const int32_t *__restrict__ cell_count,
                    const sloam_pose *__restrict__ pose_est,
                    double *__restrict__ qscratch, float *__restrict__ pscratch,
                    FitRec *__restrict__ fit, sloam_cell_plane *__restrict__ cells,
                    sloam_point *__restrict__ cell_features, sloam_point *__restrict__ kept_points,
                    int32_t *__restrict__ kept_offsets, SelKey *__restrict__ members2,
                    int32_t *__restrict__ tied, int32_t *__restrict__ n_tied) {
  // the replay instance sorts whole cells with one thread: give it room for most cells in
  // shared memory (a global-memory sort is ~10x slower per access)
  constexpr int kCap = REPLAY ? 5120 : 1;
  // the selecting instance packs kCellsPerCta independent cells (warps) into a CTA: the SM holds
  // at most 32 CTAs, and with one-warp CTAs that would cap it at 32 of its 64 warp slots
  constexpr int kW = REPLAY ? 1 : kCellsPerCta;
  const int wslot = REPLAY ? 0 : (int)(threadIdx.x >> 5);
  const int tid = threadIdx.x & 31;
  __shared__ SelKey s_list[kCap];
  __shared__ int s_lq[REPLAY ? 64 : 1], s_rq[REPLAY ? 64 : 1];  // stopper queues of warp_partition
  // the selecting instance keeps only what its passes re-read in shared memory: the z keys of
  // the cell (the radix passes) and the r kept records (rank sort); the member records
  // themselves are streamed from global memory once, by the compaction
  constexpr int kZCap = REPLAY ? 1 : 2 * kSelCap, kKeepCap = REPLAY ? 1 : 128;
  __shared__ uint32_t s_z_[kW][kZCap];
  __shared__ SelKey s_keep_[kW][kKeepCap];
  __shared__ SelKey s_tmp_[kW][kKeepCap];  // rank-sort target
  __shared__ int s_hist_[kW][256];
  __shared__ int s_warp_[kW][1];
  __shared__ int s_misc_[kW][8];
  uint32_t *s_z = s_z_[wslot];
  SelKey *s_keep = s_keep_[wslot], *s_tmp = s_tmp_[wslot];
  int *s_hist = s_hist_[wslot], *s_warp = s_warp_[wslot], *s_misc = s_misc_[wslot];

  const int B = dp->B, Fg = dp->p.numGroundFeatures;
  auto body = [&](const int k, const int cell) {
  const int n_c = cell_count[(size_t)k * kMaxCells + cell];
  const int lane = tid, warp = 0;
  const sloam_point *gk = ground + (size_t)k * stride;
  sloam_cell_plane *out = cells + (size_t)k * B + cell;
  sloam_point *fout = cell_features + ((size_t)k * B + cell) * Fg;

  // retained count r (sloam.cpp:366-384) of every cell -> this cell's offsets
  const double retainNum = 1.0 / dp->p.groundRetainThresh;
  auto kept_of = [&](int n) {
    if (n > 0 && retainNum < (double)n) { const int b = (int)((double)n / retainNum); return b < n ? b : n; }
    return n;
  };
  if (warp == 0) {  // prefix over the preceding cells, one warp, loads in parallel
    int off_all = 0, off_kept = 0;
    for (int c = lane; c < cell; c += 32) {
      const int n = cell_count[(size_t)k * kMaxCells + c];
      off_all += n; off_kept += kept_of(n);
    }
    off_all = warp_sum(off_all);
    off_kept = warp_sum(off_kept);
    if (lane == 0) {
      s_misc[0] = off_all; s_misc[1] = off_kept;
      if (kept_offsets) {
        kept_offsets[(size_t)k * (B + 1) + cell] = off_kept;
        if (cell == B - 1) kept_offsets[(size_t)k * (B + 1) + B] = off_kept + kept_of(n_c);
      }
    }
  }
  CELL_SYNC();
  const int off_all = s_misc[0], off_kept = s_misc[1];
  const bool do_sort = n_c > 0 && retainNum < (double)n_c;
  const int r = kept_of(n_c);

  if (n_c == 0 || r < Fg || r < 3) {
    // Plane::Plane: features.size() < numGroundFeatures -> invalid (plane.cpp:7-10);
    // n < 3 is out-of-range in the reference's ThinU access: declared invalid.
    // (kept point lists of such cells are still emitted below when requested.)
    if (tid == 0) {
      sloam_cell_plane c;
      for (int i = 0; i < 4; ++i) c.model.plane[i] = 0.0;
      for (int i = 0; i < 3; ++i) c.model.centroid[i] = 0.0;
      c.n_cell = n_c; c.n_kept = r; c.is_valid = 0; c.accepted = 0;
      *out = c;
      fit[(size_t)k * B + cell].valid = 0;
    }
    for (int f = tid; f < Fg; f += kGThreads) fout[f] = sloam_point{0.f, 0.f, 0.f, 0.f};
    if (kept_points == nullptr || n_c == 0) return;
  }

  // ---- the members (z key, input index), in input order: contiguous in `members` ----
  SelKey *src = members + (size_t)k * stride + off_all;
  SelKey *list = src;
  uint32_t zlo = 0xFFFFFFFFu, zhi = 0u;  // key range of this lane's members (selecting instance)
  const bool zc = !REPLAY && n_c <= kZCap;  // z keys cached in shared memory
  if (REPLAY && n_c <= kCap) {
    list = s_list;
    for (int i = tid; i < n_c; i += kGThreads) s_list[i] = src[i];
  } else if (zc) {
    for (int i = tid; i < n_c; i += kGThreads) { const uint32_t z = src[i].z; s_z[i] = z; zlo = min(zlo, z); zhi = max(zhi, z); }
  }
  CELL_SYNC();
  // key range of the cell (selecting instance): the radix select below works on key - kmin,
  // whose significant bits are few (a cell's z values span centimetres to metres), instead of
  // on the raw key, whose top two bytes are nearly the same for every member
  uint32_t kmin = 0u, krange = 0xFFFFFFFFu;
  if (!REPLAY && do_sort) {
    uint32_t lo = zlo, hi = zhi;  // gathered while the keys were copied to shared memory
    if (!zc)
      for (int i = tid; i < n_c; i += kGThreads) { const uint32_t z = list[i].z; lo = min(lo, z); hi = max(hi, z); }
    if (kGThreads == 32) {
      lo = __reduce_min_sync(kFull, lo); hi = __reduce_max_sync(kFull, hi);
    } else {
      for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(kFull, lo, o)); hi = max(hi, __shfl_xor_sync(kFull, hi, o)); }
    }
    kmin = lo; krange = hi - lo;
  }

  // ---- select the r lowest (z, j) keys ----
  // kept records: in place in shared memory, or in the second member array for oversized
  // cells (the first one stays intact for a possible replay)
  SelKey *keep = list;
  if (do_sort) {
    if (REPLAY) { if (n_c > kCap) keep = members2 + (size_t)k * stride + off_all; }
    else keep = r <= kKeepCap ? s_keep : members2 + (size_t)k * stride + off_all;
  }
  if (REPLAY && do_sort) {
    // std::sort(cell, p1.z < p2.z) replayed on the whole cell, then the first r are kept
    // (only the first r positions of the result are needed: sort_prefix; the z field is turned
    // back into the float's bits first so that the comparator is a plain float compare)
    for (int i = tid; i < n_c; i += kGThreads) {
      SelKey e = list[i];
      e.z = __float_as_uint(key_to_float(e.z));
      keep[i] = e;
    }
    CELL_SYNC();
    // the sorted prefix goes straight to the fit kernel's input (members2, the cell's region);
    // when keep already IS that region (oversized cell) the ranks are written through scratch
    {
      SelKey *kout = members2 + (size_t)k * stride + off_all;
      SelKey *dst = keep != kout ? kout : reinterpret_cast<SelKey *>(qscratch + ((size_t)k * stride + off_all) * 3);
      warp_sort_prefix(keep, n_c, r, dst, s_lq, s_rq);
      if (dst != kout) {
        for (int i = tid; i < r; i += 32) keep[i] = dst[i];
        CELL_SYNC();
      } else {
        keep = kout;
      }
    }
  } else if (do_sort) {
    // MSD radix select on (z key - kmin) for the r-th smallest (rank r-1): digits of up to 8
    // bits from the highest significant bit of the range downwards, so the first pass already
    // spreads the members over the bins (few shared-memory atomic conflicts) and a range of
    // 2^22 keys -- 1 m of spread at |z| = 2 m -- takes three passes
    uint32_t prefix = 0, pmask = 0;
    int want = r - 1;  // 0-based rank among members matching the prefix
    for (int top = 32 - __clz(krange | 1u); top > 0;) {
      const int shift = top > 8 ? top - 8 : 0;
      const uint32_t dmask = (1u << (top - shift)) - 1u;
      top = shift;
      for (int b = tid; b < 256; b += kGThreads) s_hist[b] = 0;
      CELL_SYNC();
      for (int i = tid; i < n_c; i += kGThreads) {
        const uint32_t z = (zc ? s_z[i] : list[i].z) - kmin;
        if ((z & pmask) == prefix) atomicAdd(&s_hist[(z >> shift) & dmask], 1);
      }
      CELL_SYNC();
      if (warp == 0) {  // bin that holds rank `want`: 8 bins per lane + a warp scan
        int loc[8], s = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { loc[q] = s_hist[lane * 8 + q]; s += loc[q]; }
        int inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(kFull, inc, o);
          if (lane >= o) inc += t;
        }
        const int excl = inc - s;
        if (excl <= want && want < inc) {  // exactly one lane
          int acc = excl, d = 0;
          for (; d < 7; ++d) { if (acc + loc[d] > want) break; acc += loc[d]; }
          s_misc[2] = lane * 8 + d; s_misc[3] = want - acc; s_misc[4] = s_hist[lane * 8 + d];
        }
      }
      CELL_SYNC();
      prefix |= (uint32_t)s_misc[2] << shift;
      pmask |= dmask << shift;
      want = s_misc[3];
      const int in_bin = s_misc[4];
      CELL_SYNC();
      if (kGThreads == 32 && top > 0 && in_bin <= 32) {
        // At most a warp's worth of members left in the bin that holds the r-th smallest: finish
        // among them directly instead of by further passes over the whole cell.  The candidates
        // are collected in list order (one pass, ballots), every lane ranks its candidate among
        // them with ties by list position, and the lane with rank `want` holds the pivot key.
        int nc = 0;
        for (int base = 0; base < n_c; base += 32) {
          const int i = base + lane;
          uint32_t z = 0;
          bool hit = false;
          if (i < n_c) { z = (zc ? s_z[i] : list[i].z) - kmin; hit = (z & pmask) == prefix; }
          const unsigned bh = __ballot_sync(kFull, hit);
          if (hit) s_hist[nc + __popc(bh & ((1u << lane) - 1u))] = (int)z;  // the histogram is consumed
          nc += __popc(bh);
        }
        __syncwarp();
        const uint32_t ck = lane < nc ? (uint32_t)s_hist[lane] : 0xFFFFFFFFu;
        int rank = 0;
        for (int m = 0; m < nc; ++m) {
          const uint32_t km = __shfl_sync(kFull, ck, m);
          rank += (km < ck || (km == ck && m < lane)) ? 1 : 0;
        }
        const unsigned bp = __ballot_sync(kFull, lane < nc && rank == want);  // exactly one lane
        const uint32_t pk = __shfl_sync(kFull, ck, __ffs(bp) - 1);
        const int lower = __popc(__ballot_sync(kFull, lane < nc && ck < pk));
        prefix = pk;           // the full (offset) key of the r-th smallest
        want -= lower;         // its position among the members with that key, in list order
        top = 0;
        __syncwarp();
      }
    }
    const uint32_t pivot = kmin + prefix;  // z key of the r-th smallest
    const int tie_quota = want + 1;  // members with z == pivot to keep, lowest j first
    // order-preserving compaction (the list is in j order, so ties come lowest j first)
    int kept = 0, ties_seen = 0;
    for (int base = 0; base < n_c; base += kGThreads) {
      const int i = base + tid;
      SelKey e = {0, 0};
      bool lt = false, eq = false;
      if (i < n_c) { e = list[i]; lt = e.z < pivot; eq = e.z == pivot; }
      int tot_eq, tot_keep, eq_before, pos;
      bool take;
      if (kGThreads == 32) {  // one warp: ballots instead of shared-memory scans and barriers
        const unsigned lt_mask = (1u << lane) - 1u;
        const unsigned be = __ballot_sync(kFull, eq);
        eq_before = ties_seen + __popc(be & lt_mask);
        take = lt || (eq && eq_before < tie_quota);
        const unsigned bt = __ballot_sync(kFull, take);
        pos = kept + __popc(bt & lt_mask);
        tot_eq = __popc(be); tot_keep = __popc(bt);
        __syncwarp();  // every lane has read its list entry before any slot is overwritten
      } else {
        eq_before = ties_seen + block_excl_scan(eq ? 1 : 0, s_warp, &tot_eq);
        take = lt || (eq && eq_before < tie_quota);
        pos = kept + block_excl_scan(take ? 1 : 0, s_warp, &tot_keep);
      }
      if (take) keep[pos] = e;  // pos <= i, and every slot < base is already consumed
      kept += tot_keep; ties_seen += tot_eq;
      CELL_SYNC();
    }
    // ---- sort the r kept keys: rank sort in place via shared ranks ----
    // (r is a few hundred; O(r^2 / threads))
    SelKey *tmp = r <= kKeepCap ? s_tmp : reinterpret_cast<SelKey *>(qscratch + ((size_t)k * stride + off_all) * 3);
    // Exact ties among the kept points or across the cut (more members equal to the pivot than
    // were taken): with more than 16 points in the cell the reference's result depends on
    // libstdc++'s unstable sort -> the cell is listed for the replay instance (whose outputs
    // replace everything this instance writes for the cell).  Detected inside the rank loop.
    bool tie = ties_seen > tie_quota;
    for (int i = tid; i < r; i += kGThreads) {
      const SelKey e = keep[i];
      int rank = 0, same = 0;
      for (int j = 0; j < r; ++j) {
        rank += sel_less(keep[j], e);
        same += keep[j].z == e.z;
      }
      tie |= same > 1;  // (keys equal <=> floats equal, -0 / +0 aside)
      tmp[rank] = e;
    }
    int any_tie;
    if (kGThreads == 32) { any_tie = __any_sync(kFull, tie); __syncwarp(); }
    else any_tie = __syncthreads_or(tie ? 1 : 0);
    for (int i = tid; i < r; i += kGThreads) keep[i] = tmp[i];
    CELL_SYNC();
    if (any_tie && n_c > 16 && tied != nullptr && tid == 0) tied[atomicAdd(n_tied, 1)] = (k << 8) | cell;
  }

  if (kept_points) {
    sloam_point *kp = kept_points + (size_t)k * stride + off_kept;
    for (int i = tid; i < r; i += kGThreads) st_point(kp + i, ld_point(gk + keep[i].j));
  }
  if (n_c == 0 || r < Fg || r < 3) return;
  // ---- hand the r retained records, in order, to ground_fit_kernel (members2, the cell's region)
  {
    SelKey *kout = members2 + (size_t)k * stride + off_all;
    if (keep != kout)
      for (int i = tid; i < r; i += kGThreads) kout[i] = keep[i];
    if (tid == 0) {
      FitRec *rec = fit + (size_t)k * B + cell;
      rec->n_cell = n_c; rec->n_kept = r; rec->off_all = off_all; rec->valid = 1;
    }
  }
  };  // body
  if (!REPLAY) {
    const int cell = (int)blockIdx.x * kCellsPerCta + wslot;
    if (cell < B) body((int)blockIdx.y, cell);
  } else {
    const int total = *n_tied;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
      body(tied[item] >> 8, tied[item] & 0xFF);
      CELL_SYNC();
    }
  }
}

#undef CELL_SYNC

// ---- plane fit of every cell with enough retained points: one warp per (keyframe, cell) ----
// Reads the r retained records of the cell (members2, written by the select / replay
// instances of ground_cells_kernel), gathers the points, and runs Plane::computeModel up to the
// QR preconditioner; plane_finish_kernel finishes the 3 x 3 SVD.  Kept apart from the selection
// so that the selection runs without the fit's registers and shared memory (twice the warps
// per SM), and so that cells redone by the replay instance are fitted here like all others.
#ifndef SLOAM_FIT_MIN
#define SLOAM_FIT_MIN 32  // 64 registers: fit + finish 156 -> 137 us per 1024 OS1-64 keyframes (24 CTAs: 80 registers)
#endif
__global__ void __launch_bounds__(32, SLOAM_FIT_MIN)
ground_fit_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ ground, int stride,
                  const SelKey *__restrict__ members2, double *__restrict__ qscratch, float *__restrict__ pscratch,
                  FitRec *__restrict__ fit, sloam_point *__restrict__ cell_features) {
  __shared__ double s_qr[3 * kQrCap];
  __shared__ float s_pts[3 * kQrCap];
  __shared__ double s_red[4];
  const int B = dp->B, Fg = dp->p.numGroundFeatures;
  const int k = blockIdx.y, cell = blockIdx.x;
  const int lane = threadIdx.x;
  FitRec *rec = fit + (size_t)k * B + cell;
  if (!rec->valid) return;
  const int n_c = rec->n_cell, r = rec->n_kept, off_all = rec->off_all;
  const sloam_point *gk = ground + (size_t)k * stride;
  const SelKey *keep = members2 + (size_t)k * stride + off_all;
  sloam_point *fout = cell_features + ((size_t)k * B + cell) * Fg;
  (void)n_c;
#define CELL_SYNC() __syncwarp()
  constexpr int kGThreads = 32;
  const int warp = 0;
  double *s_hist = s_red;  // scratch of the scale reduction (one warp: one entry)
  // ---- Plane::computeModel on the r retained points, in order ----
  const int n = r;
  // stage the retained points (x | y | z planes) so the serial float sum reads shared memory
  float *P = n <= kQrCap ? s_pts : pscratch + ((size_t)k * stride + off_all) * 3;
  for (int i = threadIdx.x; i < n; i += kGThreads) {
    const sloam_point p = ld_point(gk + keep[i].j);
    P[i] = p.x; P[n + i] = p.y; P[2 * n + i] = p.z;
  }
  CELL_SYNC();
  if (threadIdx.x < 3) {  // computeCentroid: float32 sequential sums (utils.h:14-28)
    const float *a = P + threadIdx.x * n;
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += a[i];
    s_red[threadIdx.x] = (double)(float)((double)acc / (double)n);
  }
  CELL_SYNC();
  const float cxf = (float)s_red[0], cyf = (float)s_red[1], czf = (float)s_red[2];
  // adjoint matrix A^T (n x 3), column-major in Q; entries (double)(float diff) (plane.cpp:106-108)
  double *Q = n <= kQrCap ? s_qr : qscratch + ((size_t)k * stride + off_all) * 3;
  double lmax = 0.0;
  for (int i = threadIdx.x; i < n; i += kGThreads) {
    const double dx = (double)(P[i] - cxf), dy = (double)(P[n + i] - cyf), dz = (double)(P[2 * n + i] - czf);
    Q[i] = dx; Q[n + i] = dy; Q[2 * n + i] = dz;
    lmax = fmax(lmax, fmax(fabs(dx), fmax(fabs(dy), fabs(dz))));
  }
  // scale = max |coeff| (JacobiSVD::compute)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lmax = fmax(lmax, __shfl_xor_sync(kFull, lmax, o));
  if (lane == 0) s_hist[3] = lmax;
  CELL_SYNC();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int w = 0; w < kGThreads / 32; ++w) m = fmax(m, s_hist[3]);
    s_red[3] = (m == 0.0) ? 1.0 : m;
  }
  CELL_SYNC();
  const double scale = s_red[3];
  for (int i = threadIdx.x; i < 3 * n; i += kGThreads) Q[i] = Q[i] / scale;
  CELL_SYNC();

  if (warp != 0) return;  // the tiny QR + Jacobi runs on one warp
  double Wm[9], U[9];
  if (n > 3) {
    // ColPivHouseholderQR of the n x 3 matrix (Eigen/src/QR/ColPivHouseholderQR.h)
    auto col_sq = [&](int c, int from) {
      double s = 0.0;
      for (int rr = from + lane; rr < n; rr += 32) s += Q[c * n + rr] * Q[c * n + rr];
      return warp_sum_d(s);
    };
    double nUpd[3], nDir[3];
    for (int c = 0; c < 3; ++c) nUpd[c] = nDir[c] = sqrt(col_sq(c, 0));
    const double downdate_thr = sqrt(DBL_EPSILON);
    int colmap[3] = {0, 1, 2};  // physical column of logical column
    int transp[3];
    for (int kk = 0; kk < 3; ++kk) {
      int big = kk;
      for (int j = kk + 1; j < 3; ++j) if (nUpd[j] > nUpd[big]) big = j;
      transp[kk] = big;
      if (big != kk) {
        const int tc = colmap[kk]; colmap[kk] = colmap[big]; colmap[big] = tc;
        const double tu = nUpd[kk]; nUpd[kk] = nUpd[big]; nUpd[big] = tu;
        const double td = nDir[kk]; nDir[kk] = nDir[big]; nDir[big] = td;
      }
      double *ck2 = Q + colmap[kk] * n;
      double tailSq = 0.0;
      for (int rr = kk + 1 + lane; rr < n; rr += 32) tailSq += ck2[rr] * ck2[rr];
      tailSq = warp_sum_d(tailSq);
      const double c0 = ck2[kk];
      double tau, beta;
      __syncwarp();
      if (tailSq <= DBL_MIN) {
        tau = 0.0; beta = c0;
        for (int rr = kk + 1 + lane; rr < n; rr += 32) ck2[rr] = 0.0;
      } else {
        beta = sqrt(c0 * c0 + tailSq);
        if (c0 >= 0.0) beta = -beta;
        for (int rr = kk + 1 + lane; rr < n; rr += 32) ck2[rr] = ck2[rr] / (c0 - beta);
        tau = (beta - c0) / beta;
      }
      if (lane == 0) ck2[kk] = beta;
      __syncwarp();
      if (tau != 0.0) {
        for (int j = kk + 1; j < 3; ++j) {
          double *cj = Q + colmap[j] * n;
          double tmp = 0.0;
          for (int rr = kk + 1 + lane; rr < n; rr += 32) tmp += ck2[rr] * cj[rr];
          tmp = warp_sum_d(tmp);
          tmp += cj[kk];
          __syncwarp();
          if (lane == 0) cj[kk] -= tau * tmp;
          for (int rr = kk + 1 + lane; rr < n; rr += 32) cj[rr] -= tau * ck2[rr] * tmp;
          __syncwarp();
        }
      }
      for (int j = kk + 1; j < 3; ++j) {  // LAPACK-style norm downdate
        if (nUpd[j] != 0.0) {
          double temp = fabs(Q[colmap[j] * n + kk]) / nUpd[j];
          temp = (1.0 + temp) * (1.0 - temp);
          temp = temp < 0.0 ? 0.0 : temp;
          const double r2 = nUpd[j] / nDir[j];
          const double temp2 = temp * r2 * r2;
          if (temp2 <= downdate_thr) {
            nDir[j] = sqrt(col_sq(colmap[j], kk + 1));
            nUpd[j] = nDir[j];
          } else {
            nUpd[j] *= sqrt(temp);
          }
        }
      }
    }
    int perm[3] = {0, 1, 2};
    for (int kk = 0; kk < 3; ++kk) { const int t = perm[kk]; perm[kk] = perm[transp[kk]]; perm[transp[kk]] = t; }
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        U[i * 3 + j] = (i == perm[j]) ? 1.0 : 0.0;
        Wm[i * 3 + j] = (j <= i) ? Q[colmap[i] * n + j] : 0.0;  // R^T
      }
  } else {  // n == 3: square, no preconditioner
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Wm[i * 3 + j] = Q[i * n + j];
        U[i * 3 + j] = (i == j) ? 1.0 : 0.0;
      }
  }
  // hand the 3x3 problem to plane_finish_kernel (one thread per cell: the Jacobi sweeps and
  // the acceptance test are scalar fp64 code, 32 cells per warp instead of one)
  if (lane < 9) { rec->W[lane] = Wm[lane]; rec->U[lane] = U[lane]; }
  if (lane == 0) { rec->c[0] = cxf; rec->c[1] = cyf; rec->c[2] = czf; }
  for (int f = lane; f < Fg; f += 32) st_point(fout + f, ld_point(gk + keep[f].j));  // features.resize(numGroundFeatures)
#undef CELL_SYNC
}

// Steps 2-4 of JacobiSVD on the QR-preconditioned 3x3, plane assembly (plane.cpp:115-127)
// and the acceptance test (sloam.cpp:401-409): one thread per (keyframe, cell).
__global__ void plane_finish_kernel(const DevParams *__restrict__ dp, int K, const FitRec *__restrict__ fit,
                                    const sloam_pose *__restrict__ pose_est,
                                    sloam_cell_plane *__restrict__ cells) {
  const int B = dp->B;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= K * B) return;
  const FitRec rec = fit[g];
  if (!rec.valid) return;  // the invalid record was written by ground_cells_kernel
  double Wm[9], U[9], nrm[3];
  for (int i = 0; i < 9; ++i) { Wm[i] = rec.W[i]; U[i] = rec.U[i]; }
  jacobi_svd3_last_u(Wm, U, nrm);
  sloam_cell_plane c;
  const double cx = (double)rec.c[0], cy = (double)rec.c[1], cz = (double)rec.c[2];
  c.model.plane[0] = nrm[0]; c.model.plane[1] = nrm[1]; c.model.plane[2] = nrm[2];
  c.model.plane[3] = -(nrm[0] * cx + nrm[1] * cy + nrm[2] * cz);  // plane.cpp:117
  c.model.centroid[0] = cx; c.model.centroid[1] = cy; c.model.centroid[2] = cz;
  c.n_cell = rec.n_cell; c.n_kept = rec.n_kept; c.is_valid = 1;
  c.accepted = plane_accept(pose_est[g / B], c.model.plane, c.model.centroid, dp->p.ground_angle_tol) ? 1 : 0;
  cells[g] = c;
}

// accepted planes of each keyframe, compact, in (radius bin, theta bin) order
__global__ void planes_compact_kernel(const DevParams *__restrict__ dp, const sloam_cell_plane *__restrict__ cells,
                                      sloam_plane *__restrict__ planes_acc, int32_t *__restrict__ acc_cell,
                                      int32_t *__restrict__ n_acc) {
  const int B = dp->B;
  const int k = blockIdx.x, lane = threadIdx.x;
  int cnt = 0;
  for (int base = 0; base < B; base += 32) {
    const int c = base + lane;
    const bool a = c < B && cells[(size_t)k * B + c].accepted;
    const unsigned b = __ballot_sync(kFull, a);
    if (a) {
      const int pos = cnt + __popc(b & ((1u << lane) - 1u));
      planes_acc[(size_t)k * B + pos] = cells[(size_t)k * B + c].model;
      acc_cell[(size_t)k * B + pos] = c;
    }
    cnt += __popc(b);
  }
  if (lane == 0) n_acc[k] = cnt;
}

int launch_ground_tag(sloam_ctx *c, int K, const sloam_point *ground, const int32_t *ground_count, int stride);

int launch_planes_compact(sloam_ctx *c, int K, const sloam_cell_plane *cells) {
  planes_compact_kernel<<<K, 32, 0, c->stream>>>(c->dp, cells, c->ws.planes_acc, c->ws.planes_acc_cell,
                                                 c->ws.n_planes_acc);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

int launch_ground_planes(sloam_ctx *c, int K, const sloam_point *ground, const int32_t *ground_count,
                         int stride, const sloam_pose *pose_est, sloam_cell_plane *cells,
                         sloam_point *cell_features, sloam_point *kept_points, int32_t *kept_offsets,
                         bool strided) {
  // strided (fused pipeline): `ground` is the INPUT cloud (the records index it), the split kernel
  // left records + tags + per-tile cell counts; else `ground` is a ground cloud tagged by
  // ground_tag_kernel and the records index that cloud
  Workspace &w = c->ws;
  dim3 grid((unsigned)c->hp.B, (unsigned)K);
  const int k1_tiles = (c->hp.N + kSplitTile - 1) / kSplitTile;
  SelKey *members = reinterpret_cast<SelKey *>(w.gscratch3);
  PROF_BEGIN(c, P_GROUND_BIN);
  if (strided) {
    ground_offsets_kernel<<<K, 256, 0, c->stream>>>(c->dp, w.cell_count, w.seg_tab);
    SB_LAUNCH_CHECK(c);
    ground_scatter_kernel<<<(K * k1_tiles + kScatWarps - 1) / kScatWarps, kScatWarps * 32, 0, c->stream>>>(
        c->dp, K, reinterpret_cast<const uint2 *>(w.gscratch), w.ground_cell, w.tile_count, w.seg_tab, members);
    SB_LAUNCH_CHECK(c);
  } else {
    const size_t bin_smem = sizeof(int) * ((size_t)(kBinItems * kBinWarps + 1) * c->hp.B + k1_tiles + 1);
    static size_t bin_set[64] = {};  // per device, only grows (see vertex_wide_kernel)
    if (bin_smem > 48 * 1024 && bin_smem > bin_set[c->device & 63]) {
      SB_CUDA(c, cudaFuncSetAttribute(ground_bin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bin_smem));
      bin_set[c->device & 63] = bin_smem;
    }
    ground_bin_kernel<<<K, kBinThreads, bin_smem, c->stream>>>(c->dp, ground, ground_count, stride, w.ground_cell,
                                                               w.cell_count, nullptr, members);
    SB_LAUNCH_CHECK(c);
  }
  PROF_END(c, P_GROUND_BIN);
  if (c->zero_valid & 8u) c->zero_valid &= ~8u;
  else SB_CUDA(c, cudaMemsetAsync(w.n_tied_cells, 0, sizeof(int32_t), c->stream));
  PROF_BEGIN(c, P_GROUND_CELLS);
  const dim3 sgrid((unsigned)((c->hp.B + kCellsPerCta - 1) / kCellsPerCta), (unsigned)K);
  ground_cells_kernel<false><<<sgrid, 32 * kCellsPerCta, 0, c->stream>>>(
      c->dp, ground, ground_count, stride, members, w.cell_count, pose_est,
      w.qscratch, w.pscratch, w.fit_rec, cells, cell_features, kept_points, kept_offsets,
      reinterpret_cast<SelKey *>(w.gscratch2), w.tied_cells, w.n_tied_cells);
  PROF_END(c, P_GROUND_CELLS);
  SB_LAUNCH_CHECK(c);
  // cells with exact z ties again, with the std::sort replay (usually an empty list)
  PROF_BEGIN(c, P_GROUND_REPLAY);
  ground_cells_kernel<true><<<c->sm_count * 4, kGThreads, 0, c->stream>>>(
      c->dp, ground, ground_count, stride, members, w.cell_count, pose_est,
      w.qscratch, w.pscratch, w.fit_rec, cells, cell_features, kept_points, kept_offsets,
      reinterpret_cast<SelKey *>(w.gscratch2), w.tied_cells, w.n_tied_cells);
  PROF_END(c, P_GROUND_REPLAY);
  SB_LAUNCH_CHECK(c);
  PROF_BEGIN(c, P_PLANE_FIT);
  ground_fit_kernel<<<grid, 32, 0, c->stream>>>(c->dp, ground, stride, reinterpret_cast<const SelKey *>(w.gscratch2), w.qscratch,
                                                w.pscratch, w.fit_rec, cell_features);
  SB_LAUNCH_CHECK(c);
  plane_finish_kernel<<<(K * c->hp.B + 127) / 128, 128, 0, c->stream>>>(c->dp, K, w.fit_rec, pose_est, cells);
  SB_LAUNCH_CHECK(c);
  const int rc_pc = launch_planes_compact(c, K, cells);
  PROF_END(c, P_PLANE_FIT);
  return rc_pc;
}

}  // namespace sb

using namespace sb;

extern "C" int sloam_b200_ground_planes_dev(sloam_ctx *c, int K, const sloam_point *ground,
                                            const int32_t *ground_count, int ground_stride,
                                            const sloam_pose *pose_est, sloam_cell_plane *cells,
                                            sloam_point *cell_features, sloam_point *kept_points,
                                            int32_t *kept_offsets) {
  if (!c || K <= 0 || K > c->max_k || !ground || !ground_count || !pose_est || !cells || !cell_features ||
      ground_stride <= 0 || ground_stride > c->hp.N)
    return set_err(c, SLOAM_E_INVALID, "ground_planes: bad arguments (ground_stride must be <= H*W)");
  int rc = launch_ground_tag(c, K, ground, ground_count, ground_stride);
  if (rc != SLOAM_OK) return rc;
  return launch_ground_planes(c, K, ground, ground_count, ground_stride, pose_est, cells, cell_features,
                              kept_points, kept_offsets, false);
}
