// k3_trellis.cu -- stage a6 + a7: tree instance detection.
//
// Replaces Instance::computeGraph (sloam/src/segmentation/trellis.cpp:134-140):
//   findClusters -> PCL OrganizedConnectedComponentSegmentation with an
//   EuclideanClusterComparator(1.0 m) over the organized H x W tree cloud
//   (trellis.cpp:15-29), i.e. connected components of the grid graph whose
//   edges are (left, up) neighbour pairs with ||pa - pb|| < threshold; the
//   PCL label of a component is its rank by the raster index of its first
//   pixel.
//   findTrees -> per cluster with > 80 points, per scan line from the bottom
//   row up, a TreeVertex from the cluster's points of that row
//   (trellis.cpp:104-132, computeVertexProperties :63-102).
//
// GPU formulation: run-based connected components.
//   cc_rows_kernel   one pass over the tree pixels (32-pixel words of the tree bit mask, one
//                    warp per non-zero word): three bit planes per pixel -- valid (finite
//                    point), start (not connected to its left neighbour: the pixel starts a
//                    row RUN), up (connected to the pixel above).
//   cc_label_kernel  one CTA per keyframe, everything in shared memory: run ids by a prefix
//                    count of the start bits (raster order), lock-free union-find over RUNS
//                    (a few thousand per scan instead of 65 k pixels; the smaller id wins, so a
//                    root is the first run of its component and the PCL label is the number of
//                    roots before it), component sizes, the big components in label order
//                    (= tree order, deterministic when there are more than max_trees), and
//                    one work item per (component, row) with its member count, binned by size.
//   vertex_kernel    sub-warp groups: 4 / 2 / 1 items per warp for rows of <= 8 / 16 / 32
//                    members (one member per lane, order statistics by counting), a warp loop
//                    for 33..128, one CTA per wider row.
//   tree_compact     > 16 / <= 56 vertex rules, bottom row first.
#include <type_traits>

#include "common.cuh"
#include "dev_stdsort.h"
#include "dev_warpsort.cuh"

namespace sb {

#ifndef SLOAM_VTX_WARPS
#define SLOAM_VTX_WARPS 8
#endif
constexpr int kVtxWarps = SLOAM_VTX_WARPS;
constexpr int kVtxCap = 128;   // members per (cluster,row) handled by the warp path
// cc_label_kernel runs with 512 threads per keyframe, or with 1024 when the image is so large
// that shared memory, not threads, limits the CTAs per SM (OS1-128: 749 -> 471 us per 512
// keyframes; OS1-64 is 3 % slower with 1024, so it keeps 512 and four CTAs per SM)
constexpr int kLblThreads = 512, kLblThreadsMax = 1024;
constexpr unsigned kSlotMask = 0x7FFu;  // slot code (slot + 1, 0 = none) in the low 11 bits of packed words

// work item of the vertex stage: the members of one component in one row
struct __align__(16) VItem {
  int32_t pix0;       // first pixel of the first run
  int32_t first_run;  // run id (within the keyframe) of the first run
  uint16_t n;         // members
  uint16_t span;      // last run id - first run id (0: one run, members are contiguous pixels)
  uint16_t slot;      // big-component slot (tree order)
  uint16_t row;
};
// class lists: [0] n <= 8, [1] n <= 16, [2] n <= 32, [3] n <= kVtxCap, [4] wider, [5] exact z ties (replay)
constexpr int kClsWide = 4, kClsTied = 5;

__device__ __forceinline__ int uf_find(const int32_t *parent, int x) {
  int p = *((volatile const int32_t *)&parent[x]);
  while (p != x) {
    x = p;
    p = *((volatile const int32_t *)&parent[x]);
  }
  return x;
}
__device__ __forceinline__ void uf_union(int32_t *parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&parent[a], b);  // attach the larger root under the smaller
    if (old == a) return;
    a = old;
  }
}

// tree_bits: one bit per pixel, [K][Nw] words (Nw = ceil(N / 32)).  Bit set = the pixel may
// hold a tree point (the split kernel sets it for mask == 255; for caller-supplied clouds it
// is isfinite(x)).  Pixels whose bit is clear are never read by the kernels below -- in the
// fused pipeline their tree point is not even written.
__global__ void tree_bits_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ tree,
                                 uint32_t *__restrict__ bits) {
  const int N = dp->N, Nw = (N + 31) >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  const bool v = i < N && isfinite(tree[(size_t)k * N + i].x);
  const unsigned b = __ballot_sync(kFull, v);
  if ((threadIdx.x & 31) == 0 && i < N) bits[(size_t)k * Nw + (i >> 5)] = b;
}

// dense organized cloud from the sparse one: NaN points (intensity 0) where the bit is clear
__global__ void tree_fill_kernel(const DevParams *__restrict__ dp, const uint32_t *__restrict__ bits,
                                 sloam_point *__restrict__ tree) {
  const int N = dp->N, Nw = (N + 31) >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k = blockIdx.y;
  if (i >= N) return;
  if (!((bits[(size_t)k * Nw + (i >> 5)] >> (i & 31)) & 1u)) {
    const float qnan = __int_as_float(0x7fc00000);
    st_point(tree + (size_t)k * N + i, sloam_point{qnan, qnan, qnan, 0.f});
  }
}

// ---- 1. bit planes: valid / run start / connected upwards -------------------
// A warp takes 32 consecutive words of the batch's tree bit mask (1024 pixels) with one
// coalesced load and then works on the SET bits only, 32 tree pixels per step, one per lane:
// the j-th set bit of the chunk is found through the prefix popcounts of the words (binary
// search with shuffles, then the bit inside the word).  A forest scan has 10-20 % tree pixels
// scattered over most words, so walking words with one lane per pixel left two thirds of the
// lanes idle.  The result planes live in per-word accumulators in shared memory and are written
// as planes[k][word] = (valid, start, up, 0) for the non-zero words at the end.
#ifndef SLOAM_CCROWS_MIN
#define SLOAM_CCROWS_MIN 8
#endif
constexpr int kRowsWarps = 8;
// position of the (n + 1)-th set bit of a non-zero word (n < popc(word))
__device__ __forceinline__ int nth_set_bit(uint32_t word, int n) {
  int pos = 0;
#pragma unroll
  for (int width = 16; width >= 1; width >>= 1) {
    const int c = __popc((word >> pos) & ((1u << width) - 1u));
    if (n >= c) { n -= c; pos += width; }
  }
  return pos;
}
__global__ void __launch_bounds__(kRowsWarps * 32, SLOAM_CCROWS_MIN)
cc_rows_kernel(const DevParams *__restrict__ dp, int K, const sloam_point *__restrict__ tree,
               const uint32_t *__restrict__ bits, uint4 *__restrict__ planes) {
  __shared__ uint32_t s_acc[kRowsWarps][3][32];
  const int N = dp->N, W = dp->p.img_w, Nw = (N + 31) >> 5;
  const float cut = dp->cluster_sq_cut;  // dist < cluster_dist_thresh  <=>  squared dist < cut
  const unsigned magic_w = dp->magic_w;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t(*acc)[32] = s_acc[warp];
  const long long total_words = (long long)K * Nw;
  const long long warp0 = (long long)blockIdx.x * kRowsWarps + warp;
  const long long nwarps = (long long)gridDim.x * kRowsWarps;
  for (long long base = warp0 * 32; base < total_words; base += nwarps * 32) {
    const long long gw = base + lane;
    const uint32_t my_word = gw < total_words ? bits[gw] : 0u;
    if (__ballot_sync(kFull, my_word != 0u) == 0u) continue;
    // keyframe / word index of this lane's word, prefix popcounts of the chunk
    const int my_k = (int)(gw / Nw), my_wi = (int)(gw - (long long)my_k * Nw);
    const int cnt = __popc(my_word);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    const int pre = inc - cnt;
    const int total = __shfl_sync(kFull, inc, 31);
    acc[0][lane] = my_word; acc[1][lane] = 0u; acc[2][lane] = my_word;
    __syncwarp();
    // point of the last pixel of the previous step (left neighbour of this step's first pixel)
    float carry_x = 0.f, carry_y = 0.f, carry_z = 0.f;
    int carry_pix = -2, carry_k = -1;
    for (int j0 = 0; j0 < total; j0 += 32) {
      const int j = j0 + lane;
      const bool act = j < total;
      // word holding the j-th set bit: the last word whose prefix is <= j
      int w = 0;
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1) {
        const int cand = w + step;
        const int pc = __shfl_sync(kFull, pre, cand & 31);
        if (cand < 32 && pc <= j) w = cand;
      }
      const uint32_t word = __shfl_sync(kFull, my_word, w);
      const int pre_w = __shfl_sync(kFull, pre, w);
      const int k = __shfl_sync(kFull, my_k, w), wi = __shfl_sync(kFull, my_wi, w);
      const int b = act ? nth_set_bit(word, j - pre_w) : 0;
      const int i = wi * 32 + b;
      const bool bit = act && i < N;
      const uint32_t *bk = bits + (size_t)k * Nw;
      const size_t g = (size_t)k * N + i;
      const int row = fast_div_w(i, magic_w), col = i - row * W;
      // own point and the point above are loaded together (guarded by their bits only)
      const int u = i - W;
      const bool bit_u = bit && row > 0 && ((bk[u >> 5] >> (u & 31)) & 1u);
      sloam_point p{0.f, 0.f, 0.f, 0.f}, qu = p;
      if (bit) p = ld_point(tree + g);
      if (bit_u) qu = ld_point(tree + g - W);
      // PCL skips a pixel iff !isfinite(x); EuclideanClusterComparator::compare is
      // dist < threshold in float (NaN compares false)
      const bool valid = bit && isfinite(p.x);
      // left neighbour = the previous set bit when it is pixel i - 1 of the same keyframe
      float lx = __shfl_up_sync(kFull, p.x, 1), ly = __shfl_up_sync(kFull, p.y, 1), lz = __shfl_up_sync(kFull, p.z, 1);
      int lpix = __shfl_up_sync(kFull, i, 1), lk = __shfl_up_sync(kFull, k, 1);
      if (lane == 0) { lx = carry_x; ly = carry_y; lz = carry_z; lpix = carry_pix; lk = carry_k; }
      bool bit_l = bit && col > 0 && lpix == i - 1 && lk == k;
      if (bit && col > 0 && !bit_l && j == 0 && wi > 0 && b == 0 && (bk[wi - 1] >> 31)) {
        // first set bit of the chunk: its left neighbour belongs to the previous chunk
        const sloam_point q = ld_point(tree + g - 1);
        lx = q.x; ly = q.y; lz = q.z;
        bit_l = true;
      }
      const bool left_ok = valid && bit_l && sqnorm3f(p.x - lx, p.y - ly, p.z - lz) < cut;
      const bool up_ok = valid && bit_u && sqnorm3f(p.x - qu.x, p.y - qu.y, p.z - qu.z) < cut;
      // OR the three bits of the pixels of each word (contiguous lanes) into the accumulators
      // The planes start from the tree bits themselves (valid and up: nearly all of them stay
      // set; start: clear) and only the exceptions touch shared memory: a non-finite point, a
      // pixel that starts a run, a pixel without an upward link.
      if (bit) {
        if (!valid) {
          atomicAnd(&acc[0][w], ~(1u << b));
          atomicAnd(&acc[2][w], ~(1u << b));
        } else {
          if (!left_ok) atomicOr(&acc[1][w], 1u << b);
          if (!up_ok) atomicAnd(&acc[2][w], ~(1u << b));
        }
      }
      carry_x = __shfl_sync(kFull, p.x, 31); carry_y = __shfl_sync(kFull, p.y, 31); carry_z = __shfl_sync(kFull, p.z, 31);
      carry_pix = __shfl_sync(kFull, bit ? i : -2, 31); carry_k = __shfl_sync(kFull, k, 31);
      __syncwarp();
    }
    if (my_word != 0u) planes[gw] = make_uint4(acc[0][lane], acc[1][lane], acc[2][lane], 0u);
    __syncwarp();
  }
}

// ---- 2. per keyframe: runs, union-find, labels, work items -------------------
// exclusive scan of one int per thread over the CTA; *total = sum.  s_tmp: kLblWarps + 1 ints
__device__ __forceinline__ int cta_scan_excl(int v, int *s_tmp, int *total) {
  const int kLblWarps = (int)(blockDim.x >> 5);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_tmp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int t = lane < kLblWarps ? s_tmp[lane] : 0;
    int ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int q = __shfl_up_sync(kFull, ti, o);
      if (lane >= o) ti += q;
    }
    if (lane < kLblWarps) s_tmp[lane] = ti - t;
    if (lane == kLblWarps - 1) s_tmp[kLblWarps] = ti;
  }
  __syncthreads();
  const int off = s_tmp[warp];
  *total = s_tmp[kLblWarps];
  __syncthreads();  // s_tmp may be reused by the next call
  return off + inc - v;
}

// run id of valid pixel i: number of run starts at or before it, minus one
__device__ __forceinline__ int run_of(const uint32_t *s_start, const int32_t *s_wbase, int i) {
  const int w = i >> 5, b = i & 31;
  return s_wbase[w] + __popc(s_start[w] & ((2u << b) - 1u)) - 1;
}
// 32 bits of a bit plane starting at bit position pos (pos may be negative or run past the end)
__device__ __forceinline__ uint32_t bits_at(const uint32_t *s_plane, int Nw, int pos) {
  if (pos <= -32) return 0u;
  const int w = pos >> 5;  // floor division also for negative pos
  const int b = pos & 31;
  const uint32_t lo = (w >= 0 && w < Nw) ? s_plane[w] : 0u;
  const uint32_t hi = (w + 1 >= 0 && w + 1 < Nw) ? s_plane[w + 1] : 0u;
  return b ? ((lo >> b) | (hi << (32 - b))) : lo;
}

__global__ void __launch_bounds__(kLblThreadsMax)
cc_label_kernel(const DevParams *__restrict__ dp, const uint32_t *__restrict__ bits,
                const uint4 *__restrict__ planes, int Rs, int32_t *__restrict__ g_par,
                uint32_t *__restrict__ g_siz, int32_t *__restrict__ g_len, int32_t *__restrict__ run_pix0,
                int32_t *__restrict__ run_info, int32_t *__restrict__ run_label, int32_t *__restrict__ wbase_out,
                int32_t *__restrict__ n_roots, int32_t *__restrict__ n_big, int32_t *__restrict__ big_rank,
                int32_t *__restrict__ kf_flags, uint32_t *__restrict__ slot_rows, VItem *__restrict__ items,
                int32_t *__restrict__ item_pool, int32_t *__restrict__ vlists, long long list_cap,
                int32_t *__restrict__ n_lists) {
  extern __shared__ __align__(16) unsigned char lbl_smem[];
  const int N = dp->N, W = dp->p.img_w, H = dp->p.img_h, T = dp->p.max_trees, Nw = (N + 31) >> 5;
  const int Hw = (H + 31) >> 5;
  const int min_pts = dp->p.min_cluster_points, min_vtx = dp->p.min_vertex_points;
  const int item_shift = dp->vw_row_bits + dp->vw_slot_bits;
  const int k = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t *s_start = reinterpret_cast<uint32_t *>(lbl_smem);          // [Nw]
  int32_t *s_wbase = reinterpret_cast<int32_t *>(s_start + Nw);        // [Nw]
  uint32_t *s_rows = reinterpret_cast<uint32_t *>(s_wbase + Nw);       // [T * Hw]
  int32_t *s_runs = reinterpret_cast<int32_t *>(s_rows + (size_t)T * Hw);  // 3 x [Rs]
  __shared__ int s_tmp[kLblThreadsMax / 32 + 1];
  const int kLblThreads = (int)blockDim.x, kLblWarps = kLblThreads >> 5;  // 512 or 1024 (run_cc)
  __shared__ int s_nitems, s_pool;
  __shared__ int s_ccnt[kClsWide + 1], s_cbase[kClsWide + 1];
  const uint32_t *bk = bits + (size_t)k * Nw;
  const uint4 *pk = planes + (size_t)k * Nw;

  // -- run ids: prefix count of the start bits, in raster order
  // (every thread takes a block of consecutive words and scans it serially: one CTA-wide scan,
  // three barriers, whatever the image size)
  int R = 0;
  {
    const int per = (Nw + kLblThreads - 1) / kLblThreads;
    const int w_lo = min((int)threadIdx.x * per, Nw), w_hi = min(w_lo + per, Nw);
    int mine = 0;
    for (int w = w_lo; w < w_hi; ++w) {
      const uint32_t s = bk[w] != 0u ? pk[w].y : 0u;
      s_start[w] = s;
      mine += __popc(s);
    }
    int run = cta_scan_excl(mine, s_tmp, &R);
    for (int w = w_lo; w < w_hi; ++w) {
      s_wbase[w] = run;
      run += __popc(s_start[w]);
    }
  }
  __syncthreads();
  if (wbase_out)
    for (int w = threadIdx.x; w < Nw; w += kLblThreads) wbase_out[(size_t)k * Nw + w] = s_wbase[w];
  for (int i = threadIdx.x; i < T * Hw; i += kLblThreads) s_rows[i] = 0u;
  if (threadIdx.x == 0) { s_nitems = 0; s_pool = 0; }
  if (threadIdx.x <= kClsWide) s_ccnt[threadIdx.x] = 0;
  // per-run arrays: shared memory, or the keyframe's slice of global scratch for scans with
  // more runs than fit (correct, slower; a forest scan has a few thousand runs)
  int32_t *par = s_runs;
  uint32_t *siz = reinterpret_cast<uint32_t *>(s_runs + Rs);
  int32_t *len = s_runs + 2 * (size_t)Rs;
  if (R > Rs) { par = g_par + (size_t)k * N; siz = g_siz + (size_t)k * N; len = g_len + (size_t)k * N; }
  for (int r = threadIdx.x; r < R; r += kLblThreads) { par[r] = r; siz[r] = 0u; len[r] = 0; }
  __syncthreads();

  // -- run lengths, first pixels, and the unions of vertically connected runs
  int32_t *pix0_k = run_pix0 + (size_t)k * N;
  for (int w = threadIdx.x; w < Nw; w += kLblThreads) {
    if (bk[w] == 0u) continue;
    const uint4 pl = pk[w];
    const uint32_t v = pl.x, s = pl.y, u = pl.z;
    const int base = s_wbase[w];
    // segments of runs inside this word: one per start bit, plus a run continuing from the
    // previous word when bit 0 is valid without being a start
    uint32_t begins = s | (v & 1u);
    while (begins) {
      const int b = __ffs(begins) - 1;
      begins &= begins - 1;
      const int rid = base + __popc(s & ((2u << b) - 1u)) - 1;
      int cnt = __ffs(~(v >> b)) - 1;  // consecutive valid bits from b (zeros are shifted in at the top)
      if (cnt < 0) cnt = 32;           // b == 0 and the whole word is valid
      const uint32_t later = s & ~((2u << b) - 1u);
      if (later) cnt = min(cnt, __ffs(later) - 1 - b);
      atomicAdd(&len[rid], cnt);
      if ((s >> b) & 1u) pix0_k[rid] = w * 32 + b;
    }
    // a union is needed where an up link is not implied by the previous pixel's: the pixel or
    // the one above starts a run, or the previous pixel has no up link (bit 0: always)
    const uint32_t su = bits_at(s_start, Nw, w * 32 - W);
    uint32_t flagged = u & (s | su | ~(u << 1));
    while (flagged) {
      const int b = __ffs(flagged) - 1;
      flagged &= flagged - 1;
      const int i = w * 32 + b;
      uf_union(par, base + __popc(s & ((2u << b) - 1u)) - 1, run_of(s_start, s_wbase, i - W));
    }
  }
  __syncthreads();
  // -- flatten, component sizes at the root
  // (finds and stores in separate phases, so that no find ever reads an entry another thread is
  // rewriting -- harmless here, parents only move towards the root, but a reported race)
  for (int r0 = 0; r0 < R; r0 += kLblThreads) {
    const int r = r0 + threadIdx.x;
    const int root = r < R ? uf_find(par, r) : 0;
    __syncthreads();
    if (r < R) par[r] = root;
    __syncthreads();
  }
  for (int r = threadIdx.x; r < R; r += kLblThreads) atomicAdd(&siz[par[r]], (uint32_t)len[r]);
  __syncthreads();
  // -- PCL label of every root (number of roots before it) and slot of every big component
  // (number of big roots before it = tree order); siz[root] := label << 11 | slot code
  // (both counts in one scan: roots in the low 20 bits, big roots above; a block of consecutive
  // runs per thread)
  int n_root = 0, n_bigc = 0;
  {
    const int per = (R + kLblThreads - 1) / kLblThreads;
    const int r_lo = min((int)threadIdx.x * per, R), r_hi = min(r_lo + per, R);
    int mine = 0;
    for (int r = r_lo; r < r_hi; ++r)
      if (par[r] == r) mine += 1 + (((int)siz[r] > min_pts) ? (1 << 20) : 0);
    int tot;
    int run = cta_scan_excl(mine, s_tmp, &tot);
    n_root = tot & 0xFFFFF;
    n_bigc = tot >> 20;
    for (int r = r_lo; r < r_hi; ++r) {
      if (par[r] != r) continue;
      const bool is_big = (int)siz[r] > min_pts;
      const int label = run & 0xFFFFF, slot = run >> 20;
      const bool keep = is_big && slot < T;  // more big components than max_trees: the first T in label order
      siz[r] = ((uint32_t)label << 11) | (keep ? (uint32_t)(slot + 1) : 0u);
      if (keep) big_rank[(size_t)k * T + slot] = label;
      run += 1 + (is_big ? (1 << 20) : 0);
    }
  }
  if (threadIdx.x == 0) {
    n_roots[k] = n_root;
    n_big[k] = min(n_bigc, T);
    if (n_bigc > T) atomicOr(&kf_flags[k], 1);
  }
  __syncthreads();
  // -- per-run tables: length and slot code (shared + global), label (global)
  int32_t *info_k = run_info + (size_t)k * N, *label_k = run_label + (size_t)k * N;
  for (int r = threadIdx.x; r < R; r += kLblThreads) {
    const uint32_t rootw = siz[par[r]];
    const int packed = (len[r] << 11) | (int)(rootw & kSlotMask);
    len[r] = packed;
    info_k[r] = packed;
    label_k[r] = (int)(rootw >> 11);
  }
  __syncthreads();
  // -- work items: one per (big component, row).  One warp per row; a lane per run of the row
  // scans the row's runs: it is the head of its (slot, row) group when no earlier run of the
  // row has the same slot, and then sums the members of the later ones.
  VItem *items_k = items + ((size_t)k << item_shift);
  for (int row = warp; row < H; row += kLblWarps) {
    const int p0 = row * W, p1 = p0 + W;
    const int rb = s_wbase[p0 >> 5] + __popc(s_start[p0 >> 5] & ((1u << (p0 & 31)) - 1u));
    const int re = p1 >= N ? R : s_wbase[p1 >> 5] + __popc(s_start[p1 >> 5] & ((1u << (p1 & 31)) - 1u));
    for (int c0 = rb; c0 < re; c0 += 32) {
      const int r = c0 + lane;
      int sc = 0, n_tot = 0, last = r;
      bool head = false;
      if (r < re) {
        const int inf = len[r];
        sc = inf & (int)kSlotMask;
        n_tot = inf >> 11;
        head = sc != 0;
      }
      if (__any_sync(kFull, head)) {
        // inside the chunk: lanes with the same slot find each other with match_any; the lowest
        // one is the head and sums the others' lengths (a handful of shuffles)
        const int own = n_tot;
        const unsigned grp = __match_any_sync(kFull, sc != 0 ? sc : -(lane + 1));
        head = head && lane == __ffs(grp) - 1;
        const int gsize = __popc(grp);
        const int gmax = __reduce_max_sync(kFull, gsize);
        unsigned rest = grp & ~(1u << lane);
        for (int it = 1; it < gmax; ++it) {
          const int src = rest ? __ffs(rest) - 1 : lane;
          const int v = __shfl_sync(kFull, own, src);
          if (rest) { n_tot += v; last = max(last, c0 + src); rest &= rest - 1; }
        }
        // rows with more than 32 runs: the same slot may also sit in another chunk
        if (re - rb > 32) {
          for (int j = rb; j < re; ++j) {
            if (j >= c0 && j < c0 + 32) continue;
            const int infj = len[j];
            if ((infj & (int)kSlotMask) == sc) {
              if (j < r) head = false;
              else { n_tot += infj >> 11; last = max(last, j); }
            }
          }
        }
      }
      int cls = -1;
      if (head && n_tot > min_vtx)
        cls = n_tot <= 8 ? 0 : (n_tot <= 16 ? 1 : (n_tot <= 32 ? 2 : (n_tot <= kVtxCap ? 3 : kClsWide)));
      const unsigned hb = __ballot_sync(kFull, cls >= 0);
      if (hb == 0u) continue;
      // item index and first slot in the keyframe's vertex-point pool (n slots per item: every
      // member may be kept), both from shared-memory counters: one atomic per chunk
      int psum = cls >= 0 ? n_tot : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, psum, o);
        if (lane >= o) psum += t;
      }
      int ibase = 0, pbase = 0;
      if (lane == 31) { ibase = atomicAdd(&s_nitems, __popc(hb)); pbase = atomicAdd(&s_pool, psum); }
      ibase = __shfl_sync(kFull, ibase, 31);
      pbase = __shfl_sync(kFull, pbase, 31);
      const int idx = ibase + __popc(hb & ((1u << lane) - 1u));
      if (cls >= 0) {
        VItem it;
        it.pix0 = pix0_k[r];
        it.first_run = r;
        it.n = (uint16_t)n_tot;
        it.span = (uint16_t)(last - r);
        it.slot = (uint16_t)(sc - 1);
        it.row = (uint16_t)row;
        items_k[idx] = it;
        item_pool[((size_t)k << item_shift) + idx] = pbase + psum - n_tot;
        atomicOr(&s_rows[(sc - 1) * Hw + (row >> 5)], 1u << (row & 31));
        atomicAdd(&s_ccnt[cls], 1);
      }
    }
  }
  __syncthreads();
  // -- the items of the keyframe join the batch's class lists: one global atomic per class
  if (threadIdx.x <= kClsWide) {
    const int cnt = s_ccnt[threadIdx.x];
    s_cbase[threadIdx.x] = cnt ? atomicAdd(&n_lists[threadIdx.x], cnt) : 0;
    s_ccnt[threadIdx.x] = 0;
  }
  __syncthreads();
  const int nitems = s_nitems;
  for (int i0 = 0; i0 < nitems; i0 += kLblThreads) {
    const int idx = i0 + threadIdx.x;
    int cls = -1;
    if (idx < nitems) {
      const int n_tot = items_k[idx].n;
      cls = n_tot <= 8 ? 0 : (n_tot <= 16 ? 1 : (n_tot <= 32 ? 2 : (n_tot <= kVtxCap ? 3 : kClsWide)));
    }
#pragma unroll
    for (int c = 0; c <= kClsWide; ++c) {
      const unsigned cb = __ballot_sync(kFull, cls == c);
      if (cb == 0u) continue;
      const int leader = __ffs(cb) - 1;
      int lb = 0;
      if (lane == leader) lb = atomicAdd(&s_ccnt[c], __popc(cb));
      lb = __shfl_sync(kFull, lb, leader);
      if (cls == c)
        vlists[(size_t)c * list_cap + s_cbase[c] + lb + __popc(cb & ((1u << lane) - 1u))] = (k << item_shift) | idx;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < T * Hw; i += kLblThreads) slot_rows[(size_t)k * T * Hw + i] = s_rows[i];
}

// ---- 3. one vertex per (cluster, row) ----------------------------------------
struct VtxSmem {
  float x[kVtxCap], y[kVtxCap], z[kVtxCap], w[kVtxCap];
  int16_t col[kVtxCap];
  int16_t order[kVtxCap];
};
// small rows: one member per lane, the groups of a warp side by side
struct __align__(16) GrpSmem {
  float x[32], y[32], z[32], w[32];
  int8_t ord[32], ord2[32];
};

// lexicographic (z, y, x, col): the order the reference's three std::sort calls
// produce (SURVEY B-3); col is unique so this is a total order
__device__ __forceinline__ bool key_less(float za, float ya, float xa, int ca, float zb, float yb,
                                         float xb, int cb) {
  if (za != zb) return za < zb;
  if (ya != yb) return ya < yb;
  if (xa != xb) return xa < xb;
  return ca < cb;
}

// The rare tie path: computeVertexProperties' std::sort by x, by y, by z on members that
// start in column order, replayed by ONE thread (the algorithm is sequential).
template <class Idx>
__device__ __noinline__ void replay_three_sorts(Idx *order, int n, const float *x, const float *y, const float *z) {
  for (int m = 0; m < n; ++m) order[m] = (Idx)m;
  const float *keys[3] = {x, y, z};
  for (int a = 0; a < 3; ++a) {  // one instance of the sort code
    StdSort<Idx> srt(order, keys[a]);
    srt.sort(n);
  }
}
// Same with (key, index) packed in one 64-bit word per element (one shared-memory load per
// access instead of two dependent ones); `pack` is scratch for n words.  A sort whose keys
// are all distinct has one possible result, so the replay only has to start at the first axis
// that has ties after it: `first` = 0 (x: start from column order), 1 (y: `init` = the members
// in x order) or 2 (z: `init` = the members in y order).
__device__ __noinline__ void replay_sorts_packed(int16_t *order, unsigned long long *pack, int n, int first,
                                                 const int16_t *init, const float *x, const float *y,
                                                 const float *z) {
  const float *keys[3] = {x, y, z};
  for (int m = 0; m < n; ++m) {
    const unsigned idx = first == 0 ? (unsigned)m : (unsigned)init[m];
    pack[m] = ((unsigned long long)__float_as_uint(keys[first][idx]) << 32) | idx;
  }
  for (int a = first; a < 3; ++a) {
    if (a > first)
      for (int m = 0; m < n; ++m) {
        const unsigned idx = (unsigned)(pack[m] & 0xFFFFFFFFull);
        pack[m] = ((unsigned long long)__float_as_uint(keys[a][idx]) << 32) | idx;
      }
    StdSortPacked srt{pack, PackedLess{}};
    srt.sort(n);
  }
  for (int m = 0; m < n; ++m) order[m] = (int16_t)(pack[m] & 0xFFFFFFFFull);
}

// The same three sorts with the whole warp (dev_warpsort.cuh): records (key bits, member) in
// `a`, sorted into `b` and back; lq / rq are the 64-entry stopper queues of the partition.
static __device__ void replay_sorts_warp(int16_t *order, SelKey *a, SelKey *b, int *lq, int *rq, int n, int first,
                                  const int16_t *init, const float *x, const float *y, const float *z) {
  const int lane = threadIdx.x & 31;
  const float *keys[3] = {x, y, z};
  for (int m = lane; m < n; m += 32) {
    const unsigned idx = first == 0 ? (unsigned)m : (unsigned)init[m];
    a[m] = SelKey{__float_as_uint(keys[first][idx]), idx};
  }
  __syncwarp();
  for (int ax = first; ax < 3; ++ax) {
    if (ax > first) {
      for (int m = lane; m < n; m += 32) a[m].z = __float_as_uint(keys[ax][a[m].j]);
      __syncwarp();
    }
    warp_sort_prefix(a, n, n, b, lq, rq);  // std::sort(a, a + n): every position is needed
    SelKey *t = a; a = b; b = t;
  }
  for (int m = lane; m < n; m += 32) order[m] = (int16_t)a[m].j;
  __syncwarp();
}

// pixel of member m of an item whose members are spread over several runs (rare: a row of a
// trunk split by a gap): walk the runs of the row between the first and the last one
__device__ __forceinline__ int member_pixel(const VItem &it, int m, const int32_t *__restrict__ pix0_k,
                                            const int32_t *__restrict__ info_k) {
  int acc = 0;
  for (int j = it.first_run; j <= it.first_run + (int)it.span; ++j) {
    const int inf = info_k[j];
    if ((inf & (int)kSlotMask) != (int)it.slot + 1) continue;
    const int ln = inf >> 11;
    if (m < acc + ln) return pix0_k[j] + (m - acc);
    acc += ln;
  }
  return it.pix0;  // not reached: n is the sum of the matching runs' lengths
}

// G lanes per item (G = 8, 16, 32): the warp handles 32 / G items side by side.  Returns (per
// lane) true when the lane's item has exact z ties among more than 16 members and must be
// redone by the replay kernel (nothing was written for it).
template <int G>
__device__ __forceinline__ bool vertex_group(const DevParams *dp, GrpSmem &s, bool active, const VItem &it, int k,
                                             const sloam_point *__restrict__ tree, const int32_t *__restrict__ run_pix0,
                                             const int32_t *__restrict__ run_info, sloam_vertex *__restrict__ slot_vertices,
                                             sloam_point *__restrict__ pool, int base) {
  const int N = dp->N, H = dp->p.img_h, T = dp->p.max_trees;
  const int lane = threadIdx.x & 31, gl = lane & (G - 1), g0 = lane & ~(G - 1);
  const unsigned gmask = G == 32 ? kFull : (((1u << G) - 1u) << g0);
  const unsigned lt = (1u << lane) - 1u;
  const int n = active ? (int)it.n : 0;
  const float inf = __int_as_float(0x7f800000);
  // -- members: one per lane, in column order (trellis.cpp:113-118)
  sloam_point p{inf, inf, inf, 0.f};
  if (gl < n) {
    const int pixel = it.span == 0 ? it.pix0 + gl
                                   : member_pixel(it, gl, run_pix0 + (size_t)k * N, run_info + (size_t)k * N);
    p = ld_point(tree + (size_t)k * N + pixel);
  }
  __syncwarp();  // the previous item's readers are done
  s.x[lane] = p.x; s.y[lane] = p.y; s.z[lane] = p.z; s.w[lane] = p.intensity;  // lanes >= n pad with +inf
  __syncwarp();
  // -- order statistics by counting: member m counts the members strictly below it on each
  // axis; without ties that is its rank (float counts: FSET + FADD per comparison)
  const int middle = n >> 1;  // (int)(n / 2.0), trellis.cpp:66
  float fx = 0.f, fy = 0.f, fz = 0.f;
  {
    const float4 *X4 = reinterpret_cast<const float4 *>(s.x + g0), *Y4 = reinterpret_cast<const float4 *>(s.y + g0),
                 *Z4 = reinterpret_cast<const float4 *>(s.z + g0);
    const int n4 = (n + 3) >> 2;
#pragma unroll
    for (int j4 = 0; j4 < G / 4; ++j4) {
      if (j4 < n4) {
        const float4 xv = X4[j4], yv = Y4[j4], zv = Z4[j4];
        fx += (flt(xv.x, p.x) + flt(xv.y, p.x)) + (flt(xv.z, p.x) + flt(xv.w, p.x));
        fy += (flt(yv.x, p.y) + flt(yv.y, p.y)) + (flt(yv.z, p.y) + flt(yv.w, p.y));
        fz += (flt(zv.x, p.z) + flt(zv.y, p.z)) + (flt(zv.z, p.z) + flt(zv.w, p.z));
      }
    }
  }
  const int lx = (int)fx, ly = (int)fy, lz = (int)fz;
  const bool mem = gl < n;
  // z order.  Without ties the counts of an item's n members are a permutation of 0 .. n-1; two
  // members with the same count are an exact z tie.  One REDUX over the warp ORs the bit
  // (item's field + count) of every member: an item is tie-free iff its field has n bits set.
  // Items with a tie do not write the order here (their members would collide on a slot).
  const unsigned occ = __reduce_or_sync(kFull, mem ? (1u << (g0 + lz)) : 0u);
  const unsigned field = G == 32 ? occ : ((occ >> g0) & ((1u << G) - 1u));
  const bool ztie = __popc(field) != n;
  if (mem && !ztie) s.ord[g0 + lz] = (int8_t)gl;
  unsigned bx = __ballot_sync(kFull, mem && lx == middle) & gmask;
  unsigned by = __ballot_sync(kFull, mem && ly == middle) & gmask;
  unsigned bz = __ballot_sync(kFull, mem && lz == middle) & gmask;
  __syncwarp();
  // a tie group straddling a median leaves no member with count == middle: stable ranks
  // (value, then index) decide, like a stable sort would (the median VALUE is what matters)
  if (__any_sync(kFull, active && (bx == 0u || by == 0u || bz == 0u))) {
    int rx = lx, ry = ly, rz = lz;
    for (int j = 0; j < G; ++j) {
      if (j < n && j < gl) {
        rx += s.x[g0 + j] == p.x; ry += s.y[g0 + j] == p.y; rz += s.z[g0 + j] == p.z;
      }
    }
    const unsigned cx = __ballot_sync(kFull, mem && rx == middle) & gmask;
    const unsigned cy = __ballot_sync(kFull, mem && ry == middle) & gmask;
    const unsigned cz = __ballot_sync(kFull, mem && rz == middle) & gmask;
    if (bx == 0u) bx = cx;
    if (by == 0u) by = cy;
    if (bz == 0u) bz = cz;
  }
  const float med0 = __shfl_sync(kFull, p.x, bx ? __ffs(bx) - 1 : lane);
  const float med1 = __shfl_sync(kFull, p.y, by ? __ffs(by) - 1 : lane);
  const float med2 = __shfl_sync(kFull, p.z, bz ? __ffs(bz) - 1 : lane);
  bool redo = false;
  if (__any_sync(kFull, ztie)) {
    // Exact z ties: the order of the tied points is whatever the reference's three std::sort
    // calls (by x, by y, by z; trellis.cpp:71-82) leave behind.  Up to 16 points libstdc++
    // sorts by insertion (stable), so the result is the lexicographic (z, y, x, column) order
    // (members are in column order, so the column comparison is the member index).  Beyond
    // that its introsort is not stable: the item goes to the replay kernel.
    if (ztie && n > 16) redo = true;
    __syncwarp();
    if (ztie && n <= 16 && mem) {
      int rk = 0;
      for (int j = 0; j < n; ++j) rk += key_less(s.z[g0 + j], s.y[g0 + j], s.x[g0 + j], j, p.z, p.y, p.x, gl);
      s.ord[g0 + rk] = (int8_t)gl;
    }
    __syncwarp();
  }
  if (redo) active = false;
  // -- keep points within max_dist_to_centroid of the median, in z order (trellis.cpp:89-93)
  const float cut = dp->centroid_sq_cut;  // dist < max_dist_to_centroid  <=>  squared dist < cut
  bool keep = false;
  int m = 0;
  if (active && mem) {
    m = s.ord[g0 + gl];
    keep = sqnorm3f(s.x[g0 + m] - med0, s.y[g0 + m] - med1, s.z[g0 + m] - med2) < cut;
  }
  const unsigned kb = __ballot_sync(kFull, keep) & gmask;
  const int kept = __popc(kb);
  if (keep) s.ord2[g0 + __popc(kb & lt)] = (int8_t)m;
  __syncwarp();
  if (active && kept > 1 && gl < kept) {
    const int q = s.ord2[g0 + gl];
    sloam_point o; o.x = s.x[g0 + q]; o.y = s.y[g0 + q]; o.z = s.z[g0 + q]; o.intensity = s.w[g0 + q];
    st_point(pool + (size_t)k * N + base + gl, o);
  }
  if (active && gl == 0) {
    float radius = 0.f;
    int n_points = 0, point_begin = 0, is_valid = 0;
    if (kept > 1) {  // trellis.cpp:95-100
      const int a = s.ord2[g0], b = s.ord2[g0 + kept - 1];
      radius = dist3f(s.x[g0 + a], s.y[g0 + a], s.z[g0 + a], s.x[g0 + b], s.y[g0 + b], s.z[g0 + b]);
      n_points = kept; point_begin = base; is_valid = 1;
    }
    float4 *dst = reinterpret_cast<float4 *>(slot_vertices + ((size_t)k * T + it.slot) * H + it.row);
    dst[0] = make_float4(med0, med1, med2, radius);
    dst[1] = make_float4(__int_as_float(n_points), __int_as_float(point_begin), __int_as_float((int)it.row),
                         __int_as_float(is_valid));
  }
  return redo;
}

// REPLAY = false: returns true when the item has exact z ties among more than 16 points and
// must be redone by the REPLAY = true instance (nothing was written)
template <bool REPLAY>
__device__ bool build_vertex(const DevParams *dp, VtxSmem &s, unsigned long long *pack, int16_t *perm, int n,
                             int row, sloam_vertex *out, sloam_point *pool, int base,
                             SelKey *pack2 = nullptr, int *lq = nullptr, int *rq = nullptr) {
  const int lane = threadIdx.x & 31;
  const int middle = (int)(n / 2.0);  // trellis.cpp:66
  // Order statistics by counting: member m counts the members strictly below it on each
  // axis.  Without ties that count is the rank: the member whose count equals `middle` holds
  // the median, and the z counts are the z order.  Exact ties are rare (float coordinates):
  // a z tie is detected by counting equal members and handled by the full-key pass below; a
  // tie group straddling the median leaves no member with count == middle, which is caught
  // after the loop and redone with the stable (value, index) ranks.
  // The member arrays are padded to a multiple of 4 with +inf and read as float4.
  float med[3] = {0.f, 0.f, 0.f};
  unsigned any_ztie = 0, any_xtie = 0, any_ytie = 0, found = 0;
  const int n4 = (n + 3) & ~3;
  if (lane < n4 - n) {
    const float inf = __int_as_float(0x7f800000);
    s.x[n + lane] = inf; s.y[n + lane] = inf; s.z[n + lane] = inf;
  }
  __syncwarp();
  const float4 *X4 = reinterpret_cast<const float4 *>(s.x), *Y4 = reinterpret_cast<const float4 *>(s.y),
               *Z4 = reinterpret_cast<const float4 *>(s.z);
  for (int m = lane; m < ((n + 31) & ~31); m += 32) {
    int lx = 0, ly = 0, lz = 0, ez = 0, ex = 0, ey = 0;
    float xm = 0.f, ym = 0.f, zm = 0.f;
    if (m < n) {
      xm = s.x[m]; ym = s.y[m]; zm = s.z[m];
      float fx = 0.f, fy = 0.f, fz = 0.f, fe = 0.f;
      for (int j4 = 0; j4 < (n4 >> 2); ++j4) {
        const float4 xv = X4[j4], yv = Y4[j4], zv = Z4[j4];
        if (REPLAY) {
          ex += (xv.x == xm) + (xv.y == xm) + (xv.z == xm) + (xv.w == xm);
          ey += (yv.x == ym) + (yv.y == ym) + (yv.z == ym) + (yv.w == ym);
        }
        // counts accumulate as floats (exact: n <= kVtxCap): one FSET (1.0f / 0.0f) on the ALU
        // pipe + one FADD on the FMA pipe per comparison, instead of compare + add + predicated
        // move (two ALU + one FMA) for an integer count
        fx += (flt(xv.x, xm) + flt(xv.y, xm)) + (flt(xv.z, xm) + flt(xv.w, xm));
        fy += (flt(yv.x, ym) + flt(yv.y, ym)) + (flt(yv.z, ym) + flt(yv.w, ym));
        fz += (flt(zv.x, zm) + flt(zv.y, zm)) + (flt(zv.z, zm) + flt(zv.w, zm));
        fe += (feq(zv.x, zm) + feq(zv.y, zm)) + (feq(zv.z, zm) + feq(zv.w, zm));
      }
      lx = (int)fx; ly = (int)fy; lz = (int)fz; ez = (int)fe;
    }
    any_ztie |= __ballot_sync(kFull, m < n && ez > 1);
    if (REPLAY) {
      any_xtie |= __ballot_sync(kFull, m < n && ex > 1);
      any_ytie |= __ballot_sync(kFull, m < n && ey > 1);
      // x / y order, used only when that axis is tie-free (tied members would collide on a slot)
      if (m < n && ex == 1) perm[lx] = (int16_t)m;
      if (m < n && ey == 1) perm[kVtxCap + ly] = (int16_t)m;
    }
    const unsigned bx = __ballot_sync(kFull, m < n && lx == middle);
    const unsigned by = __ballot_sync(kFull, m < n && ly == middle);
    const unsigned bz = __ballot_sync(kFull, m < n && lz == middle);
    if (bx) { med[0] = __shfl_sync(kFull, xm, __ffs(bx) - 1); found |= 1u; }
    if (by) { med[1] = __shfl_sync(kFull, ym, __ffs(by) - 1); found |= 2u; }
    if (bz) { med[2] = __shfl_sync(kFull, zm, __ffs(bz) - 1); found |= 4u; }
    __syncwarp();
    // final order when z has no ties (the common case); a member with a tied z skips the
    // store (two of them would hit one slot) -- the tie paths below rewrite the whole order
    if (m < n && ez == 1) s.order[lz] = (int16_t)m;
  }
  if (found != 7u) {  // ties around a median: stable ranks (value, then index)
    for (int m = lane; m < ((n + 31) & ~31); m += 32) {
      int rx = 0, ry = 0, rz = 0;
      float xm = 0.f, ym = 0.f, zm = 0.f;
      if (m < n) {
        xm = s.x[m]; ym = s.y[m]; zm = s.z[m];
        for (int j = 0; j < n; ++j) {
          const float xj = s.x[j], yj = s.y[j], zj = s.z[j];
          rx += (xj < xm) || (xj == xm && j < m);
          ry += (yj < ym) || (yj == ym && j < m);
          rz += (zj < zm) || (zj == zm && j < m);
        }
      }
      const unsigned bx = __ballot_sync(kFull, m < n && rx == middle);
      const unsigned by = __ballot_sync(kFull, m < n && ry == middle);
      const unsigned bz = __ballot_sync(kFull, m < n && rz == middle);
      if (bx) med[0] = __shfl_sync(kFull, xm, __ffs(bx) - 1);
      if (by) med[1] = __shfl_sync(kFull, ym, __ffs(by) - 1);
      if (bz) med[2] = __shfl_sync(kFull, zm, __ffs(bz) - 1);
    }
  }
  __syncwarp();
  if (any_ztie) {
    // Exact z ties: see vertex_group.  Beyond 16 points one lane replays libstdc++'s introsort
    // (dev_stdsort.h).  That code lives in a second kernel that only sees the (rare) tied
    // items: with it inside the main instance every item ran 25 % slower.
    if (n > 16) {
      if (!REPLAY) return true;
      __syncwarp();
      const int first = any_ytie ? (any_xtie ? 0 : 1) : 2;
      const int16_t *init = first == 1 ? perm : perm + kVtxCap;
      if (pack2 != nullptr) replay_sorts_warp(s.order, reinterpret_cast<SelKey *>(pack), pack2, lq, rq, n, first, init, s.x, s.y, s.z);
      else if (lane == 0) replay_sorts_packed(s.order, pack, n, first, init, s.x, s.y, s.z);
    } else
    for (int m = lane; m < n; m += 32) {
      const float xm = s.x[m], ym = s.y[m], zm = s.z[m];
      const int cm = s.col[m];
      int rk = 0;
      for (int j = 0; j < n; ++j) rk += key_less(s.z[j], s.y[j], s.x[j], s.col[j], zm, ym, xm, cm);
      s.order[rk] = (int16_t)m;
    }
    __syncwarp();
  }
  // keep points within max_dist_to_centroid of the median, in z order (trellis.cpp:89-93)
  const float cut = dp->centroid_sq_cut;  // dist < max_dist_to_centroid  <=>  squared dist < cut
  int kept = 0;
  for (int sidx = lane; sidx < ((n + 31) & ~31); sidx += 32) {
    bool keep = false;
    int m = 0;
    if (sidx < n) {
      m = s.order[sidx];
      keep = sqnorm3f(s.x[m] - med[0], s.y[m] - med[1], s.z[m] - med[2]) < cut;
    }
    const unsigned b = __ballot_sync(kFull, keep);
    __syncwarp();  // every lane has read its order[] entry before any lane overwrites one
    if (keep) s.order[kept + __popc(b & ((1u << lane) - 1u))] = (int16_t)m;  // in-place: kept+pos <= sidx
    kept += __popc(b);
    __syncwarp();
  }
  sloam_vertex v;
  v.cx = med[0]; v.cy = med[1]; v.cz = med[2];
  v.radius = 0.f; v.n_points = 0; v.point_begin = 0; v.row = row; v.is_valid = 0;
  if (kept > 1) {  // trellis.cpp:95-100
    const int a = s.order[0], b = s.order[kept - 1];
    v.radius = dist3f(s.x[a], s.y[a], s.z[a], s.x[b], s.y[b], s.z[b]);
    for (int q = lane; q < kept; q += 32) {
      const int m = s.order[q];
      sloam_point p; p.x = s.x[m]; p.y = s.y[m]; p.z = s.z[m]; p.intensity = s.w[m];
      st_point(pool + base + q, p);
    }
    v.n_points = kept; v.point_begin = base; v.is_valid = 1;
  }
  if (lane == 0) *out = v;
  return false;
}

// members of an item into the warp's arrays (column order); n <= kVtxCap
__device__ __forceinline__ void gather_members(const DevParams *dp, VtxSmem &s, const VItem &it, int k,
                                               const sloam_point *__restrict__ tree,
                                               const int32_t *__restrict__ run_pix0,
                                               const int32_t *__restrict__ run_info) {
  const int N = dp->N, W = dp->p.img_w;
  const int lane = threadIdx.x & 31;
  const sloam_point *tk = tree + (size_t)k * N;
  const int rowbase = (int)it.row * W;
  if (it.span == 0) {
    for (int m = lane; m < (int)it.n; m += 32) {
      const sloam_point p = ld_point(tk + it.pix0 + m);
      s.x[m] = p.x; s.y[m] = p.y; s.z[m] = p.z; s.w[m] = p.intensity;
      s.col[m] = (int16_t)(it.pix0 + m - rowbase);
    }
  } else {
    const int32_t *pix0_k = run_pix0 + (size_t)k * N, *info_k = run_info + (size_t)k * N;
    int acc = 0;
    for (int j = it.first_run; j <= it.first_run + (int)it.span; ++j) {
      const int inf = info_k[j];
      if ((inf & (int)kSlotMask) != (int)it.slot + 1) continue;
      const int ln = inf >> 11, p0 = pix0_k[j];
      for (int t = lane; t < ln; t += 32) {
        const sloam_point p = ld_point(tk + p0 + t);
        const int m = acc + t;
        s.x[m] = p.x; s.y[m] = p.y; s.z[m] = p.z; s.w[m] = p.intensity;
        s.col[m] = (int16_t)(p0 + t - rowbase);
      }
      acc += ln;
    }
  }
  __syncwarp();
}

#ifndef SLOAM_VTX_MIN
#define SLOAM_VTX_MIN 5
#endif
// One persistent launch walks the four size classes one after the other (independent lists).
__global__ void __launch_bounds__(kVtxWarps * 32, SLOAM_VTX_MIN)
vertex_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ tree,
              const VItem *__restrict__ items, int32_t *__restrict__ vlists, long long list_cap,
              int32_t *__restrict__ n_lists, const int32_t *__restrict__ run_pix0,
              const int32_t *__restrict__ run_info, sloam_vertex *__restrict__ slot_vertices,
              sloam_point *__restrict__ pool, const int32_t *__restrict__ item_pool) {
  __shared__ __align__(16) VtxSmem sm[kVtxWarps];
  const int N = dp->N, H = dp->p.img_h, T = dp->p.max_trees;
  const int item_shift = dp->vw_row_bits + dp->vw_slot_bits;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gwarp = blockIdx.x * kVtxWarps + warp, nwarps = gridDim.x * kVtxWarps;
  int32_t *tied = vlists + (size_t)kClsTied * list_cap;
  GrpSmem &gs = *reinterpret_cast<GrpSmem *>(&sm[warp]);
  auto run_class = [&](auto gtag, int cls) {
    constexpr int G = decltype(gtag)::value;
    constexpr int per = 32 / G;
    const int32_t *list = vlists + (size_t)cls * list_cap;
    const int count = n_lists[cls];
    // the item of the NEXT round is fetched while the current one is processed (list entry ->
    // item record -> points is a chain of three dependent loads)
    VItem nit;
    int ngid = 0, nbase = 0;
    auto fetch = [&](int ws) {
      const int li = ws * per + lane / G;
      nit.pix0 = 0; nit.first_run = 0; nit.n = 0; nit.span = 0; nit.slot = 0; nit.row = 0;
      ngid = -1;
      nbase = 0;
      if (ws * per < count && li < count) {
        ngid = list[li];
        nit = items[ngid];
        nbase = item_pool[ngid];
      }
    };
    fetch(gwarp);
    for (int ws = gwarp; ws * per < count; ws += nwarps) {
      const VItem it = nit;
      const int gid = ngid, base = nbase;
      const bool active = gid >= 0;
      fetch(ws + nwarps);
      const int k = active ? (gid >> item_shift) : 0;
      const bool redo = vertex_group<G>(dp, gs, active, it, k, tree, run_pix0, run_info, slot_vertices, pool, base);
      const unsigned rb = __ballot_sync(kFull, redo && (lane & (G - 1)) == 0);
      if (rb) {
        int tb = 0;
        if (lane == 0) tb = atomicAdd(&n_lists[kClsTied], __popc(rb));
        tb = __shfl_sync(kFull, tb, 0);
        if (redo && (lane & (G - 1)) == 0) tied[tb + __popc(rb & ((1u << lane) - 1u))] = gid;
      }
    }
  };
  run_class(std::integral_constant<int, 8>{}, 0);
  run_class(std::integral_constant<int, 16>{}, 1);
  run_class(std::integral_constant<int, 32>{}, 2);
  {  // 33 .. kVtxCap members: one warp per item, the members in shared memory
    VtxSmem &s = sm[warp];
    const int32_t *list = vlists + (size_t)3 * list_cap;
    const int count = n_lists[3];
    for (int li = gwarp; li < count; li += nwarps) {
      const int gid = list[li];
      const VItem it = items[gid];
      const int k = gid >> item_shift;
      __syncwarp();
      gather_members(dp, s, it, k, tree, run_pix0, run_info);
      sloam_vertex *out = slot_vertices + ((size_t)k * T + it.slot) * H + it.row;
      if (build_vertex<false>(dp, s, nullptr, nullptr, it.n, it.row, out, pool + (size_t)k * N, item_pool[gid]) && lane == 0)
        tied[atomicAdd(&n_lists[kClsTied], 1)] = gid;
      __syncwarp();
    }
  }
}

// the items with exact z ties among more than 16 members, with the std::sort replay
__global__ void __launch_bounds__(kVtxWarps * 32)
vertex_replay_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ tree,
                     const VItem *__restrict__ items, const int32_t *__restrict__ vlists, long long list_cap,
                     const int32_t *__restrict__ n_lists, const int32_t *__restrict__ run_pix0,
                     const int32_t *__restrict__ run_info, sloam_vertex *__restrict__ slot_vertices,
                     sloam_point *__restrict__ pool, const int32_t *__restrict__ item_pool) {
  __shared__ __align__(16) VtxSmem sm[kVtxWarps];
  __shared__ unsigned long long s_pack[kVtxWarps * kVtxCap];  // replay scratch
  __shared__ int16_t s_perm[kVtxWarps * 2 * kVtxCap];         // members in x and in y order
  __shared__ SelKey s_pack2[kVtxWarps * kVtxCap];             // second record array of the warp sort
  __shared__ int s_lq[kVtxWarps][64], s_rq[kVtxWarps][64];    // its stopper queues
  const int N = dp->N, H = dp->p.img_h, T = dp->p.max_trees;
  const int item_shift = dp->vw_row_bits + dp->vw_slot_bits;
  const int warp = threadIdx.x >> 5;
  VtxSmem &s = sm[warp];
  const int32_t *list = vlists + (size_t)kClsTied * list_cap;
  const int count = n_lists[kClsTied];
  for (int li = blockIdx.x * kVtxWarps + warp; li < count; li += gridDim.x * kVtxWarps) {
    const int gid = list[li];
    const VItem it = items[gid];
    const int k = gid >> item_shift;
    __syncwarp();
    gather_members(dp, s, it, k, tree, run_pix0, run_info);
    sloam_vertex *out = slot_vertices + ((size_t)k * T + it.slot) * H + it.row;
    build_vertex<true>(dp, s, s_pack + warp * kVtxCap, s_perm + warp * 2 * kVtxCap, it.n, it.row, out,
                       pool + (size_t)k * N, item_pool[gid], s_pack2 + warp * kVtxCap, s_lq[warp], s_rq[warp]);
    __syncwarp();
  }
}

// wide path: one CTA per item with more than kVtxCap members -- same algorithm with a
// block-wide rank sort; members live in dynamic smem.
__global__ void vertex_wide_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ tree,
                                   const VItem *__restrict__ items, const int32_t *__restrict__ vlists,
                                   long long list_cap, const int32_t *__restrict__ n_lists,
                                   const int32_t *__restrict__ run_pix0, const int32_t *__restrict__ run_info,
                                   sloam_vertex *__restrict__ slot_vertices, sloam_point *__restrict__ pool,
                                   const int32_t *__restrict__ item_pool) {
  extern __shared__ float dsm[];
  const int N = dp->N, W = dp->p.img_w, H = dp->p.img_h, T = dp->p.max_trees;
  const int item_shift = dp->vw_row_bits + dp->vw_slot_bits;
  float *sx = dsm, *sy = dsm + W, *sz = dsm + 2 * W, *sw = dsm + 3 * W;
  int *scol = (int *)(dsm + 4 * W);
  int *sorder = scol + W;
  int *skeep = sorder + W;
  __shared__ int s_kept, s_base;
  __shared__ float s_med[3];
  const int32_t *list = vlists + (size_t)kClsWide * list_cap;
  const int total = n_lists[kClsWide];
  for (int li = blockIdx.x; li < total; li += gridDim.x) {
    const int gid = list[li];
    const VItem it = items[gid];
    const int k = gid >> item_shift, row = it.row, n = it.n;
    const sloam_point *tk = tree + (size_t)k * N;
    const int32_t *pix0_k = run_pix0 + (size_t)k * N, *info_k = run_info + (size_t)k * N;
    sloam_vertex *out = slot_vertices + ((size_t)k * T + it.slot) * H + row;
    __syncthreads();
    {  // members in column order (trellis.cpp:113-118)
      int acc = 0;
      for (int j = it.first_run; j <= it.first_run + (int)it.span; ++j) {
        const int inf = info_k[j];
        if ((inf & (int)kSlotMask) != (int)it.slot + 1) continue;
        const int ln = inf >> 11, p0 = pix0_k[j];
        for (int t = threadIdx.x; t < ln; t += blockDim.x) {
          const sloam_point p = ld_point(tk + p0 + t);
          const int m = acc + t;
          sx[m] = p.x; sy[m] = p.y; sz[m] = p.z; sw[m] = p.intensity; scol[m] = p0 + t - row * W;
        }
        acc += ln;
      }
    }
    __syncthreads();
    const int middle = (int)(n / 2.0);
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
      int rx = 0, ry = 0, rz = 0, rk = 0;
      const float xm = sx[m], ym = sy[m], zm = sz[m];
      for (int j = 0; j < n; ++j) {
        rx += (sx[j] < xm) || (sx[j] == xm && j < m);
        ry += (sy[j] < ym) || (sy[j] == ym && j < m);
        rz += (sz[j] < zm) || (sz[j] == zm && j < m);
        rk += key_less(sz[j], sy[j], sx[j], scol[j], zm, ym, xm, scol[m]);
      }
      sorder[rk] = m;
      if (rx == middle) s_med[0] = xm;
      if (ry == middle) s_med[1] = ym;
      if (rz == middle) s_med[2] = zm;
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // exact z ties among > 16 points: replay the three std::sort calls
      bool tie = false;
      for (int q = 0; q + 1 < n && !tie; ++q) tie = sz[sorder[q]] == sz[sorder[q + 1]];
      if (tie) {
        replay_three_sorts(sorder, n, sx, sy, sz);
      }
    }
    __syncthreads();
    const float maxd = dp->p.max_dist_to_centroid;
    for (int q = threadIdx.x; q < n; q += blockDim.x) {
      const int m = sorder[q];
      skeep[q] = dist3f(sx[m], sy[m], sz[m], s_med[0], s_med[1], s_med[2]) < maxd;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int kept = 0;
      for (int q = 0; q < n; ++q) if (skeep[q]) sorder[kept++] = sorder[q];
      s_kept = kept;
      sloam_vertex v;
      v.cx = s_med[0]; v.cy = s_med[1]; v.cz = s_med[2];
      v.radius = 0.f; v.n_points = 0; v.point_begin = 0; v.row = row; v.is_valid = 0;
      if (kept > 1) {
        const int a = sorder[0], b = sorder[kept - 1];
        v.radius = dist3f(sx[a], sy[a], sz[a], sx[b], sy[b], sz[b]);
        s_base = item_pool[gid];
        v.n_points = kept; v.point_begin = s_base; v.is_valid = 1;
      }
      *out = v;
    }
    __syncthreads();
    if (s_kept > 1)
      for (int q = threadIdx.x; q < s_kept; q += blockDim.x) {
        const int m = sorder[q];
        sloam_point p; p.x = sx[m]; p.y = sy[m]; p.z = sz[m]; p.intensity = sw[m];
        pool[(size_t)k * N + s_base + q] = p;
      }
  }
}

// ---- 4. trees: keep clusters with enough vertices, bottom row first --------
__global__ void tree_compact_kernel(const DevParams *__restrict__ dp, const int32_t *__restrict__ n_big,
                                    const int32_t *__restrict__ big_rank, const uint32_t *__restrict__ slot_rows,
                                    const sloam_vertex *__restrict__ slot_vertices,
                                    sloam_tree *__restrict__ trees, int32_t *__restrict__ n_trees,
                                    sloam_vertex *__restrict__ vertices) {
  extern __shared__ int32_t sm[];
  const int H = dp->p.img_h, T = dp->p.max_trees, Hw = (H + 31) >> 5;
  const int minv = dp->p.min_tree_vertices, maxv = dp->p.max_tree_vertices;
  const int k = blockIdx.x;
  const int nb = n_big[k];
  int32_t *s_nv = sm;        // [T] vertices of the slot (0 when rejected)
  int32_t *s_idx = sm + T;   // [T] tree index
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int top0 = ((H + 31) & ~31) - 1;  // rows from the bottom up (trellis.cpp:111), 32 per step
  // a row holds a vertex record iff the label kernel emitted a work item for it (row bit)
  auto has_vertex = [&](const uint32_t *rows, const sloam_vertex *sv, int r) {
    return r >= 0 && r < H && ((rows[r >> 5] >> (r & 31)) & 1u) && sv[r].is_valid;
  };
  for (int s = warp; s < nb; s += nwarps) {
    const uint32_t *rows = slot_rows + ((size_t)k * T + s) * Hw;
    const sloam_vertex *sv = slot_vertices + ((size_t)k * T + s) * H;
    int cnt = 0;
    for (int top = top0; top >= 0; top -= 32) cnt += __popc(__ballot_sync(kFull, has_vertex(rows, sv, top - lane)));
    if (lane == 0) s_nv[s] = cnt > minv ? min(cnt, maxv) : 0;  // trellis.cpp:124-127
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int s = 0; s < nb; ++s) { s_idx[s] = t; t += s_nv[s] > 0; }
    n_trees[k] = t;
  }
  __syncthreads();
  for (int s = warp; s < nb; s += nwarps) {
    const int nv = s_nv[s];
    if (nv == 0) continue;
    const int t = s_idx[s];
    const uint32_t *rows = slot_rows + ((size_t)k * T + s) * Hw;
    const sloam_vertex *sv = slot_vertices + ((size_t)k * T + s) * H;
    sloam_vertex *dst = vertices + ((size_t)k * T + t) * maxv;
    int cnt = 0, npts = 0;
    for (int top = top0; top >= 0; top -= 32) {
      const int r = top - lane;
      const bool v = has_vertex(rows, sv, r);
      const unsigned b = __ballot_sync(kFull, v);
      const int pos = cnt + __popc(b & ((1u << lane) - 1u));
      int np = 0;
      if (v && pos < nv) { const sloam_vertex vv = sv[r]; dst[pos] = vv; np = vv.n_points; }
      npts += warp_sum(np);
      cnt += __popc(b);
    }
    if (lane == 0) {
      sloam_tree tr;
      tr.tree_id = big_rank[(size_t)k * T + s];
      tr.n_vertices = nv;
      tr.vertex_begin = t * maxv;
      tr.n_points = npts;
      trees[(size_t)k * T + t] = tr;
    }
  }
}

// ---- labels for the find_clusters stage entry -------------------------------
__global__ void cc_labels_kernel(const DevParams *__restrict__ dp, int K, const uint32_t *__restrict__ bits,
                                 const uint4 *__restrict__ planes, const int32_t *__restrict__ wbase,
                                 const int32_t *__restrict__ run_label, uint32_t *__restrict__ labels) {
  const int N = dp->N, Nw = (N + 31) >> 5;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)K * N) return;
  const int k = (int)(g / N);
  const int i = (int)(g - (long long)k * N);
  const size_t w = (size_t)k * Nw + (i >> 5);
  uint32_t lab = 0xFFFFFFFFu;
  if ((bits[w] >> (i & 31)) & 1u) {
    const uint4 pl = planes[w];
    if ((pl.x >> (i & 31)) & 1u)
      lab = (uint32_t)run_label[(size_t)k * N + wbase[w] + __popc(pl.y & ((2u << (i & 31)) - 1u)) - 1];
  }
  labels[g] = lab;
}

// runs per keyframe that cc_label_kernel keeps in shared memory (more: global scratch, slower).
// A trunk is one run per scan line, so a context sized for many trees expects many runs; the
// smaller the arrays, the more keyframes an SM works on at once.
static int label_smem_runs(const sloam_ctx *c) {
  const long long by_image = c->hp.N / 16, by_trees = (long long)c->hp.p.max_trees * c->hp.p.img_h / 2;
  const long long rs = std::max(by_image, by_trees);
  const int Nw = (c->hp.N + 31) / 32, Hw = (c->hp.p.img_h + 31) / 32;
  // what fits next to the bit planes and the row masks in 200 KB
  const long long fixed = 8ll * Nw + 4ll * c->hp.p.max_trees * Hw;
  const long long fit = (200 * 1024 - fixed) / 12;
  long long r = std::max(1024ll, std::min(std::min(rs, 16384ll), fit));
  // four CTAs (the thread limit) instead of three per SM when that still leaves room for one
  // run per 32 pixels: a batch of 512 keyframes is then one wave of CTAs instead of two
  // (not for a context sized for several times more trees than that leaves room for: a dense
  // forest really has that many runs, and runs beyond the capacity live in global memory)
  const long long four = (54 * 1024 - fixed) / 12;
  if (four < r && four >= c->hp.N / 32 && by_trees <= 2 * four) r = four;
  return (int)r;
}

static int run_cc(sloam_ctx *c, int K, const sloam_point *tree, bool bits_ready, bool want_labels) {
  Workspace &w = c->ws;
  const sloam_params &p = c->hp.p;
  const int N = c->hp.N, Nw = (N + 31) / 32, Hw = (p.img_h + 31) / 32, T = p.max_trees;
  const long long total_words = (long long)K * Nw;
  const bool pre_zeroed = (c->zero_valid & 2u) != 0;  // zeroed with the rest of the counters (pipeline.cu)
  c->zero_valid &= ~(2u | 4u);
  if (!pre_zeroed) {
    SB_CUDA(c, cudaMemsetAsync(w.kf_flags, 0, sizeof(int32_t) * K, c->stream));
    SB_CUDA(c, cudaMemsetAsync(w.n_vlists, 0, sizeof(int32_t) * 8, c->stream));
  }
  if (!bits_ready) {  // caller-supplied cloud: derive the bits from the points
    const dim3 blocks((unsigned)((N + 255) / 256), (unsigned)K);
    tree_bits_kernel<<<blocks, 256, 0, c->stream>>>(c->dp, tree, w.tree_bits);
    SB_LAUNCH_CHECK(c);
  }
  const unsigned rgrid = (unsigned)std::min<long long>((total_words + 255) / 256, (long long)c->sm_count * 8);
  PROF_BEGIN(c, P_CC_ROWS);
  cc_rows_kernel<<<rgrid, kRowsWarps * 32, 0, c->stream>>>(c->dp, K, tree, w.tree_bits, reinterpret_cast<uint4 *>(w.cc_planes));
  PROF_END(c, P_CC_ROWS);
  SB_LAUNCH_CHECK(c);
  const int Rs = label_smem_runs(c);
  const size_t smem = sizeof(uint32_t) * 2 * (size_t)Nw + sizeof(uint32_t) * (size_t)T * Hw + sizeof(int32_t) * 3 * (size_t)Rs;
  static size_t lbl_set[64] = {};  // opt-in shared memory per device, only grows
  if (smem > 48 * 1024 && smem > lbl_set[c->device & 63]) {
    SB_CUDA(c, cudaFuncSetAttribute(cc_label_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lbl_set[c->device & 63] = smem;
  }
  const long long list_cap = (long long)c->max_k * T * p.img_h;
  PROF_BEGIN(c, P_CC_LABEL);
  // four 512-thread CTAs per SM fit up to 54 KB each (label_smem_runs sizes the run arrays for
  // that when it can); a larger image gets 1024 threads per keyframe
  const int lbl_threads = smem <= 54 * 1024 ? kLblThreads : kLblThreadsMax;
  cc_label_kernel<<<K, lbl_threads, smem, c->stream>>>(
      c->dp, w.tree_bits, reinterpret_cast<const uint4 *>(w.cc_planes), Rs, w.run_par, reinterpret_cast<uint32_t *>(w.run_siz),
      w.run_len, w.run_pix0, w.run_info, w.run_label, want_labels ? w.cc_wbase : nullptr, w.n_roots, w.n_big,
      w.big_rank, w.kf_flags, w.slot_rows, reinterpret_cast<VItem *>(w.vitems), w.vitem_pool, w.vlists, list_cap,
      w.n_vlists);
  PROF_END(c, P_CC_LABEL);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

int launch_compute_graph(sloam_ctx *c, int K, const sloam_point *tree, sloam_tree *trees,
                         int32_t *n_trees, sloam_vertex *vertices, sloam_point *vertex_points,
                         bool bits_ready) {
  Workspace &w = c->ws;
  const sloam_params &p = c->hp.p;
  const int T = p.max_trees, W = p.img_w;
  int rc = run_cc(c, K, tree, bits_ready, false);
  if (rc != SLOAM_OK) return rc;
  const long long list_cap = (long long)c->max_k * T * p.img_h;
  const VItem *items = reinterpret_cast<const VItem *>(w.vitems);
  // persistent grid: exactly the CTAs that are resident at once (a partial second wave would
  // double the time of its work items)
  static int vtx_occ = 0;
  if (vtx_occ == 0) {
    SB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&vtx_occ, vertex_kernel, kVtxWarps * 32, 0));
    if (vtx_occ < 1) vtx_occ = 1;
  }
  const int vgrid = c->sm_count * vtx_occ;
  PROF_BEGIN(c, P_VERTEX);
  vertex_kernel<<<vgrid, kVtxWarps * 32, 0, c->stream>>>(c->dp, tree, items, w.vlists, list_cap, w.n_vlists, w.run_pix0,
                                                         w.run_info, w.slot_vertices, vertex_points, w.vitem_pool);
  PROF_END(c, P_VERTEX);
  SB_LAUNCH_CHECK(c);
  // the tied items again, with the std::sort replay (usually an empty list: the CTAs exit)
  PROF_BEGIN(c, P_VERTEX_REPLAY);
  vertex_replay_kernel<<<c->sm_count, kVtxWarps * 32, 0, c->stream>>>(c->dp, tree, items, w.vlists, list_cap, w.n_vlists,
                                                                      w.run_pix0, w.run_info, w.slot_vertices,
                                                                      vertex_points, w.vitem_pool);
  SB_LAUNCH_CHECK(c);
  const size_t wide_smem = sizeof(float) * 4 * W + sizeof(int) * 3 * W;
  // opt-in shared memory: the attribute belongs to (function, device) and must only grow -- a
  // context with a wider image than an earlier one needs more, a narrower one must not lower it
  static size_t wide_set[64] = {};
  if (wide_smem > 48 * 1024 && wide_smem > wide_set[c->device & 63]) {
    SB_CUDA(c, cudaFuncSetAttribute(vertex_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem));
    wide_set[c->device & 63] = wide_smem;
  }
  vertex_wide_kernel<<<c->sm_count, 256, wide_smem, c->stream>>>(c->dp, tree, items, w.vlists, list_cap, w.n_vlists,
                                                                 w.run_pix0, w.run_info, w.slot_vertices,
                                                                 vertex_points, w.vitem_pool);
  PROF_END(c, P_VERTEX_REPLAY);
  SB_LAUNCH_CHECK(c);
  PROF_BEGIN(c, P_TREE_COMPACT);
  tree_compact_kernel<<<K, 256, sizeof(int32_t) * 2 * T, c->stream>>>(c->dp, w.n_big, w.big_rank, w.slot_rows,
                                                                       w.slot_vertices, trees, n_trees,
                                                                       vertices);
  PROF_END(c, P_TREE_COMPACT);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

// materialise the dense tree cloud of the last fused run (see sloam_ctx::tree_sparse)
int launch_tree_fill(sloam_ctx *c, int K) {
  const dim3 blocks((unsigned)((c->hp.N + 255) / 256), (unsigned)K);
  tree_fill_kernel<<<blocks, 256, 0, c->stream>>>(c->dp, c->ws.tree_bits, c->ws.tree);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sloam_b200_find_clusters_dev(sloam_ctx *c, int K, const sloam_point *tree, uint32_t *labels,
                                 int32_t *n_clusters) {
  if (!c || K <= 0 || K > c->max_k || !tree || !labels) return set_err(c, SLOAM_E_INVALID, "find_clusters: bad arguments");
  int rc = run_cc(c, K, tree, false, true);
  if (rc != SLOAM_OK) return rc;
  const long long total = (long long)K * c->hp.N;
  cc_labels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(
      c->dp, K, c->ws.tree_bits, reinterpret_cast<const uint4 *>(c->ws.cc_planes), c->ws.cc_wbase, c->ws.run_label, labels);
  SB_LAUNCH_CHECK(c);
  if (n_clusters)
    SB_CUDA(c, cudaMemcpyAsync(n_clusters, c->ws.n_roots, sizeof(int32_t) * K, cudaMemcpyDeviceToDevice, c->stream));
  return SLOAM_OK;
}

int sloam_b200_compute_graph_dev(sloam_ctx *c, int K, const sloam_point *tree, sloam_tree *trees,
                                 int32_t *n_trees, sloam_vertex *vertices, sloam_point *vertex_points) {
  if (!c || K <= 0 || K > c->max_k || !tree || !trees || !n_trees || !vertices || !vertex_points)
    return set_err(c, SLOAM_E_INVALID, "compute_graph: bad arguments");
  return launch_compute_graph(c, K, tree, trees, n_trees, vertices, vertex_points, false);
}

}  // extern "C"
