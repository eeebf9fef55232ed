// k1_project.cu -- stage a1 + a2: spherical projection, closest-point-wins
// range image, mask gather, dense tree cloud and order-preserving ground
// compaction in ONE pass over the points.
//
// Replaces Segmentation::_doProjection (inference.cpp:80-165) and the two
// Segmentation::maskCloud calls of SLOAMNode::run (inference.cpp:230-273,
// sloamNode.cpp:212,215).  Also tags every ground point with its polar cell
// (sloam::binGroundPoints, sloam.cpp:339-358) while it is in registers, so
// stage a3 never re-reads the cloud to bin it.
//
// HBM-bound streaming kernel by design (issue-bound as measured).  Algorithmic bytes per keyframe
// of the fused path: 16N points + 1N mask + 4N range image + N/8 tree bits + 16T tree points +
// 9G ground records and cell tags (T tree-labelled, G ground-labelled points).  Persistent CTAs
// (one per resident slot) take tiles of kSplitTile consecutive points of one keyframe from a
// global counter; the points of the next tile are fetched by one TMA bulk copy (cp.async.bulk +
// mbarrier) into the other half of a shared-memory double buffer while the current tile is
// processed.  The tile stays in shared memory for the whole iteration: the phases re-read the
// points they need instead of holding 4 x float4 per thread in registers across barriers, and
// a thread carries one packed word (pixel index + theta bin) and the mask byte per point.
//
// Ground layout: the ground points of tile t are written, in input order, to slots
// [t * kSplitTile, t * kSplitTile + tile_count[t]) of the keyframe's ground array
// ("tile-strided") -- as 8-byte (z key, point index) records on the fused path, as points for
// the stage entries.  The order of the slots is the input order, which is all the ground
// stage needs (it bins and sorts by (z, slot)), so no CTA ever waits for another one: the
// single-pass compaction with a look-back scan that this replaces serialised the tiles of
// a keyframe.  Callers that want the contiguous cloud of Segmentation::maskCloud (the stage
// entries, the intermediates) get it from ground_compact_kernel.
#include "common.cuh"

#ifndef SLOAM_K1_EXP
#define SLOAM_K1_EXP 0  // timing experiments only (wrong results): 1 no mask gather, 2 no range atomics,
#endif                  // 4 no ground stores, 8 no tree stores, 16 no ground cell arithmetic, 32 no projection

namespace sb {

constexpr int kThreads = 256;
constexpr int kRounds = kSplitTile / kThreads;  // points per thread and tile (1024 / 256 = 4)

// Cheap atan2 / asin for the ESTIMATE only (the decision is made exact by the margins
// below plus the fp64 fallback).  Degree-7 polynomials in t^2 fitted on Chebyshev nodes,
// evaluated with explicit FMAs; maximum absolute error measured over 2e6 samples in fp32:
// atan on [0, 1] 1.8e-7 rad, asin on [0, 0.72] 1.0e-7 rad.  With the fast division (2 ulp)
// and the pi/2, pi reflections (float(pi) is off by 8.7e-8) the yaw estimate stays within
// 6.5e-7 rad of the true angle -- inside the 7.2e-7 rad the margins were derived for with
// CUDA's atan2f (3 ulp).  0/0, inf/inf give NaN, which fails the margin test -> exact path.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  const float t = __fdividef(mn, mx), u = t * t;
  float p = -0.0048311425f;
  p = __fmaf_rn(p, u, 0.0247566803f);
  p = __fmaf_rn(p, u, -0.0602189711f);
  p = __fmaf_rn(p, u, 0.0996791101f);
  p = __fmaf_rn(p, u, -0.140401334f);
  p = __fmaf_rn(p, u, 0.1997368011f);
  p = __fmaf_rn(p, u, -0.333323027f);
  p = __fmaf_rn(p, u, 0.9999999582f);
  float a = p * t;
  if (ay > ax) a = 1.57079633f - a;
  if (x < 0.f) a = 3.14159265f - a;
  return y < 0.f ? -a : a;
}
__device__ __forceinline__ float fast_asinf(float t) {
  if (!(fabsf(t) <= 0.72f)) return asinf(t);  // outside any lidar's vertical field of view: rare
  const float u = t * t;
  float p = 0.1204409277f;
  p = __fmaf_rn(p, u, -0.1073193451f);
  p = __fmaf_rn(p, u, 0.08617726686f);
  p = __fmaf_rn(p, u, 0.01356594868f);
  p = __fmaf_rn(p, u, 0.04694133235f);
  p = __fmaf_rn(p, u, 0.07485245047f);
  p = __fmaf_rn(p, u, 0.1666700012f);
  p = __fmaf_rn(p, u, 0.9999999924f);
  return p * t;
}

// fp32 estimate of project_pixel(): returns the pixel index when both image
// coordinates are provably (margins mx, my, in pixels) on the same side of every
// pixel boundary as the bit-exact evaluation, else -1.  Error budget (W = 2048):
// atan2 estimate <= 7.2e-7 rad (CUDA atan2f 3 ulp / fast_atan2f 6.5e-7), asin <= 4 ulp; fp32 evaluation of
// 0.5 (yaw / pi + 1) W deviates from the exact pipeline by < 9e-4 px, of
// (1 - (pitch + |fov_down|) / fov) H by < 2e-4 px.  NaN / out-of-image -> -1.
__device__ __forceinline__ int project_pixel_fast(const ProjGeom &g, float x, float y, float z, float mx,
                                                  float my, float kx, float ky, float cy, float *yaw_out) {
  // no-return beams (NaN x or y): yaw and pitch are NaN, both coordinates take the
  // clamp of inference.cpp:120,125 (std::min keeps its first argument) -> last pixel
  *yaw_out = __int_as_float(0x7fc00000);
  if (x != x || y != y) return (int)((g.Hf - 1.0f) * g.Wf + (g.Wf - 1.0f));
  const float yaw = -fast_atan2f(y, x);
  *yaw_out = yaw;
  // z / range via rsqrt (<= 2 ulp): the estimate only has to be inside the margins
  const float pitch = fast_asinf(z * rsqrtf(x * x + y * y + z * z));
  // one FMA per coordinate: 0.5 (yaw / pi + 1) W = yaw (W / 2 pi) + W / 2 and
  // (1 - (pitch + |fov_down|) / fov) H = pitch (-H / fov) + H (1 - |fov_down| / fov), constants
  // rounded once (relative 6e-8: < 7e-5 px at W = 2048, < 2e-5 px at H = 128) -- a smaller
  // deviation from the exact pipeline than the four-operation form the budget above was made for
  const float px = __fmaf_rn(yaw, kx, 0.5f * g.Wf);
  const float py = __fmaf_rn(pitch, ky, cy);
  const float fx = floorf(px), fy = floorf(py);
  const float tx = px - fx, ty = py - fy;  // exact
  const bool ok = tx > mx && tx < 1.0f - mx && ty > my && ty < 1.0f - my &&
                  fx >= 0.0f && fx <= g.Wf - 1.0f && fy >= 0.0f && fy <= g.Hf - 1.0f;
  return ok ? (int)(fy * g.Wf + fx) : -1;
}

// Same idea for the polar ground cell.  The theta bin comes from the yaw estimate of the
// projection with a margin of 1e-3 bins (255 = next to a bin edge or unknown: decide exactly);
// it is computed for every point while the yaw is at hand and travels in the spare byte of the
// pixel word, so that no per-point float has to stay alive until the ground phase.
__device__ __forceinline__ unsigned theta_bin_fast(const GroundGeom &g, float theta) {
  const float tb_f = __fmaf_rn(theta, g.inv_theta_step_f, 3.14159265f * g.inv_theta_step_f);
  const float fl = floorf(tb_f);
  const float tt = tb_f - fl;  // exact
  // margins: atan2 estimate + fp32 evaluation < 2e-5 bins for up to 255 bins; NaN fails
  if (!(tt > 1e-3f && tt < 0.999f) || g.TB > 255) return 255u;
  int tb = (int)fl;
  tb = tb < g.TB - 1 ? tb : g.TB - 1;
  tb = tb > 0 ? tb : 0;
  return (unsigned)tb;
}
// The radius tests (sloam.cpp:344) and the radial bin are exact threshold comparisons on
// r2 = x*x + y*y (proj_math.h): pow_2 rounds the exact fp64 square of a float to float == the
// fp32 product, so r2 is the argument of euclideanDist2D's sqrtf (utils.h:9-12) -- no square
// root, no double arithmetic, no fallback for the radial coordinate (up to four radial bins;
// more: estimate with the exact ground_cell_of() next to a bin edge).
__device__ __forceinline__ int ground_cell_fast(const GroundGeom &g, float x, float y, unsigned tbv) {
  const float r2 = x * x + y * y;
  int rb = 0;
  if (g.r2_bins >= 0) {
    rb = ground_radial_bin_of_r2(g, r2);  // the thresholds are warp-uniform values
    if (rb < 0) return -1;
  } else if (!(r2 >= g.r2_in_lo && r2 < g.r2_in_hi)) {
    return -1;
  }
  if (tbv == 255u) return ground_cell_of(g, x, y);
  if (g.r2_bins < 0) {
    const float rb_f = sqrtf(r2) * g.inv_radial_step_f;
    const float flr = floorf(rb_f), tr = rb_f - flr;
    if (!(tr > 1e-3f && tr < 0.999f)) return ground_cell_of(g, x, y);
    rb = (int)flr;
    rb = rb < g.RB - 1 ? rb : g.RB - 1;
    rb = rb > 0 ? rb : 0;
  }
  return rb * g.TB + (int)tbv;
}

#ifndef SLOAM_K1_MIN_CTAS
// round 1 (separate scan + flush barriers): 4 -> 427 us, 5 -> 384 us, 6 -> 390 us per 1000 VLP-16 keyframes.
// round 2, points held in registers across the phases: 5 CTAs / 48 registers spill 120 bytes -> 400 us,
// 4 CTAs / 64 registers -> 367 us per 512 OS1-64 keyframes; points re-read from shared memory
// (one packed word per point in registers): 4 CTAs 581 us, 5 CTAs / 48 registers 568 us per 1024
// keyframes (641 us before); 6 do not fit (42 KB of shared memory per CTA)
#define SLOAM_K1_MIN_CTAS 5
#endif
// FUSED (the production path, with DO_PROJECT and DO_SPLIT): the ground points are not copied.
// Every ground point becomes one 8-byte record (z key, point index) in the tile-strided layout
// described above (slot order = input order) next to its cell tag, and seg_tab[k][cell][tile]
// receives the number of points of each cell in the tile, from which ground_scatter_kernel
// (k2_ground.cu) bins every tile independently.
template <bool DO_PROJECT, bool DO_SPLIT, bool FUSED>
__global__ void __launch_bounds__(kThreads, SLOAM_K1_MIN_CTAS)
project_split_kernel(const DevParams *__restrict__ dp, int K, const sloam_point *__restrict__ points,
                     const uint8_t *__restrict__ mask, int32_t *__restrict__ pix_io,
                     unsigned *__restrict__ range_bits, sloam_point *__restrict__ tree,
                     sloam_point *__restrict__ ground, int32_t *__restrict__ ground_count,
                     uint8_t *__restrict__ ground_cell, int32_t *__restrict__ cell_count,
                     int32_t *__restrict__ tile_count, int ground_stride,
                     uint32_t *__restrict__ tree_bits, int sparse_tree,
                     uint2 *__restrict__ ground_recs, uint32_t *__restrict__ seg_tab,
                     int *__restrict__ tile_ctr) {
  __shared__ int s_pix[DO_PROJECT ? kSplitTile : 1];  // exact pixel indices of the queued points
  // per-tile cell counts and (round, warp) ground counts, double-buffered by tile parity: a
  // tile's counts are flushed / read while the next tile already fills the other buffer, which
  // saves the two block barriers that would otherwise fence their reuse
  __shared__ int s_hist2[DO_SPLIT ? 2 * kMaxCells : 1];
  __shared__ uint16_t s_slow[DO_PROJECT ? kSplitTile : 1];
  __shared__ int s_nslow;
  static_assert(kRounds * (kThreads / 32) == 32, "one (round, warp) count per lane");
  __shared__ int s_cnt2[2][32];
  // input staging: the points of the NEXT tile stream into shared memory (one TMA bulk copy,
  // completion on an mbarrier) while the current tile is being processed, so the kernel
  // always has a full tile of loads in flight per CTA
  __shared__ __align__(128) sloam_point s_in[2][kSplitTile];
  __shared__ __align__(8) unsigned long long s_mbar[2];
  // tiles are handed out by a global counter (tile_ctr, zero at launch), not by a fixed stride:
  // a CTA that becomes resident late -- behind the CTAs of another stream's kernel, e.g. the
  // result gather of the previous batch -- then takes fewer tiles instead of a full share
  __shared__ int s_tile_id[2];

  const int N = dp->N;
  const int tiles = (N + kSplitTile - 1) / kSplitTile;
  const int total_tiles = K * tiles;
  // floor(2^32 / tiles) + 1 gives the exact quotient while tile_id * tiles < 2^32; else divide
  const unsigned magic_tiles = (tiles > 1 && (unsigned long long)total_tiles * (unsigned)tiles < 0x100000000ull)
                                   ? (unsigned)(0x100000000ull / (unsigned)tiles) + 1u : 0u;
  const ProjGeom pg = dp->pg;
  const GroundGeom gg = dp->gg;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float qnan = __int_as_float(0x7fc00000);
#ifdef SLOAM_K1_EXPERIMENT_NO_EXACT  // timing experiment only (wrong pixels next to boundaries)
  const float mx = -1.0f, my = -1.0f;
#else
  const float mx = pg.Wf * 2.5e-6f + 1e-3f, my = 2e-3f;
#endif
  // coefficients of the fp32 pixel estimate (project_pixel_fast), from double
  const float kx = (float)(0.5 * (double)pg.Wf / 3.14159265358979323846);
  const float ky = (float)(-(double)pg.Hf / (double)pg.fov);
  const float cy = (float)((double)pg.Hf * (1.0 - (double)pg.fov_down_abs / (double)pg.fov));
  auto issue_tile = [&](int tile_id, int buf) {  // one thread
    const int kk = tile_id / tiles, tt = tile_id - kk * tiles;
    const int n_pts = min(kSplitTile, N - tt * kSplitTile);
    const unsigned bytes = (unsigned)(n_pts * sizeof(sloam_point));
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_mbar[buf]);
    const unsigned dst = (unsigned)__cvta_generic_to_shared(&s_in[buf][0]);
    const sloam_point *src = points + (size_t)kk * N + (size_t)tt * kSplitTile;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
  };
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_mbar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int first = atomicAdd(tile_ctr, 1);
    s_tile_id[0] = first;
    if (first < total_tiles) issue_tile(first, 0);
  }
  int fetched = total_tiles;  // (thread 0) the tile after the next one, claimed one iteration ahead
  if (threadIdx.x == 0) fetched = atomicAdd(tile_ctr, 1);
  __syncthreads();

  // cell counts of a finished tile -> global (ground stage input)
  auto flush_hist = [&](const int *hist, int fk, int ftile) {
    for (int c = threadIdx.x; c < kMaxCells; c += kThreads) {
      const int h = hist[c];
      if (h) atomicAdd(&cell_count[(size_t)fk * kMaxCells + c], h);
      // per-tile cell counts: ground_scatter_kernel (k2_ground.cu) bins every tile independently
      if (FUSED && c < dp->B) seg_tab[((size_t)fk * dp->B + c) * tiles + ftile] = (unsigned)h;
    }
  };
  int iter = 0, prev_k = -1, prev_tile = 0;
  for (;; ++iter) {
  const int buf = iter & 1;
  const int tile_id = s_tile_id[buf];  // written before the last barrier of the previous iteration
  if (tile_id >= total_tiles) break;
  const int k = magic_tiles ? (int)__umulhi((unsigned)tile_id, magic_tiles) : tile_id / tiles, tile = tile_id - k * tiles;
  const size_t kbase = (size_t)k * N;
  int *const s_hist = s_hist2 + (DO_SPLIT ? buf * kMaxCells : 0);
  int *const s_cnt = s_cnt2[buf];
  uint32_t *tree_bits_k = tree_bits ? tree_bits + (size_t)k * ((N + 31) >> 5) : nullptr;
  if (threadIdx.x == 0) s_nslow = 0;
  if (DO_SPLIT)
    for (int c = threadIdx.x; c < kMaxCells; c += kThreads) s_hist[c] = 0;
  {  // wait for this tile's points
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&s_mbar[buf]);
    const unsigned phase = (unsigned)((iter >> 1) & 1);
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
          : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // every warp is past its last read of the other buffer (the ground points of the previous
    // tile are re-read from it at the very end of an iteration): start the copy of the next tile
    s_tile_id[buf ^ 1] = fetched;
    if (fetched < total_tiles) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue_tile(fetched, buf ^ 1);
      fetched = atomicAdd(tile_ctr, 1);  // not needed before the next iteration
    }
  }
  // the previous tile of this CTA is complete behind that barrier: hand over its cell counts
  if (DO_SPLIT && prev_k >= 0) flush_hist(s_hist2 + (buf ^ 1) * kMaxCells, prev_k, prev_tile);
  prev_k = k; prev_tile = tile;

  // ---- phase A: project every point of the tile.
  // The pixel index is an integer derived from atan2f/asinf; the bit-exact evaluation
  // (proj_math.h, fp64) is ~230 DP instructions, so it only runs for the ~1.5 % of points
  // whose fast fp32 estimate lies within a proven error margin of a pixel boundary.
  // Those are queued and re-projected densely in phase B instead of diverging here.
  // The points themselves are NOT kept in registers: the tile stays in shared memory until the
  // end of the iteration and the later phases re-read what they need (one LDS.128).  What a
  // thread carries per point is one word -- the pixel index in the low 24 bits (kSlowPix =
  // queued) and the theta bin of the ground grid in the high 8 (255 = decide exactly) -- and
  // the mask byte, fetched as soon as the pixel is known so that the gather's latency runs
  // under phase B and its barriers.
  constexpr unsigned kSlowPix = 0xFFFFFFu;
  unsigned pixr[kRounds];
  unsigned char mk[kRounds];
  const sloam_point *s_pts = &s_in[buf][0];
#pragma unroll
  for (int j = 0; j < kRounds; ++j) {
    const int i = tile * kSplitTile + j * kThreads + threadIdx.x;
    pixr[j] = 0u;
    mk[j] = 0;
    if (i < N) {
      if (DO_PROJECT) {
        const sloam_point p = ld_point(s_pts + j * kThreads + threadIdx.x);
        float yaw;
        const int pix = (SLOAM_K1_EXP & 32) ? ((yaw = p.x), i) : project_pixel_fast(pg, p.x, p.y, p.z, mx, my, kx, ky, cy, &yaw);
        const unsigned tbv = DO_SPLIT ? theta_bin_fast(gg, -yaw) : 0u;
        pixr[j] = (tbv << 24) | (pix < 0 ? kSlowPix : (unsigned)pix);
        if (pix < 0) {
          s_slow[atomicAdd(&s_nslow, 1)] = j * kThreads + threadIdx.x;
        } else {
          if (DO_SPLIT) mk[j] = (SLOAM_K1_EXP & 1) ? (unsigned char)((pix & 3) == 0 ? 1 : ((pix & 31) == 1 ? 255 : 0))
                                                   : mask[kbase + pix];  // inference.cpp:242-243
          // closest point wins (inference.cpp:135,160-162): minimum over the bit pattern of the
          // non-negative SQUARED range (sqrtf is monotone, so the same point wins); the one
          // square root per pixel is taken by range_finalize_kernel.  NaN ranges never write.
          const float range_sq = p.x * p.x + p.y * p.y + p.z * p.z;
          if (!(SLOAM_K1_EXP & 2) && range_bits != nullptr && range_sq == range_sq)
            atomicMin(&range_bits[kbase + pix], __float_as_uint(range_sq));
        }
      } else {
        pixr[j] = (unsigned)pix_io[kbase + i];
      }
    }
  }
  if (DO_PROJECT) {
    __syncthreads();
    // ---- phase B: exact projection of the queued points, one per thread
    const int ns = s_nslow;
    for (int q = threadIdx.x; q < ns; q += kThreads) {
      const int idx = s_slow[q];
      const sloam_point p = ld_point(s_pts + idx);
      float range;
      s_pix[idx] = project_pixel(pg, p.x, p.y, p.z, &range);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kRounds; ++j) {
      const int i = tile * kSplitTile + j * kThreads + threadIdx.x;
      if (i < N) {
        if ((pixr[j] & kSlowPix) == kSlowPix) {  // a queued point: same steps as above with the exact pixel
          const int pix = s_pix[j * kThreads + threadIdx.x];
          pixr[j] = (pixr[j] & 0xFF000000u) | (unsigned)pix;
          if (DO_SPLIT) mk[j] = mask[kbase + pix];
          const sloam_point p = ld_point(s_pts + j * kThreads + threadIdx.x);
          const float range_sq = p.x * p.x + p.y * p.y + p.z * p.z;
          if (!(SLOAM_K1_EXP & 2) && range_bits != nullptr && range_sq == range_sq)
            atomicMin(&range_bits[kbase + pix], __float_as_uint(range_sq));
        }
        // FUSED: the pixel indices (proj_xs / proj_ys, which the reference keeps only for
        // maskCloud, inference.cpp:131-132) feed the mask gather and are not stored;
        // sloam_b200_get_intermediates recomputes them on demand
        if (!FUSED) pix_io[kbase + i] = (int)(pixr[j] & kSlowPix);
      }
    }
  }

  // ---- phase C: dense tree cloud, order-preserving ground compaction
  if (!DO_SPLIT) { __syncthreads(); continue; }
  // the (round, warp) ballot counts give every ground point its slot in input order
  unsigned bal[kRounds];
  unsigned gmask = 0;  // bit j: point j of this thread is ground
#pragma unroll
  for (int j = 0; j < kRounds; ++j) {
    const int i = tile * kSplitTile + j * kThreads + threadIdx.x;
    const bool in = i < N;
    unsigned char m = mk[j];
    if (!DO_PROJECT && in) m = mask[kbase + pixr[j]];  // inference.cpp:242-243
    // dense mode (:247-251): the point or a NaN point with intensity 0.  With sparse_tree
    // (fused pipeline) only the tree points are written; tree_bits says which pixels hold one
    // and the NaN points are materialised on demand (pipeline.cu).
    const bool is_t = in && (m == 255);
    if (!(SLOAM_K1_EXP & 8) && in && (is_t || !sparse_tree)) {
      sloam_point t;
      if (is_t) t = ld_point(s_pts + j * kThreads + threadIdx.x);
      else { t.x = qnan; t.y = qnan; t.z = qnan; t.intensity = 0.f; }
      st_point(tree + kbase + i, t);
    }
    if (tree_bits != nullptr) {
      const unsigned tb = __ballot_sync(kFull, is_t);
      if (lane == 0 && in) tree_bits_k[i >> 5] = tb;
    }
    const bool is_g = in && (m == 1);
    bal[j] = __ballot_sync(kFull, is_g);
    gmask |= (is_g ? 1u : 0u) << j;
    if (lane == 0) s_cnt[j * (kThreads / 32) + warp] = __popc(bal[j]);
  }
  __syncthreads();
  // every warp sums the counts of the units before its own (one count per lane, one REDUX per
  // round) instead of waiting for a scan by one warp behind a second barrier
  int unit_base[kRounds];
  {
    const int cnt_l = s_cnt[lane];
#pragma unroll
    for (int j = 0; j < kRounds; ++j) unit_base[j] = __reduce_add_sync(kFull, lane < j * (kThreads / 32) + warp ? cnt_l : 0);
    if (warp == 0) {
      const int tot = __reduce_add_sync(kFull, cnt_l);
      if (lane == 0) {  // ground points of this tile
        tile_count[(size_t)k * tiles + tile] = tot;
        if (tot) atomicAdd(&ground_count[k], tot);
      }
    }
  }
  // ground points go straight to their slots: the ground lanes of a warp own consecutive
  // slots, so the stores of a warp form contiguous runs
  sloam_point *gout = FUSED ? nullptr : ground + (size_t)k * ground_stride + (size_t)tile * kSplitTile;
  uint2 *rout = FUSED ? ground_recs + kbase + (size_t)tile * kSplitTile : nullptr;
  uint8_t *cout = ground_cell + (size_t)k * ground_stride + (size_t)tile * kSplitTile;
#pragma unroll
  for (int j = 0; j < kRounds; ++j) {
    if ((gmask >> j) & 1u) {
      const sloam_point p = ld_point(s_pts + j * kThreads + threadIdx.x);
      const int slot = unit_base[j] + __popc(bal[j] & ((1u << lane) - 1u));
      const unsigned tbv = DO_PROJECT ? (pixr[j] >> 24) : theta_bin_fast(gg, fast_atan2f(p.y, p.x));
      const int cell = (SLOAM_K1_EXP & 16) ? (slot & 31) : ground_cell_fast(gg, p.x, p.y, tbv);
      // FUSED: the point itself is not copied -- an 8-byte (z key, point index) record is all the
      // ground stage sorts; it reads the few retained points from the input cloud
      if (SLOAM_K1_EXP & 4) { if (cell >= 0) atomicAdd(&s_hist[cell], 1); continue; }
      if (FUSED) rout[slot] = make_uint2(float_key(p.z), (unsigned)(tile * kSplitTile + j * kThreads + threadIdx.x));
      else st_point(gout + slot, p);
      cout[slot] = (uint8_t)(cell < 0 ? 255 : cell);
      if (cell >= 0) atomicAdd(&s_hist[cell], 1);
    }
  }
  }  // tile loop
  if (DO_SPLIT && prev_k >= 0) {  // the last tile of this CTA
    __syncthreads();
    flush_hist(s_hist2 + ((iter - 1) & 1) * kMaxCells, prev_k, prev_tile);
  }
}

// Contiguous ground cloud (Segmentation::maskCloud's output) from the tile-strided one:
// grid (tiles, K); a CTA copies its tile's run to the offset given by the counts before it.
__global__ void ground_compact_kernel(const DevParams *__restrict__ dp, const sloam_point *__restrict__ src,
                                      int src_stride, const int32_t *__restrict__ tile_count,
                                      sloam_point *__restrict__ dst, int dst_stride) {
  __shared__ int s_off;
  const int tiles = (dp->N + kSplitTile - 1) / kSplitTile;
  const int k = blockIdx.y, tile = blockIdx.x;
  const int32_t *tc = tile_count + (size_t)k * tiles;
  if (threadIdx.x < 32) {
    int off = 0;
    for (int t = threadIdx.x; t < tile; t += 32) off += tc[t];
    off = warp_sum(off);
    if (threadIdx.x == 0) s_off = off;
  }
  __syncthreads();
  const int n = tc[tile];
  const sloam_point *s = src + (size_t)k * src_stride + (size_t)tile * kSplitTile;
  sloam_point *d = dst + (size_t)k * dst_stride + s_off;
  for (int i = threadIdx.x; i < n; i += blockDim.x) st_point(d + i, ld_point(s + i));
}

// squared range of the closest point -> range; empty pixels (still 0xFFFFFFFF) become 0
// (inference.cpp:135,150-158)
__device__ __forceinline__ unsigned range_final(unsigned v) {
  return v == 0xFFFFFFFFu ? 0u : __float_as_uint(sqrtf(__uint_as_float(v)));
}
__global__ void range_finalize_kernel(unsigned *__restrict__ range_bits, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    uint4 v = *reinterpret_cast<uint4 *>(range_bits + i);
    v.x = range_final(v.x); v.y = range_final(v.y); v.z = range_final(v.z); v.w = range_final(v.w);
    *reinterpret_cast<uint4 *>(range_bits + i) = v;
  } else {
    for (long long j = i; j < n; ++j) range_bits[j] = range_final(range_bits[j]);
  }
}

// Segmentation::_destaggerCloud (inference.cpp:200-228), applied by maskCloud to the organized
// (dense) cloud only (:256-259): the output starts as a copy of the masked cloud
// (pcl::copyPointCloud, :255); then, in raster order, x/y/z of pixel (row, col) are copied to
// (row, col + 32) on EVEN rows (Ouster column stagger) and to itself on odd rows.  The bound
// test is `im_col > W` (:209), so col + 32 == W passes and lands on pixel (row + 1, 0) -- which
// the next (odd) row then overwrites with its own point; columns beyond are dropped.  Net
// effect: out(row, col).xyz = in(row, col - 32).xyz for even rows and col >= 32, everything
// else unchanged (columns 0..31 of even rows keep their own point, and it also appears at
// col + 32); intensity is never moved.  For an odd H the reference's last write of the last
// row is past the end of the cloud (undefined): dropped here and in the oracle.
__device__ __forceinline__ int destagger_source(int i, int W, unsigned magic_w) {
  const int row = fast_div_w(i, magic_w), col = i - row * W;
  return ((row & 1) == 0 && col >= 32) ? i - 32 : i;
}
// dense: src = masked organized cloud (NaN points included), dst = destaggered cloud
__global__ void destagger_dense_kernel(const DevParams *__restrict__ dp, int K, const sloam_point *__restrict__ src,
                                       sloam_point *__restrict__ dst) {
  const int N = dp->N;
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long long)K * N) return;
  const int k = (int)(g / N), i = (int)(g - (long long)k * N);
  const int s = destagger_source(i, dp->p.img_w, dp->magic_w);
  sloam_point p = ld_point(src + (size_t)k * N + s);
  if (s != i) p.intensity = src[g].intensity;
  st_point(dst + g, p);
}
// sparse (fused pipeline): only the pixels whose bit is set hold a point
__global__ void destagger_sparse_kernel(const DevParams *__restrict__ dp, int K, const sloam_point *__restrict__ src,
                                        const uint32_t *__restrict__ sbits, sloam_point *__restrict__ dst,
                                        uint32_t *__restrict__ dbits) {
  const int N = dp->N, W = dp->p.img_w, Nw = (N + 31) >> 5;
  const unsigned magic_w = dp->magic_w;
  const int lane = threadIdx.x & 31;
  const long long total = (long long)K * Nw;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long gw = warp0; gw < total; gw += nwarps) {
    const int k = (int)(gw / Nw), wi = (int)(gw - (long long)k * Nw);
    const uint32_t *bk = sbits + (size_t)k * Nw;
    const int i = wi * 32 + lane;
    bool sb = false;
    int s = i;
    if (i < N) {
      s = destagger_source(i, W, magic_w);
      sb = (bk[s >> 5] >> (s & 31)) & 1u;
    }
    const unsigned ob = __ballot_sync(kFull, sb);
    if (lane == 0) dbits[gw] = ob;
    if (sb) {
      sloam_point p = ld_point(src + (size_t)k * N + s);
      if (s != i) p.intensity = ((bk[i >> 5] >> (i & 31)) & 1u) ? src[(size_t)k * N + i].intensity : 0.f;
      st_point(dst + (size_t)k * N + i, p);
    }
  }
}

// Polar cell tags for a ground cloud that did not come through the split
// kernel (stage entry sloam_b200_ground_planes_dev on caller-supplied clouds).
__global__ void ground_tag_kernel(const DevParams *__restrict__ dp, int K,
                                  const sloam_point *__restrict__ ground,
                                  const int32_t *__restrict__ ground_count, int stride,
                                  uint8_t *__restrict__ ground_cell, int32_t *__restrict__ cell_count) {
  __shared__ int s_hist[kMaxCells];
  const int k = blockIdx.y;
  for (int c = threadIdx.x; c < kMaxCells; c += blockDim.x) s_hist[c] = 0;
  __syncthreads();
  const int n = ground_count[k];
  const GroundGeom gg = dp->gg;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const sloam_point p = ground[(size_t)k * stride + i];
    const int cell = ground_cell_of(gg, p.x, p.y);
    ground_cell[(size_t)k * stride + i] = (uint8_t)(cell < 0 ? 255 : cell);
    if (cell >= 0) atomicAdd(&s_hist[cell], 1);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < kMaxCells; c += blockDim.x)
    if (s_hist[c]) atomicAdd(&cell_count[(size_t)k * kMaxCells + c], s_hist[c]);
}

int launch_ground_tag(sloam_ctx *c, int K, const sloam_point *ground, const int32_t *ground_count,
                      int stride) {
  SB_CUDA(c, cudaMemsetAsync(c->ws.cell_count, 0, sizeof(int32_t) * (size_t)K * kMaxCells, c->stream));
  dim3 grid((unsigned)std::min(64, (stride + 255) / 256), (unsigned)K);
  ground_tag_kernel<<<grid, 256, 0, c->stream>>>(c->dp, K, ground, ground_count, stride,
                                                 c->ws.ground_cell, c->ws.cell_count);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

int launch_project_split(sloam_ctx *c, int K, bool do_project, bool do_split,
                         const sloam_point *points, const uint8_t *mask, int32_t *pix,
                         float *range_image, sloam_point *tree, sloam_point *ground,
                         int32_t *ground_count, uint32_t *tree_bits, bool sparse_tree) {
  // fused pipeline: projection + split, sparse tree cloud, ground points as per-tile cell records
  const bool fused = do_project && do_split && sparse_tree && ground == nullptr;
  const int N = c->hp.N;
  const int tiles = (N + kSplitTile - 1) / kSplitTile;
  const long long total = (long long)K * N;
  unsigned *rb = reinterpret_cast<unsigned *>(range_image);
  // destagger: the kernel writes the masked cloud into scratch, a second pass shifts it
  const bool destagger = do_split && c->hp.p.do_destagger != 0;
  sloam_point *tree_final = tree;
  uint32_t *bits_final = tree_bits;
  if (destagger) {
    if (!c->ws.tree2) return set_err(c, SLOAM_E_INVALID, "do_destagger was not enabled when the context was created");
    tree = c->ws.tree2;
    if (tree_bits) tree_bits = c->ws.tree_bits2;
  }
  if (do_project && rb) SB_CUDA(c, cudaMemsetAsync(rb, 0xFF, sizeof(unsigned) * total, c->stream));
  // The kernel always writes the tile-strided ground layout.  The fused pipeline consumes it
  // as is (ground == ws.ground); a stage entry gets the contiguous cloud by compaction.
  const bool strided_out = fused;
  bool counters_prezeroed = false;
  if (do_split) {
    if ((c->zero_valid & 1u) && ground_count == c->ws.ground_count) {
      c->zero_valid &= ~1u;  // zeroed with the rest of the counters (pipeline.cu)
      counters_prezeroed = true;
    } else {
      SB_CUDA(c, cudaMemsetAsync(ground_count, 0, sizeof(int32_t) * (size_t)K, c->stream));
      SB_CUDA(c, cudaMemsetAsync(c->ws.cell_count, 0, sizeof(int32_t) * (size_t)K * kMaxCells, c->stream));
    }
  }
  // persistent CTAs: exactly the resident ones, tiles handed out by a counter.  The counter is the
  // first word of the per-run zero block; outside a fused run it is cleared here.
  int *tile_ctr = c->ws.zero_begin;
  if (!counters_prezeroed) SB_CUDA(c, cudaMemsetAsync(tile_ctr, 0, sizeof(int), c->stream));
  static int occ[4] = {0, 0, 0, 0};
  const int which = fused ? 3 : ((do_project && do_split) ? 0 : (do_project ? 1 : 2));
  if (occ[which] == 0) {
    if (which == 0) SB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], project_split_kernel<true, true, false>, kThreads, 0));
    else if (which == 1) SB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], project_split_kernel<true, false, false>, kThreads, 0));
    else if (which == 2) SB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], project_split_kernel<false, true, false>, kThreads, 0));
    else SB_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[3], project_split_kernel<true, true, true>, kThreads, 0));
    if (occ[which] < 1) occ[which] = 1;
  }
  const long long all_tiles = (long long)K * tiles;
  const unsigned grid = (unsigned)std::min<long long>(all_tiles, (long long)c->sm_count * occ[which]);
#define SB_K1_ARGS c->dp, K, points, mask, pix, rb, tree, c->ws.ground, ground_count, c->ws.ground_cell, \
                   c->ws.cell_count, c->ws.tile_count, N, tree_bits, sparse_tree ? 1 : 0,                 \
                   reinterpret_cast<uint2 *>(c->ws.gscratch), c->ws.seg_tab, tile_ctr
  PROF_BEGIN(c, P_SPLIT);
  if (fused) project_split_kernel<true, true, true><<<grid, kThreads, 0, c->stream>>>(SB_K1_ARGS);
  else if (do_project && do_split) project_split_kernel<true, true, false><<<grid, kThreads, 0, c->stream>>>(SB_K1_ARGS);
  else if (do_project) project_split_kernel<true, false, false><<<grid, kThreads, 0, c->stream>>>(SB_K1_ARGS);
  else project_split_kernel<false, true, false><<<grid, kThreads, 0, c->stream>>>(SB_K1_ARGS);
#undef SB_K1_ARGS
  PROF_END(c, P_SPLIT);
  SB_LAUNCH_CHECK(c);
  if (destagger) {
    if (sparse_tree) {
      const long long words = (long long)K * ((N + 31) / 32);
      const unsigned dgrid = (unsigned)std::min<long long>((words + 7) / 8, (long long)c->sm_count * 8);
      destagger_sparse_kernel<<<dgrid, 256, 0, c->stream>>>(c->dp, K, tree, tree_bits, tree_final, bits_final);
    } else {
      destagger_dense_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c->stream>>>(c->dp, K, tree, tree_final);
    }
    SB_LAUNCH_CHECK(c);
  }
  if (do_split && !strided_out) {
    ground_compact_kernel<<<dim3((unsigned)tiles, (unsigned)K), 256, 0, c->stream>>>(
        c->dp, c->ws.ground, N, c->ws.tile_count, ground, N);
    SB_LAUNCH_CHECK(c);
  }
  if (do_project && rb) {
    const long long nvec = (total + 3) / 4;
    PROF_BEGIN(c, P_RANGE_FIN);
    range_finalize_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, c->stream>>>(rb, total);
    PROF_END(c, P_RANGE_FIN);
    SB_LAUNCH_CHECK(c);
  }
  return SLOAM_OK;
}

int launch_ground_compact(sloam_ctx *c, int K, sloam_point *dst) {
  const int N = c->hp.N, tiles = (N + kSplitTile - 1) / kSplitTile;
  ground_compact_kernel<<<dim3((unsigned)tiles, (unsigned)K), 256, 0, c->stream>>>(c->dp, c->ws.ground, N,
                                                                                   c->ws.tile_count, dst, N);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sloam_b200_project_dev(sloam_ctx *c, int K, const sloam_point *points, int32_t *pix,
                           float *range_image) {
  if (!c || K <= 0 || K > c->max_k || !points || !pix) return set_err(c, SLOAM_E_INVALID, "project: bad arguments");
  return launch_project_split(c, K, true, false, points, nullptr, pix, range_image, nullptr, nullptr, nullptr,
                              nullptr, false);
}

int sloam_b200_mask_cloud_dev(sloam_ctx *c, int K, const sloam_point *points, const int32_t *pix,
                              const uint8_t *mask, sloam_point *tree, sloam_point *ground,
                              int32_t *ground_count) {
  if (!c || K <= 0 || K > c->max_k || !points || !pix || !mask || !tree || !ground || !ground_count)
    return set_err(c, SLOAM_E_INVALID, "mask_cloud: bad arguments");
  return launch_project_split(c, K, false, true, points, mask, const_cast<int32_t *>(pix), nullptr,
                              tree, ground, ground_count, nullptr, false);
}

int sloam_b200_project_split_dev(sloam_ctx *c, int K, const sloam_point *points, const uint8_t *mask,
                                 int32_t *pix, float *range_image, sloam_point *tree,
                                 sloam_point *ground, int32_t *ground_count) {
  if (!c || K <= 0 || K > c->max_k || !points || !mask || !pix || !tree || !ground || !ground_count)
    return set_err(c, SLOAM_E_INVALID, "project_split: bad arguments");
  return launch_project_split(c, K, true, true, points, mask, pix, range_image, tree, ground, ground_count,
                              nullptr, false);
}

}  // extern "C"
