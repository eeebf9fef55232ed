// k4_cylinder.cu -- stages a8 + a9 + a10: nearest ground plane per tree and the
// Cylinder landmark model.
//
// Replaces the landmark loop of sloam::computeModels (sloam/src/core/sloam.cpp:
// 418-435) and Cylinder::Cylinder / computeModel / groundBasedRoot / filter
// (sloam/src/objects/cylinder.cpp:3-173) including the PCL 1.10 line RANSAC it
// calls (SACSegmentation, SACMODEL_LINE, SAC_RANSAC, threshold 0.25,
// optimizeCoefficients).
//
// One warp per tree.  The tree's <= 64 vertex medians are staged in shared
// memory with a TMA bulk copy (cp.async.bulk + mbarrier); hypotheses are scored
// 32 at a time, one per lane (float32, same operation order as PCL's
// Vector4f code, --fmad=false, so inlier counts are bit-exact); the sequential
// semantics of PCL's loop (strict first maximum, adaptive k, skip counting) are
// replayed over the batch results; the winning model's inlier set is a 64-bit
// ballot mask.  Not HBM-bound: ~12 V bytes in, 88 + 16 F_t bytes out per tree;
// the stress mode (4096 hypotheses/tree) is FP32-ALU bound.
#include <float.h>

#include "common.cuh"

namespace sb {

constexpr int kCylWarps = 4;
constexpr int kMaxV = 64;

struct CylSmem {
  alignas(16) sloam_vertex vtx[kMaxV];  // 32 B each, bulk-copied
  float mx[kMaxV], my[kMaxV], mz[kMaxV];
  float radii[kMaxV];
  alignas(8) unsigned long long mbar;
};

// ---- PCL SampleConsensusModelLine pieces (float32) --------------------------
struct LineModel { float px, py, pz, dx, dy, dz; };

// computeModelCoefficients: false when the two samples coincide
__device__ __forceinline__ bool line_from_samples(const CylSmem &s, int i0, int i1, LineModel &m) {
  const float ax = s.mx[i0], ay = s.my[i0], az = s.mz[i0];
  const float bx = s.mx[i1], by = s.my[i1], bz = s.mz[i1];
  if (fabsf(ax - bx) <= FLT_EPSILON && fabsf(ay - by) <= FLT_EPSILON && fabsf(az - bz) <= FLT_EPSILON)
    return false;
  float x = bx - ax, y = by - ay, z = bz - az;
  const float sq = x * x + (y * y + z * z);  // tail<3>().normalize()
  if (sq > 0.f) { const float n = sqrtf(sq); x = x / n; y = y / n; z = z / n; }
  m.px = ax; m.py = ay; m.pz = az; m.dx = x; m.dy = y; m.dz = z;
  return true;
}

// countWithinDistance / selectWithinDistance: line_dir re-normalised as a Vector4f
// ((x^2 + z^2) + (y^2 + w^2), SSE2 packet reduction), cross3, squared norm -> double
struct LineScorer {
  float px, py, pz, dx, dy, dz;
  double sqr_thr;
  __device__ __forceinline__ LineScorer(const LineModel &m, double thr) {
    px = m.px; py = m.py; pz = m.pz;
    float x = m.dx, y = m.dy, z = m.dz;
    const float sq = (x * x + z * z) + (y * y + 0.0f * 0.0f);
    if (sq > 0.f) { const float n = sqrtf(sq); x = x / n; y = y / n; z = z / n; }
    dx = x; dy = y; dz = z;
    sqr_thr = thr * thr;
  }
  __device__ __forceinline__ bool inlier(float x, float y, float z) const {
    const float ax = px - x, ay = py - y, az = pz - z;
    const float cx = ay * dz - az * dy;
    const float cy = az * dx - ax * dz;
    const float cz = ax * dy - ay * dx;
    return (double)((cx * cx + cz * cz) + (cy * cy + 0.0f)) < sqr_thr;
  }
};

// pcl::computeRoots (float)
__device__ void compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.f) d = 0.f;
  const float sd = sqrtf(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}
__device__ void compute_roots(const float m[9], float roots[3]) {
  const float c0 = m[0] * m[4] * m[8] + 2.0f * m[1] * m[2] * m[5] - m[0] * m[5] * m[5] -
                   m[4] * m[2] * m[2] - m[8] * m[1] * m[1];
  const float c1 = m[0] * m[4] - m[1] * m[1] + m[0] * m[8] - m[2] * m[2] + m[4] * m[8] - m[5] * m[5];
  const float c2 = m[0] + m[4] + m[8];
  if (fabsf(c0) < FLT_EPSILON) { compute_roots2(c2, c1, roots); return; }
  const float s_inv3 = 1.0f / 3.0f;
  const float s_sqrt3 = sqrtf(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.f) a_over_3 = 0.f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.f) q = 0.f;
  const float rho = sqrtf(-a_over_3);
  const float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
  const float cos_theta = cosf(theta), sin_theta = sinf(theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) { const float t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
  if (roots[1] >= roots[2]) {
    const float t = roots[1]; roots[1] = roots[2]; roots[2] = t;
    if (roots[0] >= roots[1]) { const float u = roots[0]; roots[0] = roots[1]; roots[1] = u; }
  }
  if (roots[0] <= 0.f) compute_roots2(c2, c1, roots);
}

// optimizeModelCoefficients: PCA refit on the inliers (mask), sequential float sums
__device__ void refit_line(const CylSmem &s, int V, unsigned long long mask, int n_inl,
                           const LineModel &in, LineModel &out) {
  out = in;
  if (n_inl <= 2) return;
  float cx = 0.f, cy = 0.f, cz = 0.f;
  for (int i = 0; i < V; ++i)
    if ((mask >> i) & 1ull) { cx += s.mx[i]; cy += s.my[i]; cz += s.mz[i]; }
  const float nf = (float)n_inl;
  cx = cx / nf; cy = cy / nf; cz = cz / nf;
  float C[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = 0; i < V; ++i)
    if ((mask >> i) & 1ull) {
      float x = s.mx[i] - cx, y = s.my[i] - cy, z = s.mz[i] - cz;
      C[4] += y * y; C[5] += y * z; C[8] += z * z;
      const float sx = x;
      x = x * sx; y = y * sx; z = z * sx;
      C[0] += x; C[1] += y; C[2] += z;
    }
  C[3] = C[1]; C[6] = C[2]; C[7] = C[5];
  float scale = 0.f;
  for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(C[i]));
  if (scale <= FLT_MIN) scale = 1.0f;
  float S[9];
  for (int i = 0; i < 9; ++i) S[i] = C[i] / scale;
  float ev[3];
  compute_roots(S, ev);
  for (int i = 0; i < 3; ++i) ev[i] = ev[i] * scale;
  const float lam = ev[2] / scale;  // computeCorrespondingEigenVector(mat, ev[2])
  S[0] -= lam; S[4] -= lam; S[8] -= lam;
  float v1[3], v2[3], v3[3];
  v1[0] = S[1] * S[5] - S[2] * S[4]; v1[1] = S[2] * S[3] - S[0] * S[5]; v1[2] = S[0] * S[4] - S[1] * S[3];
  v2[0] = S[1] * S[8] - S[2] * S[7]; v2[1] = S[2] * S[6] - S[0] * S[8]; v2[2] = S[0] * S[7] - S[1] * S[6];
  v3[0] = S[4] * S[8] - S[5] * S[7]; v3[1] = S[5] * S[6] - S[3] * S[8]; v3[2] = S[3] * S[7] - S[4] * S[6];
  const float l1 = v1[0] * v1[0] + (v1[1] * v1[1] + v1[2] * v1[2]);
  const float l2 = v2[0] * v2[0] + (v2[1] * v2[1] + v2[2] * v2[2]);
  const float l3 = v3[0] * v3[0] + (v3[1] * v3[1] + v3[2] * v3[2]);
  const float *v; float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  else { v = v3; l = l3; }
  const float nl = sqrtf(l);
  out.px = cx; out.py = cy; out.pz = cz;
  out.dx = v[0] / nl; out.dy = v[1] / nl; out.dz = v[2] / nl;
}

__device__ __forceinline__ unsigned long long inlier_mask(const CylSmem &s, int V, const LineModel &m,
                                                          double thr) {
  const int lane = threadIdx.x & 31;
  const LineScorer sc(m, thr);
  const bool a = lane < V && sc.inlier(s.mx[lane], s.my[lane], s.mz[lane]);
  const bool b = lane + 32 < V && sc.inlier(s.mx[lane + 32], s.my[lane + 32], s.mz[lane + 32]);
  const unsigned lo = __ballot_sync(kFull, a), hi = __ballot_sync(kFull, b);
  return ((unsigned long long)hi << 32) | lo;
}

#ifndef SLOAM_CYL_MIN
#define SLOAM_CYL_MIN 8  // 64 registers: 141 -> 123 us per 1000 VLP-16 keyframes (6: 131 us)
#endif
__global__ void __launch_bounds__(kCylWarps * 32, SLOAM_CYL_MIN)
cylinder_kernel(const DevParams *__restrict__ dp, const sloam_tree *__restrict__ trees,
                const int32_t *__restrict__ n_trees, const sloam_vertex *__restrict__ vertices,
                const sloam_point *__restrict__ vpoints, int vstride, int pstride,
                const sloam_plane *__restrict__ planes, const int32_t *__restrict__ n_planes,
                const int32_t *__restrict__ pairs, const int32_t *__restrict__ pair_off,
                sloam_tree_model *__restrict__ models, sloam_point *__restrict__ features) {
  __shared__ CylSmem sm[kCylWarps];
  const sloam_params &P = dp->p;
  const int T = P.max_trees, B = dp->B, Ft = P.featuresPerTree;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = blockIdx.y;
  const int t = blockIdx.x * kCylWarps + warp;
  CylSmem &s = sm[warp];
  if (t >= n_trees[k]) return;  // warp-uniform
  const sloam_tree tr = trees[(size_t)k * T + t];
  // the draw tables cover up to max_tree_vertices samples (ctx.cu upload_tables); computeGraph never
  // emits more (trellis.cpp:125-127), caller-supplied trees are cut like it cuts them
  const int vcap = P.max_tree_vertices < kMaxV ? P.max_tree_vertices : kMaxV;
  const int V = tr.n_vertices < vcap ? tr.n_vertices : vcap;
  const sloam_vertex *vsrc = vertices + (size_t)k * vstride + tr.vertex_begin;
  const sloam_point *psrc = vpoints + (size_t)k * pstride;
  sloam_tree_model *out = models + (size_t)k * T + t;
  sloam_point *fout = features + ((size_t)k * T + t) * Ft;

  // ---- stage the vertex records: one TMA bulk copy per tree (V * 32 B) ----
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(&s.mbar);
  const unsigned dst = (unsigned)__cvta_generic_to_shared(&s.vtx[0]);
  const unsigned bytes = (unsigned)(V * sizeof(sloam_vertex));
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (lane == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst), "l"(vsrc), "r"(bytes), "r"(mbar) : "memory");
  }
  {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
          : "=r"(done) : "r"(mbar), "r"(0u) : "memory");
    }
  }
  for (int i = lane; i < V; i += 32) {
    s.mx[i] = s.vtx[i].cx; s.my[i] = s.vtx[i].cy; s.mz[i] = s.vtx[i].cz;
  }
  __syncwarp();

  sloam_tree_model m;
  for (int i = 0; i < 3; ++i) { m.model.root[i] = 0.0; m.model.ray[i] = 0.0; }
  m.model.radius = 0.0;
  m.id = tr.tree_id;  // vertices[2].treeId: every vertex of a tree carries the cluster label
  m.is_valid = 0; m.plane_index = -1; m.n_inliers = 0; m.best_hypothesis = -1;
  m.n_hypotheses = 0; m.n_refit_inliers = 0; m.reserved = 0;
  for (int f = lane; f < Ft; f += 32) fout[f] = sloam_point{0.f, 0.f, 0.f, 0.f};

  const int np = n_planes[k];
  if (np == 0 || V < 3) {  // computeModels returns before the landmark loop (sloam.cpp:414)
    if (lane == 0) *out = m;
    return;
  }
  // ---- nearest accepted plane to vertices[1].coords (sloam.cpp:420-431) ----
  const double ax = (double)s.mx[1], ay = (double)s.my[1], az = (double)s.mz[1];
  double best_d = 100000.0;
  int best_g = 0x7fffffff;
  for (int g = lane; g < np; g += 32) {
    const sloam_plane &pl = planes[(size_t)k * B + g];
    const double num = fabs(pl.plane[0] * ax + pl.plane[1] * ay + pl.plane[2] * az + pl.plane[3]);
    const double d = num / sqrt(pl.plane[0] * pl.plane[0] + (pl.plane[1] * pl.plane[1] + pl.plane[2] * pl.plane[2]));
    if (d < best_d) { best_d = d; best_g = g; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(kFull, best_d, o);
    const int og = __shfl_xor_sync(kFull, best_g, o);
    if (od < best_d || (od == best_d && og < best_g)) { best_d = od; best_g = og; }
  }
  if (best_g == 0x7fffffff) best_g = 0;  // every distance >= 100000 (or NaN): bestPlane = planes[0]
  m.plane_index = best_g;
  const sloam_plane gp = planes[(size_t)k * B + best_g];

  // ---- Cylinder::Cylinder gate (cylinder.cpp:6-16) ----
  {
    const float cxf = (float)gp.centroid[0], cyf = (float)gp.centroid[1];
    const float ddx = s.mx[1] - cxf, ddy = s.my[1] - cyf;
    const float d2 = sqrtf((float)((double)ddx * (double)ddx) + (float)((double)ddy * (double)ddy));
    if (!((double)d2 < P.maxLidarDist)) {
      if (lane == 0) *out = m;
      return;
    }
  }
  // ---- computeModel (cylinder.cpp:75-173) ----
  m.model.root[0] = ax; m.model.root[1] = ay; m.model.root[2] = az;
  bool model_ok = true;
  {
    const float dx = s.mx[V - 2] - s.mx[1], dy = s.my[V - 2] - s.my[1], dz = s.mz[V - 2] - s.mz[1];
    if ((double)sqnorm3f(dx, dy, dz) < P.min_tree_height_sq) model_ok = false;  // :109
  }
  LineModel best_model = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (model_ok) {
    const int32_t *tab = pairs + 2 * (size_t)pair_off[V];
    const int n_draws = pair_off[V + 1] - pair_off[V];
    const int fixed = P.ransac_fixed_hypotheses;
    const int max_it = P.ransac_max_iterations;
    const unsigned max_skip = (unsigned)max_it * 10u;
    const double log_prob = log(1.0 - P.ransac_probability);
    const double inv_n = 1.0 / (double)V;
    int iterations = 0, best = -2147483647, best_hyp = -1, best_draw = -1;
    unsigned skipped = 0;
    int bad_run = 0;  // consecutive not-good draws inside one getSamples call
    double kk = 1.0;
    bool stop = false;
    for (int d0 = 0; d0 < n_draws && !stop; d0 += 32) {
      const int d = d0 + lane;
      int status = 0, cnt = 0;  // 0 not good, 1 degenerate, 2 hypothesis
      if (d < n_draws) {
        const int i0 = tab[2 * d], i1 = tab[2 * d + 1];
        // isSampleGood (PCL 1.10): all three coordinates differ
        if (s.mx[i0] != s.mx[i1] && s.my[i0] != s.my[i1] && s.mz[i0] != s.mz[i1]) {
          LineModel lm;
          if (line_from_samples(s, i0, i1, lm)) {
            status = 2;
            const LineScorer sc(lm, P.ransac_threshold);
            for (int i = 0; i < V; ++i) cnt += sc.inlier(s.mx[i], s.my[i], s.mz[i]) ? 1 : 0;
          } else status = 1;
        }
      } else status = -1;
      if (fixed > 0) {
        // stress mode: exactly `fixed` hypotheses, strict first maximum
        const unsigned hb = __ballot_sync(kFull, status == 2);
        const int hyp = iterations + __popc(hb & ((1u << lane) - 1u));
        int c = (status == 2 && hyp < fixed) ? cnt : -2147483647;
        int h = (status == 2 && hyp < fixed) ? hyp : 0x7fffffff;
        int dd = d;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const int oc = __shfl_xor_sync(kFull, c, o), oh = __shfl_xor_sync(kFull, h, o);
          const int od = __shfl_xor_sync(kFull, dd, o);
          if (oc > c || (oc == c && oh < h)) { c = oc; h = oh; dd = od; }
        }
        if (c > best) { best = c; best_hyp = h; best_draw = dd; }
        iterations += __popc(hb);
        if (iterations >= fixed) { iterations = fixed; stop = true; }
      } else {
        // replay RandomSampleConsensus::computeModel over the batch, in draw order
        for (int e = 0; e < 32 && !stop; ++e) {
          const int st = __shfl_sync(kFull, status, e);
          const int c = __shfl_sync(kFull, cnt, e);
          if (st < 0) { stop = true; break; }
          if (!((double)iterations < kk && skipped < max_skip)) { stop = true; break; }
          if (st == 0) { if (++bad_run >= 1000) stop = true; continue; }
          bad_run = 0;
          if (st == 1) { ++skipped; continue; }
          if (c > best) {
            best = c; best_hyp = iterations; best_draw = d0 + e;
            const double w = (double)best * inv_n;
            double p_no = 1.0 - w * w;
            p_no = fmax(DBL_EPSILON, p_no);
            p_no = fmin(1.0 - DBL_EPSILON, p_no);
            kk = log_prob / log(p_no);
          }
          ++iterations;
          if (iterations > max_it) stop = true;
        }
      }
    }
    m.n_hypotheses = iterations;
    m.best_hypothesis = best_hyp;
    if (best_draw < 0) model_ok = false;
    else {
      m.n_inliers = best;
      line_from_samples(s, tab[2 * best_draw], tab[2 * best_draw + 1], best_model);
    }
  }
  if (model_ok) {
    const unsigned long long mask = inlier_mask(s, V, best_model, P.ransac_threshold);
    LineModel refined;
    refit_line(s, V, mask, __popcll(mask), best_model, refined);  // every lane computes the same
    const unsigned long long mask2 = inlier_mask(s, V, refined, P.ransac_threshold);
    m.n_refit_inliers = __popcll(mask2);
    if (mask2 == 0ull) model_ok = false;  // cylinder.cpp:126-131
    m.model.ray[0] = (double)refined.dx; m.model.ray[1] = (double)refined.dy; m.model.ray[2] = (double)refined.dz;
  }
  if (!model_ok) {
    m.model.radius = -1.0;
    m.model.ray[0] = m.model.ray[1] = m.model.ray[2] = 0.0;
  } else {
    // ---- radius statistic (cylinder.cpp:141-163): ascending radii of the vertices with more
    // than 3 points, by rank (every lane ranks its vertices against all others; only the
    // sorted VALUES matter, so ties may land in any order) ----
    int nr = 0;
    for (int base = 0; base < V; base += 32) {
      const int i = base + lane;
      const bool use = i < V && s.vtx[i].n_points > 3;
      const float r = use ? s.vtx[i].radius : 0.f;
      int rank = 0;
      if (use)
        for (int j = 0; j < V; ++j) {
          const float rj = s.vtx[j].radius;
          rank += (s.vtx[j].n_points > 3) && (rj < r || (rj == r && j < i));
        }
      if (use) s.radii[rank] = r;
      nr += __popc(__ballot_sync(kFull, use));
    }
    __syncwarp();
    double radius = -1.0;
    if (nr > 0) {
      int d = 0;
      while (d < nr && !(s.radii[d] > 0.f)) ++d;
      int middle = (d + 1) + (nr - d) / 2;
      if (middle > nr - 1) middle = nr - 1;
      radius = (double)s.radii[middle];
      if (radius == 0.0) radius = -1.0;
      else if (radius < P.defaultTreeRadius) radius = P.defaultTreeRadius;
    }
    m.model.radius = radius;
    __syncwarp();
    // ---- features: first F_t vertex points bottom-up, intensity = tree id (:87-91,:172).
    // Feature f is point f - start[i] of the vertex i whose running point count covers f;
    // the running counts go to shared memory (radii scratch), the lanes take f = lane, +32, ...
    int *start = reinterpret_cast<int *>(s.radii);
    {
      int carry = 0;
      for (int base = 0; base < V; base += 32) {
        const int i = base + lane;
        const int v = i < V ? s.vtx[i].n_points : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(kFull, inc, o);
          if (lane >= o) inc += t;
        }
        if (i < V) start[i] = carry + inc - v;
        carry += __shfl_sync(kFull, inc, 31);
      }
      __syncwarp();
      const int total = carry < Ft ? carry : Ft;
      for (int f = lane; f < total; f += 32) {
        int i = 0;
        while (i + 1 < V && start[i + 1] <= f) ++i;  // last vertex with start <= f (empty ones skipped)
        sloam_point p = psrc[s.vtx[i].point_begin + (f - start[i])];
        p.intensity = (float)tr.tree_id;
        fout[f] = p;
      }
    }
  }
  // ---- groundBasedRoot (cylinder.cpp:41-55) + rayPlaneIntersection (utils.h:41-52) ----
  bool validZ = false;
  {
    const double *g = gp.plane;
    const double nn = sqrt(g[0] * g[0] + (g[1] * g[1] + g[2] * g[2]));
    const float dist = (float)(fabs(g[0] * m.model.root[0] + g[1] * m.model.root[1] + g[2] * m.model.root[2] + g[3]) / nn);
    if ((double)dist < P.root_plane_max_dist) {
      const float denom = (float)(g[0] * m.model.ray[0] + (g[1] * m.model.ray[1] + g[2] * m.model.ray[2]));
      if (fabsf(denom) > 0.001f) {
        const double dcx = gp.centroid[0] - m.model.root[0], dcy = gp.centroid[1] - m.model.root[1],
                     dcz = gp.centroid[2] - m.model.root[2];
        const float tt = (float)((dcx * g[0] + (dcy * g[1] + dcz * g[2])) / (double)denom);
        if (tt >= 0.001f)
          for (int i = 0; i < 3; ++i) m.model.root[i] = m.model.root[i] + (double)tt * m.model.ray[i];
      }
      validZ = true;
    }
  }
  const bool validNorm = sqrt(m.model.root[0] * m.model.root[0] +
                              (m.model.root[1] * m.model.root[1] + m.model.root[2] * m.model.root[2])) > 0.01;
  const bool validRadius = m.model.radius > 0.0;
  bool validTree = false;
  if (m.model.radius != -1.0) {  // filter, cylinder.cpp:57-73
    const double *g = gp.plane;
    const double rn = sqrt(dot3d(m.model.ray, m.model.ray)), un = sqrt(g[0] * g[0] + (g[1] * g[1] + g[2] * g[2]));
    const double theta = (180.0 / 3.14159265) * acos((m.model.ray[0] * g[0] + (m.model.ray[1] * g[1] + m.model.ray[2] * g[2])) / (rn * un));
    validTree = m.model.radius < P.maxTreeRadius && (theta <= P.maxAxisTheta || theta >= 180.0 - P.maxAxisTheta);
  }
  m.is_valid = (validZ && validRadius && validTree && validNorm) ? 1 : 0;
  if (lane == 0) *out = m;
}

// valid cylinders of each keyframe, compact, in tree order (computeModels :433-434)
__global__ void cylinders_compact_kernel(const DevParams *__restrict__ dp, const int32_t *__restrict__ n_trees,
                                         const sloam_tree_model *__restrict__ models,
                                         sloam_cylinder *__restrict__ lm_cyl, int32_t *__restrict__ lm_src,
                                         int32_t *__restrict__ n_lm) {
  const int T = dp->p.max_trees;
  const int k = blockIdx.x, lane = threadIdx.x;
  const int nt = n_trees[k];
  int cnt = 0;
  for (int base = 0; base < nt; base += 32) {
    const int t = base + lane;
    const bool v = t < nt && models[(size_t)k * T + t].is_valid;
    const unsigned b = __ballot_sync(kFull, v);
    if (v) {
      const int pos = cnt + __popc(b & ((1u << lane) - 1u));
      lm_cyl[(size_t)k * T + pos] = models[(size_t)k * T + t].model;
      lm_src[(size_t)k * T + pos] = t;
    }
    cnt += __popc(b);
  }
  if (lane == 0) n_lm[k] = cnt;
}

int launch_cylinders_strided(sloam_ctx *c, int K, const sloam_tree *trees, const int32_t *n_trees,
                             const sloam_vertex *vertices, int vstride, const sloam_point *vpoints,
                             int pstride, const sloam_plane *planes_acc, const int32_t *n_planes_acc,
                             sloam_tree_model *models, sloam_point *features);

int launch_cylinders(sloam_ctx *c, int K, const sloam_tree *trees, const int32_t *n_trees,
                     const sloam_vertex *vertices, const sloam_point *vpoints,
                     const sloam_plane *planes_acc, const int32_t *n_planes_acc,
                     sloam_tree_model *models, sloam_point *features) {
  const sloam_params &p = c->hp.p;
  return launch_cylinders_strided(c, K, trees, n_trees, vertices, p.max_trees * p.max_tree_vertices, vpoints,
                                  c->hp.N, planes_acc, n_planes_acc, models, features);
}

int launch_cylinders_strided(sloam_ctx *c, int K, const sloam_tree *trees, const int32_t *n_trees,
                             const sloam_vertex *vertices, int vstride, const sloam_point *vpoints,
                             int pstride, const sloam_plane *planes_acc, const int32_t *n_planes_acc,
                             sloam_tree_model *models, sloam_point *features) {
  const sloam_params &p = c->hp.p;
  dim3 grid((unsigned)((p.max_trees + kCylWarps - 1) / kCylWarps), (unsigned)K);
  PROF_BEGIN(c, P_CYLINDER);
  cylinder_kernel<<<grid, kCylWarps * 32, 0, c->stream>>>(
      c->dp, trees, n_trees, vertices, vpoints, vstride, pstride, planes_acc,
      n_planes_acc, c->ws.ransac_pairs, c->ws.ransac_pairs_offset, models, features);
  SB_LAUNCH_CHECK(c);
  cylinders_compact_kernel<<<K, 32, 0, c->stream>>>(c->dp, n_trees, models, c->ws.lm_cyl, c->ws.lm_src, c->ws.n_lm);
  PROF_END(c, P_CYLINDER);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

int launch_planes_compact(sloam_ctx *c, int K, const sloam_cell_plane *cells);

}  // namespace sb

using namespace sb;

extern "C" int sloam_b200_cylinders_dev(sloam_ctx *c, int K, const sloam_tree *trees, const int32_t *n_trees,
                                        const sloam_vertex *vertices, const sloam_point *vertex_points,
                                        const sloam_cell_plane *cells, sloam_tree_model *models,
                                        sloam_point *features) {
  if (!c || K <= 0 || K > c->max_k || !trees || !n_trees || !vertices || !vertex_points || !cells || !models || !features)
    return set_err(c, SLOAM_E_INVALID, "cylinders: bad arguments");
  const int rc = launch_planes_compact(c, K, cells);
  if (rc != SLOAM_OK) return rc;
  return launch_cylinders(c, K, trees, n_trees, vertices, vertex_points, c->ws.planes_acc,
                          c->ws.n_planes_acc, models, features);
}
