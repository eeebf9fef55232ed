// synth.cu -- synthetic forest generator entry points (see synth.h).
#include <algorithm>
#include <vector>

#include "synth.h"

namespace sb {

static void build_scene(const sloam_synth_config &c, SynthScene *sc, std::vector<SynthTree> *trees) {
  // ground: slope about a seeded horizontal axis through the origin at z = 0
  const double slope = (double)c.ground_slope_deg * 0.017453292519943295;
  const double dir = 6.283185307179586 * hash_uniform(c.seed, 11, 0, 0);
  sc->gn[0] = (float)(sin(slope) * cos(dir));
  sc->gn[1] = (float)(sin(slope) * sin(dir));
  sc->gn[2] = (float)cos(slope);
  sc->gd = 0.0f;
  trees->clear();
  const double min_gap = 1.3;  // surface gap > CC threshold (1.0 m) + margin
  for (uint64_t attempt = 0; attempt < 200000 && (int)trees->size() < c.n_trees; ++attempt) {
    const double u = hash_uniform(c.seed, 21, attempt, 0), v = hash_uniform(c.seed, 21, attempt, 1);
    const double r2min = (double)c.tree_r_min * c.tree_r_min, r2max = (double)c.tree_r_max * c.tree_r_max;
    const double r = sqrt(r2min + u * (r2max - r2min));
    double ang = 6.283185307179586 * v;
    SynthTree t;
    t.cx = (float)(r * cos(ang));
    t.cy = (float)(r * sin(ang));
    t.cz = (float)(-(sc->gn[0] * t.cx + sc->gn[1] * t.cy + sc->gd) / sc->gn[2]);
    t.radius = (float)(c.trunk_radius_min +
                       (c.trunk_radius_max - c.trunk_radius_min) * hash_uniform(c.seed, 21, attempt, 2));
    const double tilt = (double)c.max_tilt_deg * 0.017453292519943295 * hash_uniform(c.seed, 21, attempt, 3);
    const double tdir = 6.283185307179586 * hash_uniform(c.seed, 21, attempt, 4);
    t.ax = (float)(sin(tilt) * cos(tdir));
    t.ay = (float)(sin(tilt) * sin(tdir));
    t.az = (float)cos(tilt);
    t.height = (float)(10.0 + 6.0 * hash_uniform(c.seed, 21, attempt, 5));
    bool ok = true;
    for (const SynthTree &o : *trees) {
      const double dx = o.cx - t.cx, dy = o.cy - t.cy;
      // trunks lean up to max_tilt over ~16 m of height: keep generous clearance
      if (sqrt(dx * dx + dy * dy) < min_gap + o.radius + t.radius + 2.0 * 16.0 * sin((double)c.max_tilt_deg * 0.0174533))
        ok = false;
    }
    if (ok) trees->push_back(t);
  }
  sc->n_trees = (int)trees->size();
}

__global__ void synth_kernel(sloam_synth_config c, SynthScene sc, const SynthTree *trees, int64_t k0,
                             int K, sloam_point *points, unsigned long long *best) {
  const int N = c.img_h * c.img_w;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)K * N) return;
  const int k = (int)(gid / N), i = (int)(gid % N);
  sloam_pose T;
  synth_gt_pose(c, k0 + k, &T);
  sloam_point p;
  const int label = synth_beam(c, sc, trees, T, k0 + k, i / c.img_w, i % c.img_w, &p);
  points[gid] = p;
  if (label != 0) {
    ProjGeom g;
    const float fu = (float)((double)c.fov_up_deg / 180.0 * 3.14159265358979323846);
    const float fd = (float)((double)c.fov_down_deg / 180.0 * 3.14159265358979323846);
    g.fov_down_abs = fabsf(fd); g.fov = fabsf(fd) + fabsf(fu);
    g.Wf = (float)c.img_w; g.Hf = (float)c.img_h;
    float range;
    const int pix = project_pixel(g, p.x, p.y, p.z, &range);
    if (range == range) {
      const unsigned long long key = ((unsigned long long)__float_as_uint(range) << 8) | (unsigned)label;
      atomicMin(&best[(long long)k * N + pix], key);
    }
  }
}

// No-return beams (NaN or all-zero points) all project to one "landing" pixel
// through the clamps of inference.cpp:119-127.  In the reference that pixel ends
// up labelled 0 (its range-image entry is 0/NaN, so _makeTensor/_mask mark it
// invalid, inference.cpp:183-186,295-297); the ground-truth mask does the same.
__host__ __device__ inline int landing_pixel(const sloam_synth_config &c) {
  return c.nan_no_return ? c.img_h * c.img_w - 1 : (c.img_h - 1) * c.img_w + c.img_w / 2;
}

__global__ void synth_mask_kernel(sloam_synth_config c, const unsigned long long *best, long long n,
                                  uint8_t *mask) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= n) return;
  const unsigned long long b = best[gid];
  const int i = (int)(gid % ((long long)c.img_h * c.img_w));
  mask[gid] = (b == ~0ull || i == landing_pixel(c)) ? 0 : (uint8_t)(b & 0xFF);
}

}  // namespace sb

using namespace sb;

extern "C" {

void sloam_synth_default_config(sloam_synth_config *c, int img_h, int img_w, int n_trees) {
  c->img_h = img_h; c->img_w = img_w;
  c->fov_up_deg = 22.5f; c->fov_down_deg = -22.5f;
  c->n_trees = n_trees;
  c->tree_r_min = 3.0f;
  c->tree_r_max = img_w >= 2048 ? 18.0f : 12.0f;
  c->trunk_radius_min = 0.10f; c->trunk_radius_max = 0.28f;
  c->max_tilt_deg = 3.0f;
  c->sensor_height = 3.4f;
  c->ground_slope_deg = 2.0f;
  c->ground_noise = 0.02f; c->range_noise = 0.01f;
  c->max_range = 60.0f;
  c->step_per_keyframe = 0.5f;
  c->azimuth_offset_cols = 0.37f;
  c->guess_sigma_t = 0.05f; c->guess_sigma_r = 0.0087f;
  c->nan_no_return = 1;
  c->seed = 20260000ull;
}

int sloam_synth_scene(const sloam_synth_config *c, sloam_cylinder *trees_out) {
  SynthScene sc;
  std::vector<SynthTree> trees;
  build_scene(*c, &sc, &trees);
  for (size_t i = 0; i < trees.size(); ++i) {
    sloam_cylinder m;
    m.root[0] = trees[i].cx; m.root[1] = trees[i].cy; m.root[2] = trees[i].cz;
    m.ray[0] = trees[i].ax; m.ray[1] = trees[i].ay; m.ray[2] = trees[i].az;
    m.radius = trees[i].radius;
    trees_out[i] = m;
  }
  return (int)trees.size();
}

void sloam_synth_pose(const sloam_synth_config *c, int64_t k, sloam_pose *gt, sloam_pose *guess) {
  sloam_pose T;
  synth_gt_pose(*c, k, &T);
  if (gt) *gt = T;
  if (!guess) return;
  // guess = GT o exp(N(0, sigma)) : small rotation vector + translation in the sensor frame
  double w[3], dt[3];
  for (int a = 0; a < 3; ++a) {
    w[a] = (double)c->guess_sigma_r * hash_normal(c->seed, 31, (uint64_t)k, a);
    dt[a] = (double)c->guess_sigma_t * hash_normal(c->seed, 32, (uint64_t)k, a);
  }
  const double th = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double dq[4] = {0, 0, 0, 1};
  if (th > 0) {
    const double s = sin(0.5 * th) / th;
    dq[0] = s * w[0]; dq[1] = s * w[1]; dq[2] = s * w[2]; dq[3] = cos(0.5 * th);
  }
  // q = T.q * dq
  const double *a = T.q;
  guess->q[3] = a[3] * dq[3] - a[0] * dq[0] - a[1] * dq[1] - a[2] * dq[2];
  guess->q[0] = a[3] * dq[0] + a[0] * dq[3] + a[1] * dq[2] - a[2] * dq[1];
  guess->q[1] = a[3] * dq[1] + a[1] * dq[3] + a[2] * dq[0] - a[0] * dq[2];
  guess->q[2] = a[3] * dq[2] + a[2] * dq[3] + a[0] * dq[1] - a[1] * dq[0];
  double rt[3];
  q_rotate(T.q, dt, rt);
  for (int i = 0; i < 3; ++i) guess->t[i] = T.t[i] + rt[i];
}

int sloam_synth_generate_host(const sloam_synth_config *c, int64_t k0, int K, sloam_point *points,
                              uint8_t *mask) {
  SynthScene sc;
  std::vector<SynthTree> trees;
  build_scene(*c, &sc, &trees);
  const int N = c->img_h * c->img_w;
  ProjGeom g;
  const float fu = (float)((double)c->fov_up_deg / 180.0 * 3.14159265358979323846);
  const float fd = (float)((double)c->fov_down_deg / 180.0 * 3.14159265358979323846);
  g.fov_down_abs = fabsf(fd); g.fov = fabsf(fd) + fabsf(fu);
  g.Wf = (float)c->img_w; g.Hf = (float)c->img_h;
  std::vector<unsigned long long> best((size_t)N);
  for (int k = 0; k < K; ++k) {
    std::fill(best.begin(), best.end(), ~0ull);
    sloam_pose T;
    synth_gt_pose(*c, k0 + k, &T);
    for (int i = 0; i < N; ++i) {
      sloam_point p;
      const int label = synth_beam(*c, sc, trees.data(), T, k0 + k, i / c->img_w, i % c->img_w, &p);
      points[(size_t)k * N + i] = p;
      if (label != 0) {
        float range;
        const int pix = project_pixel(g, p.x, p.y, p.z, &range);
        if (range == range) {
          uint32_t rb; memcpy(&rb, &range, 4);
          const unsigned long long key = ((unsigned long long)rb << 8) | (unsigned)label;
          best[pix] = std::min(best[pix], key);
        }
      }
    }
    for (int i = 0; i < N; ++i)
      mask[(size_t)k * N + i] = (best[i] == ~0ull || i == landing_pixel(*c)) ? 0 : (uint8_t)(best[i] & 0xFF);
  }
  return SLOAM_OK;
}

int sloam_synth_generate_dev(sloam_ctx *ctx, const sloam_synth_config *c, int64_t k0, int K,
                             sloam_point *points, uint8_t *mask) {
  if (!ctx || !c || K <= 0) return SLOAM_E_INVALID;
  SynthScene sc;
  std::vector<SynthTree> trees;
  build_scene(*c, &sc, &trees);
  const long long n = (long long)K * c->img_h * c->img_w;
  SynthTree *dtrees = nullptr;
  unsigned long long *best = nullptr;
  SB_CUDA(ctx, cudaMallocAsync((void **)&dtrees, sizeof(SynthTree) * std::max<size_t>(trees.size(), 1), ctx->stream));
  SB_CUDA(ctx, cudaMallocAsync((void **)&best, sizeof(unsigned long long) * n, ctx->stream));
  SB_CUDA(ctx, cudaMemcpyAsync(dtrees, trees.data(), sizeof(SynthTree) * trees.size(),
                               cudaMemcpyHostToDevice, ctx->stream));
  SB_CUDA(ctx, cudaMemsetAsync(best, 0xFF, sizeof(unsigned long long) * n, ctx->stream));
  const int threads = 256;
  const unsigned blocks = (unsigned)((n + threads - 1) / threads);
  synth_kernel<<<blocks, threads, 0, ctx->stream>>>(*c, sc, dtrees, k0, K, points, best);
  SB_LAUNCH_CHECK(ctx);
  synth_mask_kernel<<<blocks, threads, 0, ctx->stream>>>(*c, best, n, mask);
  SB_LAUNCH_CHECK(ctx);
  SB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // trees vector must outlive the copy
  SB_CUDA(ctx, cudaFreeAsync(dtrees, ctx->stream));
  SB_CUDA(ctx, cudaFreeAsync(best, ctx->stream));
  return SLOAM_OK;
}

}  // extern "C"
