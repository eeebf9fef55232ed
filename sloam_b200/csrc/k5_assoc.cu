// k5_assoc.cu -- stages a11..a13: brute-force landmark data association.
//
// Replaces the argmin loops of sloam::matchFeatures<Cylinder> and
// sloam::matchModels (sloam/src/core/sloam.cpp:257-328) with
// Cylinder::distance(model) (sloam/src/objects/cylinder.cpp:175-194) and
// Cylinder::project (cylinder.cpp:205-211).
//
// One thread per detection, the map streamed through shared memory in tiles of
// pre-evaluated sample points (the three points of each axis at heights 0, 3,
// 6 m), fp64 throughout with the reference's operation order so the argmin
// (strict <, first minimum) is bit-exact.  Large maps are split over
// gridDim.y with a second pass picking the best partial (ties -> lower index).
// FP64-ALU bound: 56 (T + M) bytes in, 12 T bytes out.
#include "common.cuh"

namespace sb {

constexpr int kAssocThreads = 128;
constexpr int kMapTile = 64;

struct Samples { double p[9]; };  // points at heights 0, 3, 6

// root + ((h - root.z) / ray.z) * ray for h in {0, 3, 6}  (cylinder.cpp:184-190)
__device__ __forceinline__ void axis_samples(const double root[3], const double ray[3], Samples &s) {
#pragma unroll
  for (int h = 0; h < 3; ++h) {
    const double t = ((double)(3 * h) - root[2]) / ray[2];
    s.p[3 * h + 0] = root[0] + t * ray[0];
    s.p[3 * h + 1] = root[1] + t * ray[1];
    s.p[3 * h + 2] = root[2] + t * ray[2];
  }
}

__global__ void __launch_bounds__(kAssocThreads)
assoc_kernel(int T, const sloam_cylinder *__restrict__ det, const int32_t *__restrict__ n_det, int det_stride,
             const sloam_pose *__restrict__ tf, const sloam_cylinder *__restrict__ map,
             const int32_t *__restrict__ n_map, int map_stride, int map_shared, int splits,
             int32_t *__restrict__ best_index, double *__restrict__ best_dist) {
  __shared__ Samples s_map[kMapTile];
  const int k = blockIdx.z, split = blockIdx.y;
  const int nd = min(n_det[k], det_stride);
  const int nm = min(map_shared ? n_map[0] : n_map[k], map_stride);
  const int i = blockIdx.x * kAssocThreads + threadIdx.x;
  if (blockIdx.x * kAssocThreads >= nd) return;
  const sloam_cylinder *mk = map_shared ? map : map + (size_t)k * map_stride;
  Samples me;
  const bool active = i < nd;
  if (active) {
    sloam_cylinder c = det[(size_t)k * det_stride + i];
    if (tf) {  // Cylinder::project: root' = T root ; ray' = T (root + ray) - root'
      const sloam_pose T0 = tf[k];
      double other[3] = {c.root[0] + c.ray[0], c.root[1] + c.ray[1], c.root[2] + c.ray[2]};
      double r2[3], o2[3];
      pose_apply(T0, c.root, r2);
      pose_apply(T0, other, o2);
      for (int a = 0; a < 3; ++a) { c.root[a] = r2[a]; c.ray[a] = o2[a] - r2[a]; }
    }
    axis_samples(c.root, c.ray, me);
  }
  const int per = (nm + splits - 1) / splits;
  const int m0 = split * per, m1 = min(nm, m0 + per);
  double bd = INFINITY;
  int bi = -1;
  for (int base = m0; base < m1; base += kMapTile) {
    __syncthreads();
    if (threadIdx.x < kMapTile && base + threadIdx.x < m1) {
      const sloam_cylinder c = mk[base + threadIdx.x];
      axis_samples(c.root, c.ray, s_map[threadIdx.x]);
    }
    __syncthreads();
    if (active) {
      const int cnt = min(kMapTile, m1 - base);
      for (int j = 0; j < cnt; ++j) {
        const Samples &mo = s_map[j];
        double d = 0.0;
#pragma unroll
        for (int h = 0; h < 3; ++h) {
          // (modelPoint - tgtPoint).norm() with model = map object, tgt = detection
          const double dx = mo.p[3 * h] - me.p[3 * h], dy = mo.p[3 * h + 1] - me.p[3 * h + 1],
                       dz = mo.p[3 * h + 2] - me.p[3 * h + 2];
          d += sqrt(dx * dx + (dy * dy + dz * dz));
        }
        d = d / 3.0;
        if (d < bd) { bd = d; bi = base + j; }
      }
    }
  }
  if (active) {
    const size_t o = ((size_t)k * T + i) * splits + split;
    best_index[o] = bi;
    best_dist[o] = bd;
  }
}

__global__ void assoc_reduce_kernel(int T, const int32_t *__restrict__ n_det, int splits,
                                    const int32_t *__restrict__ pi, const double *__restrict__ pd,
                                    int32_t *__restrict__ best_index, double *__restrict__ best_dist,
                                    int out_stride) {
  const int k = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_det[k]) return;
  double bd = INFINITY;
  int bi = -1;
  for (int s = 0; s < splits; ++s) {  // ascending map ranges: strict < keeps the first minimum
    const size_t o = ((size_t)k * T + i) * splits + s;
    if (pd[o] < bd) { bd = pd[o]; bi = pi[o]; }
  }
  best_index[(size_t)k * out_stride + i] = bi;
  best_dist[(size_t)k * out_stride + i] = bd;
}

// matchFeatures<Plane> (sloam.cpp:257-286): Plane::project moves the centroid
// (plane.cpp:174), Plane::distance(model) is the centroid distance (plane.cpp:131-134).
// One thread per current plane; the few dozen map planes are read through L1.
__global__ void __launch_bounds__(kAssocThreads)
assoc_planes_kernel(const sloam_plane *__restrict__ det, const int32_t *__restrict__ n_det, int det_stride,
                    const sloam_pose *__restrict__ tf, const sloam_plane *__restrict__ map,
                    const int32_t *__restrict__ n_map, int map_stride, int32_t *__restrict__ best_index,
                    double *__restrict__ best_dist) {
  const int k = blockIdx.y;
  const int i = blockIdx.x * kAssocThreads + threadIdx.x;
  if (i >= min(n_det[k], det_stride)) return;
  const int nm = min(n_map[k], map_stride);
  double c2[3];
  const double *cen = det[(size_t)k * det_stride + i].centroid;
  if (tf) pose_apply(tf[k], cen, c2);
  else { c2[0] = cen[0]; c2[1] = cen[1]; c2[2] = cen[2]; }
  double bd = INFINITY;
  int bi = -1;
  for (int j = 0; j < nm; ++j) {
    const double *pc = map[(size_t)k * map_stride + j].centroid;
    const double dx = pc[0] - c2[0], dy = pc[1] - c2[1], dz = pc[2] - c2[2];
    const double d = sqrt(dx * dx + (dy * dy + dz * dz));
    if (d < bd) { bd = d; bi = j; }
  }
  best_index[(size_t)k * det_stride + i] = bi;
  best_dist[(size_t)k * det_stride + i] = bd;
}

int launch_associate(sloam_ctx *c, int K, const sloam_cylinder *det, const int32_t *n_det, int det_stride,
                     int det_cap, const sloam_pose *tf, const sloam_cylinder *map, const int32_t *n_map,
                     int map_stride, int map_shared, int map_cap, int32_t *best_index, double *best_dist) {
  // split large maps so that the grid fills the GPU
  int splits = 1;
  const int det_tiles = (det_cap + kAssocThreads - 1) / kAssocThreads;
  if (map_cap > 4096) {
    splits = std::min(64, std::max(1, (c->sm_count * 4) / std::max(1, det_tiles * K)));
    splits = std::min(splits, (map_cap + kMapTile - 1) / kMapTile);
  }
  if (splits > 1) {
    const size_t need = (size_t)K * det_cap * splits;
    if (c->assoc_part_cap < need) {
      if (c->assoc_part_i) { cudaFree(c->assoc_part_i); cudaFree(c->assoc_part_d); }
      SB_CUDA(c, cudaMalloc((void **)&c->assoc_part_i, need * sizeof(int32_t)));
      SB_CUDA(c, cudaMalloc((void **)&c->assoc_part_d, need * sizeof(double)));
      c->assoc_part_cap = need;
    }
  }
  dim3 grid((unsigned)det_tiles, (unsigned)splits, (unsigned)K);
  if (splits == 1) {
    assoc_kernel<<<grid, kAssocThreads, 0, c->stream>>>(det_cap, det, n_det, det_stride, tf, map, n_map,
                                                        map_stride, map_shared, 1, best_index, best_dist);
    SB_LAUNCH_CHECK(c);
  } else {
    assoc_kernel<<<grid, kAssocThreads, 0, c->stream>>>(det_cap, det, n_det, det_stride, tf, map, n_map,
                                                        map_stride, map_shared, splits, c->assoc_part_i,
                                                        c->assoc_part_d);
    SB_LAUNCH_CHECK(c);
    dim3 g2((unsigned)((det_cap + 127) / 128), (unsigned)K);
    assoc_reduce_kernel<<<g2, 128, 0, c->stream>>>(det_cap, n_det, splits, c->assoc_part_i, c->assoc_part_d,
                                                   best_index, best_dist, det_cap);
    SB_LAUNCH_CHECK(c);
  }
  return SLOAM_OK;
}

}  // namespace sb

using namespace sb;

extern "C" int sloam_b200_associate_dev(sloam_ctx *c, int K, const sloam_cylinder *det, const int32_t *n_det,
                                        int det_stride, const sloam_pose *tf, const sloam_cylinder *map,
                                        const int32_t *n_map, int map_stride, int map_shared,
                                        int32_t *best_index, double *best_dist) {
  if (!c || K <= 0 || !det || !n_det || !map || !n_map || !best_index || !best_dist || det_stride <= 0 ||
      map_stride <= 0)
    return set_err(c, SLOAM_E_INVALID, "associate: bad arguments");
  // outputs are [K][det_stride]
  return launch_associate(c, K, det, n_det, det_stride, det_stride, tf, map, n_map, map_stride, map_shared,
                          map_stride, best_index, best_dist);
}

extern "C" int sloam_b200_associate_planes_dev(sloam_ctx *c, int K, const sloam_plane *det, const int32_t *n_det,
                                               int det_stride, const sloam_pose *tf, const sloam_plane *map,
                                               const int32_t *n_map, int map_stride, int32_t *best_index,
                                               double *best_dist) {
  if (!c || K <= 0 || !det || !n_det || !map || !n_map || !best_index || !best_dist || det_stride <= 0 || map_stride <= 0)
    return set_err(c, SLOAM_E_INVALID, "associate_planes: bad arguments");
  dim3 grid((unsigned)((det_stride + kAssocThreads - 1) / kAssocThreads), (unsigned)K);
  assoc_planes_kernel<<<grid, kAssocThreads, 0, c->stream>>>(det, n_det, det_stride, tf, map, n_map, map_stride,
                                                             best_index, best_dist);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}
