// k7_map.cu -- SURVEY 8(f)-1/2: the semantic map (MapManager) and the sequential,
// state-carrying keyframe step (SLOAMNode::run minus ROS), device-resident.
//
// Replaces MapManager::getSubmap / updateMap / getMap
// (sloam/src/core/mapManager.cpp:8-71) and the call sequence of SLOAMNode::run
// (sloam/src/core/sloamNode.cpp:186-282): getSubmap -> run -> maskCloud x2 ->
// computeGraph -> RunSloam -> updateMap, with firstScan_ / prevGPlanes_
// (sloam/include/core/sloam.h:99-106) kept on the device between calls.
//
// getSubmap: pcl::KdTreeFLANN kNN(100) around (pose.x, pose.y, 1) is an exact
// search; here a brute-force pass (float squared L2 in FLANN's summation order)
// + 8-bit radix select + rank sort of the <= 100 winners, ties by lower index,
// then the "last 200 landmarks" filter in neighbour order.  One CTA: the map is
// at most ~1e5 roots (1.6 MB, L2-resident); this path is sequential by nature.
#include <cstdlib>
#include "common.cuh"

namespace sb {

constexpr int kMapThreads = 1024;
constexpr int kMapKnnMax = 128;

struct MapState {
  sloam_cylinder *models = nullptr;
  float4 *roots = nullptr;       // x, y, z of model.root as float (landmarks_), w unused
  int32_t *ids = nullptr;
  int32_t *hits = nullptr;
  uint32_t *keys = nullptr;      // scratch: distance keys of the current query
  int32_t *size = nullptr;       // device counter
  int32_t *matches_map = nullptr;  // [kMapKnnMax] submap index -> map index
  int32_t *n_matches_map = nullptr;
  int capacity = 0;
  int knn = 100, recent = 200;   // mapManager.cpp:54,60
};

__device__ __forceinline__ int map_block_scan(int v, int *s_warp, int *total) {
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
  for (int w = 0; w < kMapThreads / 32; ++w) {
    const int c = s_warp[w];
    if (w < warp) off += c;
    tot += c;
  }
  __syncthreads();
  *total = tot;
  return off + inc - v;
}

__global__ void __launch_bounds__(kMapThreads)
map_submap_kernel(MapState m, const sloam_pose *__restrict__ pose, sloam_cylinder *__restrict__ submap,
                  int32_t *__restrict__ n_submap, int submap_cap) {
  __shared__ int s_hist[256];
  __shared__ int s_warp[kMapThreads / 32];
  __shared__ int s_misc[4];
  __shared__ unsigned long long s_cand[kMapKnnMax];  // (key << 32) | index
  __shared__ unsigned long long s_sorted[kMapKnnMax];
  const int n = *m.size;
  if (n == 0) {  // mapManager.cpp:42
    if (threadIdx.x == 0) { *n_submap = 0; *m.n_matches_map = 0; }
    return;
  }
  const float qx = (float)pose->t[0], qy = (float)pose->t[1], qz = 1.0f;  // :51-53
  for (int i = threadIdx.x; i < n; i += kMapThreads) {
    const float4 r = m.roots[i];
    const float dx = r.x - qx, dy = r.y - qy, dz = r.z - qz;
    m.keys[i] = __float_as_uint((dx * dx + dy * dy) + dz * dz);  // >= 0: bit order == value order
  }
  __syncthreads();
  const int kk = min(m.knn, n);
  // radix select of the kk-th smallest key
  uint32_t prefix = 0, pmask = 0;
  int want = kk - 1;
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (threadIdx.x < 256) s_hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kMapThreads) {
      const uint32_t z = m.keys[i];
      if ((z & pmask) == prefix) atomicAdd(&s_hist[(z >> shift) & 0xFF], 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0, d = 0;
      for (; d < 255; ++d) { if (acc + s_hist[d] > want) break; acc += s_hist[d]; }
      s_misc[0] = d; s_misc[1] = want - acc;
    }
    __syncthreads();
    prefix |= (uint32_t)s_misc[0] << shift;
    pmask |= 0xFFu << shift;
    want = s_misc[1];
    __syncthreads();
  }
  const uint32_t pivot = prefix;
  const int tie_quota = want + 1;
  if (threadIdx.x == 0) s_misc[2] = 0;
  __syncthreads();
  int ties_seen = 0;
  for (int base = 0; base < n; base += kMapThreads) {
    const int i = base + threadIdx.x;
    uint32_t key = 0xFFFFFFFFu;
    bool lt = false, eq = false;
    if (i < n) { key = m.keys[i]; lt = key < pivot; eq = key == pivot; }
    int tot_eq;
    const int eq_before = ties_seen + map_block_scan(eq ? 1 : 0, s_warp, &tot_eq);
    if (lt || (eq && eq_before < tie_quota)) {
      const int slot = atomicAdd(&s_misc[2], 1);
      s_cand[slot] = ((unsigned long long)key << 32) | (unsigned)i;
    }
    ties_seen += tot_eq;
  }
  __syncthreads();
  // sort the kk winners by (distance, index)
  if (threadIdx.x < kk) {
    const unsigned long long e = s_cand[threadIdx.x];
    int rank = 0;
    for (int j = 0; j < kk; ++j) rank += s_cand[j] < e;
    s_sorted[rank] = e;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // :57-66, in neighbour order
    int cnt = 0;
    for (int j = 0; j < kk && cnt < submap_cap; ++j) {
      const int map_idx = (int)(s_sorted[j] & 0xFFFFFFFFu);
      if (n - map_idx < m.recent) {
        m.matches_map[cnt] = map_idx;
        submap[cnt] = m.models[map_idx];
        ++cnt;
      }
    }
    *n_submap = cnt;
    *m.n_matches_map = cnt;
  }
}

// updateMap (:8-28).  Appends keep the observation order; overwrites are applied in
// observation order by one thread (two observations may hit the same landmark).
__global__ void __launch_bounds__(kMapThreads)
map_update_kernel(MapState m, sloam_kf_result *__restrict__ res, const sloam_cylinder *__restrict__ tm,
                  const int32_t *__restrict__ tm_id, const int32_t *__restrict__ matches,
                  int32_t *__restrict__ overflow) {
  __shared__ int s_warp[kMapThreads / 32];
  const int code = SLOAM_KF_CODE(res->status);
  const bool ran = code == SLOAM_KF_OK || code == SLOAM_KF_NOT_CONVERGED;
  const int n_obs = ran ? res->n_landmarks : 0;  // out.tm is empty when RunSloam bailed out
  const int size0 = *m.size;
  const int nmm = *m.n_matches_map;
  int appended = 0;
  for (int base = 0; base < n_obs; base += kMapThreads) {
    const int i = base + threadIdx.x;
    const bool app = i < n_obs && matches[i] == -1;
    int tot;
    const int pos = size0 + appended + map_block_scan(app ? 1 : 0, s_warp, &tot);
    if (app) {
      if (pos < m.capacity) {
        const sloam_cylinder c = tm[i];
        m.models[pos] = c;
        m.roots[pos] = make_float4((float)c.root[0], (float)c.root[1], (float)c.root[2], 0.f);
        m.ids[pos] = tm_id[i];
        m.hits[pos] = 1;
      } else {
        atomicOr(overflow, 1);
      }
    }
    appended += tot;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < n_obs; ++i) {
      const int mt = matches[i];
      if (mt == -1) continue;
      const int idx = (mt >= 0 && mt < nmm) ? m.matches_map[mt] : 0;  // std::map::operator[] default
      if (idx < min(size0 + appended, m.capacity)) {
        const sloam_cylinder c = tm[i];
        m.models[idx] = c;
        m.roots[idx] = make_float4((float)c.root[0], (float)c.root[1], (float)c.root[2], 0.f);
        m.ids[idx] = tm_id[i];
        m.hits[idx] += 1;
      }
    }
    *m.size = min(size0 + appended, m.capacity);
    *m.n_matches_map = 0;  // matchesMap.clear()
    if (size0 + appended > m.capacity) res->status |= SLOAM_KF_FLAG_MAP_CAPACITY;  // reported, not silent
  }
}

}  // namespace sb

using namespace sb;

// sequential state hanging off the context (opaque to common.cuh)
struct sloam_seq_state {
  MapState map;
  void *arena = nullptr;
  int32_t *overflow = nullptr;
  // SLOAMNode::run state
  bool core_first_scan = true;
  sloam_plane *prev_planes = nullptr;  // device [max_prev_planes]
  int32_t *n_prev = nullptr;           // device
  // single-keyframe device staging
  sloam_point *points = nullptr; uint8_t *mask = nullptr; sloam_pose *pose = nullptr; uint8_t *first = nullptr;
  sloam_cylinder *submap = nullptr; int32_t *n_submap = nullptr;
  sloam_kf_result *res = nullptr; int32_t *matches = nullptr; sloam_cylinder *tm = nullptr; int32_t *tm_id = nullptr;
  sloam_plane *planes = nullptr; int32_t *n_planes = nullptr;
  // The device part of a step (submap query, ~25 kernels of the fused path with K = 1, map
  // update, state copies) is launch-bound: it is captured into a CUDA graph on the second call
  // (the first one runs eagerly and does every lazy initialisation) and replayed afterwards.
  cudaGraphExec_t graph = nullptr;
  int graph_epoch = -1, steps = 0;
  int64_t graph_launches = 0;  // kernels per replay (for sloam_b200_kernel_launches)
  bool graph_off = false;      // capture failed once, or SLOAM_B200_NO_GRAPH is set
};

static sloam_seq_state *seq_of(sloam_ctx *c) { return static_cast<sloam_seq_state *>(c->seq); }

extern "C" {

int sloam_b200_run_keyframes_dev(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out);

int sloam_b200_map_init(sloam_ctx *c, int capacity) {
  if (!c || capacity <= 0) return SLOAM_E_INVALID;
  if (c->seq) sloam_b200_map_free(c);
  sloam_seq_state *s = new (std::nothrow) sloam_seq_state();
  if (!s) return SLOAM_E_NOMEM;
  const sloam_params &p = c->hp.p;
  const size_t N = (size_t)c->hp.N, T = p.max_trees, M = p.max_map_models, PP = p.max_prev_planes, C = capacity;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
  const size_t o_models = take(C * sizeof(sloam_cylinder)), o_roots = take(C * 16), o_ids = take(C * 4),
               o_hits = take(C * 4), o_keys = take(C * 4), o_size = take(4), o_mm = take(kMapKnnMax * 4),
               o_nmm = take(4), o_ovf = take(4), o_prev = take(PP * sizeof(sloam_plane)), o_nprev = take(4),
               o_pts = take(N * 16), o_mask = take(N), o_pose = take(sizeof(sloam_pose)), o_first = take(1),
               o_sub = take(M * sizeof(sloam_cylinder)), o_nsub = take(4), o_res = take(sizeof(sloam_kf_result)),
               o_match = take(T * 4), o_tm = take(T * sizeof(sloam_cylinder)), o_tmid = take(T * 4),
               o_planes = take(PP * sizeof(sloam_plane)), o_npl = take(4);
  if (cudaMalloc(&s->arena, off) != cudaSuccess) { delete s; return SLOAM_E_NOMEM; }
  char *b = (char *)s->arena;
  cudaMemsetAsync(s->arena, 0, off, c->stream);
  s->map.models = (sloam_cylinder *)(b + o_models); s->map.roots = (float4 *)(b + o_roots);
  s->map.ids = (int32_t *)(b + o_ids); s->map.hits = (int32_t *)(b + o_hits);
  s->map.keys = (uint32_t *)(b + o_keys); s->map.size = (int32_t *)(b + o_size);
  s->map.matches_map = (int32_t *)(b + o_mm); s->map.n_matches_map = (int32_t *)(b + o_nmm);
  s->map.capacity = capacity;
  s->overflow = (int32_t *)(b + o_ovf);
  s->prev_planes = (sloam_plane *)(b + o_prev); s->n_prev = (int32_t *)(b + o_nprev);
  s->points = (sloam_point *)(b + o_pts); s->mask = (uint8_t *)(b + o_mask); s->pose = (sloam_pose *)(b + o_pose);
  s->first = (uint8_t *)(b + o_first); s->submap = (sloam_cylinder *)(b + o_sub); s->n_submap = (int32_t *)(b + o_nsub);
  s->res = (sloam_kf_result *)(b + o_res); s->matches = (int32_t *)(b + o_match); s->tm = (sloam_cylinder *)(b + o_tm);
  s->tm_id = (int32_t *)(b + o_tmid); s->planes = (sloam_plane *)(b + o_planes); s->n_planes = (int32_t *)(b + o_npl);
  c->seq = s;
  return SLOAM_OK;
}

void sloam_b200_map_free(sloam_ctx *c) {
  if (!c || !c->seq) return;
  sloam_seq_state *s = seq_of(c);
  cudaStreamSynchronize(c->stream);
  if (s->graph) cudaGraphExecDestroy(s->graph);
  if (s->arena) cudaFree(s->arena);
  delete s;
  c->seq = nullptr;
}

int sloam_b200_map_get_submap_dev(sloam_ctx *c, const sloam_pose *pose, sloam_cylinder *submap, int32_t *n_submap) {
  if (!c || !c->seq || !pose || !submap || !n_submap) return set_err(c, SLOAM_E_INVALID, "map_get_submap: bad arguments / map_init not called");
  sloam_seq_state *s = seq_of(c);
  map_submap_kernel<<<1, kMapThreads, 0, c->stream>>>(s->map, pose, submap, n_submap, c->hp.p.max_map_models);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

int sloam_b200_map_update_dev(sloam_ctx *c, sloam_kf_result *res, const sloam_cylinder *tm, const int32_t *tm_id,
                              const int32_t *matches) {
  if (!c || !c->seq || !res || !tm || !tm_id || !matches) return set_err(c, SLOAM_E_INVALID, "map_update: bad arguments");
  sloam_seq_state *s = seq_of(c);
  map_update_kernel<<<1, kMapThreads, 0, c->stream>>>(s->map, res, tm, tm_id, matches, s->overflow);
  SB_LAUNCH_CHECK(c);
  return SLOAM_OK;
}

int sloam_b200_map_dump_host(sloam_ctx *c, sloam_cylinder *models, int32_t *hits, int cap) {
  if (!c || !c->seq) return SLOAM_E_INVALID;
  sloam_seq_state *s = seq_of(c);
  int32_t n = 0;
  SB_CUDA(c, cudaMemcpyAsync(&n, s->map.size, 4, cudaMemcpyDeviceToHost, c->stream));
  SB_CUDA(c, cudaStreamSynchronize(c->stream));
  const int m = n < cap ? n : cap;
  if (m > 0 && models) SB_CUDA(c, cudaMemcpyAsync(models, s->map.models, sizeof(sloam_cylinder) * m, cudaMemcpyDeviceToHost, c->stream));
  if (m > 0 && hits) SB_CUDA(c, cudaMemcpyAsync(hits, s->map.hits, 4 * (size_t)m, cudaMemcpyDeviceToHost, c->stream));
  SB_CUDA(c, cudaStreamSynchronize(c->stream));
  return n;
}

// SLOAMNode::run (sloamNode.cpp:186-282) for one keyframe, host buffers in, state on the device.
// Returns the bool of run() in result->success; the node-level "Discarding msg" early exit
// (:199-203) coincides with RunSloam's empty-map guard (same return value, no map change).
int sloam_b200_sequence_step_host(sloam_ctx *c, const sloam_point *points, const uint8_t *mask,
                                  const sloam_pose *pose_estimate, sloam_kf_result *result,
                                  int32_t *matches, sloam_cylinder *tm, int32_t *tm_id) {
  if (!c || !c->seq || !points || !mask || !pose_estimate || !result)
    return set_err(c, SLOAM_E_INVALID, "sequence_step: bad arguments / map_init not called");
  if (c->max_k < 1) return SLOAM_E_INVALID;
  sloam_seq_state *s = seq_of(c);
  const sloam_params &p = c->hp.p;
  const size_t N = (size_t)c->hp.N, T = p.max_trees, PP = p.max_prev_planes;
  cudaStream_t st = c->stream;
  const uint8_t first = s->core_first_scan ? 1 : 0;
  SB_CUDA(c, cudaMemcpyAsync(s->points, points, N * 16, cudaMemcpyHostToDevice, st));
  SB_CUDA(c, cudaMemcpyAsync(s->mask, mask, N, cudaMemcpyHostToDevice, st));
  SB_CUDA(c, cudaMemcpyAsync(s->pose, pose_estimate, sizeof(sloam_pose), cudaMemcpyHostToDevice, st));
  SB_CUDA(c, cudaMemcpyAsync(s->first, &first, 1, cudaMemcpyHostToDevice, st));
  // ---- device part: getSubmap -> RunSloam -> updateMap -> state for the next call ----
  auto device_part = [&]() -> int {
    int rc = sloam_b200_map_get_submap_dev(c, s->pose, s->submap, s->n_submap);  // :197
    if (rc != SLOAM_OK) return rc;
    sloam_batch_in in{};
    in.points = s->points; in.mask = s->mask; in.pose_est = s->pose; in.first_scan = s->first;
    in.map_models = s->submap; in.n_map_models = s->n_submap; in.map_shared = 0;
    in.prev_planes = s->prev_planes; in.n_prev_planes = s->n_prev;
    sloam_batch_out out{};
    out.results = s->res; out.matches = s->matches; out.tm = s->tm; out.tm_id = s->tm_id;
    out.planes = s->planes; out.n_planes = s->n_planes; out.range_image = nullptr;
    rc = sloam_b200_run_keyframes_dev(c, 1, &in, &out);  // :208-235
    if (rc != SLOAM_OK) return rc;
    rc = sloam_b200_map_update_dev(c, s->res, s->tm, s->tm_id, s->matches);  // :236, also after a false return
    if (rc != SLOAM_OK) return rc;
    // prevGPlanes_ for the next call (unchanged copy when RunSloam bailed out)
    SB_CUDA(c, cudaMemcpyAsync(s->prev_planes, s->planes, PP * sizeof(sloam_plane), cudaMemcpyDeviceToDevice, st));
    SB_CUDA(c, cudaMemcpyAsync(s->n_prev, s->n_planes, 4, cudaMemcpyDeviceToDevice, st));
    return SLOAM_OK;
  };
  static const bool no_graph = getenv("SLOAM_B200_NO_GRAPH") != nullptr;
  if (s->graph && s->graph_epoch != c->epoch) { cudaGraphExecDestroy(s->graph); s->graph = nullptr; }
  int rc = SLOAM_OK;
  if (s->graph) {
    SB_CUDA(c, cudaGraphLaunch(s->graph, st));
    c->launches += s->graph_launches;
  } else if (s->steps == 0 || no_graph || s->graph_off || c->prof_on) {
    rc = device_part();  // eager: first call (lazy initialisation), or graphs disabled
    if (rc != SLOAM_OK) return rc;
  } else {
    const int64_t l0 = c->launches;
    cudaGraph_t g = nullptr;
    bool ok = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
    if (ok) {
      rc = device_part();
      ok = cudaStreamEndCapture(st, &g) == cudaSuccess && rc == SLOAM_OK && g != nullptr;
    }
    if (ok) ok = cudaGraphInstantiate(&s->graph, g, 0) == cudaSuccess;
    if (g) cudaGraphDestroy(g);
    if (ok) {
      s->graph_epoch = c->epoch;
      s->graph_launches = c->launches - l0;
      SB_CUDA(c, cudaGraphLaunch(s->graph, st));
    } else {  // not capturable in this configuration: stay eager from now on
      (void)cudaGetLastError();
      s->graph = nullptr;
      s->graph_off = true;
      c->launches = l0;
      rc = device_part();
      if (rc != SLOAM_OK) return rc;
    }
  }
  ++s->steps;
  SB_CUDA(c, cudaMemcpyAsync(result, s->res, sizeof(sloam_kf_result), cudaMemcpyDeviceToHost, st));
  if (matches) SB_CUDA(c, cudaMemcpyAsync(matches, s->matches, T * 4, cudaMemcpyDeviceToHost, st));
  if (tm) SB_CUDA(c, cudaMemcpyAsync(tm, s->tm, T * sizeof(sloam_cylinder), cudaMemcpyDeviceToHost, st));
  if (tm_id) SB_CUDA(c, cudaMemcpyAsync(tm_id, s->tm_id, T * 4, cudaMemcpyDeviceToHost, st));
  SB_CUDA(c, cudaStreamSynchronize(st));
  s->core_first_scan = false;  // sloam.cpp:472 (the first-scan branch has no early exit)
  return SLOAM_OK;
}

}  // extern "C"
