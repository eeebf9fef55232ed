// pipeline.cu -- the fused batched path: Segmentation::run (projection) +
// maskCloud x2 + Instance::computeGraph + sloam::RunSloam for K keyframes
// (the call sequence of SLOAMNode::run, sloam/src/core/sloamNode.cpp:208-236,
// minus the segmentation network whose H x W {0,1,255} mask is an input).
#include "common.cuh"

namespace sb {

int launch_project_split(sloam_ctx *c, int K, bool do_project, bool do_split, const sloam_point *points,
                         const uint8_t *mask, int32_t *pix, float *range_image, sloam_point *tree,
                         sloam_point *ground, int32_t *ground_count, uint32_t *tree_bits, bool sparse_tree);
int launch_ground_planes(sloam_ctx *c, int K, const sloam_point *ground, const int32_t *ground_count,
                         int stride, const sloam_pose *pose_est, sloam_cell_plane *cells,
                         sloam_point *cell_features, sloam_point *kept_points, int32_t *kept_offsets,
                         bool strided);
int launch_compute_graph(sloam_ctx *c, int K, const sloam_point *tree, sloam_tree *trees, int32_t *n_trees,
                         sloam_vertex *vertices, sloam_point *vertex_points, bool bits_ready);
int launch_tree_fill(sloam_ctx *c, int K);
int launch_ground_compact(sloam_ctx *c, int K, sloam_point *dst);
int launch_cylinders(sloam_ctx *c, int K, const sloam_tree *trees, const int32_t *n_trees,
                     const sloam_vertex *vertices, const sloam_point *vpoints, const sloam_plane *planes_acc,
                     const int32_t *n_planes_acc, sloam_tree_model *models, sloam_point *features);
int launch_sloam_core(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out);
int launch_ground_tag(sloam_ctx *c, int K, const sloam_point *ground, const int32_t *ground_count, int stride);
int launch_cylinders_strided(sloam_ctx *c, int K, const sloam_tree *trees, const int32_t *n_trees,
                             const sloam_vertex *vertices, int vstride, const sloam_point *vpoints,
                             int pstride, const sloam_plane *planes_acc, const int32_t *n_planes_acc,
                             sloam_tree_model *models, sloam_point *features);

static int check_batch(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out) {
  if (!c) return SLOAM_E_INVALID;
  if (K <= 0 || K > c->max_k) return set_err(c, SLOAM_E_INVALID, "run_keyframes: K out of range");
  if (!in || !out || !in->points || !in->mask || !in->pose_est || !in->first_scan || !in->map_models ||
      !in->n_map_models || !in->prev_planes || !in->n_prev_planes || !out->results || !out->matches ||
      !out->tm || !out->tm_id || !out->planes || !out->n_planes)
    return set_err(c, SLOAM_E_INVALID, "run_keyframes: null buffer");
  return SLOAM_OK;
}

static int run_dev(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out, bool allow_split = true);

// sub-batch [k0, k0 + n) of a device batch
static void slice_batch(const sloam_ctx *c, const sloam_batch_in *in, const sloam_batch_out *out, int k0,
                        sloam_batch_in *din, sloam_batch_out *dout) {
  const sloam_params &p = c->hp.p;
  const size_t N = (size_t)c->hp.N, T = (size_t)p.max_trees, M = (size_t)p.max_map_models,
               PP = (size_t)p.max_prev_planes, k = (size_t)k0;
  const bool shared = in->map_shared != 0;
  *din = *in;
  din->points = in->points + k * N;
  din->mask = in->mask + k * N;
  din->pose_est = in->pose_est + k;
  din->first_scan = in->first_scan + k;
  din->map_models = in->map_models + (shared ? 0 : k * M);
  din->n_map_models = in->n_map_models + (shared ? 0 : k);
  din->prev_planes = in->prev_planes + k * PP;
  din->n_prev_planes = in->n_prev_planes + k;
  *dout = *out;
  dout->results = out->results + k;
  dout->matches = out->matches + k * T;
  dout->tm = out->tm + k * T;
  dout->tm_id = out->tm_id + k * T;
  dout->planes = out->planes + k * PP;
  dout->n_planes = out->n_planes + k;
  dout->range_image = out->range_image ? out->range_image + k * N : nullptr;
}

// fused run cut into sub-batches that run concurrently on the lanes' streams
static int run_lanes(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out) {
  const int n = c->n_lanes, per = (K + n - 1) / n;
  SB_CUDA(c, cudaEventRecord(c->ev_lane_start, c->stream));
  for (int l = 0; l < n; ++l) {
    const int k0 = l * per, kn = K - k0 < per ? K - k0 : per;
    if (kn <= 0) break;
    sloam_ctx *lc = c->lane[l];
    SB_CUDA(c, cudaStreamWaitEvent(lc->stream, c->ev_lane_start, 0));
    sloam_batch_in din;
    sloam_batch_out dout;
    slice_batch(c, in, out, k0, &din, &dout);
    const int rc = run_dev(lc, kn, &din, &dout, false);
    if (rc != SLOAM_OK) return set_err(c, rc, lc->err);
    SB_CUDA(c, cudaEventRecord(c->ev_lane_done[l], lc->stream));
    SB_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_lane_done[l], 0));
  }
  c->last_k = c->lane[0]->last_k;
  return SLOAM_OK;
}

static int run_dev_body(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out);

static int run_dev(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out, bool allow_split) {
  // with kernel profiling on, a run is one lane and one stream, so that the event pairs
  // time each kernel group alone
  if (allow_split && c->n_lanes > 1 && !c->prof_on && K >= 64 * c->n_lanes) return run_lanes(c, K, in, out);
  const int rc = run_dev_body(c, K, in, out);
  c->zero_valid = 0;  // the pre-zeroed counters belong to this run only (also after an error)
  return rc;
}

static int run_dev_body(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out) {
  Workspace &w = c->ws;
  // every counter / flag array of the run in one memset (they are contiguous in the arena)
  SB_CUDA(c, cudaMemsetAsync(w.zero_begin, 0, (size_t)((char *)w.zero_end - (char *)w.zero_begin), c->stream));
  c->zero_valid = 1u | 2u | 4u | 8u;
  float *range = out->range_image;  // optional output
  // sparse tree cloud: only the tree-labelled points and the bit mask are written (the NaN
  // points of the dense cloud are ~90 % of its bytes and nothing downstream needs them)
  // ... and the ground points are not copied either: ground == nullptr selects the record
  // layout (k1_project.cu, FUSED), the cell stage reads the retained points from in->points
  int rc = launch_project_split(c, K, true, true, in->points, in->mask, w.pix, range, w.tree, nullptr,
                                w.ground_count, w.tree_bits, true);
  c->tree_sparse = true;
  c->ground_strided = true;
  c->last_points = in->points;
  c->last_mask = in->mask;
  if (rc != SLOAM_OK) return rc;
  // fork: ground cells + plane fits (K2, main stream) and the tree detector (K3, side
  // stream) both depend only on K1 and are latency-bound, so they run concurrently
  cudaStream_t main_stream = c->stream;
  const bool fork = !c->prof_on;
  if (fork) {
    SB_CUDA(c, cudaEventRecord(c->ev_fork, main_stream));
    SB_CUDA(c, cudaStreamWaitEvent(c->side, c->ev_fork, 0));
    c->stream = c->side;
  }
  rc = launch_compute_graph(c, K, w.tree, w.trees, w.n_trees, w.vertices, w.vertex_points, true);
  c->stream = main_stream;
  if (rc != SLOAM_OK) return rc;
  if (fork) SB_CUDA(c, cudaEventRecord(c->ev_join, c->side));
  rc = launch_ground_planes(c, K, in->points, w.ground_count, c->hp.N, in->pose_est, w.cells,
                            w.cell_features, nullptr, nullptr, true);
  if (rc != SLOAM_OK) return rc;
  if (fork) SB_CUDA(c, cudaStreamWaitEvent(main_stream, c->ev_join, 0));  // join
  rc = launch_cylinders(c, K, w.trees, w.n_trees, w.vertices, w.vertex_points, w.planes_acc,
                        w.n_planes_acc, w.tree_models, w.tree_features);
  if (rc != SLOAM_OK) return rc;
  c->kf_flags_valid = true;  // written by the tree detector of this run
  {  // a result gather still in flight may be reading the output buffers (comm.cu)
    sloam_ctx *root = c->parent ? c->parent : c;
    if (root->gather_pending) SB_CUDA(c, cudaStreamWaitEvent(c->stream, root->ev_gather_done, 0));
  }
  rc = launch_sloam_core(c, K, in, out);
  c->kf_flags_valid = false;
  c->last_k = K;
  if (c->prof_on && c->prof_n < sloam_ctx::kProfRuns) ++c->prof_n;  // next fused run -> next event slots
  return rc;
}

}  // namespace sb

using namespace sb;

extern "C" {

int sloam_b200_run_keyframes_dev(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out) {
  const int rc = check_batch(c, K, in, out);
  if (rc != SLOAM_OK) return rc;
  return run_dev(c, K, in, out);
}

}  // extern "C"

namespace sb {
// x, y, z packed (12 bytes per point) -> sloam_point with intensity 0
__global__ void expand_xyz_kernel(const float *__restrict__ xyz, sloam_point *__restrict__ pts, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  st_point(pts + i, sloam_point{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f});
}
}  // namespace sb

// host entry; xyz != nullptr: the cloud comes as packed x, y, z instead of in->points
static int run_keyframes_host_impl(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out,
                                   const float *xyz) {
  sloam_batch_in in_chk;
  if (xyz && in) { in_chk = *in; in_chk.points = reinterpret_cast<const sloam_point *>(xyz); }
  int rc = check_batch(c, K, xyz && in ? &in_chk : in, out);
  if (rc != SLOAM_OK) return rc;
  const sloam_params &p = c->hp.p;
  // host buffers: the counts can be checked against the capacities before anything is copied
  for (int k = 0; k < K; ++k) {
    const int nm = in->n_map_models[in->map_shared ? 0 : k], npv = in->n_prev_planes[k];
    if (nm < 0 || nm > p.max_map_models || npv < 0 || npv > p.max_prev_planes)
      return set_err(c, SLOAM_E_INVALID, "run_keyframes_host: n_map_models / n_prev_planes exceed max_map_models / max_prev_planes");
  }
  const size_t N = (size_t)c->hp.N, T = (size_t)p.max_trees, M = (size_t)p.max_map_models,
               PP = (size_t)p.max_prev_planes, Kc = (size_t)c->max_k;
  // device staging for the largest batch, allocated once
  struct Part { size_t off, bytes; };
  size_t off = 0;
  auto part = [&](size_t bytes) { Part q{off, bytes}; off = (off + bytes + 255) / 256 * 256; return q; };
  const Part d_points = part(Kc * N * sizeof(sloam_point)), d_mask = part(Kc * N), d_pose = part(Kc * sizeof(sloam_pose)),
             d_first = part(Kc), d_map = part(Kc * M * sizeof(sloam_cylinder)), d_nmap = part(Kc * 4),
             d_prev = part(Kc * PP * sizeof(sloam_plane)), d_nprev = part(Kc * 4),
             d_res = part(Kc * sizeof(sloam_kf_result)), d_match = part(Kc * T * 4),
             d_tm = part(Kc * T * sizeof(sloam_cylinder)), d_tmid = part(Kc * T * 4),
             d_planes = part(Kc * PP * sizeof(sloam_plane)), d_npl = part(Kc * 4),
             d_range = part(Kc * N * 4), d_xyz = part(xyz ? Kc * N * 12 : 0);
  if (c->stage_dev_bytes < off) {
    if (c->stage_dev) cudaFree(c->stage_dev);
    c->stage_dev = nullptr; c->stage_dev_bytes = 0;
    SB_CUDA(c, cudaMalloc(&c->stage_dev, off));
    c->stage_dev_bytes = off;
  }
  char *base = (char *)c->stage_dev;
  cudaStream_t s = c->stream;
  // The batch is cut into chunks: the host->device copies of all chunks are queued on a copy
  // stream up front, the compute stream runs chunk j as soon as its inputs have landed, so
  // the PCIe transfer of chunk j+1.. overlaps the kernels of chunk j (pinned host buffers).
  constexpr int kMaxChunks = 8;
  if (!c->copy_stream) {
    SB_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int j = 0; j < kMaxChunks; ++j) SB_CUDA(c, cudaEventCreateWithFlags(&c->ev_chunk[j], cudaEventDisableTiming));
    SB_CUDA(c, cudaEventCreateWithFlags(&c->ev_stage_free, cudaEventDisableTiming));
  }
  int n_chunks = K / 96;
  n_chunks = n_chunks < 1 ? 1 : (n_chunks > kMaxChunks ? kMaxChunks : n_chunks);
  const int per = (K + n_chunks - 1) / n_chunks;
  const bool shared = in->map_shared != 0;
  // the copy stream must not overwrite the staging area while earlier work still reads it
  SB_CUDA(c, cudaEventRecord(c->ev_stage_free, s));
  SB_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_stage_free, 0));
#define H2D(partv, src, elem, k0, n)                                                                        \
  SB_CUDA(c, cudaMemcpyAsync(base + partv.off + (size_t)(k0) * (elem), (const char *)(src) + (size_t)(k0) * (elem), \
                             (size_t)(n) * (elem), cudaMemcpyHostToDevice, c->copy_stream))
  if (shared) {
    H2D(d_map, in->map_models, M * sizeof(sloam_cylinder), 0, 1);
    H2D(d_nmap, in->n_map_models, 4, 0, 1);
  }
  for (int j = 0; j < n_chunks; ++j) {
    const int k0 = j * per, n = (K - k0 < per) ? K - k0 : per;
    if (n <= 0) { n_chunks = j; break; }
    if (xyz) H2D(d_xyz, xyz, N * 12, k0, n);
    else H2D(d_points, in->points, N * sizeof(sloam_point), k0, n);
    H2D(d_mask, in->mask, N, k0, n);
    H2D(d_pose, in->pose_est, sizeof(sloam_pose), k0, n);
    H2D(d_first, in->first_scan, 1, k0, n);
    if (!shared) {
      H2D(d_map, in->map_models, M * sizeof(sloam_cylinder), k0, n);
      H2D(d_nmap, in->n_map_models, 4, k0, n);
    }
    H2D(d_prev, in->prev_planes, PP * sizeof(sloam_plane), k0, n);
    H2D(d_nprev, in->n_prev_planes, 4, k0, n);
    SB_CUDA(c, cudaEventRecord(c->ev_chunk[j], c->copy_stream));
  }
#undef H2D
  for (int j = 0; j < n_chunks; ++j) {
    const int k0 = j * per, n = (K - k0 < per) ? K - k0 : per;
    const size_t k0s = (size_t)k0, ns = (size_t)n;
    SB_CUDA(c, cudaStreamWaitEvent(s, c->ev_chunk[j], 0));
    if (xyz) {
      const long long npts = (long long)n * (long long)N;
      expand_xyz_kernel<<<(unsigned)((npts + 255) / 256), 256, 0, s>>>(
          reinterpret_cast<const float *>(base + d_xyz.off) + k0s * N * 3,
          reinterpret_cast<sloam_point *>(base + d_points.off) + k0s * N, npts);
      SB_LAUNCH_CHECK(c);
    }
    sloam_batch_in din = *in;
    din.points = (const sloam_point *)(base + d_points.off) + k0s * N;
    din.mask = (const uint8_t *)(base + d_mask.off) + k0s * N;
    din.pose_est = (const sloam_pose *)(base + d_pose.off) + k0s;
    din.first_scan = (const uint8_t *)(base + d_first.off) + k0s;
    din.map_models = (const sloam_cylinder *)(base + d_map.off) + (shared ? 0 : k0s * M);
    din.n_map_models = (const int32_t *)(base + d_nmap.off) + (shared ? 0 : k0s);
    din.prev_planes = (const sloam_plane *)(base + d_prev.off) + k0s * PP;
    din.n_prev_planes = (const int32_t *)(base + d_nprev.off) + k0s;
    sloam_batch_out dout;
    dout.results = (sloam_kf_result *)(base + d_res.off) + k0s;
    dout.matches = (int32_t *)(base + d_match.off) + k0s * T;
    dout.tm = (sloam_cylinder *)(base + d_tm.off) + k0s * T;
    dout.tm_id = (int32_t *)(base + d_tmid.off) + k0s * T;
    dout.planes = (sloam_plane *)(base + d_planes.off) + k0s * PP;
    dout.n_planes = (int32_t *)(base + d_npl.off) + k0s;
    dout.range_image = out->range_image ? (float *)(base + d_range.off) + k0s * N : nullptr;
    rc = run_dev(c, n, &din, &dout, false);
    if (rc != SLOAM_OK) return rc;
#define D2H(dst, partv, elem) \
  SB_CUDA(c, cudaMemcpyAsync((char *)(dst) + k0s * (elem), base + partv.off + k0s * (elem), ns * (elem), cudaMemcpyDeviceToHost, s))
    D2H(out->results, d_res, sizeof(sloam_kf_result));
    D2H(out->matches, d_match, T * 4);
    D2H(out->tm, d_tm, T * sizeof(sloam_cylinder));
    D2H(out->tm_id, d_tmid, T * 4);
    D2H(out->planes, d_planes, PP * sizeof(sloam_plane));
    D2H(out->n_planes, d_npl, 4);
    if (out->range_image) D2H(out->range_image, d_range, N * 4);
#undef D2H
  }
  SB_CUDA(c, cudaStreamSynchronize(s));
  return SLOAM_OK;
}

extern "C" {

int sloam_b200_run_keyframes_host(sloam_ctx *c, int K, const sloam_batch_in *in, const sloam_batch_out *out) {
  return run_keyframes_host_impl(c, K, in, out, nullptr);
}

int sloam_b200_run_keyframes_host_xyz(sloam_ctx *c, int K, const float *points_xyz, const sloam_batch_in *in,
                                      const sloam_batch_out *out) {
  if (!points_xyz) return set_err(c, SLOAM_E_INVALID, "run_keyframes_host_xyz: null cloud");
  return run_keyframes_host_impl(c, K, in, out, points_xyz);
}

int sloam_b200_run_sloam_dev(sloam_ctx *c, int K, const sloam_point *ground, const int32_t *ground_count,
                             int ground_stride, const sloam_tree *trees, const int32_t *n_trees,
                             const sloam_vertex *vertices, int vertex_stride,
                             const sloam_point *vertex_points, int point_stride, const sloam_batch_in *in,
                             const sloam_batch_out *out) {
  if (!c || K <= 0 || K > c->max_k || !ground || !ground_count || !trees || !n_trees || !vertices ||
      !vertex_points || !in || !out || !in->pose_est || !in->first_scan || !in->map_models ||
      !in->n_map_models || !in->prev_planes || !in->n_prev_planes || !out->results || !out->matches ||
      !out->tm || !out->tm_id || !out->planes || !out->n_planes || ground_stride <= 0 ||
      ground_stride > c->hp.N)
    return set_err(c, SLOAM_E_INVALID, "run_sloam: bad arguments (ground_stride must be <= H*W)");
  Workspace &w = c->ws;
  int rc = launch_ground_tag(c, K, ground, ground_count, ground_stride);
  if (rc != SLOAM_OK) return rc;
  rc = launch_ground_planes(c, K, ground, ground_count, ground_stride, in->pose_est, w.cells,
                            w.cell_features, nullptr, nullptr, false);
  if (rc != SLOAM_OK) return rc;
  rc = launch_cylinders_strided(c, K, trees, n_trees, vertices, vertex_stride, vertex_points, point_stride,
                                w.planes_acc, w.n_planes_acc, w.tree_models, w.tree_features);
  if (rc != SLOAM_OK) return rc;
  // the core reads the per-keyframe counts from the workspace for its result records
  SB_CUDA(c, cudaMemcpyAsync(w.ground_count, ground_count, sizeof(int32_t) * K, cudaMemcpyDeviceToDevice, c->stream));
  SB_CUDA(c, cudaMemcpyAsync(w.n_trees, n_trees, sizeof(int32_t) * K, cudaMemcpyDeviceToDevice, c->stream));
  return launch_sloam_core(c, K, in, out);
}

void *sloam_b200_dev_alloc(sloam_ctx *c, uint64_t bytes) {
  if (!c) return nullptr;
  void *p = nullptr;
  cudaSetDevice(c->device);
  if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) { c->err = "dev_alloc: cudaMalloc failed"; return nullptr; }
  return p;
}

void sloam_b200_dev_free(sloam_ctx *c, void *p) {
  if (!c || !p) return;
  cudaStreamSynchronize(c->stream);
  cudaFree(p);
}

int sloam_b200_copy_h2d(sloam_ctx *c, void *dst, const void *src, uint64_t bytes) {
  if (!c || (!dst && bytes) || (!src && bytes)) return SLOAM_E_INVALID;
  if (bytes) SB_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  return SLOAM_OK;
}

int sloam_b200_copy_d2h(sloam_ctx *c, void *dst, const void *src, uint64_t bytes) {
  if (!c || (!dst && bytes) || (!src && bytes)) return SLOAM_E_INVALID;
  if (bytes) SB_CUDA(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  SB_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLOAM_OK;
}

int sloam_b200_get_intermediates(sloam_ctx *c, sloam_intermediates *o) {
  if (!c || !o) return SLOAM_E_INVALID;
  if (c->n_lanes > 1 && c->lane[0] && c->lane[0]->last_k > 0 && c->last_k == c->lane[0]->last_k) {
    SB_CUDA(c, cudaStreamSynchronize(c->stream));
    const int rc = sloam_b200_get_intermediates(c->lane[0], o);
    if (rc == SLOAM_OK) SB_CUDA(c, cudaStreamSynchronize(c->lane[0]->stream));
    return rc;
  }
  const Workspace &w = c->ws;
  if (c->tree_sparse && c->last_k > 0) {  // make ws.tree the dense organized cloud of stage a2
    const int rc = launch_tree_fill(c, c->last_k);
    if (rc != SLOAM_OK) return rc;
    c->tree_sparse = false;
  }
  const sloam_point *ground_dense = w.ground;
  if (c->ground_strided && c->last_k > 0) {
    // the fused run kept neither the pixel indices nor a ground cloud: redo the projection and
    // the label split of stages a1 + a2 (the input buffers of that run must still be alive);
    // this also makes ws.tree dense
    int rc = launch_project_split(c, c->last_k, true, false, c->last_points, nullptr, w.pix, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, false);
    if (rc != SLOAM_OK) return rc;
    rc = launch_project_split(c, c->last_k, false, true, c->last_points, c->last_mask, w.pix, nullptr, w.tree,
                              reinterpret_cast<sloam_point *>(w.qscratch), w.ground_count, nullptr, false);
    if (rc != SLOAM_OK) return rc;
    c->tree_sparse = false;
  }
  if (c->ground_strided) ground_dense = reinterpret_cast<const sloam_point *>(w.qscratch);
  o->pix = w.pix; o->tree = w.tree; o->ground = const_cast<sloam_point *>(ground_dense); o->ground_count = w.ground_count;
  o->cells = w.cells; o->cell_features = w.cell_features; o->trees = w.trees; o->n_trees = w.n_trees;
  o->vertices = w.vertices; o->vertex_points = w.vertex_points; o->tree_models = w.tree_models;
  o->tree_features = w.tree_features;
  return SLOAM_OK;
}

}  // extern "C"
