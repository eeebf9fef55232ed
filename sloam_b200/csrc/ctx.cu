// ctx.cu -- context life cycle, parameters and scratch arena of the C ABI.
#include <algorithm>
#include <cstring>
#include <random>
#include <vector>

#include "common.cuh"

namespace sb {

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Bump {
  char *base;
  size_t off = 0;
  template <typename T>
  void take(T *&ptr, size_t count) {
    off = align_up(off, 256);
    if (base) ptr = reinterpret_cast<T *>(base + off);
    off += sizeof(T) * count;
  }
};

// PCL sampling stream (SampleConsensusModel::drawIndexSample, seed 12345,
// uniform_int<>(0, INT_MAX) over mt19937 == mt() >> 1; the shuffled index
// vector persists across draws).  A fresh model is built per tree
// (sloam/src/objects/cylinder.cpp:116-124), so the draw sequence depends only
// on the number of vertices V.  Precomputed on the host per V.
static void ransac_table(int V, int n_draws, std::vector<int32_t> &out) {
  std::mt19937 rng(12345u);
  std::vector<int> shuf(V);
  for (int i = 0; i < V; ++i) shuf[i] = i;
  for (int t = 0; t < n_draws; ++t) {
    for (unsigned i = 0; i < 2; ++i) {
      const unsigned r = (unsigned)(rng() >> 1);
      std::swap(shuf[i], shuf[i + (r % ((unsigned)V - i))]);
    }
    out.push_back(shuf[0]);
    out.push_back(shuf[1]);
  }
}

int ransac_draws_per_tree(const sloam_params &p) {
  // hypotheses + head-room for rejected (not "good") draws
  const int hyp = p.ransac_fixed_hypotheses > 0 ? p.ransac_fixed_hypotheses : p.ransac_max_iterations + 1;
  return hyp + 64;
}

static void layout(sloam_ctx *c, Bump &b) {
  Workspace &w = c->ws;
  const sloam_params &p = c->hp.p;
  const size_t K = (size_t)c->max_k, N = (size_t)c->hp.N, B = (size_t)c->hp.B;
  const size_t T = (size_t)p.max_trees, H = (size_t)p.img_h;
  const size_t tiles = (N + kSplitTile - 1) / kSplitTile;
  // counters and flags that every fused run starts from zero: contiguous, one memset
  b.take(w.zero_begin, 64);
  b.take(w.ground_count, K);
  b.take(w.cell_count, K * kMaxCells);
  b.take(w.n_tied_cells, 4);
  b.take(w.n_vlists, 8);
  b.take(w.kf_flags, K);
  b.take(w.zero_end, 64);
  b.take(w.pix, K * N);
  b.take(w.tree, K * N);
  b.take(w.ground, K * N);
  b.take(w.tree_bits, K * ((N + 31) / 32));
  if (p.do_destagger) {
    b.take(w.tree2, K * N);
    b.take(w.tree_bits2, K * ((N + 31) / 32));
  }
  b.take(w.ground_cell, K * N);
  b.take(w.tile_count, K * tiles);
  b.take(w.range_image, K * N);
  b.take(w.cells, K * B);
  b.take(w.cell_features, K * B * (size_t)p.numGroundFeatures);
  b.take(w.planes_acc, K * B);
  b.take(w.planes_acc_cell, K * B);
  b.take(w.n_planes_acc, K);
  b.take(w.gscratch, K * N);
  b.take(w.gscratch2, K * N);
  b.take(w.gscratch3, K * N);
  b.take(w.seg_tab, K * tiles * B);
  b.take(w.tied_cells, K * kMaxCells);
  b.take(w.qscratch, K * N * 3);
  b.take(w.pscratch, K * N * 3);
  b.take(w.fit_rec, K * B);
  b.take(w.cc_planes, 4 * K * ((N + 31) / 32));
  b.take(w.cc_wbase, K * ((N + 31) / 32));
  b.take(w.run_par, K * N);
  b.take(w.run_siz, K * N);
  b.take(w.run_len, K * N);
  b.take(w.run_pix0, K * N);
  b.take(w.run_info, K * N);
  b.take(w.run_label, K * N);
  b.take(w.n_big, K);
  b.take(w.n_roots, K);
  b.take(w.big_rank, K * T);
  b.take(w.slot_rows, K * T * ((H + 31) / 32));
  b.take(w.slot_vertices, K * T * H);
  {
    sloam_point *items16 = nullptr;  // 16-byte records
    b.take(items16, K << (c->hp.vw_row_bits + c->hp.vw_slot_bits));
    w.vitems = items16;
  }
  b.take(w.vitem_pool, K << (c->hp.vw_row_bits + c->hp.vw_slot_bits));
  b.take(w.vlists, 6 * K * T * H);
  b.take(w.trees, K * T);
  b.take(w.n_trees, K);
  b.take(w.vertices, K * T * (size_t)p.max_tree_vertices);
  b.take(w.vertex_points, K * N);
  b.take(w.tree_models, K * T);
  b.take(w.tree_features, K * T * (size_t)p.featuresPerTree);
  b.take(w.ransac_pairs, (size_t)(p.max_tree_vertices + 1) * ransac_draws_per_tree(p) * 2);
  b.take(w.ransac_pairs_offset, (size_t)p.max_tree_vertices + 2);
  b.take(w.lm_cyl, K * T);
  b.take(w.lm_src, K * T);
  b.take(w.n_lm, K);
  b.take(w.assoc_idx, K * T);
  b.take(w.assoc_dist, K * T);
  b.take(w.res_tree_feat, K * T * (size_t)p.featuresPerTree * 3);
  b.take(w.res_tree_obj, K * T * (size_t)p.featuresPerTree);
  b.take(w.res_plane_feat, K * B * (size_t)p.numGroundFeatures * 3);
  b.take(w.res_plane_obj, K * B * (size_t)p.numGroundFeatures);
  b.take(w.n_tree_res, K);
  b.take(w.n_plane_res, K);
  b.take(w.optim_flags, K * 2);
  b.take(w.kf_mode, K);
  b.take(w.lm_x, K * 16);
  b.take(w.lm_info, K * 4);
  b.take(w.curr_pose, K);
  b.take(w.results, K);
  b.take(w.matches, K * T);
  b.take(w.tm, K * T);
  b.take(w.tm_id, K * T);
  b.take(w.planes_out, K * (size_t)p.max_prev_planes);
  b.take(w.n_planes_out, K);
}

static int validate(const sloam_params &p, std::string &why) {
  if (p.img_h <= 0 || p.img_w <= 0) { why = "img_h/img_w must be positive"; return -1; }
  if ((long long)p.img_h * p.img_w >= (1ll << 24)) { why = "image larger than 2^24 pixels"; return -1; }
  if (p.groundRadiiBins <= 0 || p.groundThetaBins <= 0 ||
      p.groundRadiiBins * p.groundThetaBins > kMaxCells - 1) { why = "ground bins out of range"; return -1; }
  if (!(p.groundRetainThresh > 0.0 && p.groundRetainThresh <= 1.0)) {
    why = "groundRetainThresh must be in (0,1] (the reference erases past end() above 1, SURVEY B-5)";
    return -1;
  }
  if (p.numGroundFeatures <= 0 || p.featuresPerTree <= 0) { why = "feature counts must be positive"; return -1; }
  if (p.max_tree_vertices < 3 || p.max_tree_vertices > 64) { why = "max_tree_vertices must be in [3,64]"; return -1; }
  if (p.min_tree_vertices < 2) { why = "min_tree_vertices must be >= 2 (cylinder.cpp:11,78 index vertices[2])"; return -1; }
  if (p.max_trees <= 0 || p.max_map_models <= 0) { why = "capacities must be positive"; return -1; }
  if (p.max_trees > 2046) { why = "max_trees must be <= 2046 (11-bit slot codes, k3_trellis.cu)"; return -1; }
  if (p.max_prev_planes < p.groundRadiiBins * p.groundThetaBins) { why = "max_prev_planes < number of ground cells"; return -1; }
  if (p.ransac_fixed_hypotheses < 0 || p.ransac_max_iterations <= 0) { why = "ransac counts"; return -1; }
  return 0;
}

static void derive(DevParams &d) {
  const sloam_params &p = d.p;
  d.N = p.img_h * p.img_w;
  d.B = p.groundRadiiBins * p.groundThetaBins;
  // Segmentation constructor, inference.cpp:7-9: double expressions stored to float members
  d.fov_up = (float)((double)p.fov_up_deg / 180.0 * 3.14159265358979323846);
  d.fov_down = (float)((double)p.fov_down_deg / 180.0 * 3.14159265358979323846);
  d.fov = fabsf(d.fov_down) + fabsf(d.fov_up);
  d.pg.fov_down_abs = fabsf(d.fov_down);
  d.pg.fov = d.fov;
  d.pg.Wf = (float)p.img_w;
  d.pg.Hf = (float)p.img_h;
  // sloam.cpp:348-352
  d.gg.max_dist = p.maxGroundLidarDist;
  d.gg.min_dist = p.minGroundLidarDist;
  d.gg.radial_step = p.maxGroundLidarDist / (double)p.groundRadiiBins;
  d.gg.theta_step = 2 * 3.14159265 / (double)p.groundThetaBins;
  d.gg.RB = p.groundRadiiBins;
  d.gg.TB = p.groundThetaBins;
  d.gg.inv_radial_step_f = (float)(1.0 / d.gg.radial_step);
  d.gg.inv_theta_step_f = (float)(1.0 / d.gg.theta_step);
  ground_geom_thresholds(d.gg);  // exact r^2 thresholds of the radius tests and the radial bin (proj_math.h)
  auto sq_cut = [](float t) {
    if (!(t > 0.f)) return 0.f;  // sqrtf(s) < t never holds for s >= 0
    float s = t * t;
    while (s > 0.f && sqrtf(s) >= t) s = nextafterf(s, 0.f);
    while (sqrtf(s) < t) s = nextafterf(s, INFINITY);
    return s;
  };
  d.cluster_sq_cut = sq_cut(p.cluster_dist_thresh);
  d.centroid_sq_cut = sq_cut(p.max_dist_to_centroid);
  d.magic_w = (unsigned)((1ull << 32) / (unsigned)p.img_w) + 1u;
  auto bits_for = [](int n) { int b = 1; while ((1 << b) < n) ++b; return b; };  // values 0..n-1
  d.vw_row_bits = bits_for(p.img_h);
  d.vw_slot_bits = bits_for(p.max_trees);
}

static int upload_tables(sloam_ctx *c) {
  const sloam_params &p = c->hp.p;
  const int draws = ransac_draws_per_tree(p);
  std::vector<int32_t> pairs, offs(p.max_tree_vertices + 2, 0);
  for (int V = 0; V <= p.max_tree_vertices; ++V) {
    offs[V] = (int32_t)(pairs.size() / 2);
    if (V >= 2) ransac_table(V, draws, pairs);
    else pairs.resize(pairs.size() + (size_t)2 * draws, 0);
  }
  offs[p.max_tree_vertices + 1] = (int32_t)(pairs.size() / 2);
  SB_CUDA(c, cudaMemcpyAsync(c->ws.ransac_pairs, pairs.data(), pairs.size() * sizeof(int32_t),
                             cudaMemcpyHostToDevice, c->stream));
  SB_CUDA(c, cudaMemcpyAsync(c->ws.ransac_pairs_offset, offs.data(), offs.size() * sizeof(int32_t),
                             cudaMemcpyHostToDevice, c->stream));
  SB_CUDA(c, cudaMemcpyAsync(c->dp, &c->hp, sizeof(DevParams), cudaMemcpyHostToDevice, c->stream));
  SB_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLOAM_OK;
}

}  // namespace sb

using namespace sb;

extern "C" {

const char *sloam_b200_version(void) { return "sloam_b200 0.1 (sm_100a)"; }

void sloam_b200_default_params(sloam_params *p) {
  std::memset(p, 0, sizeof *p);
  p->img_h = 64; p->img_w = 1024;
  p->fov_up_deg = 22.5f; p->fov_down_deg = -22.5f;  // sloamNode.cpp:77,81
  p->do_destagger = 0;                              // params/sim.yaml:5
  // params/sloam.yaml over the code defaults of sloamNode.cpp:57-128
  p->scansPerSweep = 1;
  p->minTreeModels = 5; p->minGroundModels = 36;
  p->maxLidarDist = 20; p->maxGroundLidarDist = 25; p->minGroundLidarDist = 5;
  p->twoStepOptim = 1;
  p->groundRadiiBins = 2; p->groundThetaBins = 18;
  p->groundRetainThresh = 0.05;
  p->groundMatchThresh = 2.0; p->roughTreeMatchThresh = 3.0;
  p->treeMatchThresh = 0.5;
  p->maxTreeRadius = 0.3; p->maxAxisTheta = 10; p->maxFocusOutlierDistance = 0.5;
  p->AddNewTreeThreshDist = 1.5;
  p->featuresPerTree = 20; p->numGroundFeatures = 5;
  p->defaultTreeRadius = 0.2;
  p->max_dist_to_centroid = 0.2f;
  p->cluster_dist_thresh = 1.0f;
  p->min_cluster_points = 80; p->min_vertex_points = 3;
  p->min_tree_vertices = 16; p->max_tree_vertices = 56;
  p->ransac_threshold = 0.25; p->ransac_max_iterations = 50; p->ransac_probability = 0.99;
  p->ransac_fixed_hypotheses = 0;
  p->min_tree_height_sq = 1.5; p->root_plane_max_dist = 2.0;
  p->plane_match_thresh = 1.0; p->ground_angle_tol = 0.1; p->huber_delta = 0.1;
  p->lm_max_iterations = 50;
  p->max_trees = 512; p->max_map_models = 512; p->max_prev_planes = 64;
}

int sloam_b200_create(const sloam_params *p, int device, int max_keyframes, sloam_ctx **out) {
  // batch keyframe index and 32-pixel word index are packed 16 + 16 bits (k3_trellis.cu)
  if (!p || !out || max_keyframes <= 0 || max_keyframes > 65535) return SLOAM_E_INVALID;
  if ((long long)p->img_h * p->img_w > (1ll << 21) || p->img_w > 2048) return SLOAM_E_INVALID;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || device < 0 || device >= n_dev)
    return SLOAM_E_NODEVICE;  // no CPU fallback
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SLOAM_E_NODEVICE;
  if (prop.major < 10) return SLOAM_E_NODEVICE;  // built for sm_100a only
  std::string why;
  if (validate(*p, why) != 0) return SLOAM_E_INVALID;
  sloam_ctx *c = new (std::nothrow) sloam_ctx();
  if (!c) return SLOAM_E_NOMEM;
  c->device = device;
  c->max_k = max_keyframes;
  c->sm_count = prop.multiProcessorCount;
  c->hp.p = *p;
  derive(c->hp);
  {  // the packed (keyframe, slot, row) work items must fit 31 bits
    int kb = 1;
    while ((1 << kb) < max_keyframes) ++kb;
    if (c->hp.vw_row_bits + c->hp.vw_slot_bits + kb > 31) { delete c; return SLOAM_E_INVALID; }
  }
  if (cudaSetDevice(device) != cudaSuccess) { delete c; return SLOAM_E_CUDA; }
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return SLOAM_E_CUDA; }
  c->own_stream = true;
  if (cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return SLOAM_E_CUDA;
  }
  Bump dry{nullptr};
  layout(c, dry);
  c->arena_bytes = align_up(dry.off, 256);
  if (cudaMalloc(&c->arena, c->arena_bytes) != cudaSuccess) {
    cudaStreamDestroy(c->stream);
    delete c;
    return SLOAM_E_NOMEM;
  }
  Bump real{(char *)c->arena};
  layout(c, real);
  if (cudaMalloc((void **)&c->dp, sizeof(DevParams)) != cudaSuccess) {
    cudaFree(c->arena); cudaStreamDestroy(c->stream); delete c;
    return SLOAM_E_NOMEM;
  }
  const int rc = upload_tables(c);
  if (rc != SLOAM_OK) { sloam_b200_destroy(c); return rc; }
  *out = c;
  return SLOAM_OK;
}

void sloam_b200_map_free(sloam_ctx *c);
int sloam_b200_comm_destroy(sloam_ctx *c);

void sloam_b200_destroy(sloam_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  for (sloam_ctx *&l : c->lane) { if (l) sloam_b200_destroy(l); l = nullptr; }
  if (c->ev_lane_start) cudaEventDestroy(c->ev_lane_start);
  for (cudaEvent_t e : c->ev_lane_done) if (e) cudaEventDestroy(e);
  sloam_b200_map_free(c);
  sloam_b200_comm_destroy(c);
  cudaDeviceSynchronize();
  if (c->ev_gather_done) cudaEventDestroy(c->ev_gather_done);
  if (c->arena) cudaFree(c->arena);
  if (c->dp) cudaFree(c->dp);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->stage_dev) cudaFree(c->stage_dev);
  if (c->assoc_part_i) cudaFree(c->assoc_part_i);
  if (c->assoc_part_d) cudaFree(c->assoc_part_d);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->side) cudaStreamDestroy(c->side);
  for (cudaEvent_t e : c->prof_ev) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_chunk) if (e) cudaEventDestroy(e);
  if (c->ev_stage_free) cudaEventDestroy(c->ev_stage_free);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  delete c;
}

int sloam_b200_set_params(sloam_ctx *c, const sloam_params *p) {
  if (!c || !p) return SLOAM_E_INVALID;
  std::string why;
  if (validate(*p, why) != 0) return set_err(c, SLOAM_E_INVALID, why);
  const sloam_params &o = c->hp.p;
  // anything that sizes the arena must not grow
  if (p->img_h * p->img_w > o.img_h * o.img_w || p->img_h > o.img_h ||
      p->groundRadiiBins * p->groundThetaBins > o.groundRadiiBins * o.groundThetaBins ||
      p->numGroundFeatures > o.numGroundFeatures || p->featuresPerTree > o.featuresPerTree ||
      p->max_trees > o.max_trees || p->max_tree_vertices > o.max_tree_vertices ||
      p->max_prev_planes > o.max_prev_planes || (p->do_destagger && !o.do_destagger && !c->ws.tree2) ||
      ransac_draws_per_tree(*p) > ransac_draws_per_tree(o))
    return set_err(c, SLOAM_E_INVALID, "set_params: capacities/image size must not grow; create a new context");
  // keep the arena layout of the creation-time parameters: only values change
  sloam_params np = *p;
  c->hp.p = np;
  ++c->epoch;
  derive(c->hp);
  for (sloam_ctx *l : c->lane)
    if (l) { const int rc = sloam_b200_set_params(l, p); if (rc != SLOAM_OK) return set_err(c, rc, "set_params: lane"); }
  return upload_tables(c);
}

int sloam_b200_get_params(const sloam_ctx *c, sloam_params *p) {
  if (!c || !p) return SLOAM_E_INVALID;
  *p = c->hp.p;
  return SLOAM_OK;
}

int sloam_b200_set_stream(sloam_ctx *c, void *s) {
  if (!c) return SLOAM_E_INVALID;
  ++c->epoch;
  if (c->own_stream && c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
  if (s) { c->stream = (cudaStream_t)s; c->own_stream = false; }
  else {
    SB_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  return SLOAM_OK;
}

int sloam_b200_sync(sloam_ctx *c) {
  if (!c) return SLOAM_E_INVALID;
  SB_CUDA(c, cudaStreamSynchronize(c->stream));
  return SLOAM_OK;
}

const char *sloam_b200_last_error(const sloam_ctx *c) { return c ? c->err.c_str() : "null context"; }
int64_t sloam_b200_kernel_launches(const sloam_ctx *c) {
  if (!c) return 0;
  int64_t n = c->launches;
  for (const sloam_ctx *l : c->lane) if (l) n += l->launches;
  return n;
}

int sloam_b200_set_lanes(sloam_ctx *c, int n) {
  if (!c || n < 1 || n > 4) return SLOAM_E_INVALID;
  cudaSetDevice(c->device);
  SB_CUDA(c, cudaStreamSynchronize(c->stream));
  for (sloam_ctx *&l : c->lane) { if (l) sloam_b200_destroy(l); l = nullptr; }
  c->n_lanes = 1;
  if (n == 1) return SLOAM_OK;
  if (!c->ev_lane_start) {
    SB_CUDA(c, cudaEventCreateWithFlags(&c->ev_lane_start, cudaEventDisableTiming));
    for (cudaEvent_t &e : c->ev_lane_done) SB_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  const int per = (c->max_k + n - 1) / n;
  for (int l = 0; l < n; ++l) {
    const int rc = sloam_b200_create(&c->hp.p, c->device, per, &c->lane[l]);
    if (rc != SLOAM_OK) {
      for (sloam_ctx *&q : c->lane) { if (q) sloam_b200_destroy(q); q = nullptr; }
      return set_err(c, rc, "set_lanes: could not create a lane context");
    }
    c->lane[l]->parent = c;
  }
  c->n_lanes = n;
  return SLOAM_OK;
}
int64_t sloam_b200_workspace_bytes(const sloam_ctx *c) { return c ? (int64_t)c->arena_bytes : 0; }

static const char *const kProfNames[P_COUNT] = {
    "project_split_kernel", "range_finalize_kernel", "ground_offsets+ground_scatter_kernel", "ground_cells_kernel<0>",
    "ground_cells_kernel<1> (tie replay)", "ground_fit+plane_finish+planes_compact", "cc_rows_kernel", "cc_label_kernel",
    "vertex_kernel", "vertex_replay+vertex_wide (tie replay)", "tree_compact_kernel", "cylinder_kernel+compact",
    "assoc_kernel (sensor frame)", "build_matches_kernel", "lm_kernel", "finish_kernel",
    "assoc_kernel+matches (map frame)"};

int sloam_b200_profile_enable(sloam_ctx *c, int on) {
  if (!c) return SLOAM_E_INVALID;
  cudaSetDevice(c->device);
  if (on)
    for (cudaEvent_t &e : c->prof_ev)
      if (!e) SB_CUDA(c, cudaEventCreate(&e));
  c->prof_on = on != 0;
  c->prof_n = 0;
  for (unsigned &m : c->prof_seen) m = 0;
  for (sloam_ctx *l : c->lane)
    if (l) { const int rc = sloam_b200_profile_enable(l, on); if (rc != SLOAM_OK) return rc; }
  return SLOAM_OK;
}

// Summed time of kernel group `id` over the fused runs since the last enable/read.  With lanes
// the kernels of the sub-batches overlap: a run counts from the earliest start to the latest
// end over the lanes (event timestamps are device-wide).
static int prof_sum(sloam_ctx *c, int id, double *total_ms, int *runs_out) {
  sloam_ctx *src[4] = {c, nullptr, nullptr, nullptr};
  int ns = 1;
  if (c->n_lanes > 1 && c->lane[0] && c->lane[0]->prof_n > 0) {
    ns = c->n_lanes;
    for (int l = 0; l < ns; ++l) src[l] = c->lane[l];
  }
  int runs = 0;
  for (int l = 0; l < ns; ++l) runs = std::max(runs, std::min(src[l]->prof_n, (int)sloam_ctx::kProfRuns));
  double total = 0.0;
  int counted = 0;
  for (int i = 0; i < runs; ++i) {
    float span = -1.f;
    for (int a = 0; a < ns; ++a)
      for (int b = 0; b < ns; ++b) {
        if (src[a]->prof_n <= i || src[b]->prof_n <= i) continue;
        if (!((src[a]->prof_seen[i] >> id) & 1u) || !((src[b]->prof_seen[i] >> id) & 1u)) continue;
        float ms = 0.f;
        SB_CUDA(c, cudaEventElapsedTime(&ms, src[a]->prof_ev[((size_t)i * sloam_ctx::kProfIds + id) * 2],
                                        src[b]->prof_ev[((size_t)i * sloam_ctx::kProfIds + id) * 2 + 1]));
        if (ms > span) span = ms;
      }
    if (span >= 0.f) { total += span; ++counted; }
  }
  *total_ms = total;
  *runs_out = counted;
  return SLOAM_OK;
}

static int prof_sync_all(sloam_ctx *c) {
  SB_CUDA(c, cudaStreamSynchronize(c->stream));
  SB_CUDA(c, cudaStreamSynchronize(c->side));
  for (int l = 0; l < c->n_lanes && c->n_lanes > 1; ++l)
    if (c->lane[l]) {
      SB_CUDA(c, cudaStreamSynchronize(c->lane[l]->stream));
      SB_CUDA(c, cudaStreamSynchronize(c->lane[l]->side));
    }
  return SLOAM_OK;
}

static void prof_reset(sloam_ctx *c) {
  c->prof_n = 0;
  for (unsigned &m : c->prof_seen) m = 0;
  for (sloam_ctx *l : c->lane)
    if (l) { l->prof_n = 0; for (unsigned &m : l->prof_seen) m = 0; }
}

int sloam_b200_profile_read(sloam_ctx *c, double *split_kernel_ms, int32_t *launches) {
  if (!c || !split_kernel_ms || !launches) return SLOAM_E_INVALID;
  int rc = prof_sync_all(c);
  if (rc != SLOAM_OK) return rc;
  int runs = 0;
  rc = prof_sum(c, P_SPLIT, split_kernel_ms, &runs);
  *launches = runs;
  prof_reset(c);
  return rc;
}

int sloam_b200_profile_read_kernels(sloam_ctx *c, sloam_prof_kernel *out, int cap, int32_t *n_out) {
  if (!c || !out || !n_out || cap < 0) return SLOAM_E_INVALID;
  int rc = prof_sync_all(c);
  if (rc != SLOAM_OK) return rc;
  int n = 0;
  for (int id = 0; id < P_COUNT && n < cap; ++id) {
    double ms = 0.0;
    int runs = 0;
    rc = prof_sum(c, id, &ms, &runs);
    if (rc != SLOAM_OK) return rc;
    if (runs == 0) continue;
    std::memset(&out[n], 0, sizeof out[n]);
    std::strncpy(out[n].name, kProfNames[id], sizeof(out[n].name) - 1);
    out[n].ms = ms;
    out[n].launches = runs;
    ++n;
  }
  *n_out = n;
  prof_reset(c);
  return SLOAM_OK;
}

}  // extern "C"
