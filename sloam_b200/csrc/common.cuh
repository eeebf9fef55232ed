// common.cuh -- context, scratch layout and small device helpers shared by the
// sm_100a kernels of the SLOAM hot path.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include "../../include/sloam_b200.h"
#include "proj_math.h"

#define SLOAM_HD_FN __host__ __device__ __forceinline__

#include "dev_geom.h"

namespace sb {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kMaxCells = 256;        // groundRadiiBins * groundThetaBins upper bound
constexpr int kSplitTile = 1024;      // points per CTA in the project/split kernel

// Device copy of the parameters plus derived constants.
struct DevParams {
  sloam_params p;
  int N;          // img_h * img_w
  int B;          // ground cells
  float fov_up, fov_down, fov;  // radians, float like the reference members
  ProjGeom pg;
  GroundGeom gg;
  // sqrtf(s) < t  <=>  s < sq_cut(t): the smallest float whose correctly rounded square
  // root reaches t (sqrtf is monotone), so distance tests need no square root
  float cluster_sq_cut;   // t = cluster_dist_thresh (trellis.cpp clustering tolerance)
  float centroid_sq_cut;  // t = max_dist_to_centroid
  unsigned magic_w;       // floor(2^32 / img_w) + 1: i / img_w == umulhi(i, magic_w) for i < 2^21
  // (keyframe, big-cluster slot, row) work items of the vertex kernels in one int32:
  // row in the low vw_row_bits, slot in the next vw_slot_bits, keyframe above (create checks
  // that max_keyframes fits)
  int vw_row_bits, vw_slot_bits;
};

// hand-over record between the per-cell QR (one CTA per cell) and the per-cell
// 3x3 Jacobi + acceptance (one THREAD per cell): k2_ground.cu
struct FitRec {
  double W[9], U[9];
  float c[3];
  int32_t n_cell, n_kept, valid;
  int32_t off_all;  // first member slot of the cell (select kernel -> fit kernel)
};

// Scratch owned by the context, sized for max_keyframes.
struct Workspace {
  // K1
  int32_t *zero_begin = nullptr, *zero_end = nullptr;  // bounds of the block zeroed per fused run
  int32_t *pix = nullptr;            // [K][N]
  sloam_point *tree = nullptr;       // [K][N]
  sloam_point *ground = nullptr;     // [K][N]
  int32_t *ground_count = nullptr;   // [K]
  uint32_t *tree_bits = nullptr;     // [K][ceil(N/32)] bit i: pixel i may hold a tree point
  sloam_point *tree2 = nullptr;      // [K][N] masked cloud before the destagger pass (do_destagger only)
  uint32_t *tree_bits2 = nullptr;    // [K][ceil(N/32)]  "
  uint8_t *ground_cell = nullptr;    // [K][N] polar cell of each ground point (255 = none)
  int32_t *cell_count = nullptr;     // [K][kMaxCells]
  int32_t *tile_count = nullptr;     // [K][tiles] ground points of each K1 tile (tile-strided ground layout)
  float *range_image = nullptr;      // [K][N] (when the caller passes none)
  // K2
  sloam_cell_plane *cells = nullptr; // [K][B]
  sloam_point *cell_features = nullptr; // [K][B][Fg]
  sloam_plane *planes_acc = nullptr; // [K][B] accepted planes, compact, sensor frame
  int32_t *planes_acc_cell = nullptr;// [K][B] cell index of each accepted plane
  int32_t *n_planes_acc = nullptr;   // [K]
  unsigned long long *gscratch2 = nullptr; // [K][N] kept records of oversized cells (gscratch stays intact)
  int32_t *tied_cells = nullptr;     // [K][kMaxCells] (keyframe << 8 | cell) of cells with exact z ties
  int32_t *n_tied_cells = nullptr;   // [1]
  unsigned long long *gscratch = nullptr; // [K][N] fused: (z key, point index) records from the split kernel, tile-strided, input order
  unsigned long long *gscratch3 = nullptr; // [K][N] (z key, index) member lists, contiguous per cell
  uint32_t *seg_tab = nullptr;       // [K][B][tiles] points of each cell in each K1 tile, then their first member slot
  double *qscratch = nullptr;        // [K][N][3] QR workspace of oversized cells
  float *pscratch = nullptr;         // [K][N][3] point staging of oversized cells
  FitRec *fit_rec = nullptr;         // [K][B]
  // K3 (k3_trellis.cu)
  uint32_t *cc_planes = nullptr;     // [K][ceil(N/32)] uint4 (valid, run start, connected upwards, 0) bit planes
  int32_t *cc_wbase = nullptr;       // [K][ceil(N/32)] run id base of each word (find_clusters entry only)
  int32_t *run_par = nullptr;        // [K][N] union-find over runs: global fallback of the shared-memory arrays
  int32_t *run_siz = nullptr;        // [K][N]  "
  int32_t *run_len = nullptr;        // [K][N]  "
  int32_t *run_pix0 = nullptr;       // [K][N] first pixel of each run
  int32_t *run_info = nullptr;       // [K][N] length << 11 | slot code of each run
  int32_t *run_label = nullptr;      // [K][N] PCL label of each run's component
  int32_t *n_big = nullptr;          // [K] big components kept (<= max_trees)
  int32_t *n_roots = nullptr;        // [K] components
  int32_t *big_rank = nullptr;       // [K][max_trees] PCL label of each big component
  uint32_t *slot_rows = nullptr;     // [K][max_trees][ceil(H/32)] rows of a big component that hold a vertex record
  sloam_vertex *slot_vertices = nullptr; // [K][max_trees][H] one candidate vertex per (component, row)
  void *vitems = nullptr;            // [K << (row bits + slot bits)] VItem work items of the vertex stage
  int32_t *vitem_pool = nullptr;     // [K << (row bits + slot bits)] first vertex-point slot of each item
  int32_t *vlists = nullptr;         // [6][K*max_trees*H] item ids by size class (+ wide, tied)
  int32_t *n_vlists = nullptr;       // [8] lengths of the class lists
  int32_t *kf_flags = nullptr;       // [K] bit 0: more big components than max_trees (first max_trees kept)
  sloam_tree *trees = nullptr;       // [K][max_trees]
  int32_t *n_trees = nullptr;        // [K]
  sloam_vertex *vertices = nullptr;  // [K][max_trees*max_tree_vertices]
  sloam_point *vertex_points = nullptr; // [K][N]
  // K4
  sloam_tree_model *tree_models = nullptr; // [K][max_trees]
  sloam_point *tree_features = nullptr;    // [K][max_trees][Ft]
  int32_t *ransac_pairs = nullptr;   // draw tables per V (device copy)
  int32_t *ransac_pairs_offset = nullptr; // [max_tree_vertices+1]
  // K5/K6
  sloam_cylinder *lm_cyl = nullptr;  // [K][max_trees] valid cylinders, compact (sensor frame)
  int32_t *lm_src = nullptr;         // [K][max_trees] tree index of each compact cylinder
  int32_t *n_lm = nullptr;           // [K]
  int32_t *assoc_idx = nullptr;      // [K][max_trees]
  double *assoc_dist = nullptr;      // [K][max_trees]
  double *res_tree_feat = nullptr;   // [K][max_trees*Ft][3] matched tree features (sensor frame)
  sloam_cylinder *res_tree_obj = nullptr; // [K][max_trees*Ft] matched map cylinder per feature
  double *res_plane_feat = nullptr;  // [K][B*Fg][3]
  sloam_plane *res_plane_obj = nullptr;   // [K][B*Fg]
  int32_t *n_tree_res = nullptr;     // [K]
  int32_t *n_plane_res = nullptr;    // [K]
  uint8_t *optim_flags = nullptr;    // [K][2] treeCheck, groundCheck
  uint8_t *kf_mode = nullptr;        // [K] 0 optimise, 1 first scan, 2 bailed out
  double *lm_x = nullptr;            // [K][2][8] solver output parameters
  int32_t *lm_info = nullptr;        // [K][2][2] iterations, termination
  sloam_pose *curr_pose = nullptr;   // [K]
  sloam_kf_result *results = nullptr;// [K]
  int32_t *matches = nullptr;        // [K][max_trees]
  sloam_cylinder *tm = nullptr;      // [K][max_trees]
  int32_t *tm_id = nullptr;          // [K][max_trees]
  sloam_plane *planes_out = nullptr; // [K][max_prev_planes]
  int32_t *n_planes_out = nullptr;   // [K]
};

}  // namespace sb

struct sloam_ctx {
  int device = 0;
  int max_k = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // side stream: the tree detector (K3) runs concurrently with the ground stage (K2);
  // both only depend on the split kernel (K1)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  sb::DevParams hp;              // host copy
  sb::DevParams *dp = nullptr;   // device copy
  sb::Workspace ws;
  void *arena = nullptr;         // one allocation backing the workspace
  size_t arena_bytes = 0;
  int64_t launches = 0;
  std::string err;
  // host staging for the *_host entry points
  void *pinned = nullptr;
  size_t pinned_bytes = 0;
  void *stage_dev = nullptr;
  size_t stage_dev_bytes = 0;
  int last_k = 0;
  // the fused path writes only the tree-labelled points of ws.tree (+ ws.tree_bits); the NaN
  // points of the dense cloud are filled in when the intermediates are asked for
  bool tree_sparse = false;
  // likewise ws.ground is tile-strided after a fused run; ground_dense (ws.qscratch) receives
  // the contiguous cloud on demand
  bool ground_strided = false;
  bool kf_flags_valid = false;  // ws.kf_flags belongs to the run in flight (set by the tree detector of a fused run)
  const sloam_point *last_points = nullptr;  // inputs of the last fused run (intermediates on demand)
  const uint8_t *last_mask = nullptr;
  // optional event pairs around the split kernel of fused runs (sloam_b200_profile_*)
  // run_keyframes_host: copy stream + per-chunk events (H2D of chunk j+1 overlaps compute of j)
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_chunk[8] = {};
  cudaEvent_t ev_stage_free = nullptr;
  // lanes: sub-contexts that run sub-batches of a fused run concurrently (sloam_b200_set_lanes)
  int n_lanes = 1;
  sloam_ctx *lane[4] = {};
  cudaEvent_t ev_lane_start = nullptr, ev_lane_done[4] = {};
  // set by zero_counters(): the launchers named by the bits may skip their own memsets once
  unsigned zero_valid = 0;  // 1 split kernel, 2 connected components, 4 vertex stage, 8 ground cells
  int epoch = 0;  // bumped by set_params / set_stream: invalidates captured CUDA graphs
  // optional event pairs around every kernel group of the fused runs (sloam_b200_profile_*):
  // prof_ev[(run * kProfIds + id) * 2 + {0 begin, 1 end}], run = fused run since enable/read
  bool prof_on = false;
  int prof_n = 0;
  static constexpr int kProfRuns = 64;
  static constexpr int kProfIds = 24;
  cudaEvent_t prof_ev[2 * kProfRuns * kProfIds] = {};
  unsigned prof_seen[kProfRuns] = {};  // bit id: the pair of that run was recorded
  // partial results of split association (large maps), grown on demand
  int32_t *assoc_part_i = nullptr;
  double *assoc_part_d = nullptr;
  size_t assoc_part_cap = 0;
  void *seq = nullptr;  // sloam_seq_state (k7_map.cu): semantic map + sequential state
  // multi-GPU gather (comm.cu): communicator + side stream; the output kernels of a fused run
  // wait for a gather in flight, which may still be reading the output buffers
  void *comm = nullptr;
  cudaEvent_t ev_gather_done = nullptr;
  bool gather_pending = false;
  sloam_ctx *parent = nullptr;  // lane contexts: the context that owns them
};

namespace sb {

// kernel groups timed by sloam_b200_profile_* (names: ctx.cu kProfNames, same order)
enum ProfId {
  P_SPLIT = 0, P_RANGE_FIN, P_GROUND_BIN, P_GROUND_CELLS, P_GROUND_REPLAY, P_PLANE_FIT, P_CC_ROWS, P_CC_LABEL,
  P_VERTEX, P_VERTEX_REPLAY, P_TREE_COMPACT, P_CYLINDER, P_ASSOC_1, P_BUILD_MATCHES, P_LM, P_FINISH, P_ASSOC_2,
  P_COUNT
};
static_assert(P_COUNT <= sloam_ctx::kProfIds, "profile id table too small");

// event on the stream the NEXT / PREVIOUS kernel of the group is launched on (c->stream is
// switched to the side stream for the tree detector); no synchronisation is added
inline void prof_mark(sloam_ctx *c, int id, int end) {
  if (!c->prof_on || c->prof_n >= sloam_ctx::kProfRuns) return;
  cudaEventRecord(c->prof_ev[((size_t)c->prof_n * sloam_ctx::kProfIds + id) * 2 + end], c->stream);
  if (end) c->prof_seen[c->prof_n] |= 1u << id;
}
#define PROF_BEGIN(c, id) sb::prof_mark((c), (id), 0)
#define PROF_END(c, id) sb::prof_mark((c), (id), 1)

inline int set_err(sloam_ctx *c, int code, const std::string &m) {
  if (c) c->err = m;
  return code;
}

#define SB_CUDA(ctx, call)                                                        \
  do {                                                                            \
    cudaError_t e__ = (call);                                                     \
    if (e__ != cudaSuccess)                                                       \
      return sb::set_err(ctx, SLOAM_E_CUDA,                                       \
                         std::string(#call) + ": " + cudaGetErrorString(e__));    \
  } while (0)

#define SB_LAUNCH_CHECK(ctx)                                                      \
  do {                                                                            \
    (ctx)->launches++;                                                            \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess)                                                       \
      return sb::set_err(ctx, SLOAM_E_CUDA,                                       \
                         std::string("kernel launch (") + __FILE__ + ":" + std::to_string(__LINE__) + \
                             "): " + cudaGetErrorString(e__));                    \
  } while (0)

// 16-byte vector access to points (sloam_point is four floats; every point array
// of the ABI must be 16-byte aligned, which cudaMalloc / torch allocations are)
__device__ __forceinline__ sloam_point ld_point(const sloam_point *p) {
  const float4 v = *reinterpret_cast<const float4 *>(p);
  sloam_point r; r.x = v.x; r.y = v.y; r.z = v.z; r.intensity = v.w;
  return r;
}
__device__ __forceinline__ void st_point(sloam_point *p, const sloam_point &v) {
  *reinterpret_cast<float4 *>(p) = make_float4(v.x, v.y, v.z, v.intensity);
}

// Comparison as 1.0f / 0.0f (PTX set -> one FSET.BF): see the counting loop of k3_trellis.cu.
// NaN compares false, like `<` and `==`.
__device__ __forceinline__ float flt(float a, float b) {
  float d;
  asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ float feq(float a, float b) {
  float d;
  asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}

// ---- float helpers with the reference's operation order -----------------
// Eigen Vector3f::norm / squaredNorm: x^2 + (y^2 + z^2)  (Redux.h unroller)
SLOAM_HD_FN float sqnorm3f(float dx, float dy, float dz) { return dx * dx + (dy * dy + dz * dz); }
__device__ __forceinline__ int vw_pack(const DevParams *dp, int k, int slot, int row) {
  return (k << (dp->vw_row_bits + dp->vw_slot_bits)) | (slot << dp->vw_row_bits) | row;
}
__device__ __forceinline__ void vw_unpack(const DevParams *dp, int w, int &k, int &slot, int &row) {
  row = w & ((1 << dp->vw_row_bits) - 1);
  slot = (w >> dp->vw_row_bits) & ((1 << dp->vw_slot_bits) - 1);
  k = w >> (dp->vw_row_bits + dp->vw_slot_bits);
}

// i / W and i % W without an integer division (i < 2^21, W <= 2^11: the error term
// i * (magic * W - 2^32) stays below 2^32, so the high word is the exact quotient)
__device__ __forceinline__ int fast_div_w(int i, unsigned magic_w) { return (int)__umulhi((unsigned)i, magic_w); }

SLOAM_HD_FN float dist3f(float ax, float ay, float az, float bx, float by, float bz) {
  return sqrtf(sqnorm3f(ax - bx, ay - by, az - bz));
}

// order-preserving map float -> uint32 (for sorting / selection keys)
__device__ __forceinline__ uint32_t float_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

}  // namespace sb
