/*
 * sloam_b200.h -- C ABI of the B200-native SLOAM per-keyframe hot path.
 *
 * This is the drop-in boundary.  Every entry point is
 * `extern "C"`, takes plain pointers and sizes, returns an int status
 * (0 = ok, negative = error, see SLOAM_E_*), and never throws or aborts.
 * Each one names the reference interface it replaces (file:line relative to
 * the KumarRobotics/sloam tree).
 *
 * Conventions
 *   - A context is bound to one GPU and one CUDA stream, owns all scratch
 *     memory, and is NOT thread-safe (the reference core is single-threaded
 *     and non-reentrant: sloam/include/core/sloam.h:99-106).
 *   - All *_dev entry points take DEVICE pointers and are asynchronous on
 *     the context stream; call sloam_b200_sync() before reading results.
 *     The *_host entry points take HOST pointers, stage through context-owned
 *     pinned/device buffers and return after the results are on the host.
 *   - Everything is batched over K keyframes (leading dimension K); a single
 *     reference call is K = 1.
 *   - There is no CPU fallback: if no CUDA device is usable, create() fails.
 */
#ifndef SLOAM_B200_H
#define SLOAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
#define SLOAM_OK 0
#define SLOAM_E_INVALID (-1)   /* bad argument / capacity exceeded        */
#define SLOAM_E_CUDA (-2)      /* CUDA runtime error (see last_error)     */
#define SLOAM_E_NOMEM (-3)
#define SLOAM_E_NODEVICE (-4)  /* no usable sm_100 device: no CPU fallback */

/* per-keyframe status of RunSloam (sloam/src/core/sloam.cpp:453-532) */
#define SLOAM_KF_OK 0            /* returned true                                  */
#define SLOAM_KF_EMPTY_MAP 1     /* sloam.cpp:476-480, returned false              */
#define SLOAM_KF_NO_MODELS 2     /* sloam.cpp:482-486, returned false              */
#define SLOAM_KF_NOT_CONVERGED 3 /* joint OptimizePose did not converge (:241,:507)*/
/* The code above is the low byte of sloam_kf_result::status (SLOAM_KF_CODE).  Capacity
 * flags are OR-ed on top; the reference has no capacities (its vectors grow), so a set flag
 * means the result was computed on truncated data:
 *   TREE_CAPACITY  more clusters above min_cluster_points than max_trees; the first
 *                  max_trees in PCL label order were kept (deterministic)
 *   MAP_CAPACITY   sequential mode: the semantic map is full, new landmarks were dropped */
#define SLOAM_KF_CODE(status) ((status) & 0xFF)
#define SLOAM_KF_FLAG_TREE_CAPACITY 0x100
#define SLOAM_KF_FLAG_MAP_CAPACITY 0x200

/* ------------------------------------------------------------------- types */

/* PointT = pcl::PointXYZI (definitions.h:42).  PCL pads it to 32 B; on the
 * device it is one float4. */
typedef struct sloam_point {
  float x, y, z, intensity;
} sloam_point;

/* SE3 = Sophus::SE3d (definitions.h:33): translation + unit quaternion.
 * q is stored x,y,z,w (Eigen coefficient order). */
typedef struct sloam_pose {
  double t[3];
  double q[4];
} sloam_pose;

/* The part of CylinderParameters (cylinder.h:14-23) that association and the
 * optimiser read: 56 B. */
typedef struct sloam_cylinder {
  double root[3];
  double ray[3];
  double radius;
} sloam_cylinder;

/* PlaneParameters (plane.h:14-19): 56 B. */
typedef struct sloam_plane {
  double plane[4];
  double centroid[3];
} sloam_plane;

/* TreeVertex (definitions.h:56-66) flattened: `points` become a slice
 * [point_begin, point_begin + n_points) of a per-keyframe point pool.
 * beam and prevVertexSize are always 0 in the reference (trellis.cpp:53,58)
 * and are not stored; `row` is the scan line the vertex came from. */
typedef struct sloam_vertex {
  float cx, cy, cz; /* coords (component-wise median, trellis.cpp:74-87) */
  float radius;     /* ||first - last|| of the kept points (trellis.cpp:98) */
  int32_t n_points;
  int32_t point_begin;
  int32_t row;
  int32_t is_valid;
} sloam_vertex;

/* One landmark = std::vector<TreeVertex> (trellis.cpp:110-128). */
typedef struct sloam_tree {
  int32_t tree_id;      /* PCL cluster label (trellis.cpp:52,120) */
  int32_t n_vertices;   /* 17..56 with the reference constants     */
  int32_t vertex_begin; /* into the per-keyframe vertex array      */
  int32_t n_points;     /* sum of n_points over the vertices       */
} sloam_tree;

/* Result of the Plane constructor for one polar cell (plane.cpp:3-17,96-128)
 * plus the validity test of computeModels (sloam.cpp:401-410). */
typedef struct sloam_cell_plane {
  sloam_plane model;
  int32_t n_cell;    /* points binned into the cell (sloam.cpp:358)        */
  int32_t n_kept;    /* after bottom-k% retention (sloam.cpp:362-385)       */
  int32_t is_valid;  /* Plane::isValid (plane.cpp:7-16)                     */
  int32_t accepted;  /* isValid && angleCheck && heightCheck (sloam.cpp:409) */
} sloam_cell_plane;

/* Result of the Cylinder constructor for one tree (cylinder.cpp:3-28). */
typedef struct sloam_tree_model {
  sloam_cylinder model; /* sensor frame                                    */
  int32_t id;           /* vertices[2].treeId (cylinder.cpp:78,100)         */
  int32_t is_valid;     /* cylinder.cpp:26                                  */
  int32_t plane_index;  /* nearest accepted plane (sloam.cpp:420-431)       */
  int32_t n_inliers;    /* inliers of the winning RANSAC hypothesis         */
  int32_t best_hypothesis; /* index of the winning hypothesis (first max)   */
  int32_t n_hypotheses; /* hypotheses actually scored (adaptive exit)       */
  int32_t n_refit_inliers; /* inliers after the PCA refit                   */
  int32_t reserved;
} sloam_tree_model;

/* Per-keyframe output record of RunSloam (SloamOutput, sloam.h:48-55). */
typedef struct sloam_kf_result {
  int32_t status;  /* SLOAM_KF_* code | SLOAM_KF_FLAG_*            */
  int32_t success; /* the bool RunSloam returns                    */
  int32_t n_ground;        /* points in groundCloud                */
  int32_t n_planes;        /* accepted ground planes               */
  int32_t n_trees;         /* landmarks from computeGraph          */
  int32_t n_landmarks;     /* valid cylinders (= out.tm.size())    */
  int32_t n_tree_matches;  /* treeMatches.size()  (sloam.cpp:489)  */
  int32_t n_plane_matches; /* planeMatches.size() (sloam.cpp:490)  */
  int32_t lm_iterations[2]; /* joint: [0]; two-step: [0]=XYYaw [1]=ZRollPitch */
  int32_t lm_termination[2]; /* 0 CONVERGENCE, 1 NO_CONVERGENCE, 2 FAILURE, -1 not run */
  sloam_pose T_Map_Curr;
  sloam_pose T_Delta;
} sloam_kf_result;

/* Parameters.  FeatureModelParams (definitions.h:76-105) with the reference
 * member names, Instance::Params (trellis.h:31-40), the sensor geometry of
 * the Segmentation constructor (inference.cpp:5-14), and the literals of the
 * hot path promoted to parameters with the reference values as defaults
 * (SURVEY.md appendix D / B-14).  Fill with sloam_b200_default_params(). */
typedef struct sloam_params {
  /* sensor: Segmentation(model, fov_up, fov_down, img_w, img_h, ...) */
  int32_t img_h, img_w;
  float fov_up_deg, fov_down_deg; /* sloamNode.cpp:77,81: +22.5 / -22.5 */
  int32_t do_destagger;           /* inference.cpp:200-228             */

  /* FeatureModelParams */
  int32_t scansPerSweep;
  double minTreeModels, minGroundModels;
  double maxLidarDist, maxGroundLidarDist, minGroundLidarDist;
  int32_t twoStepOptim;
  int32_t groundRadiiBins, groundThetaBins;
  double groundRetainThresh;
  double groundMatchThresh, roughTreeMatchThresh; /* unused by the reference */
  double treeMatchThresh;
  double maxTreeRadius, maxAxisTheta, maxFocusOutlierDistance;
  double AddNewTreeThreshDist;
  int32_t featuresPerTree, numGroundFeatures;
  double defaultTreeRadius;

  /* Instance::Params + trellis.cpp literals */
  float max_dist_to_centroid;  /* trellis.h:34, YAML 0.2                 */
  float cluster_dist_thresh;   /* trellis.cpp:23   1.0                   */
  int32_t min_cluster_points;  /* trellis.cpp:109  > 80                  */
  int32_t min_vertex_points;   /* trellis.cpp:119  > 3                   */
  int32_t min_tree_vertices;   /* trellis.cpp:124  > 16                  */
  int32_t max_tree_vertices;   /* trellis.cpp:125-127  56                */

  /* cylinder.cpp literals + PCL SACSegmentation defaults */
  double ransac_threshold;     /* cylinder.cpp:122  0.25                 */
  int32_t ransac_max_iterations; /* PCL default 50                       */
  double ransac_probability;   /* PCL default 0.99                       */
  int32_t ransac_fixed_hypotheses; /* 0 = reference-faithful adaptive exit;
                                      >0 = score exactly this many, first
                                      maximum wins (BASELINE config 3)   */
  double min_tree_height_sq;   /* cylinder.cpp:109  1.5                  */
  double root_plane_max_dist;  /* cylinder.cpp:48   2.0                  */

  /* sloam.cpp literals */
  double plane_match_thresh;   /* sloam.cpp:490     1.0                  */
  double ground_angle_tol;     /* sloam.cpp:405-406 0.1 rad              */
  double huber_delta;          /* sloam.cpp:69,130,181  0.1              */
  int32_t lm_max_iterations;   /* sloam.cpp:92,150,222  50               */

  /* capacities of the flattened outputs (per keyframe) */
  int32_t max_trees;           /* landmarks per keyframe                 */
  int32_t max_map_models;      /* submap cylinders per keyframe          */
  int32_t max_prev_planes;     /* >= groundRadiiBins*groundThetaBins     */
} sloam_params;

typedef struct sloam_ctx sloam_ctx;

/* ----------------------------------------------------------- life cycle */

/* Reference values: code defaults of sloamNode.cpp:57-128 overridden by the
 * shipped sloam/params/sloam.yaml. */
void sloam_b200_default_params(sloam_params *p);

/* Replaces the constructors sloam::sloam() (sloam.cpp:6-12), Instance()
 * (trellis.cpp:13) and Segmentation() (inference.cpp:5-19, minus the ONNX
 * session).  max_keyframes = largest K any later call will pass; at most 65535, and
 * bits(img_h) + bits(max_trees) + bits(max_keyframes) <= 31 (packed work items), which
 * allows e.g. 32768 keyframes at 128 scan lines and 1024 trees.  SLOAM_E_INVALID otherwise. */
int sloam_b200_create(const sloam_params *p, int device, int max_keyframes,
                      sloam_ctx **out);
void sloam_b200_destroy(sloam_ctx *ctx);
/* setFmParams (sloam.h:67) + Instance::set_params (trellis.h:53). Capacities
 * and image size must not grow. */
int sloam_b200_set_params(sloam_ctx *ctx, const sloam_params *p);
int sloam_b200_get_params(const sloam_ctx *ctx, sloam_params *p);
/* Use an existing cudaStream_t (e.g. torch's current stream); NULL = own. */
int sloam_b200_set_stream(sloam_ctx *ctx, void *cuda_stream);
int sloam_b200_sync(sloam_ctx *ctx);
const char *sloam_b200_last_error(const sloam_ctx *ctx);
/* Kernels this library has launched on the context since creation. */
int64_t sloam_b200_kernel_launches(const sloam_ctx *ctx);
/* Bytes of device scratch owned by the context. */
int64_t sloam_b200_workspace_bytes(const sloam_ctx *ctx);
/* Device-side timing of the projection / label-split kernel (stage a1 + a2) inside
 * sloam_b200_run_keyframes_*: with profiling on, every fused run brackets that kernel
 * with CUDA events on the context's stream (no synchronisation is added).
 * sloam_b200_profile_read synchronises the stream, returns the summed kernel time in
 * milliseconds and the number of launches since the last enable/read, and resets both.
 * No counterpart in the reference (its timing is the wall clock of SLOAMNode::run,
 * sloamNode.cpp:200-233). */
/* Lanes: with n > 1 a fused run over K keyframes is cut into n sub-batches that run
 * concurrently, each on its own stream with its own scratch, so that the latency-bound tail
 * kernels of one sub-batch (pose optimiser, cylinder fits, per-cell plane fits) overlap the
 * issue-bound head kernels of another.  Results are identical to n = 1 (keyframes are
 * independent).  n = 1..4; each lane owns scratch for ceil(max_keyframes / n) keyframes in
 * addition to the context's own.  With n > 1 sloam_b200_get_intermediates describes the
 * first sub-batch.  Batches smaller than 64 keyframes per lane are not split.
 * No counterpart in the reference (one keyframe at a time, sloamNode.cpp:186). */
int sloam_b200_set_lanes(sloam_ctx *ctx, int n);
int sloam_b200_profile_enable(sloam_ctx *ctx, int on);
int sloam_b200_profile_read(sloam_ctx *ctx, double *split_kernel_ms, int32_t *launches);
/* Same for every kernel group of the fused path: one record per group that ran, `ms` summed
 * over the fused runs since the last enable/read (at most 64 runs are recorded), `launches` =
 * runs counted.  While profiling is enabled a fused run uses one lane and one stream (the
 * ground stage and the tree detector run one after the other), so that every pair times its
 * kernels alone; results are unchanged.  Synchronises and resets like profile_read. */
typedef struct sloam_prof_kernel {
  char name[48];
  double ms;
  int32_t launches;
  int32_t reserved;
} sloam_prof_kernel;
int sloam_b200_profile_read_kernels(sloam_ctx *ctx, sloam_prof_kernel *out, int cap, int32_t *n_out);
const char *sloam_b200_version(void);

/* ------------------------------------------------ stage entries (device) */

/* a1: Segmentation::_doProjection (inference.cpp:80-165).
 * points [K][N] -> pix [K][N] = proj_y*W + proj_x per point in input order
 * (the reference keeps proj_xs/proj_ys, :131-132), range_image [K][H*W]
 * closest-point-wins (:135,:160-162), empty = 0.  N must equal H*W
 * (maskCloud asserts it, :239). range_image may be NULL. */
int sloam_b200_project_dev(sloam_ctx *ctx, int K, const sloam_point *points,
                           int32_t *pix, float *range_image);

/* a2: Segmentation::maskCloud x2 (inference.cpp:230-273) as called by
 * sloamNode.cpp:212,215: ground = points whose mask pixel == 1, compacted in
 * input order; tree = organized cloud with NaN where mask pixel != 255.
 * ground [K][N] (capacity N per keyframe), ground_count [K]. */
int sloam_b200_mask_cloud_dev(sloam_ctx *ctx, int K, const sloam_point *points,
                              const int32_t *pix, const uint8_t *mask,
                              sloam_point *tree, sloam_point *ground,
                              int32_t *ground_count);

/* a1+a2 fused (one pass over the points): the production path. */
int sloam_b200_project_split_dev(sloam_ctx *ctx, int K,
                                 const sloam_point *points,
                                 const uint8_t *mask, int32_t *pix,
                                 float *range_image, sloam_point *tree,
                                 sloam_point *ground, int32_t *ground_count);

/* a3+a4+a5: sloam::binGroundPoints (sloam.cpp:330-386), Plane::Plane
 * (plane.cpp:3-17,96-128) per cell and the acceptance test of computeModels
 * (sloam.cpp:394-412).  ground [K][ground_stride], ground_count [K],
 * pose_est [K].  cells [K][B] in (radius bin, theta bin) order,
 * cell_features [K][B][numGroundFeatures] (Plane::features after the resize,
 * plane.cpp:14).  kept_points/kept_offsets are optional (may be NULL):
 * the retained point lists scgf[r][t] of binGroundPoints, kept_points
 * [K][ground_stride] and kept_offsets [K][B+1]. */
int sloam_b200_ground_planes_dev(sloam_ctx *ctx, int K,
                                 const sloam_point *ground,
                                 const int32_t *ground_count,
                                 int ground_stride, const sloam_pose *pose_est,
                                 sloam_cell_plane *cells,
                                 sloam_point *cell_features,
                                 sloam_point *kept_points,
                                 int32_t *kept_offsets);

/* a6: Instance::findClusters (trellis.cpp:15-29) = PCL organized connected
 * components.  tree [K][H*W] organized -> labels [K][H*W] (0xFFFFFFFF =
 * invalid), n_clusters [K]. */
int sloam_b200_find_clusters_dev(sloam_ctx *ctx, int K,
                                 const sloam_point *tree, uint32_t *labels,
                                 int32_t *n_clusters);

/* a6+a7: Instance::computeGraph (trellis.cpp:134-140).
 * trees [K][max_trees], n_trees [K], vertices [K][max_trees*max_tree_vertices],
 * vertex_points [K][H*W] pool. Offsets are relative to the keyframe's slice. */
int sloam_b200_compute_graph_dev(sloam_ctx *ctx, int K,
                                 const sloam_point *tree, sloam_tree *trees,
                                 int32_t *n_trees, sloam_vertex *vertices,
                                 sloam_point *vertex_points);

/* a8+a9+a10: nearest plane per tree (sloam.cpp:418-431) and Cylinder::Cylinder
 * (cylinder.cpp:3-173).  Consumes the outputs of ground_planes and
 * compute_graph.  models [K][max_trees] (one per tree, valid or not),
 * features [K][max_trees][featuresPerTree] (cylinder.cpp:87-91,172). */
int sloam_b200_cylinders_dev(sloam_ctx *ctx, int K, const sloam_tree *trees,
                             const int32_t *n_trees,
                             const sloam_vertex *vertices,
                             const sloam_point *vertex_points,
                             const sloam_cell_plane *cells,
                             sloam_tree_model *models, sloam_point *features);

/* a11-a13: brute-force nearest map cylinder (matchFeatures / matchModels,
 * sloam.cpp:257-328, Cylinder::distance cylinder.cpp:175-194).
 * det [K][det_stride] are projected with tf [K] (NULL = identity) first
 * (Cylinder::project, cylinder.cpp:205-211).  map [K][map_stride], or one
 * shared map when map_shared != 0.  best_index = first minimum (strict <),
 * -1 when there is no map; best_dist its distance. */
int sloam_b200_associate_dev(sloam_ctx *ctx, int K, const sloam_cylinder *det,
                             const int32_t *n_det, int det_stride,
                             const sloam_pose *tf, const sloam_cylinder *map,
                             const int32_t *n_map, int map_stride,
                             int map_shared, int32_t *best_index,
                             double *best_dist);

/* a12 for ground planes: the argmin of matchFeatures<Plane> (sloam.cpp:257-286) --
 * Plane::project moves the centroid (plane.cpp:174), Plane::distance(model) is the centroid
 * distance (plane.cpp:131-134).  det [K][det_stride] projected with tf [K] (NULL = identity),
 * map [K][map_stride]; best_index = first minimum (strict <), -1 without map planes. */
int sloam_b200_associate_planes_dev(sloam_ctx *ctx, int K, const sloam_plane *det,
                                    const int32_t *n_det, int det_stride, const sloam_pose *tf,
                                    const sloam_plane *map, const int32_t *n_map, int map_stride,
                                    int32_t *best_index, double *best_dist);

/* a14-a17: OptimizePose / TwoStepOptimizePose (sloam.cpp:33-255) on explicit
 * match lists.  tree_feat [K][tf_stride][3] sensor-frame features with their
 * matched cylinder tree_obj [K][tf_stride]; plane_feat/plane_obj likewise.
 * mode 0 = joint (OptimizePose), 1 = two-step; optim_trees/optim_ground [K]
 * are the treeCheck/groundCheck flags (two-step only).  out_pose [K],
 * iterations [K][2], termination [K][2]. */
int sloam_b200_optimize_pose_dev(sloam_ctx *ctx, int K, int mode,
                                 const sloam_pose *pose_est,
                                 const double *tree_feat,
                                 const sloam_cylinder *tree_obj,
                                 const int32_t *n_tree_res, int tf_stride,
                                 const double *plane_feat,
                                 const sloam_plane *plane_obj,
                                 const int32_t *n_plane_res, int pf_stride,
                                 const uint8_t *optim_trees,
                                 const uint8_t *optim_ground,
                                 sloam_pose *out_pose, int32_t *iterations,
                                 int32_t *termination);

/* ------------------------------------------- fused batched path (a1..a19) */

/* Inputs of K independent keyframes.  The state the reference carries
 * across calls (firstScan_, prevGPlanes_, sloam.h:99-106) and the submap the
 * caller supplies (SloamInput::mapModels, sloam.h:44) are explicit. */
typedef struct sloam_batch_in {
  const sloam_point *points;     /* [K][H*W]                               */
  const uint8_t *mask;           /* [K][H*W] 0 other, 1 ground, 255 tree   */
  const sloam_pose *pose_est;    /* [K] SloamInput::poseEstimate           */
  const uint8_t *first_scan;     /* [K] firstScan_                         */
  const sloam_cylinder *map_models; /* [K][max_map_models] or shared       */
  const int32_t *n_map_models;   /* [K] (or [1] when shared)               */
  int32_t map_shared;
  const sloam_plane *prev_planes; /* [K][max_prev_planes] prevGPlanes_ (map frame) */
  const int32_t *n_prev_planes;  /* [K]                                    */
} sloam_batch_in;

typedef struct sloam_batch_out {
  sloam_kf_result *results;      /* [K]                                    */
  int32_t *matches;              /* [K][max_trees] SloamOutput::matches    */
  sloam_cylinder *tm;            /* [K][max_trees] SloamOutput::tm (map frame) */
  int32_t *tm_id;                /* [K][max_trees] Cylinder::id            */
  sloam_plane *planes;           /* [K][max_prev_planes] next prevGPlanes_ */
  int32_t *n_planes;             /* [K]                                    */
  float *range_image;            /* [K][H*W] or NULL                       */
} sloam_batch_out;

/* Segmentation::run (projection part) + maskCloud x2 + computeGraph +
 * RunSloam for K keyframes (sloamNode.cpp:208-236 minus the network). */
int sloam_b200_run_keyframes_dev(sloam_ctx *ctx, int K,
                                 const sloam_batch_in *in,
                                 const sloam_batch_out *out);
/* Same with HOST buffers: H2D of the inputs, the pipeline, D2H of the
 * results, then a stream synchronise.  This is what bench.py's e2e times. */
int sloam_b200_run_keyframes_host(sloam_ctx *ctx, int K,
                                  const sloam_batch_in *in,
                                  const sloam_batch_out *out);

/* Same for a cloud packed as x, y, z (12 bytes per point, [K][H*W][3]); in->points is ignored.
 * pcl::PointXYZI is 32 bytes in host memory, so a binding repacks the cloud anyway, and no
 * output of RunSloam depends on the input intensity (tree features carry the tree id,
 * cylinder.cpp:87-91; the optimiser reads x, y, z): packing only x, y, z moves 25 % fewer
 * bytes over PCIe, which is what bounds this entry.  Intensities read as 0 in the intermediates. */
int sloam_b200_run_keyframes_host_xyz(sloam_ctx *ctx, int K, const float *points_xyz,
                                      const sloam_batch_in *in, const sloam_batch_out *out);

/* sloam::RunSloam (sloam.cpp:453-532) alone, for callers that already hold the
 * SloamInput of the reference: ground clouds [K][ground_stride] with counts,
 * landmarks as produced by compute_graph (trees [K][max_trees], vertices
 * [K][vertex_stride], vertex_points [K][point_stride]); in->points / in->mask
 * are ignored.  Device pointers. */
int sloam_b200_run_sloam_dev(sloam_ctx *ctx, int K, const sloam_point *ground,
                             const int32_t *ground_count, int ground_stride,
                             const sloam_tree *trees, const int32_t *n_trees,
                             const sloam_vertex *vertices, int vertex_stride,
                             const sloam_point *vertex_points, int point_stride,
                             const sloam_batch_in *in,
                             const sloam_batch_out *out);

/* ------------------------------------------------------------------ multi-GPU
 * SURVEY 8(e).  Keyframes are sharded over ranks (one process and one context per GPU);
 * nothing is exchanged during compute.  The one exchange is the gather of the per-keyframe
 * SloamOutput records (sloam/include/core/sloam.h:48-55) of every rank, ncclAllGather over
 * NVLink on a side stream of the context.  The reference has no counterpart (one process,
 * one keyframe at a time, sloamNode.cpp:186).
 *   comm_unique_id  rank 0 fills 128 bytes (ncclUniqueId) and hands them to the other ranks by
 *                   any means (torch.distributed, MPI, a file)
 *   comm_init       collective: every rank calls it with the same id
 *   gather_results  all-gather of results [K] (and of matches / tm / tm_id [K][max_trees] when
 *                   both `local` and `all` hold a pointer for them): all->x is [world][K]...,
 *                   rank-major.  Asynchronous: it starts when the work queued on the context
 *                   stream so far is done and overlaps whatever is queued next; a following
 *                   run_keyframes_* only makes its OUTPUT kernels wait for it, so the `local`
 *                   buffers may be the output buffers of the next run.
 *   comm_wait       makes the context stream wait for the last gather (call it before reading
 *                   `all` through the context stream or sloam_b200_sync)
 * SLOAM_E_NODEVICE when libnccl.so.2 cannot be loaded. */
#define SLOAM_COMM_ID_BYTES 128
int sloam_b200_comm_unique_id(void *id128);
int sloam_b200_comm_init(sloam_ctx *ctx, int rank, int world, const void *id128);
int sloam_b200_comm_destroy(sloam_ctx *ctx);
int sloam_b200_comm_size(const sloam_ctx *ctx);
int sloam_b200_comm_rank(const sloam_ctx *ctx);
int sloam_b200_gather_results_dev(sloam_ctx *ctx, int K, const struct sloam_batch_out *local,
                                  const struct sloam_batch_out *all);
int sloam_b200_comm_wait(sloam_ctx *ctx);

/* ---------------------------------------------- semantic map + sequential mode
 * SURVEY 8(f)-1/2.  MapManager (sloam/src/core/mapManager.cpp:8-71) and the
 * state-carrying call sequence of SLOAMNode::run (sloamNode.cpp:186-282), both
 * device-resident.  map_init allocates the map (capacity landmarks) and resets
 * firstScan_ / prevGPlanes_; the context must have been created with
 * max_keyframes >= 1. */
int sloam_b200_map_init(sloam_ctx *ctx, int capacity);
void sloam_b200_map_free(sloam_ctx *ctx);
/* MapManager::getSubmap (:41-71): kNN(100) around (pose.x, pose.y, 1) over the landmark
 * roots, then the "last 200 landmarks" filter.  pose, submap [max_map_models], n_submap:
 * device pointers.  The submap-index -> map-index table is kept for update. */
int sloam_b200_map_get_submap_dev(sloam_ctx *ctx, const sloam_pose *pose,
                                  sloam_cylinder *submap, int32_t *n_submap);
/* MapManager::updateMap (:8-28) with the outputs of run_keyframes (K = 1): device pointers.
 * res->status receives SLOAM_KF_FLAG_MAP_CAPACITY when a new landmark did not fit. */
int sloam_b200_map_update_dev(sloam_ctx *ctx, sloam_kf_result *res,
                              const sloam_cylinder *tm, const int32_t *tm_id,
                              const int32_t *matches);
/* Whole map to the host (treeModels_, treeHits_); returns the map size (getMap (:30-39) is
 * the subset with hits > 2).  Synchronises. */
int sloam_b200_map_dump_host(sloam_ctx *ctx, sloam_cylinder *models, int32_t *hits, int cap);
/* SLOAMNode::run for one keyframe: HOST points [H*W], mask [H*W], poseEstimate
 * (= prevKeyPose * initialGuess, sloamNode.cpp:192); getSubmap -> projection/split ->
 * computeGraph -> RunSloam -> updateMap on the device; result (and optionally matches /
 * tm / tm_id, [max_trees]) back on the host.  result->success is run()'s return value. */
int sloam_b200_sequence_step_host(sloam_ctx *ctx, const sloam_point *points,
                                  const uint8_t *mask, const sloam_pose *pose_estimate,
                                  sloam_kf_result *result, int32_t *matches,
                                  sloam_cylinder *tm, int32_t *tm_id);

/* ------------------------------------------------ segmentation network hook
 * SURVEY 8(f)-4.  The two elementwise passes either side of the (external) RangeNet++
 * engine.  make_tensor = Segmentation::_makeTensor (inference.cpp:167-198) for the
 * one-channel configuration: tensor [K][H*W] = (range - mean) / std for valid pixels,
 * the raw value for invalid ones (|range| < 1: the reference tests the value converted
 * to int, :183); invalid [K][H*W] flags replace the reference's index list, n_invalid [K]
 * (optional) counts them.  The reference hard-codes mean 12.97, std 12.35 (:169-170).
 * mask_from_logits = Segmentation::_mask (:275-300): logits [K][3][H*W] channel-major,
 * strict-'<' argmax, class 2 -> 255, invalid (optional) -> 0.  Device pointers. */
int sloam_b200_make_tensor_dev(sloam_ctx *ctx, int K, const float *range_image, float mean,
                               float stdv, float *tensor, uint8_t *invalid, int32_t *n_invalid);
int sloam_b200_mask_from_logits_dev(sloam_ctx *ctx, int K, const float *logits,
                                    const uint8_t *invalid, uint8_t *mask);

/* Device-memory helpers so that a host-language binding (cgo / JNI / the C++
 * classes in sloam_b200/host) needs no CUDA headers.  copy_d2h synchronises
 * the context stream before returning. */
void *sloam_b200_dev_alloc(sloam_ctx *ctx, uint64_t bytes);
void sloam_b200_dev_free(sloam_ctx *ctx, void *ptr);
int sloam_b200_copy_h2d(sloam_ctx *ctx, void *dst_dev, const void *src_host, uint64_t bytes);
int sloam_b200_copy_d2h(sloam_ctx *ctx, void *dst_host, const void *src_dev, uint64_t bytes);

/* Intermediate device buffers of the last run_keyframes call (for tests and
 * for callers that want the landmarks): pointers into context scratch. */
typedef struct sloam_intermediates {
  const int32_t *pix;                 /* [K][N]                    */
  const sloam_point *tree;            /* [K][N]                    */
  const sloam_point *ground;          /* [K][N]                    */
  const int32_t *ground_count;        /* [K]                       */
  const sloam_cell_plane *cells;      /* [K][B]                    */
  const sloam_point *cell_features;   /* [K][B][F_g]               */
  const sloam_tree *trees;            /* [K][max_trees]            */
  const int32_t *n_trees;             /* [K]                       */
  const sloam_vertex *vertices;       /* [K][max_trees*max_tree_vertices] */
  const sloam_point *vertex_points;   /* [K][N]                    */
  const sloam_tree_model *tree_models; /* [K][max_trees]           */
  const sloam_point *tree_features;   /* [K][max_trees][F_t]       */
} sloam_intermediates;
int sloam_b200_get_intermediates(sloam_ctx *ctx, sloam_intermediates *out);

/* ------------------------------------------------- synthetic forest scans */

/* Scene + sensor description of the synthetic generator.
 * Test/bench infrastructure shipped with the library so that the GPU box can
 * generate batches on the device. */
typedef struct sloam_synth_config {
  int32_t img_h, img_w;
  float fov_up_deg, fov_down_deg;
  int32_t n_trees;
  float tree_r_min, tree_r_max;      /* annulus of trunk centres [m]       */
  float trunk_radius_min, trunk_radius_max;
  float max_tilt_deg;                /* trunk tilt                          */
  float sensor_height;               /* above the ground plane             */
  float ground_slope_deg;
  float ground_noise, range_noise;   /* sigma [m]                           */
  float max_range;                   /* beams beyond this: no return        */
  float step_per_keyframe;           /* metres of forward motion            */
  float azimuth_offset_cols;         /* sensor column vs projected column   */
  float guess_sigma_t, guess_sigma_r;/* pose-guess perturbation [m], [rad]  */
  int32_t nan_no_return;             /* 1: NaN xyz, 0: zeros                */
  uint64_t seed;
} sloam_synth_config;

void sloam_synth_default_config(sloam_synth_config *c, int img_h, int img_w,
                                int n_trees);
/* Ground-truth scene of a sequence: tree axes in the map frame.  Host-side.
 * trees_out [n_trees], returns the number written. */
int sloam_synth_scene(const sloam_synth_config *c, sloam_cylinder *trees_out);
/* Ground-truth pose and perturbed pose guess of keyframe k (host). */
void sloam_synth_pose(const sloam_synth_config *c, int64_t keyframe,
                      sloam_pose *gt, sloam_pose *guess);
/* Generate keyframes [k0, k0+K): points [K][H*W] and mask [K][H*W].
 * _host writes host buffers (single thread per call); _dev launches a kernel
 * on the context stream and writes device buffers. */
int sloam_synth_generate_host(const sloam_synth_config *c, int64_t k0, int K,
                              sloam_point *points, uint8_t *mask);
int sloam_synth_generate_dev(sloam_ctx *ctx, const sloam_synth_config *c,
                             int64_t k0, int K, sloam_point *points,
                             uint8_t *mask);

#ifdef __cplusplus
}
#endif
#endif /* SLOAM_B200_H */
