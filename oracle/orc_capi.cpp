/* oracle/orc_capi.cpp -- flat C entry points of the CPU oracle for ctypes
 * (test infrastructure; see orc.h).  The flattened records are the ones of
 * include/sloam_b200.h so tests compare arrays field by field. */
#include <chrono>
#include <cstring>

#include "orc.h"

using namespace orc;

namespace {
Options make_opts(const sloam_params *p, int use_libm) {
  Options o;
  o.p = *p;
  o.use_libm = use_libm != 0;
  return o;
}
const Pt *as_pt(const sloam_point *p) { return reinterpret_cast<const Pt *>(p); }
sloam_point to_abi(const Pt &p) { return {p.x, p.y, p.z, p.intensity}; }

void flatten_landmarks(const Landmarks &lm, int max_trees, sloam_tree *trees, int32_t *n_trees,
                       sloam_vertex *vertices, sloam_point *vpoints) {
  int nt = 0, nv = 0, np = 0;
  for (const auto &tree : lm) {
    if (nt >= max_trees) break;
    sloam_tree t{};
    t.tree_id = tree.empty() ? -1 : tree[0].treeId;
    t.n_vertices = (int)tree.size();
    t.vertex_begin = nv;
    int tp = 0;
    for (const TreeVertex &v : tree) {
      sloam_vertex fv{};
      fv.cx = v.coords.x; fv.cy = v.coords.y; fv.cz = v.coords.z;
      fv.radius = (float)v.radius;
      fv.n_points = (int)v.points.size();
      fv.point_begin = np;
      fv.row = v.row;
      fv.is_valid = v.isValid ? 1 : 0;
      vertices[nv++] = fv;
      for (const Pt &p : v.points) vpoints[np++] = to_abi(p);
      tp += (int)v.points.size();
    }
    t.n_points = tp;
    trees[nt++] = t;
  }
  *n_trees = nt;
}

Landmarks unflatten_landmarks(const sloam_tree *trees, int n_trees, const sloam_vertex *vertices,
                              const sloam_point *vpoints) {
  Landmarks lm;
  for (int i = 0; i < n_trees; ++i) {
    std::vector<TreeVertex> tree;
    for (int k = 0; k < trees[i].n_vertices; ++k) {
      const sloam_vertex &fv = vertices[trees[i].vertex_begin + k];
      TreeVertex v;
      v.treeId = trees[i].tree_id;
      v.radius = fv.radius;
      v.isValid = fv.is_valid != 0;
      v.coords.x = fv.cx; v.coords.y = fv.cy; v.coords.z = fv.cz;
      v.row = fv.row;
      for (int j = 0; j < fv.n_points; ++j) {
        const sloam_point &p = vpoints[fv.point_begin + j];
        Pt q; q.x = p.x; q.y = p.y; q.z = p.z; q.intensity = p.intensity;
        v.points.push_back(q);
      }
      tree.push_back(v);
    }
    lm.push_back(tree);
  }
  return lm;
}

sloam_cell_plane cell_to_abi(const Plane &pl, bool accepted) {
  sloam_cell_plane c{};
  for (int i = 0; i < 4; ++i) c.model.plane[i] = pl.model.plane[i];
  c.model.centroid[0] = pl.model.centroid.x;
  c.model.centroid[1] = pl.model.centroid.y;
  c.model.centroid[2] = pl.model.centroid.z;
  c.n_cell = pl.n_cell;
  c.n_kept = pl.n_kept;
  c.is_valid = pl.isValid ? 1 : 0;
  c.accepted = accepted ? 1 : 0;
  return c;
}
Plane plane_from_abi(const sloam_plane &m) {
  Plane p;
  p.isValid = true;
  for (int i = 0; i < 4; ++i) p.model.plane[i] = m.plane[i];
  p.model.centroid = {m.centroid[0], m.centroid[1], m.centroid[2]};
  return p;
}
sloam_plane plane_to_abi(const Plane &p) {
  sloam_plane m;
  for (int i = 0; i < 4; ++i) m.plane[i] = p.model.plane[i];
  m.centroid[0] = p.model.centroid.x; m.centroid[1] = p.model.centroid.y; m.centroid[2] = p.model.centroid.z;
  return m;
}
Cylinder cyl_from_abi(const sloam_cylinder &m) {
  Cylinder c;
  c.isValid = true;
  c.model.root = {m.root[0], m.root[1], m.root[2]};
  c.model.ray = {m.ray[0], m.ray[1], m.ray[2]};
  c.model.radius = m.radius;
  return c;
}
sloam_cylinder cyl_to_abi(const CylinderParameters &m) {
  sloam_cylinder c;
  c.root[0] = m.root.x; c.root[1] = m.root.y; c.root[2] = m.root.z;
  c.ray[0] = m.ray.x; c.ray[1] = m.ray.y; c.ray[2] = m.ray.z;
  c.radius = m.radius;
  return c;
}
sloam_tree_model tree_model_to_abi(const Cylinder &c) {
  sloam_tree_model m{};
  m.model = cyl_to_abi(c.model);
  m.id = (int32_t)c.id;
  m.is_valid = c.isValid ? 1 : 0;
  m.plane_index = c.plane_index;
  m.n_inliers = c.n_inliers;
  m.best_hypothesis = c.best_hypothesis;
  m.n_hypotheses = c.n_hypotheses;
  m.n_refit_inliers = c.n_refit_inliers;
  return m;
}
}  // namespace

static thread_local std::vector<TreeMatch> g_last_tm;
static thread_local std::vector<PlaneMatch> g_last_gm;

extern "C" {

/* match lists of the last orc_run_keyframe call on this thread (debugging / LM tests) */
int orc_last_matches(double *tree_feat, sloam_cylinder *tree_obj, int tcap, double *plane_feat,
                     sloam_plane *plane_obj, int pcap, int32_t *n_plane) {
  const int nt = std::min<int>((int)g_last_tm.size(), tcap), np = std::min<int>((int)g_last_gm.size(), pcap);
  for (int i = 0; i < nt; ++i) {
    tree_feat[3 * i] = g_last_tm[i].feature.x; tree_feat[3 * i + 1] = g_last_tm[i].feature.y; tree_feat[3 * i + 2] = g_last_tm[i].feature.z;
    Cylinder c; c.model = g_last_tm[i].object; tree_obj[i] = cyl_to_abi(c.model);
  }
  for (int i = 0; i < np; ++i) {
    plane_feat[3 * i] = g_last_gm[i].feature.x; plane_feat[3 * i + 1] = g_last_gm[i].feature.y; plane_feat[3 * i + 2] = g_last_gm[i].feature.z;
    Plane pl; pl.model = g_last_gm[i].object; plane_obj[i] = plane_to_abi(pl);
  }
  *n_plane = np;
  return nt;
}

void orc_default_params(sloam_params *p) {
  std::memset(p, 0, sizeof *p);
  p->img_h = 64; p->img_w = 1024; p->fov_up_deg = 22.5f; p->fov_down_deg = -22.5f;
  p->do_destagger = 0;
  /* sloam/params/sloam.yaml over the code defaults of sloamNode.cpp:57-128 */
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
  p->max_dist_to_centroid = 0.2f; p->cluster_dist_thresh = 1.0f;
  p->min_cluster_points = 80; p->min_vertex_points = 3;
  p->min_tree_vertices = 16; p->max_tree_vertices = 56;
  p->ransac_threshold = 0.25; p->ransac_max_iterations = 50; p->ransac_probability = 0.99;
  p->ransac_fixed_hypotheses = 0;
  p->min_tree_height_sq = 1.5; p->root_plane_max_dist = 2.0;
  p->plane_match_thresh = 1.0; p->ground_angle_tol = 0.1; p->huber_delta = 0.1;
  p->lm_max_iterations = 50;
  p->max_trees = 512; p->max_map_models = 512; p->max_prev_planes = 64;
}

void orc_project(const sloam_params *p, int use_libm, const sloam_point *pts, int n, int32_t *pix,
                 float *range_image) {
  project(make_opts(p, use_libm), as_pt(pts), n, pix, range_image);
}

void orc_mask_cloud(const sloam_params *p, const sloam_point *pts, int n, const int32_t *pix,
                    const uint8_t *mask, sloam_point *tree, sloam_point *ground, int32_t *n_ground) {
  Cloud t, g;
  mask_cloud(make_opts(p, 0), as_pt(pts), n, pix, mask, t, g);
  for (int i = 0; i < n; ++i) tree[i] = to_abi(t[i]);
  for (size_t i = 0; i < g.size(); ++i) ground[i] = to_abi(g[i]);
  *n_ground = (int32_t)g.size();
}

/* binGroundPoints + Plane per cell + acceptance: the a3+a4+a5 stage entry */
void orc_ground_planes(const sloam_params *p, int use_libm, const sloam_point *ground, int n,
                       const sloam_pose *pose_est, sloam_cell_plane *cells,
                       sloam_point *cell_features, sloam_point *kept_points,
                       int32_t *kept_offsets) {
  const Options o = make_opts(p, use_libm);
  std::vector<Cloud> cl;
  std::vector<int> n_cell;
  bin_ground_points(o, V3{0, 0, 0}, as_pt(ground), n, cl, n_cell);
  const SE3 pose = pose_from_abi(*pose_est);
  const int B = p->groundRadiiBins * p->groundThetaBins, Fg = p->numGroundFeatures;
  int off = 0;
  for (int c = 0; c < B; ++c) {
    Plane pl = make_plane(cl[c], Fg);
    pl.n_cell = n_cell[c];
    const bool ok = pl.isValid && plane_accept(o, pose, pl);
    cells[c] = cell_to_abi(pl, ok);
    for (int f = 0; f < Fg; ++f)
      cell_features[(size_t)c * Fg + f] =
          (pl.isValid && f < (int)pl.features.size()) ? to_abi(pl.features[f]) : sloam_point{0, 0, 0, 0};
    if (kept_offsets) kept_offsets[c] = off;
    if (kept_points)
      for (const Pt &q : cl[c]) kept_points[off++] = to_abi(q);
    else
      off += (int)cl[c].size();
  }
  if (kept_offsets) kept_offsets[B] = off;
}

/* Plane(points, params) on an explicit point list (plane_test.cpp) */
void orc_plane_fit(const sloam_point *pts, int n, int numGroundFeatures, sloam_cell_plane *out) {
  Cloud c(n);
  for (int i = 0; i < n; ++i) { c[i].x = pts[i].x; c[i].y = pts[i].y; c[i].z = pts[i].z; c[i].intensity = pts[i].intensity; }
  Plane pl = make_plane(c, numGroundFeatures);
  pl.n_cell = n;
  *out = cell_to_abi(pl, pl.isValid);
}

void orc_plane_project(sloam_plane *m, const sloam_pose *tf) {
  Plane p = plane_from_abi(*m);
  plane_project(p, pose_from_abi(*tf));
  *m = plane_to_abi(p);
}
double orc_plane_distance_point(const sloam_plane *m, const sloam_point *pt) {
  Pt q; q.x = pt->x; q.y = pt->y; q.z = pt->z;
  return plane_distance_point(plane_from_abi(*m).model, q);
}

void orc_find_clusters(const sloam_params *p, const sloam_point *tree, uint32_t *labels,
                       int32_t *n_clusters) {
  std::vector<uint32_t> lab;
  std::vector<std::vector<int>> idx;
  find_clusters(make_opts(p, 0), as_pt(tree), p->img_h, p->img_w, lab, idx);
  std::memcpy(labels, lab.data(), lab.size() * sizeof(uint32_t));
  *n_clusters = idx.empty() ? 0 : (int32_t)idx.size() - 1;
}

void orc_compute_graph(const sloam_params *p, const sloam_point *tree, sloam_tree *trees,
                       int32_t *n_trees, sloam_vertex *vertices, sloam_point *vertex_points) {
  Landmarks lm;
  compute_graph(make_opts(p, 0), as_pt(tree), p->img_h, p->img_w, lm);
  flatten_landmarks(lm, p->max_trees, trees, n_trees, vertices, vertex_points);
}

/* nearest accepted plane per tree + Cylinder(): a8+a9+a10 */
void orc_cylinders(const sloam_params *p, const sloam_tree *trees, int n_trees,
                   const sloam_vertex *vertices, const sloam_point *vertex_points,
                   const sloam_cell_plane *cells, sloam_tree_model *models, sloam_point *features) {
  const Options o = make_opts(p, 0);
  const Landmarks lm = unflatten_landmarks(trees, n_trees, vertices, vertex_points);
  const int B = p->groundRadiiBins * p->groundThetaBins, Ft = p->featuresPerTree;
  std::vector<Plane> planes;
  for (int c = 0; c < B; ++c)
    if (cells[c].accepted) planes.push_back(plane_from_abi(cells[c].model));
  for (int i = 0; i < n_trees; ++i) {
    sloam_tree_model m{};
    m.plane_index = -1;
    m.best_hypothesis = -1;
    for (int f = 0; f < Ft; ++f) features[(size_t)i * Ft + f] = sloam_point{0, 0, 0, 0};
    if (!planes.empty()) {
      const Pt pos = lm[i][1].coords;
      double best = 100000; int bi = 0;
      for (size_t g = 0; g < planes.size(); ++g) {
        const double d = plane_distance_point(planes[g].model, pos);
        if (d < best) { best = d; bi = (int)g; }
      }
      Cylinder c = make_cylinder(o, lm[i], planes[bi]);
      c.plane_index = bi;
      m = tree_model_to_abi(c);
      if (c.isValid)
        for (int f = 0; f < Ft && f < (int)c.features.size(); ++f)
          features[(size_t)i * Ft + f] = to_abi(c.features[f]);
    }
    models[i] = m;
  }
}

void orc_ransac_draw_table(int n, int n_draws, int32_t *pairs) {
  std::vector<int32_t> v;
  ransac_draw_table(n, n_draws, v);
  std::memcpy(pairs, v.data(), v.size() * sizeof(int32_t));
}

double orc_cylinder_distance_model(const sloam_cylinder *a, const sloam_cylinder *b) {
  return cylinder_distance_model(cyl_from_abi(*a).model, cyl_from_abi(*b).model);
}
double orc_cylinder_distance_point(const sloam_cylinder *a, const sloam_point *pt) {
  Pt q; q.x = pt->x; q.y = pt->y; q.z = pt->z;
  return cylinder_distance_point(cyl_from_abi(*a).model, q);
}
void orc_cylinder_project(sloam_cylinder *m, const sloam_pose *tf) {
  Cylinder c = cyl_from_abi(*m);
  cylinder_project(c, pose_from_abi(*tf));
  *m = cyl_to_abi(c.model);
}

/* matchFeatures/matchModels argmin: a11-a13 */
void orc_associate(const sloam_cylinder *det, int n_det, const sloam_pose *tf,
                   const sloam_cylinder *map, int n_map, int32_t *best_index, double *best_dist) {
  std::vector<CylinderParameters> mp(n_map);
  for (int k = 0; k < n_map; ++k) mp[k] = cyl_from_abi(map[k]).model;
  for (int i = 0; i < n_det; ++i) {
    Cylinder c = cyl_from_abi(det[i]);
    if (tf) cylinder_project(c, pose_from_abi(*tf));
    double bd = INFINITY; int bi = -1;
    for (int k = 0; k < n_map; ++k) {
      const double d = cylinder_distance_model(mp[k], c.model);
      if (d < bd) { bd = d; bi = k; }
    }
    best_index[i] = bi;
    best_dist[i] = bd;
  }
}

/* OptimizePose / TwoStepOptimizePose on explicit match lists: a14-a17 */
void orc_optimize_pose(const sloam_params *p, int mode, const sloam_pose *pose_est,
                       const double *tree_feat, const sloam_cylinder *tree_obj, int n_tree,
                       const double *plane_feat, const sloam_plane *plane_obj, int n_plane,
                       int optim_trees, int optim_ground, sloam_pose *out_pose,
                       int32_t *iterations, int32_t *termination) {
  const Options o = make_opts(p, 0);
  std::vector<TreeMatch> tm(n_tree);
  std::vector<PlaneMatch> gm(n_plane);
  for (int i = 0; i < n_tree; ++i) {
    tm[i].feature = {tree_feat[3 * i], tree_feat[3 * i + 1], tree_feat[3 * i + 2]};
    tm[i].object = cyl_from_abi(tree_obj[i]).model;
  }
  for (int i = 0; i < n_plane; ++i) {
    gm[i].feature = {plane_feat[3 * i], plane_feat[3 * i + 1], plane_feat[3 * i + 2]};
    gm[i].object = plane_from_abi(plane_obj[i]).model;
  }
  const SE3 est = pose_from_abi(*pose_est);
  SE3 tf = est;
  LMSummary s[2];
  if (mode == 0) {
    SE3 T_Delta; /* sloam.cpp:497: starts as identity, stays so on failure */
    const bool ok = optimize_pose(o, est, tm, gm, T_Delta, &s[0]);
    tf = ok ? T_Delta : est;
  } else {
    two_step_optimize_pose(o, est, optim_trees != 0, optim_ground != 0, tm, gm, tf, s);
  }
  *out_pose = pose_to_abi(tf);
  for (int i = 0; i < 2; ++i) { iterations[i] = s[i].iterations; termination[i] = s[i].termination; }
}

/* One full keyframe: projection + split + computeGraph + RunSloam
 * (sloamNode.cpp:208-236 minus the network).  Optional outputs may be NULL. */
void orc_run_keyframe(const sloam_params *p, int use_libm, const sloam_point *points,
                      const uint8_t *mask, const sloam_pose *pose_est, int first_scan,
                      const sloam_cylinder *map_models, int n_map, const sloam_plane *prev_planes,
                      int n_prev, sloam_kf_result *result, int32_t *matches, sloam_cylinder *tm,
                      int32_t *tm_id, sloam_plane *planes_out, int32_t *n_planes_out,
                      /* optional intermediates */
                      int32_t *pix_out, float *range_image, sloam_cell_plane *cells_out,
                      sloam_tree *trees_out, int32_t *n_trees_out, sloam_vertex *vertices_out,
                      sloam_point *vpoints_out, sloam_tree_model *models_out) {
  const Options o = make_opts(p, use_libm);
  const int N = p->img_h * p->img_w;
  std::vector<int32_t> pix(N);
  project(o, as_pt(points), N, pix.data(), range_image);
  if (pix_out) std::memcpy(pix_out, pix.data(), sizeof(int32_t) * N);
  SloamInput in;
  Cloud tree;
  mask_cloud(o, as_pt(points), N, pix.data(), mask, tree, in.groundCloud);
  compute_graph(o, tree.data(), p->img_h, p->img_w, in.landmarks);
  if (trees_out) flatten_landmarks(in.landmarks, p->max_trees, trees_out, n_trees_out, vertices_out, vpoints_out);
  in.poseEstimate = pose_from_abi(*pose_est);
  for (int k = 0; k < n_map; ++k) in.mapModels.push_back(cyl_from_abi(map_models[k]));
  Sloam s(o);
  s.firstScan = first_scan != 0;
  for (int k = 0; k < n_prev; ++k) s.prevGPlanes.push_back(plane_from_abi(prev_planes[k]));
  SloamOutput out;
  s.RunSloam(in, out);
  g_last_tm = s.lastTreeMatches;
  g_last_gm = s.lastPlaneMatches;
  *result = s.last;
  const int T = std::min((int)out.tm.size(), p->max_trees);
  for (int i = 0; i < T; ++i) {
    matches[i] = out.matches[i];
    tm[i] = cyl_to_abi(out.tm[i].model);
    tm_id[i] = (int32_t)out.tm[i].id;
  }
  /* prevGPlanes_ after the call (unchanged input when RunSloam bailed out early) */
  int np = 0;
  for (const Plane &pl : s.prevGPlanes) if (np < p->max_prev_planes) planes_out[np++] = plane_to_abi(pl);
  *n_planes_out = np;
  if (cells_out)
    for (size_t c = 0; c < s.cellPlanes.size(); ++c) cells_out[c] = cell_to_abi(s.cellPlanes[c], s.cellAccepted[c]);
  if (models_out)
    for (size_t i = 0; i < s.treeModels.size() && (int)i < p->max_trees; ++i)
      models_out[i] = tree_model_to_abi(s.treeModels[i]);
}

/* RunSloam on explicit ground cloud + landmarks (core_test.cpp restatement) */
void orc_run_sloam(const sloam_params *p, int use_libm, const sloam_point *ground, int n_ground,
                   const sloam_tree *trees, int n_trees, const sloam_vertex *vertices,
                   const sloam_point *vpoints, const sloam_pose *pose_est, int first_scan,
                   const sloam_cylinder *map_models, int n_map, const sloam_plane *prev_planes,
                   int n_prev, sloam_kf_result *result, int32_t *matches, sloam_cylinder *tm,
                   int32_t *tm_id, sloam_plane *planes_out, int32_t *n_planes_out,
                   sloam_tree_model *models_out) {
  const Options o = make_opts(p, use_libm);
  SloamInput in;
  in.groundCloud.resize(n_ground);
  for (int i = 0; i < n_ground; ++i) {
    in.groundCloud[i].x = ground[i].x; in.groundCloud[i].y = ground[i].y;
    in.groundCloud[i].z = ground[i].z; in.groundCloud[i].intensity = ground[i].intensity;
  }
  in.landmarks = unflatten_landmarks(trees, n_trees, vertices, vpoints);
  in.poseEstimate = pose_from_abi(*pose_est);
  for (int k = 0; k < n_map; ++k) in.mapModels.push_back(cyl_from_abi(map_models[k]));
  Sloam s(o);
  s.firstScan = first_scan != 0;
  for (int k = 0; k < n_prev; ++k) s.prevGPlanes.push_back(plane_from_abi(prev_planes[k]));
  SloamOutput out;
  s.RunSloam(in, out);
  *result = s.last;
  const int T = std::min((int)out.tm.size(), p->max_trees);
  for (int i = 0; i < T; ++i) {
    matches[i] = out.matches[i];
    tm[i] = cyl_to_abi(out.tm[i].model);
    tm_id[i] = (int32_t)out.tm[i].id;
  }
  int np = 0;
  for (const Plane &pl : s.prevGPlanes) if (np < p->max_prev_planes) planes_out[np++] = plane_to_abi(pl);
  *n_planes_out = np;
  if (models_out)
    for (size_t i = 0; i < s.treeModels.size() && (int)i < p->max_trees; ++i)
      models_out[i] = tree_model_to_abi(s.treeModels[i]);
}

/* Timed loop for bench.py's cpu_baseline: runs keyframes [0,K) single-threaded,
 * returns seconds (steady_clock). */
double orc_time_keyframes(const sloam_params *p, int use_libm, int K, const sloam_point *points,
                          const uint8_t *mask, const sloam_pose *pose_est,
                          const uint8_t *first_scan, const sloam_cylinder *map_models,
                          const int32_t *n_map, int map_stride, const sloam_plane *prev_planes,
                          const int32_t *n_prev, int prev_stride, sloam_kf_result *results) {
  const int N = p->img_h * p->img_w;
  std::vector<int32_t> matches(p->max_trees), tm_id(p->max_trees);
  std::vector<sloam_cylinder> tm(p->max_trees);
  std::vector<sloam_plane> planes(p->max_prev_planes);
  int32_t npl = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (int k = 0; k < K; ++k) {
    orc_run_keyframe(p, use_libm, points + (size_t)k * N, mask + (size_t)k * N, pose_est + k,
                     first_scan[k], map_models + (size_t)k * map_stride, n_map[k],
                     prev_planes + (size_t)k * prev_stride, n_prev[k], results + k, matches.data(),
                     tm.data(), tm_id.data(), planes.data(), &npl, nullptr, nullptr, nullptr,
                     nullptr, nullptr, nullptr, nullptr, nullptr);
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

int orc_abi_sizes(int32_t *out) {
  out[0] = sizeof(sloam_params); out[1] = sizeof(sloam_kf_result); out[2] = sizeof(sloam_cell_plane);
  out[3] = sizeof(sloam_tree_model); out[4] = sizeof(sloam_vertex); out[5] = sizeof(sloam_tree);
  return 6;
}

}  // extern "C"
