// sloam_host.h -- host-side C++ mirror of the reference's core API for the hot path,
// marshalling to the C ABI of include/sloam_b200.h.
//
// Same class / member names and call order as the reference so that it drops into
// SLOAMNode::run (sloam/src/core/sloamNode.cpp:186-282):
//   seg::Segmentation::run / maskCloud      sloam/include/segmentation/inference.h:57-67
//   Instance::computeGraph / set_params     sloam/include/segmentation/trellis.h:42-65
//   Plane, Cylinder (SemanticObject<T>)     sloam/include/objects/{plane,cylinder,semanticObject}.h
//   sloam::sloam::RunSloam, setFmParams     sloam/include/core/sloam.h:57-107
//   SloamInput / SloamOutput                sloam/include/core/sloam.h:32-55
//   FeatureModelParams, TreeVertex          sloam/include/helpers/definitions.h:56-105
// PCL / Eigen / Sophus / OpenCV are not available in this image, so minimal look-alike types
// (PointT, CloudT, SE3, Vector3, Vector4, Mask) stand in; a ROS build swaps the real headers
// back in and keeps the marshalling.  All arithmetic of the path runs on the GPU: the methods
// that remain host code are accessors and the three-line model transforms / distances the
// reference's tests call on single objects.  Header-only; needs only libsloam_b200.so.
#pragma once

#include <cmath>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sloam_b200.h"

using Scalar = double;

struct PointT {  // pcl::PointXYZI stand-in, layout of sloam_point
  float x = 0.f, y = 0.f, z = 0.f, intensity = 0.f;
};
static_assert(sizeof(PointT) == sizeof(sloam_point), "PointT must match sloam_point");
using VectorType = std::vector<PointT>;
using Slash = VectorType;

struct CloudT {  // pcl::PointCloud<PointT> stand-in
  using Ptr = std::shared_ptr<CloudT>;
  VectorType points;
  uint32_t width = 0, height = 0;
  bool is_dense = false;
  size_t size() const { return points.size(); }
};
using Cloud = CloudT;

struct Vector3 {
  double v[3] = {0, 0, 0};
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double norm() const { return std::sqrt(v[0] * v[0] + (v[1] * v[1] + v[2] * v[2])); }
};
struct Vector4 {
  double v[4] = {0, 0, 0, 0};
  double &operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  double &operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};

class SE3 {  // Sophus::SE3d stand-in: unit quaternion (x,y,z,w) + translation
 public:
  SE3() { p_.q[3] = 1.0; }
  explicit SE3(const sloam_pose &p) : p_(p) {}
  Vector3 &translation() { return *reinterpret_cast<Vector3 *>(p_.t); }
  const Vector3 &translation() const { return *reinterpret_cast<const Vector3 *>(p_.t); }
  const double *unit_quaternion() const { return p_.q; }  // x y z w
  void setQuaternion(double w, double x, double y, double z) {
    const double n = std::sqrt(w * w + x * x + y * y + z * z);
    p_.q[0] = x / n; p_.q[1] = y / n; p_.q[2] = z / n; p_.q[3] = w / n;
  }
  Vector3 operator*(const Vector3 &p) const {
    const double *q = p_.q;
    double uv[3] = {q[1] * p[2] - q[2] * p[1], q[2] * p[0] - q[0] * p[2], q[0] * p[1] - q[1] * p[0]};
    for (double &u : uv) u += u;
    Vector3 r;
    r[0] = p[0] + q[3] * uv[0] + (q[1] * uv[2] - q[2] * uv[1]) + p_.t[0];
    r[1] = p[1] + q[3] * uv[1] + (q[2] * uv[0] - q[0] * uv[2]) + p_.t[1];
    r[2] = p[2] + q[3] * uv[2] + (q[0] * uv[1] - q[1] * uv[0]) + p_.t[2];
    return r;
  }
  const sloam_pose &abi() const { return p_; }

 private:
  sloam_pose p_{};
};

struct TreeVertex {  // definitions.h:56-66
  int treeId = 0;
  int beam = 0;
  int prevVertexSize = 0;
  Scalar radius = 0;
  bool isValid = false;
  PointT coords;
  Slash points;
};

struct FeatureModelParams {  // definitions.h:76-105
  int scansPerSweep = 1;
  Scalar minTreeModels = 5, minGroundModels = 36;
  Scalar maxLidarDist = 20, maxGroundLidarDist = 25, minGroundLidarDist = 5;
  bool twoStepOptim = true;
  int groundRadiiBins = 2, groundThetaBins = 18;
  Scalar groundRetainThresh = 0.05, groundMatchThresh = 2.0;
  Scalar roughTreeMatchThresh = 3.0, treeMatchThresh = 0.5;
  Scalar maxTreeRadius = 0.3, maxAxisTheta = 10, maxFocusOutlierDistance = 0.5;
  Scalar AddNewTreeThreshDist = 1.5;
  int featuresPerTree = 20, numGroundFeatures = 5;
  Scalar defaultTreeRadius = 0.2;
};

namespace sloam_b200 {

// Sensor geometry / capacities that the reference keeps in other objects
// (Segmentation ctor, Instance::Params) or hard-codes.
struct HostConfig {
  int img_h = 64, img_w = 2048;  // sloam/params/sloam.yaml:31-33
  float fov_up = 22.5f, fov_down = -22.5f;
  float max_dist_to_centroid = 0.2f;
  int max_trees = 512, max_map_models = 512;
  bool do_destagger = false;  // Segmentation ctor argument (inference.cpp:5-14); sloamNode.cpp:80 defaults to true
};

inline void fill_params(sloam_params &p, const FeatureModelParams &f, const HostConfig &h) {
  sloam_b200_default_params(&p);
  p.img_h = h.img_h; p.img_w = h.img_w; p.fov_up_deg = h.fov_up; p.fov_down_deg = h.fov_down;
  p.max_dist_to_centroid = h.max_dist_to_centroid;
  p.max_trees = h.max_trees; p.max_map_models = h.max_map_models;
  p.do_destagger = h.do_destagger ? 1 : 0;
  p.scansPerSweep = f.scansPerSweep;
  p.minTreeModels = f.minTreeModels; p.minGroundModels = f.minGroundModels;
  p.maxLidarDist = f.maxLidarDist; p.maxGroundLidarDist = f.maxGroundLidarDist;
  p.minGroundLidarDist = f.minGroundLidarDist; p.twoStepOptim = f.twoStepOptim ? 1 : 0;
  p.groundRadiiBins = f.groundRadiiBins; p.groundThetaBins = f.groundThetaBins;
  p.groundRetainThresh = f.groundRetainThresh; p.groundMatchThresh = f.groundMatchThresh;
  p.roughTreeMatchThresh = f.roughTreeMatchThresh; p.treeMatchThresh = f.treeMatchThresh;
  p.maxTreeRadius = f.maxTreeRadius; p.maxAxisTheta = f.maxAxisTheta;
  p.maxFocusOutlierDistance = f.maxFocusOutlierDistance; p.AddNewTreeThreshDist = f.AddNewTreeThreshDist;
  p.featuresPerTree = f.featuresPerTree; p.numGroundFeatures = f.numGroundFeatures;
  p.defaultTreeRadius = f.defaultTreeRadius;
  if (p.max_prev_planes < p.groundRadiiBins * p.groundThetaBins)
    p.max_prev_planes = p.groundRadiiBins * p.groundThetaBins;
}

// Device staging of the mirror: one arena per context instead of a cudaMalloc per buffer and
// call.  Buffers are bump-allocated; when the last live buffer of a call is released the arena
// rewinds, and if the call needed more than the arena holds it is re-created with that size --
// so from the second call of a given shape on there is no device allocation at all.
struct Arena {
  char *base = nullptr;
  size_t cap = 0, off = 0, want = 0;
  int live = 0;
};
inline std::map<sloam_ctx *, Arena> &arenas() {
  static std::map<sloam_ctx *, Arena> m;
  return m;
}

// RAII device buffer through the C ABI helpers (no CUDA headers on the host side).
class DevBuf {
 public:
  DevBuf(sloam_ctx *c, size_t bytes) : c_(c), bytes_(bytes) {
    Arena &a = arenas()[c];
    const size_t need = (std::max<size_t>(bytes, 16) + 255) / 256 * 256;
    a.want += need;
    if (a.live == 0 && a.cap < std::max(a.want, need)) {  // between calls: grow to what the last call asked for
      if (a.base) sloam_b200_dev_free(c, a.base);
      a.cap = std::max(std::max(a.want, need) * 5 / 4, (size_t)1 << 20);
      a.base = static_cast<char *>(sloam_b200_dev_alloc(c, a.cap));
      if (!a.base) { a.cap = 0; throw std::runtime_error("sloam_b200: device allocation failed"); }
      a.off = 0;
    }
    if (a.live == 0) a.want = need;
    if (a.off + need <= a.cap) {
      p_ = a.base + a.off;
      a.off += need;
      ++a.live;
    } else {  // first call of this shape: beyond the arena, allocated on its own
      p_ = sloam_b200_dev_alloc(c, bytes);
      own_ = true;
      if (!p_) throw std::runtime_error("sloam_b200: device allocation failed");
    }
  }
  ~DevBuf() {
    if (own_) { sloam_b200_dev_free(c_, p_); return; }
    Arena &a = arenas()[c_];
    if (--a.live == 0) a.off = 0;
  }
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  template <typename T> T *as() { return static_cast<T *>(p_); }
  void upload(const void *src, size_t bytes) { check(sloam_b200_copy_h2d(c_, p_, src, bytes)); }
  void download(void *dst, size_t bytes) { check(sloam_b200_copy_d2h(c_, dst, p_, bytes)); }
  void check(int rc) {
    if (rc != SLOAM_OK) throw std::runtime_error(std::string("sloam_b200: ") + sloam_b200_last_error(c_));
  }

 private:
  sloam_ctx *c_;
  size_t bytes_;
  void *p_ = nullptr;
  bool own_ = false;
};

// One context per parameter set (K = 1 like a reference call).
class Runtime {
 public:
  Runtime(const FeatureModelParams &f, const HostConfig &h) {
    fill_params(p_, f, h);
    const int rc = sloam_b200_create(&p_, 0, 1, &ctx_);
    if (rc != SLOAM_OK)
      throw std::runtime_error("sloam_b200_create failed (" + std::to_string(rc) +
                               "): no CPU fallback, a B200 is required");
  }
  ~Runtime() {
    auto it = arenas().find(ctx_);
    if (it != arenas().end()) {
      if (it->second.base) sloam_b200_dev_free(ctx_, it->second.base);
      arenas().erase(it);
    }
    sloam_b200_destroy(ctx_);
  }
  Runtime(const Runtime &) = delete;
  Runtime &operator=(const Runtime &) = delete;
  sloam_ctx *ctx() const { return ctx_; }
  const sloam_params &params() const { return p_; }
  void check(int rc) const {
    if (rc != SLOAM_OK) throw std::runtime_error(std::string("sloam_b200: ") + sloam_b200_last_error(ctx_));
  }

 private:
  sloam_params p_{};
  sloam_ctx *ctx_ = nullptr;
};

// The objects of the reference API are created by the thousand (one Plane per ground cell, one
// Cylinder per tree): they share one cached context per parameter set instead of owning one.
inline std::shared_ptr<Runtime> shared_runtime(const FeatureModelParams &f, const HostConfig &h) {
  arenas();  // constructed before the cache, hence destroyed after the contexts that use it
  static std::map<std::string, std::shared_ptr<Runtime>> cache;
  sloam_params p;
  fill_params(p, f, h);
  const std::string key(reinterpret_cast<const char *>(&p), sizeof p);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  if (cache.size() >= 8) cache.clear();  // parameter sweeps: bounded
  auto rt = std::make_shared<Runtime>(f, h);
  cache.emplace(key, rt);
  return rt;
}

// landmarks <-> flattened ABI records
inline void flatten(const std::vector<std::vector<TreeVertex>> &lm, std::vector<sloam_tree> &trees,
                    std::vector<sloam_vertex> &verts, std::vector<sloam_point> &pts) {
  for (const auto &tree : lm) {
    sloam_tree t{};
    t.tree_id = tree.empty() ? -1 : tree[0].treeId;
    t.n_vertices = (int)tree.size();
    t.vertex_begin = (int)verts.size();
    for (const auto &v : tree) {
      sloam_vertex fv{};
      fv.cx = v.coords.x; fv.cy = v.coords.y; fv.cz = v.coords.z;
      fv.radius = (float)v.radius;
      fv.n_points = (int)v.points.size();
      fv.point_begin = (int)pts.size();
      fv.row = -1;
      fv.is_valid = v.isValid ? 1 : 0;
      verts.push_back(fv);
      for (const auto &p : v.points) pts.push_back(sloam_point{p.x, p.y, p.z, p.intensity});
      t.n_points += fv.n_points;
    }
    trees.push_back(t);
  }
}

}  // namespace sloam_b200

// ------------------------------------------------------------------ SemanticObject
template <typename T>
class SemanticObject {  // semanticObject.h:5-19
 public:
  virtual ~SemanticObject() {}
  virtual Scalar distance(const T &model) const = 0;
  virtual Scalar distance(const PointT &point) const = 0;
  virtual void project(const SE3 &tf) = 0;
  T getModel() const { return model; }
  VectorType getFeatures() const { return features; }
  size_t id = 0;
  bool isValid = false;
  VectorType features;
  T model;
};

struct PlaneParameters {  // plane.h:14-19
  Vector4 plane;
  Vector3 centroid;
  double dist = 0;
};

class Plane : public SemanticObject<PlaneParameters> {
 public:
  Plane() {}
  // plane.cpp:3-17: the fit runs on the GPU (binGroundPoints is bypassed by using one cell
  // that keeps every point: groundRetainThresh = 1 -> sorted by z like any retained cell is not
  // wanted here, so the points go through the RunSloam-independent plane-fit entry).
  explicit Plane(const VectorType &points, const FeatureModelParams &fmParams,
                 const sloam_b200::HostConfig &hc = sloam_b200::HostConfig());
  Scalar distance(const PlaneParameters &tgt) const override {  // plane.cpp:131-134
    Vector3 d;
    for (int i = 0; i < 3; ++i) d[i] = model.centroid[i] - tgt.centroid[i];
    return d.norm();
  }
  Scalar distance(const PointT &p) const override {  // plane.cpp:136-151
    const double n = std::sqrt(model.plane[0] * model.plane[0] +
                               (model.plane[1] * model.plane[1] + model.plane[2] * model.plane[2]));
    return std::fabs(model.plane[0] * p.x + model.plane[1] * p.y + model.plane[2] * p.z + model.plane[3]) / n;
  }
  void project(const SE3 &tf) override {  // plane.cpp:153-176
    for (auto &f : features) {
      if (!std::isfinite(f.x) || !std::isfinite(f.y) || !std::isfinite(f.z)) continue;
      Vector3 v; v[0] = f.x; v[1] = f.y; v[2] = f.z;
      const Vector3 r = tf * v;
      f.x = (float)r[0]; f.y = (float)r[1]; f.z = (float)r[2];
    }
    Vector3 n; n[0] = model.plane[0]; n[1] = model.plane[1]; n[2] = model.plane[2];
    SE3 rot = tf; rot.translation() = Vector3();
    const Vector3 rn = rot * n;
    const Vector3 &t = tf.translation();
    model.plane[3] = model.plane[3] - (rn[0] * t[0] + rn[1] * t[1] + rn[2] * t[2]);
    for (int i = 0; i < 3; ++i) model.plane[i] = rn[i];
    model.centroid = tf * model.centroid;
  }
};

struct CylinderParameters {  // cylinder.h:14-23
  Vector3 root, ray;
  std::vector<double> radii;
  std::vector<TreeVertex> vertices;
  double radius = 0, lambda = 1.0;
};

class Cylinder : public SemanticObject<CylinderParameters> {
 public:
  Cylinder() {}
  explicit Cylinder(const std::vector<TreeVertex> vertices, const Plane &gplane,
                    const FeatureModelParams &fmParams,
                    const sloam_b200::HostConfig &hc = sloam_b200::HostConfig());
  Scalar distance(const CylinderParameters &tgt) const override {  // cylinder.cpp:175-194
    double d = 0.0;
    for (double h : {0.0, 3.0, 6.0}) {
      const double s = (h - model.root[2]) / model.ray[2], t = (h - tgt.root[2]) / tgt.ray[2];
      Vector3 e;
      for (int i = 0; i < 3; ++i) e[i] = (model.root[i] + s * model.ray[i]) - (tgt.root[i] + t * tgt.ray[i]);
      d += e.norm();
    }
    return d / 3.0;
  }
  Scalar distance(const PointT &p) const override {  // cylinder.cpp:196-203
    Vector3 e; e[0] = p.x - model.root[0]; e[1] = p.y - model.root[1]; e[2] = p.z - model.root[2];
    const double rr = model.ray[0] * model.ray[0] + (model.ray[1] * model.ray[1] + model.ray[2] * model.ray[2]);
    const double s = (e[0] * model.ray[0] + (e[1] * model.ray[1] + e[2] * model.ray[2])) / rr;
    Vector3 d;
    for (int i = 0; i < 3; ++i) d[i] = e[i] - s * model.ray[i];
    return d.norm() - model.radius;
  }
  void project(const SE3 &tf) override {  // cylinder.cpp:205-211
    Vector3 other;
    for (int i = 0; i < 3; ++i) other[i] = model.root[i] + model.ray[i];
    model.root = tf * model.root;
    other = tf * other;
    for (int i = 0; i < 3; ++i) model.ray[i] = other[i] - model.root[i];
  }
};

// ------------------------------------------------------------------ sloam core
template <typename T>
struct ObjectMatch {  // sloam.h:13-27 (`dist = dist` there leaves the member unset; it is set here)
  ObjectMatch(PointT ft, T obj, Scalar d) : object(obj), dist(d) { feature[0] = ft.x; feature[1] = ft.y; feature[2] = ft.z; }
  Vector3 feature;
  T object;
  double dist;
};

struct SloamInput {  // sloam.h:32-46
  SloamInput() : groundCloud(new CloudT()) {}
  SE3 poseEstimate;
  Scalar distance = 0;
  CloudT::Ptr groundCloud;
  std::vector<Cylinder> mapModels;
  std::vector<std::vector<TreeVertex>> landmarks;
};
struct SloamOutput {  // sloam.h:48-55
  std::vector<int> matches;
  std::vector<Cylinder> tm;
  SE3 T_Map_Curr;
  SE3 T_Delta;
};

// boost::multi_array<VectorType, 2> stand-in: scgf[radial bin][theta bin]
using GroundGrid = std::vector<std::vector<VectorType>>;

namespace sloam {
class sloam {  // sloam.h:57-107
 public:
  explicit sloam(const sloam_b200::HostConfig &hc = sloam_b200::HostConfig()) : hc_(hc) {}
  const FeatureModelParams &fmParams() const { return fmParams_; }
  void setFmParams(const FeatureModelParams &p) { fmParams_ = p; rt_.reset(); }
  std::vector<Plane> getPrevGroundModel() { return prevGPlanes_; }
  CloudT getPrevGroundFeatures() {  // sloam.h:96: the features of prevGPlanes_, one cloud
    CloudT c;
    for (const Plane &pl : prevGPlanes_) c.points.insert(c.points.end(), pl.features.begin(), pl.features.end());
    c.width = (uint32_t)c.points.size(); c.height = 1;
    return c;
  }

  // ---- Pose optimisation (sloam.cpp:33-255): sloam_b200_optimize_pose_dev -------------------
  // joint 6-DoF; true and tf = the optimum iff Ceres would report CONVERGENCE (:241-246)
  bool OptimizePose(const SE3 &poseEstimate, const std::vector<ObjectMatch<Cylinder>> &allTMatch,
                    const std::vector<ObjectMatch<Plane>> &allGMatch, SE3 &tf) {
    sloam_pose out;
    int32_t term[2];
    solve(0, poseEstimate, true, true, allTMatch, allGMatch, out, term);
    if (term[0] != 0) return false;
    tf = SE3(out);
    return true;
  }
  // XYYaw over the trees + ZRollPitch over the ground, each falling back to the estimate
  // unless optimised and converged; always true (:33-53)
  bool TwoStepOptimizePose(const SE3 &poseEstimate, const bool optimTrees, const bool optimGround,
                           const std::vector<ObjectMatch<Cylinder>> &allTMatch,
                           const std::vector<ObjectMatch<Plane>> &allGMatch, SE3 &tf) {
    sloam_pose out;
    int32_t term[2];
    solve(1, poseEstimate, optimTrees, optimGround, allTMatch, allGMatch, out, term);
    tf = SE3(out);
    return true;
  }
  // out = {x, y, yaw} (:55-114).  The composed pose of the two-step entry with the ground block
  // switched off is angle-axis (rx0, ry0, yaw) and translation (x, y, z0): its parts are read back.
  void OptimizeXYYaw(const SE3 &poseEstimate, const bool optimize, const std::vector<ObjectMatch<Cylinder>> &allTMatch,
                     double *out) {
    sloam_pose r;
    int32_t term[2];
    solve(1, poseEstimate, optimize, false, allTMatch, {}, r, term);
    double aa[3];
    quat_to_angle_axis(r.q, aa);
    out[0] = r.t[0]; out[1] = r.t[1]; out[2] = aa[2];
  }
  // out = {z, roll, pitch} (:116-172)
  void OptimizeZRollPitch(const SE3 &poseEstimate, const bool optimize, const std::vector<ObjectMatch<Plane>> &allGMatch,
                          double *out) {
    sloam_pose r;
    int32_t term[2];
    solve(1, poseEstimate, false, optimize, {}, allGMatch, r, term);
    double aa[3];
    quat_to_angle_axis(r.q, aa);
    out[0] = r.t[2]; out[1] = aa[0]; out[2] = aa[1];
  }

  // ---- Model estimation --------------------------------------------------------------------
  void projectModels(const SE3 &tf, std::vector<Cylinder> &landmarks, std::vector<Plane> &planes) {  // :438-451
    for (auto &p : planes) p.project(tf);
    for (auto &l : landmarks) l.project(tf);
  }
  // :330-386 through sloam_b200_ground_planes_dev (kept point lists).  The device bins around the
  // origin, as the reference calls it (:392); for another pose the cloud is shifted on the way in,
  // and every retained point is mapped back to the caller's point through its index (carried in
  // the intensity lane), so the lists hold the original points bit for bit.
  void binGroundPoints(const SE3 pose, const VectorType &points, GroundGrid &scgf) {
    auto rt = runtime(points.size());
    sloam_ctx *c = rt->ctx();
    const sloam_params &p = rt->params();
    const int N = p.img_h * p.img_w, B = p.groundRadiiBins * p.groundThetaBins, Fg = p.numGroundFeatures;
    std::vector<sloam_point> in(points.size());
    const float ox = (float)pose.translation()[0], oy = (float)pose.translation()[1];
    for (size_t i = 0; i < points.size(); ++i) {
      in[i] = sloam_point{points[i].x - ox, points[i].y - oy, points[i].z, 0.f};
      if (ox == 0.f && oy == 0.f) { in[i].x = points[i].x; in[i].y = points[i].y; }
      const uint32_t idx = (uint32_t)i;
      std::memcpy(&in[i].intensity, &idx, 4);
    }
    const int32_t n = (int32_t)points.size();
    sloam_pose ident{}; ident.q[3] = 1.0;
    using sloam_b200::DevBuf;
    DevBuf d_g(c, sizeof(sloam_point) * (size_t)N), d_n(c, 4), d_pose(c, sizeof ident), d_cells(c, sizeof(sloam_cell_plane) * B),
        d_feat(c, sizeof(sloam_point) * (size_t)B * Fg), d_kept(c, sizeof(sloam_point) * (size_t)N), d_off(c, 4 * (size_t)(B + 1));
    d_g.upload(in.data(), sizeof(sloam_point) * in.size());
    d_n.upload(&n, 4);
    d_pose.upload(&ident, sizeof ident);
    rt->check(sloam_b200_ground_planes_dev(c, 1, d_g.as<sloam_point>(), d_n.as<int32_t>(), N, d_pose.as<sloam_pose>(),
                                           d_cells.as<sloam_cell_plane>(), d_feat.as<sloam_point>(), d_kept.as<sloam_point>(),
                                           d_off.as<int32_t>()));
    std::vector<int32_t> off((size_t)B + 1);
    d_off.download(off.data(), 4 * off.size());
    std::vector<sloam_point> kept((size_t)std::max(off[B], 1));
    if (off[B] > 0) d_kept.download(kept.data(), sizeof(sloam_point) * (size_t)off[B]);
    scgf.assign((size_t)p.groundRadiiBins, std::vector<VectorType>((size_t)p.groundThetaBins));
    for (int cell = 0; cell < B; ++cell) {
      VectorType &dst = scgf[(size_t)(cell / p.groundThetaBins)][(size_t)(cell % p.groundThetaBins)];
      for (int i = off[cell]; i < off[cell + 1]; ++i) {
        uint32_t idx;
        std::memcpy(&idx, &kept[(size_t)i].intensity, 4);
        dst.push_back(points[idx]);
      }
    }
  }
  // :388-436: ground cells -> accepted planes, then one cylinder per landmark on its nearest plane
  void computeModels(SloamInput &in, std::vector<Cylinder> &landmarks, std::vector<Plane> &planes) {
    auto rt = runtime(in.groundCloud->points.size());
    sloam_ctx *c = rt->ctx();
    const sloam_params &p = rt->params();
    const int N = p.img_h * p.img_w, B = p.groundRadiiBins * p.groundThetaBins, Fg = p.numGroundFeatures,
              Ft = p.featuresPerTree, T = p.max_trees;
    std::vector<sloam_tree> trees; std::vector<sloam_vertex> verts; std::vector<sloam_point> vpts;
    sloam_b200::flatten(in.landmarks, trees, verts, vpts);
    if ((int)trees.size() > T) throw std::runtime_error("sloam_b200: more landmarks than max_trees");
    for (const sloam_tree &t : trees)
      if (t.n_vertices > p.max_tree_vertices) throw std::runtime_error("sloam_b200: a landmark has more vertices than max_tree_vertices");
    const int32_t n_ground = (int32_t)in.groundCloud->points.size(), n_trees = (int32_t)trees.size();
    const sloam_pose pose = in.poseEstimate.abi();
    using sloam_b200::DevBuf;
    DevBuf d_g(c, sizeof(sloam_point) * (size_t)N), d_n(c, 4), d_pose(c, sizeof pose), d_cells(c, sizeof(sloam_cell_plane) * B),
        d_feat(c, sizeof(sloam_point) * (size_t)B * Fg), d_trees(c, sizeof(sloam_tree) * T), d_nt(c, 4),
        d_verts(c, sizeof(sloam_vertex) * (size_t)T * p.max_tree_vertices), d_vpts(c, sizeof(sloam_point) * (size_t)N),
        d_models(c, sizeof(sloam_tree_model) * T), d_tfeat(c, sizeof(sloam_point) * (size_t)T * Ft);
    if (vpts.size() > (size_t)N) throw std::runtime_error("sloam_b200: landmark points exceed H*W");
    d_g.upload(in.groundCloud->points.data(), sizeof(sloam_point) * (size_t)n_ground);
    d_n.upload(&n_ground, 4);
    d_pose.upload(&pose, sizeof pose);
    rt->check(sloam_b200_ground_planes_dev(c, 1, d_g.as<sloam_point>(), d_n.as<int32_t>(), N, d_pose.as<sloam_pose>(),
                                           d_cells.as<sloam_cell_plane>(), d_feat.as<sloam_point>(), nullptr, nullptr));
    std::vector<sloam_cell_plane> cells((size_t)B);
    std::vector<sloam_point> feats((size_t)B * Fg);
    d_cells.download(cells.data(), sizeof(sloam_cell_plane) * B);
    d_feat.download(feats.data(), sizeof(sloam_point) * feats.size());
    for (int cell = 0; cell < B; ++cell) {
      if (!cells[(size_t)cell].accepted) continue;
      Plane pl;
      for (int a = 0; a < 4; ++a) pl.model.plane[a] = cells[(size_t)cell].model.plane[a];
      for (int a = 0; a < 3; ++a) pl.model.centroid[a] = cells[(size_t)cell].model.centroid[a];
      pl.isValid = true;
      for (int f = 0; f < Fg; ++f) { const sloam_point &q = feats[(size_t)cell * Fg + f]; PointT t; t.x = q.x; t.y = q.y; t.z = q.z; t.intensity = q.intensity; pl.features.push_back(t); }
      planes.push_back(pl);
    }
    if (planes.empty() || trees.empty()) return;  // :414
    // the flattened landmarks are laid out like compute_graph's output: vertex_begin = t * max_tree_vertices
    std::vector<sloam_vertex> vpad((size_t)T * p.max_tree_vertices);
    for (size_t t = 0; t < trees.size(); ++t) {
      for (int k = 0; k < trees[t].n_vertices; ++k) vpad[t * p.max_tree_vertices + k] = verts[(size_t)trees[t].vertex_begin + k];
      trees[t].vertex_begin = (int)(t * p.max_tree_vertices);
    }
    d_trees.upload(trees.data(), sizeof(sloam_tree) * trees.size());
    d_nt.upload(&n_trees, 4);
    d_verts.upload(vpad.data(), sizeof(sloam_vertex) * vpad.size());
    if (!vpts.empty()) d_vpts.upload(vpts.data(), sizeof(sloam_point) * vpts.size());
    rt->check(sloam_b200_cylinders_dev(c, 1, d_trees.as<sloam_tree>(), d_nt.as<int32_t>(), d_verts.as<sloam_vertex>(),
                                       d_vpts.as<sloam_point>(), d_cells.as<sloam_cell_plane>(), d_models.as<sloam_tree_model>(),
                                       d_tfeat.as<sloam_point>()));
    std::vector<sloam_tree_model> models(trees.size());
    std::vector<sloam_point> tfeat(trees.size() * (size_t)Ft);
    d_models.download(models.data(), sizeof(sloam_tree_model) * models.size());
    d_tfeat.download(tfeat.data(), sizeof(sloam_point) * tfeat.size());
    for (size_t t = 0; t < trees.size(); ++t) {
      if (!models[t].is_valid) continue;  // :433-434
      Cylinder cy;
      for (int a = 0; a < 3; ++a) { cy.model.root[a] = models[t].model.root[a]; cy.model.ray[a] = models[t].model.ray[a]; }
      cy.model.radius = models[t].model.radius;
      cy.model.vertices = in.landmarks[t];
      for (const auto &v : in.landmarks[t]) cy.model.radii.push_back(v.radius);
      cy.id = (size_t)models[t].id;
      cy.isValid = true;
      for (int f = 0; f < Ft; ++f) { const sloam_point &q = tfeat[t * Ft + f]; PointT pt; pt.x = q.x; pt.y = q.y; pt.z = q.z; pt.intensity = q.intensity; cy.features.push_back(pt); }
      landmarks.push_back(cy);
    }
  }

  // ---- Data association (:257-328) ----------------------------------------------------------
  // matchIndices[i] = index of the nearest map cylinder when closer than AddNewTreeThreshDist
  void matchModels(const std::vector<Cylinder> &currObjects, const std::vector<Cylinder> &mapObjects,
                   std::vector<int> &matchIndices) {
    std::vector<int32_t> idx; std::vector<double> dist;
    nearest(currObjects, mapObjects, nullptr, idx, dist);
    if (matchIndices.size() < currObjects.size()) matchIndices.resize(currObjects.size(), -1);
    for (size_t i = 0; i < currObjects.size(); ++i)
      if (idx[i] >= 0 && dist[i] < fmParams_.treeMatchThresh + 100 && dist[i] < fmParams_.AddNewTreeThreshDist) matchIndices[i] = idx[i];
  }
  // one ObjectMatch per feature of every current object whose nearest map object (after
  // projecting the current object with tf) is closer than distThresh
  std::vector<ObjectMatch<Cylinder>> matchFeatures(const SE3 tf, const std::vector<Cylinder> &currObjects,
                                                   const std::vector<Cylinder> &mapObjects, const Scalar distThresh) {
    std::vector<int32_t> idx; std::vector<double> dist;
    nearest(currObjects, mapObjects, &tf, idx, dist);
    std::vector<ObjectMatch<Cylinder>> matches;
    for (size_t i = 0; i < currObjects.size(); ++i)
      if (idx[i] >= 0 && dist[i] < distThresh)
        for (const PointT &f : currObjects[i].features) matches.emplace_back(f, mapObjects[(size_t)idx[i]], dist[i]);
    return matches;
  }
  std::vector<ObjectMatch<Plane>> matchFeatures(const SE3 tf, const std::vector<Plane> &currObjects,
                                                const std::vector<Plane> &mapObjects, const Scalar distThresh) {
    std::vector<ObjectMatch<Plane>> matches;
    if (currObjects.empty() || mapObjects.empty()) return matches;
    auto rt = runtime(0);
    sloam_ctx *c = rt->ctx();
    const int32_t nd = (int32_t)currObjects.size(), nm = (int32_t)mapObjects.size();
    std::vector<sloam_plane> det((size_t)nd), map((size_t)nm);
    for (int i = 0; i < nd; ++i) to_abi(currObjects[(size_t)i], det[(size_t)i]);
    for (int i = 0; i < nm; ++i) to_abi(mapObjects[(size_t)i], map[(size_t)i]);
    const sloam_pose pose = tf.abi();
    using sloam_b200::DevBuf;
    DevBuf d_det(c, sizeof(sloam_plane) * nd), d_nd(c, 4), d_map(c, sizeof(sloam_plane) * nm), d_nm(c, 4), d_tf(c, sizeof pose),
        d_idx(c, 4 * (size_t)nd), d_dist(c, 8 * (size_t)nd);
    d_det.upload(det.data(), sizeof(sloam_plane) * nd); d_nd.upload(&nd, 4);
    d_map.upload(map.data(), sizeof(sloam_plane) * nm); d_nm.upload(&nm, 4);
    d_tf.upload(&pose, sizeof pose);
    rt->check(sloam_b200_associate_planes_dev(c, 1, d_det.as<sloam_plane>(), d_nd.as<int32_t>(), nd, d_tf.as<sloam_pose>(),
                                              d_map.as<sloam_plane>(), d_nm.as<int32_t>(), nm, d_idx.as<int32_t>(),
                                              d_dist.as<double>()));
    std::vector<int32_t> idx((size_t)nd); std::vector<double> dist((size_t)nd);
    d_idx.download(idx.data(), 4 * (size_t)nd);
    d_dist.download(dist.data(), 8 * (size_t)nd);
    for (int i = 0; i < nd; ++i)
      if (idx[(size_t)i] >= 0 && dist[(size_t)i] < distThresh)
        for (const PointT &f : currObjects[(size_t)i].features) matches.emplace_back(f, mapObjects[(size_t)idx[(size_t)i]], dist[(size_t)i]);
    return matches;
  }

  // sloam.cpp:453-532, one keyframe.  Returns false exactly when the reference does.
  bool RunSloam(SloamInput &in, SloamOutput &out) {
    auto rt = runtime(in.groundCloud->points.size());
    sloam_ctx *c = rt->ctx();
    const sloam_params &p = rt->params();
    const int N = p.img_h * p.img_w, T = p.max_trees, M = p.max_map_models, PP = p.max_prev_planes;
    const int B = p.groundRadiiBins * p.groundThetaBins, Fg = p.numGroundFeatures;
    std::vector<sloam_tree> trees; std::vector<sloam_vertex> verts; std::vector<sloam_point> vpts;
    sloam_b200::flatten(in.landmarks, trees, verts, vpts);
    if ((int)trees.size() > T || (int)in.mapModels.size() > M || (int)prevGPlanes_.size() > PP)
      throw std::runtime_error("sloam_b200: capacity exceeded (max_trees / max_map_models / max_prev_planes)");
    for (const sloam_tree &t : trees)
      if (t.n_vertices > p.max_tree_vertices) throw std::runtime_error("sloam_b200: a landmark has more vertices than max_tree_vertices");
    const int32_t n_ground = (int32_t)in.groundCloud->points.size(), n_trees = (int32_t)trees.size();
    if (n_ground > N) throw std::runtime_error("sloam_b200: ground cloud larger than the context was sized for");
    std::vector<sloam_cylinder> map(std::max<size_t>(in.mapModels.size(), 1));
    for (size_t i = 0; i < in.mapModels.size(); ++i) to_abi(in.mapModels[i], map[i]);
    std::vector<sloam_plane> prev(std::max<size_t>(prevGPlanes_.size(), 1));
    for (size_t i = 0; i < prevGPlanes_.size(); ++i) to_abi(prevGPlanes_[i], prev[i]);
    const int32_t n_map = (int32_t)in.mapModels.size(), n_prev = (int32_t)prevGPlanes_.size();
    const uint8_t first = firstScan_ ? 1 : 0;
    const sloam_pose pose = in.poseEstimate.abi();
    // one packed upload and one packed download per call (every copy through the C ABI is a
    // synchronous round trip): the inputs are laid out in one host block mirroring one device block
    using sloam_b200::DevBuf;
    struct Part { size_t off, bytes; };
    size_t in_bytes = 0, out_bytes = 0;
    auto part = [](size_t &total, size_t bytes) { Part q{total, bytes}; total = (total + bytes + 255) / 256 * 256; return q; };
    const Part i_ground = part(in_bytes, sizeof(sloam_point) * (size_t)n_ground), i_ng = part(in_bytes, 4),
               i_trees = part(in_bytes, sizeof(sloam_tree) * (size_t)T), i_nt = part(in_bytes, 4),
               i_verts = part(in_bytes, sizeof(sloam_vertex) * std::max<size_t>(verts.size(), 1)),
               i_vpts = part(in_bytes, sizeof(sloam_point) * std::max<size_t>(vpts.size(), 1)), i_pose = part(in_bytes, sizeof pose),
               i_first = part(in_bytes, 1), i_map = part(in_bytes, sizeof(sloam_cylinder) * (size_t)M), i_nmap = part(in_bytes, 4),
               i_prev = part(in_bytes, sizeof(sloam_plane) * (size_t)PP), i_nprev = part(in_bytes, 4);
    const Part o_res = part(out_bytes, sizeof(sloam_kf_result)), o_npl = part(out_bytes, 4), o_match = part(out_bytes, 4 * (size_t)T),
               o_tmid = part(out_bytes, 4 * (size_t)T), o_tm = part(out_bytes, sizeof(sloam_cylinder) * (size_t)T),
               o_planes = part(out_bytes, sizeof(sloam_plane) * (size_t)PP);
    stage_.resize(std::max(in_bytes, out_bytes));
    auto put = [&](const Part &q, const void *src, size_t bytes) { if (bytes) std::memcpy(stage_.data() + q.off, src, bytes); };
    put(i_ground, in.groundCloud->points.data(), sizeof(sloam_point) * (size_t)n_ground);
    put(i_ng, &n_ground, 4);
    put(i_trees, trees.data(), sizeof(sloam_tree) * trees.size());
    put(i_nt, &n_trees, 4);
    put(i_verts, verts.data(), sizeof(sloam_vertex) * verts.size());
    put(i_vpts, vpts.data(), sizeof(sloam_point) * vpts.size());
    put(i_pose, &pose, sizeof pose);
    put(i_first, &first, 1);
    put(i_map, map.data(), sizeof(sloam_cylinder) * in.mapModels.size());
    put(i_nmap, &n_map, 4);
    put(i_prev, prev.data(), sizeof(sloam_plane) * prevGPlanes_.size());
    put(i_nprev, &n_prev, 4);
    DevBuf d_in(c, in_bytes), d_out(c, out_bytes);
    d_in.upload(stage_.data(), in_bytes);
    char *di = d_in.as<char>(), *dout = d_out.as<char>();
    sloam_batch_in bi{};
    bi.pose_est = reinterpret_cast<sloam_pose *>(di + i_pose.off); bi.first_scan = reinterpret_cast<uint8_t *>(di + i_first.off);
    bi.map_models = reinterpret_cast<sloam_cylinder *>(di + i_map.off); bi.n_map_models = reinterpret_cast<int32_t *>(di + i_nmap.off);
    bi.prev_planes = reinterpret_cast<sloam_plane *>(di + i_prev.off); bi.n_prev_planes = reinterpret_cast<int32_t *>(di + i_nprev.off);
    sloam_batch_out bo{};
    bo.results = reinterpret_cast<sloam_kf_result *>(dout + o_res.off); bo.matches = reinterpret_cast<int32_t *>(dout + o_match.off);
    bo.tm = reinterpret_cast<sloam_cylinder *>(dout + o_tm.off); bo.tm_id = reinterpret_cast<int32_t *>(dout + o_tmid.off);
    bo.planes = reinterpret_cast<sloam_plane *>(dout + o_planes.off); bo.n_planes = reinterpret_cast<int32_t *>(dout + o_npl.off);
    rt->check(sloam_b200_run_sloam_dev(c, 1, reinterpret_cast<sloam_point *>(di + i_ground.off),
                                       reinterpret_cast<int32_t *>(di + i_ng.off), std::max(n_ground, 1),
                                       reinterpret_cast<sloam_tree *>(di + i_trees.off), reinterpret_cast<int32_t *>(di + i_nt.off),
                                       reinterpret_cast<sloam_vertex *>(di + i_verts.off), (int)std::max<size_t>(verts.size(), 1),
                                       reinterpret_cast<sloam_point *>(di + i_vpts.off), (int)std::max<size_t>(vpts.size(), 1), &bi, &bo));
    d_out.download(stage_.data(), out_bytes);
    sloam_kf_result res;
    std::memcpy(&res, stage_.data() + o_res.off, sizeof res);
    last_ = res;
    int32_t npl = 0;
    std::memcpy(&npl, stage_.data() + o_npl.off, 4);
    std::vector<sloam_plane> planes(std::max(npl, 1));
    std::memcpy(planes.data(), stage_.data() + o_planes.off, sizeof(sloam_plane) * (size_t)npl);
    if (SLOAM_KF_CODE(res.status) == SLOAM_KF_EMPTY_MAP || SLOAM_KF_CODE(res.status) == SLOAM_KF_NO_MODELS) return false;  // :476-486
    std::vector<int32_t> matches(std::max(res.n_landmarks, 1)), ids(std::max(res.n_landmarks, 1));
    std::vector<sloam_cylinder> tm(std::max(res.n_landmarks, 1));
    std::memcpy(matches.data(), stage_.data() + o_match.off, 4 * (size_t)res.n_landmarks);
    std::memcpy(ids.data(), stage_.data() + o_tmid.off, 4 * (size_t)res.n_landmarks);
    std::memcpy(tm.data(), stage_.data() + o_tm.off, sizeof(sloam_cylinder) * (size_t)res.n_landmarks);
    out.matches.assign(matches.begin(), matches.begin() + res.n_landmarks);
    out.tm.clear();
    for (int i = 0; i < res.n_landmarks; ++i) {
      Cylinder cy;
      for (int a = 0; a < 3; ++a) { cy.model.root[a] = tm[i].root[a]; cy.model.ray[a] = tm[i].ray[a]; }
      cy.model.radius = tm[i].radius;
      cy.id = (size_t)ids[i];
      cy.isValid = true;
      out.tm.push_back(cy);
    }
    out.T_Map_Curr = SE3(res.T_Map_Curr);
    out.T_Delta = SE3(res.T_Delta);
    // prevGPlanes_ = planes (:471,:525): the accepted planes projected with the new pose, with
    // their features (Plane::project moves them with pcl::transformPoint: float results)
    sloam_intermediates im{};
    rt->check(sloam_b200_get_intermediates(c, &im));
    std::vector<sloam_cell_plane> cells((size_t)B);
    std::vector<sloam_point> feats((size_t)B * Fg);
    rt->check(sloam_b200_copy_d2h(c, cells.data(), im.cells, sizeof(sloam_cell_plane) * B));
    rt->check(sloam_b200_copy_d2h(c, feats.data(), im.cell_features, sizeof(sloam_point) * feats.size()));
    prevGPlanes_.clear();
    int acc = 0;
    for (int cell = 0; cell < B && acc < npl; ++cell) {
      if (!cells[(size_t)cell].accepted) continue;
      Plane pl;
      for (int a = 0; a < 4; ++a) pl.model.plane[a] = planes[acc].plane[a];
      for (int a = 0; a < 3; ++a) pl.model.centroid[a] = planes[acc].centroid[a];
      pl.isValid = true;
      for (int f = 0; f < Fg; ++f) {
        const sloam_point &q = feats[(size_t)cell * Fg + f];
        Vector3 v; v[0] = q.x; v[1] = q.y; v[2] = q.z;
        const Vector3 w = out.T_Map_Curr * v;
        PointT t; t.x = (float)w[0]; t.y = (float)w[1]; t.z = (float)w[2]; t.intensity = q.intensity;
        pl.features.push_back(t);
      }
      prevGPlanes_.push_back(pl);
      ++acc;
    }
    firstScan_ = false;
    return res.success != 0;
  }
  const sloam_kf_result &lastResult() const { return last_; }

 private:
  static void to_abi(const Cylinder &cy, sloam_cylinder &o) {
    for (int a = 0; a < 3; ++a) { o.root[a] = cy.model.root[a]; o.ray[a] = cy.model.ray[a]; }
    o.radius = cy.model.radius;
  }
  static void to_abi(const Plane &pl, sloam_plane &o) {
    for (int a = 0; a < 4; ++a) o.plane[a] = pl.model.plane[a];
    for (int a = 0; a < 3; ++a) o.centroid[a] = pl.model.centroid[a];
  }
  // ceres::QuaternionToAngleAxis on (x, y, z, w)
  static void quat_to_angle_axis(const double *q, double *aa) {
    const double s2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
    if (s2 > 0.0) {
      const double s = std::sqrt(s2), cw = q[3];
      const double two_theta = 2.0 * (cw < 0.0 ? std::atan2(-s, -cw) : std::atan2(s, cw));
      const double k = two_theta / s;
      aa[0] = q[0] * k; aa[1] = q[1] * k; aa[2] = q[2] * k;
    } else {
      aa[0] = q[0] * 2.0; aa[1] = q[1] * 2.0; aa[2] = q[2] * 2.0;
    }
  }
  // nearest map cylinder of every current cylinder (projected with tf when given)
  void nearest(const std::vector<Cylinder> &curr, const std::vector<Cylinder> &mapObjects, const SE3 *tf,
               std::vector<int32_t> &idx, std::vector<double> &dist) {
    idx.assign(curr.size(), -1);
    dist.assign(curr.size(), std::numeric_limits<double>::infinity());
    if (curr.empty() || mapObjects.empty()) return;
    auto rt = runtime(0);
    sloam_ctx *c = rt->ctx();
    const int32_t nd = (int32_t)curr.size(), nm = (int32_t)mapObjects.size();
    std::vector<sloam_cylinder> det((size_t)nd), map((size_t)nm);
    for (int i = 0; i < nd; ++i) to_abi(curr[(size_t)i], det[(size_t)i]);
    for (int i = 0; i < nm; ++i) to_abi(mapObjects[(size_t)i], map[(size_t)i]);
    using sloam_b200::DevBuf;
    DevBuf d_det(c, sizeof(sloam_cylinder) * nd), d_nd(c, 4), d_map(c, sizeof(sloam_cylinder) * nm), d_nm(c, 4),
        d_tf(c, sizeof(sloam_pose)), d_idx(c, 4 * (size_t)nd), d_dist(c, 8 * (size_t)nd);
    d_det.upload(det.data(), sizeof(sloam_cylinder) * nd); d_nd.upload(&nd, 4);
    d_map.upload(map.data(), sizeof(sloam_cylinder) * nm); d_nm.upload(&nm, 4);
    if (tf) d_tf.upload(&tf->abi(), sizeof(sloam_pose));
    rt->check(sloam_b200_associate_dev(c, 1, d_det.as<sloam_cylinder>(), d_nd.as<int32_t>(), nd,
                                       tf ? d_tf.as<sloam_pose>() : nullptr, d_map.as<sloam_cylinder>(), d_nm.as<int32_t>(),
                                       nm, 0, d_idx.as<int32_t>(), d_dist.as<double>()));
    d_idx.download(idx.data(), 4 * (size_t)nd);
    d_dist.download(dist.data(), 8 * (size_t)nd);
  }
  // OptimizePose / TwoStepOptimizePose on explicit match lists
  void solve(int mode, const SE3 &est, bool optimTrees, bool optimGround, const std::vector<ObjectMatch<Cylinder>> &tm,
             const std::vector<ObjectMatch<Plane>> &gm, sloam_pose &out, int32_t term[2]) {
    auto rt = runtime(0);
    sloam_ctx *c = rt->ctx();
    const int32_t nt = (int32_t)tm.size(), ng = (int32_t)gm.size();
    const int ts = std::max(nt, 1), gs = std::max(ng, 1);
    std::vector<double> tf((size_t)ts * 3, 0.0), gf((size_t)gs * 3, 0.0);
    std::vector<sloam_cylinder> to((size_t)ts);
    std::vector<sloam_plane> go((size_t)gs);
    for (int i = 0; i < nt; ++i) { for (int a = 0; a < 3; ++a) tf[(size_t)i * 3 + a] = tm[(size_t)i].feature[a]; to_abi(tm[(size_t)i].object, to[(size_t)i]); }
    for (int i = 0; i < ng; ++i) { for (int a = 0; a < 3; ++a) gf[(size_t)i * 3 + a] = gm[(size_t)i].feature[a]; to_abi(gm[(size_t)i].object, go[(size_t)i]); }
    const sloam_pose pose = est.abi();
    const uint8_t ot = optimTrees ? 1 : 0, og = optimGround ? 1 : 0;
    using sloam_b200::DevBuf;
    DevBuf d_pose(c, sizeof pose), d_tf(c, 8 * tf.size()), d_to(c, sizeof(sloam_cylinder) * to.size()), d_nt(c, 4),
        d_gf(c, 8 * gf.size()), d_go(c, sizeof(sloam_plane) * go.size()), d_ng(c, 4), d_ot(c, 1), d_og(c, 1),
        d_out(c, sizeof(sloam_pose)), d_it(c, 8), d_term(c, 8);
    d_pose.upload(&pose, sizeof pose);
    d_tf.upload(tf.data(), 8 * tf.size()); d_to.upload(to.data(), sizeof(sloam_cylinder) * to.size()); d_nt.upload(&nt, 4);
    d_gf.upload(gf.data(), 8 * gf.size()); d_go.upload(go.data(), sizeof(sloam_plane) * go.size()); d_ng.upload(&ng, 4);
    d_ot.upload(&ot, 1); d_og.upload(&og, 1);
    rt->check(sloam_b200_optimize_pose_dev(c, 1, mode, d_pose.as<sloam_pose>(), d_tf.as<double>(), d_to.as<sloam_cylinder>(),
                                           d_nt.as<int32_t>(), ts, d_gf.as<double>(), d_go.as<sloam_plane>(), d_ng.as<int32_t>(), gs,
                                           d_ot.as<uint8_t>(), d_og.as<uint8_t>(), d_out.as<sloam_pose>(), d_it.as<int32_t>(),
                                           d_term.as<int32_t>()));
    d_out.download(&out, sizeof out);
    d_term.download(term, 8);
  }
  // the shared context for this parameter set, wide enough for n_ground points
  std::shared_ptr<sloam_b200::Runtime> runtime(size_t n_ground) {
    if (!rt_ || (size_t)rt_->params().img_h * rt_->params().img_w < n_ground) {
      sloam_b200::HostConfig h = hc_;
      while ((size_t)h.img_h * h.img_w < n_ground) h.img_w *= 2;
      rt_ = sloam_b200::shared_runtime(fmParams_, h);
    }
    return rt_;
  }
  sloam_b200::HostConfig hc_;
  FeatureModelParams fmParams_;
  std::shared_ptr<sloam_b200::Runtime> rt_;
  std::vector<Plane> prevGPlanes_;
  bool firstScan_ = true;
  sloam_kf_result last_{};
  std::vector<char> stage_;  // host side of the packed copies, kept between calls
};
}  // namespace sloam

// ------------------------------------------------------------------ Plane / Cylinder ctors
inline Plane::Plane(const VectorType &points, const FeatureModelParams &fmParams,
                    const sloam_b200::HostConfig &hc) {
  // One polar cell covering everything, retaining every point in input order
  // (groundRetainThresh = 1 keeps all points but would sort them; the reference's Plane
  // constructor does not sort, so the cell is made "small": 1/thresh >= size).
  features = points;
  if ((int)features.size() < fmParams.numGroundFeatures || features.size() < 3) { isValid = false; return; }
  FeatureModelParams f = fmParams;
  f.groundRadiiBins = 1; f.groundThetaBins = 1;
  f.minGroundLidarDist = -1.0; f.maxGroundLidarDist = 1e30;
  f.groundRetainThresh = 1.0 / 1073741824.0;  // 1/thresh >= any size: never sorted (and one cache key for all sizes)
  sloam_b200::HostConfig h = hc;
  while ((size_t)h.img_h * h.img_w < features.size()) h.img_w *= 2;
  const auto rtp = sloam_b200::shared_runtime(f, h);  // one cached context, not one per object
  sloam_b200::Runtime &rt = *rtp;
  sloam_ctx *c = rt.ctx();
  const int N = rt.params().img_h * rt.params().img_w, Fg = f.numGroundFeatures;
  const int32_t n = (int32_t)features.size();
  sloam_pose ident{}; ident.q[3] = 1.0; ident.t[2] = 1e9;  // acceptance is not part of the ctor
  sloam_b200::DevBuf d_g(c, sizeof(sloam_point) * (size_t)N), d_n(c, 4), d_pose(c, sizeof ident),
      d_cells(c, sizeof(sloam_cell_plane)), d_feat(c, sizeof(sloam_point) * Fg);
  d_g.upload(features.data(), sizeof(sloam_point) * features.size());
  d_n.upload(&n, 4);
  d_pose.upload(&ident, sizeof ident);
  rt.check(sloam_b200_ground_planes_dev(c, 1, d_g.as<sloam_point>(), d_n.as<int32_t>(), N, d_pose.as<sloam_pose>(),
                                        d_cells.as<sloam_cell_plane>(), d_feat.as<sloam_point>(), nullptr, nullptr));
  sloam_cell_plane cell;
  d_cells.download(&cell, sizeof cell);
  isValid = cell.is_valid != 0;
  for (int a = 0; a < 4; ++a) model.plane[a] = cell.model.plane[a];
  for (int a = 0; a < 3; ++a) model.centroid[a] = cell.model.centroid[a];
  features.resize(Fg);  // plane.cpp:14
}

inline Cylinder::Cylinder(const std::vector<TreeVertex> vertices, const Plane &gplane,
                          const FeatureModelParams &fmParams, const sloam_b200::HostConfig &hc) {
  const auto rtp = sloam_b200::shared_runtime(fmParams, hc);  // one cached context, not one per object
  sloam_b200::Runtime &rt = *rtp;
  sloam_ctx *c = rt.ctx();
  const sloam_params &p = rt.params();
  std::vector<sloam_tree> trees; std::vector<sloam_vertex> verts; std::vector<sloam_point> vpts;
  sloam_b200::flatten({vertices}, trees, verts, vpts);
  if ((int)vertices.size() > p.max_tree_vertices)
    throw std::runtime_error("sloam_b200: a landmark has more vertices than max_tree_vertices");
  const int B = p.groundRadiiBins * p.groundThetaBins, Ft = p.featuresPerTree;
  std::vector<sloam_cell_plane> cells(B);
  std::memset(cells.data(), 0, sizeof(sloam_cell_plane) * B);
  for (int a = 0; a < 4; ++a) cells[0].model.plane[a] = gplane.model.plane[a];
  for (int a = 0; a < 3; ++a) cells[0].model.centroid[a] = gplane.model.centroid[a];
  cells[0].is_valid = cells[0].accepted = 1;
  const int32_t one = 1;
  sloam_b200::DevBuf d_trees(c, sizeof(sloam_tree) * p.max_trees), d_nt(c, 4),
      d_verts(c, sizeof(sloam_vertex) * std::max<size_t>(verts.size(), 1) + 0),
      d_vpts(c, sizeof(sloam_point) * std::max<size_t>(vpts.size(), 1)), d_cells(c, sizeof(sloam_cell_plane) * B),
      d_models(c, sizeof(sloam_tree_model) * p.max_trees), d_feat(c, sizeof(sloam_point) * p.max_trees * Ft);
  // the stage entry uses the compute_graph strides: one tree, so offsets are the same
  d_trees.upload(trees.data(), sizeof(sloam_tree));
  d_nt.upload(&one, 4);
  d_verts.upload(verts.data(), sizeof(sloam_vertex) * verts.size());
  d_vpts.upload(vpts.data(), sizeof(sloam_point) * vpts.size());
  d_cells.upload(cells.data(), sizeof(sloam_cell_plane) * B);
  rt.check(sloam_b200_cylinders_dev(c, 1, d_trees.as<sloam_tree>(), d_nt.as<int32_t>(), d_verts.as<sloam_vertex>(),
                                    d_vpts.as<sloam_point>(), d_cells.as<sloam_cell_plane>(),
                                    d_models.as<sloam_tree_model>(), d_feat.as<sloam_point>()));
  sloam_tree_model m;
  d_models.download(&m, sizeof m);
  std::vector<sloam_point> f(Ft);
  d_feat.download(f.data(), sizeof(sloam_point) * Ft);
  for (int a = 0; a < 3; ++a) { model.root[a] = m.model.root[a]; model.ray[a] = m.model.ray[a]; }
  model.radius = m.model.radius;
  model.vertices = vertices;
  for (const auto &v : vertices) model.radii.push_back(v.radius);
  id = (size_t)m.id;
  isValid = m.is_valid != 0;
  features.resize(Ft);
  for (int i = 0; i < Ft; ++i) { features[i].x = f[i].x; features[i].y = f[i].y; features[i].z = f[i].z; features[i].intensity = f[i].intensity; }
}

// ------------------------------------------------------------------ Instance (trellis.h)
class Instance {
 public:
  struct Params {  // trellis.h:31-40
    float beam_cluster_threshold = 0.1f;
    float max_dist_to_centroid = 0.2f;
    int min_vertex_size = 2;
    int min_landmark_size = 4;
    float min_landmark_height = 1.0f;
  };
  explicit Instance(const sloam_b200::HostConfig &hc = sloam_b200::HostConfig()) : hc_(hc) {}
  const Params &params() const { return params_; }
  void set_params(const Params &p) { params_ = p; rt_.reset(); }
  void reset_tree_id() {}

  // trellis.cpp:134-140; `cloud` is unused there too
  void computeGraph(const CloudT::Ptr /*cloud*/, const CloudT::Ptr tree_cloud,
                    std::vector<std::vector<TreeVertex>> &landmarks) {
    if (tree_cloud->size() == 0) return;  // trellis.cpp:17
    sloam_b200::HostConfig h = hc_;
    h.img_h = (int)tree_cloud->height; h.img_w = (int)tree_cloud->width;
    h.max_dist_to_centroid = params_.max_dist_to_centroid;
    if (!rt_ || rt_->params().img_h != h.img_h || rt_->params().img_w != h.img_w)
      rt_.reset(new sloam_b200::Runtime(FeatureModelParams(), h));
    sloam_ctx *c = rt_->ctx();
    const sloam_params &p = rt_->params();
    const size_t N = (size_t)p.img_h * p.img_w, T = p.max_trees, V = p.max_tree_vertices;
    if (tree_cloud->points.size() != N) throw std::runtime_error("computeGraph: cloud is not organized H x W");
    sloam_b200::DevBuf d_tree(c, sizeof(sloam_point) * N), d_trees(c, sizeof(sloam_tree) * T), d_nt(c, 4),
        d_verts(c, sizeof(sloam_vertex) * T * V), d_vpts(c, sizeof(sloam_point) * N);
    d_tree.upload(tree_cloud->points.data(), sizeof(sloam_point) * N);
    rt_->check(sloam_b200_compute_graph_dev(c, 1, d_tree.as<sloam_point>(), d_trees.as<sloam_tree>(),
                                            d_nt.as<int32_t>(), d_verts.as<sloam_vertex>(), d_vpts.as<sloam_point>()));
    int32_t nt = 0;
    d_nt.download(&nt, 4);
    std::vector<sloam_tree> trees(std::max(nt, 1));
    std::vector<sloam_vertex> verts(T * V);
    std::vector<sloam_point> vpts(N);
    d_trees.download(trees.data(), sizeof(sloam_tree) * nt);
    d_verts.download(verts.data(), sizeof(sloam_vertex) * T * V);
    d_vpts.download(vpts.data(), sizeof(sloam_point) * N);
    for (int t = 0; t < nt; ++t) {
      std::vector<TreeVertex> tree;
      for (int k = 0; k < trees[t].n_vertices; ++k) {
        const sloam_vertex &fv = verts[trees[t].vertex_begin + k];
        TreeVertex v;
        v.treeId = trees[t].tree_id;
        v.radius = fv.radius;
        v.isValid = fv.is_valid != 0;
        v.coords.x = fv.cx; v.coords.y = fv.cy; v.coords.z = fv.cz;
        for (int j = 0; j < fv.n_points; ++j) {
          const sloam_point &q = vpts[fv.point_begin + j];
          PointT pt; pt.x = q.x; pt.y = q.y; pt.z = q.z; pt.intensity = q.intensity;
          v.points.push_back(pt);
        }
        tree.push_back(v);
      }
      landmarks.push_back(tree);
    }
  }

 private:
  sloam_b200::HostConfig hc_;
  Params params_;
  std::unique_ptr<sloam_b200::Runtime> rt_;
};

// ------------------------------------------------------------------ seg::Segmentation
struct Mask {  // cv::Mat (CV_8U) stand-in: rows x cols labels, 0 other / 1 ground / 255 tree
  int rows = 0, cols = 0;
  std::vector<unsigned char> data;
  Mask() {}
  Mask(int r, int c) : rows(r), cols(c), data((size_t)r * c, 0) {}
};

namespace seg {
class Segmentation {  // inference.h:55-110
 public:
  using Ptr = std::shared_ptr<Segmentation>;
  // The model path of the reference constructor is dropped: the network is out of scope and
  // its H x W mask is supplied by the caller (setLabelSource) before run().
  Segmentation(const float fov_up, const float fov_down, const int img_w, const int img_h, const int /*img_d*/,
               bool do_destagger) {
    sloam_b200::HostConfig h;
    h.img_h = img_h; h.img_w = img_w; h.fov_up = fov_up; h.fov_down = fov_down;
    h.do_destagger = do_destagger;  // _destaggerCloud, inference.cpp:200-228
    rt_.reset(new sloam_b200::Runtime(FeatureModelParams(), h));
  }
  Segmentation(const Segmentation &) = delete;
  Segmentation operator=(const Segmentation &) = delete;
  // stands in for the ONNX session: the mask run() hands back
  void setLabelSource(const Mask &m) { labels_ = m; }

  // inference.cpp:374-448 minus the network: projection (kept for maskCloud, like proj_xs/ys) + labels
  void run(const Cloud::Ptr cloud, Mask &maskImg) {
    const sloam_params &p = rt_->params();
    const size_t N = (size_t)p.img_h * p.img_w;
    if (cloud->points.size() != N) throw std::runtime_error("Segmentation::run: cloud must hold H*W points");
    sloam_ctx *c = rt_->ctx();
    d_points_.reset(new sloam_b200::DevBuf(c, sizeof(sloam_point) * N));
    d_pix_.reset(new sloam_b200::DevBuf(c, 4 * N));
    d_points_->upload(cloud->points.data(), sizeof(sloam_point) * N);
    sloam_b200::DevBuf d_range(c, 4 * N);
    rt_->check(sloam_b200_project_dev(c, 1, d_points_->as<sloam_point>(), d_pix_->as<int32_t>(), d_range.as<float>()));
    range_image.resize(N);
    d_range.download(range_image.data(), 4 * N);
    if (labels_.rows * labels_.cols != (int)N) throw std::runtime_error("Segmentation::run: no label source set");
    maskImg = labels_;
  }

  // inference.cpp:230-273.  The two calls of SLOAMNode::run (val 1 sparse, val 255 dense) share
  // one GPU pass; any other val falls outside the reference's use.
  void maskCloud(const Cloud::Ptr cloud, Mask mask, Cloud::Ptr &outCloud, unsigned char val, bool dense = false) {
    const sloam_params &p = rt_->params();
    const size_t N = (size_t)p.img_h * p.img_w;
    if ((size_t)mask.rows * mask.cols != cloud->points.size() || !d_pix_)
      throw std::runtime_error("maskCloud: call run() first with a cloud of H*W points");  // assert :239
    if (!((val == 1 && !dense) || (val == 255 && dense)))
      throw std::runtime_error("maskCloud: only (1, sparse) and (255, dense) are part of the hot path");
    sloam_ctx *c = rt_->ctx();
    sloam_b200::DevBuf d_mask(c, N), d_tree(c, sizeof(sloam_point) * N), d_ground(c, sizeof(sloam_point) * N), d_n(c, 4);
    d_mask.upload(mask.data.data(), N);
    rt_->check(sloam_b200_mask_cloud_dev(c, 1, d_points_->as<sloam_point>(), d_pix_->as<int32_t>(), d_mask.as<uint8_t>(),
                                         d_tree.as<sloam_point>(), d_ground.as<sloam_point>(), d_n.as<int32_t>()));
    if (!outCloud) outCloud.reset(new Cloud());
    if (dense) {
      outCloud->points.resize(N);
      d_tree.download(outCloud->points.data(), sizeof(sloam_point) * N);
      outCloud->width = p.img_w; outCloud->height = p.img_h; outCloud->is_dense = true;
    } else {
      int32_t n = 0;
      d_n.download(&n, 4);
      outCloud->points.resize(n);
      d_ground.download(outCloud->points.data(), sizeof(sloam_point) * (size_t)n);
      outCloud->width = n; outCloud->height = 1; outCloud->is_dense = false;
    }
  }
  std::vector<float> range_image;  // what _doProjection returns

 private:
  std::unique_ptr<sloam_b200::Runtime> rt_;
  std::unique_ptr<sloam_b200::DevBuf> d_points_, d_pix_;
  Mask labels_;
};
}  // namespace seg

// ------------------------------------------------------------------ MapManager
// sloam/include/core/mapManager.h, sloam/src/core/mapManager.cpp:8-71 with the landmark map
// (models, hit counts, kNN index) resident on the device.  It lives in the context of the
// Runtime it is given; one map per context.
class MapManager {
 public:
  explicit MapManager(std::shared_ptr<sloam_b200::Runtime> rt, int capacity = 1 << 16) : rt_(std::move(rt)) {
    rt_->check(sloam_b200_map_init(rt_->ctx(), capacity));
    capacity_ = capacity;
  }
  ~MapManager() { sloam_b200_map_free(rt_->ctx()); }
  MapManager(const MapManager &) = delete;
  MapManager &operator=(const MapManager &) = delete;

  // mapManager.cpp:41-71: kNN(100) around (pose.x, pose.y, 1) + the "last 200 landmarks" filter
  void getSubmap(const SE3 &pose, std::vector<Cylinder> &submap) {
    sloam_ctx *c = rt_->ctx();
    const int M = rt_->params().max_map_models;
    sloam_b200::DevBuf d_pose(c, sizeof(sloam_pose)), d_sub(c, sizeof(sloam_cylinder) * (size_t)M), d_n(c, 4);
    d_pose.upload(&pose.abi(), sizeof(sloam_pose));
    rt_->check(sloam_b200_map_get_submap_dev(c, d_pose.as<sloam_pose>(), d_sub.as<sloam_cylinder>(), d_n.as<int32_t>()));
    int32_t n = 0;
    d_n.download(&n, 4);
    std::vector<sloam_cylinder> flat((size_t)n);
    if (n > 0) d_sub.download(flat.data(), sizeof(sloam_cylinder) * (size_t)n);
    for (const sloam_cylinder &m : flat) submap.push_back(from_abi(m));
  }
  // mapManager.cpp:8-28: matched observations overwrite their landmark and count a hit, the
  // others are appended
  void updateMap(std::vector<Cylinder> &obs_tms, const std::vector<int> &cyl_matches) {
    sloam_ctx *c = rt_->ctx();
    const size_t T = (size_t)rt_->params().max_trees, n = obs_tms.size();
    if (n > T || cyl_matches.size() != n) throw std::runtime_error("MapManager::updateMap: bad sizes");
    std::vector<sloam_cylinder> tm(T);
    std::vector<int32_t> ids(T, 0), matches(T, -1);
    for (size_t i = 0; i < n; ++i) {
      for (int a = 0; a < 3; ++a) { tm[i].root[a] = obs_tms[i].model.root[a]; tm[i].ray[a] = obs_tms[i].model.ray[a]; }
      tm[i].radius = obs_tms[i].model.radius;
      ids[i] = (int32_t)obs_tms[i].id;
      matches[i] = cyl_matches[i];
    }
    sloam_kf_result r{};
    r.status = SLOAM_KF_OK; r.success = 1; r.n_landmarks = (int32_t)n;
    sloam_b200::DevBuf d_r(c, sizeof r), d_tm(c, sizeof(sloam_cylinder) * T), d_id(c, 4 * T), d_m(c, 4 * T);
    d_r.upload(&r, sizeof r); d_tm.upload(tm.data(), sizeof(sloam_cylinder) * T);
    d_id.upload(ids.data(), 4 * T); d_m.upload(matches.data(), 4 * T);
    rt_->check(sloam_b200_map_update_dev(c, d_r.as<sloam_kf_result>(), d_tm.as<sloam_cylinder>(), d_id.as<int32_t>(),
                                         d_m.as<int32_t>()));
    rt_->check(sloam_b200_sync(c));
  }
  // mapManager.cpp:30-39: landmarks seen more than twice
  std::vector<Cylinder> getMap() {
    std::vector<sloam_cylinder> models((size_t)capacity_);
    std::vector<int32_t> hits((size_t)capacity_);
    const int n = sloam_b200_map_dump_host(rt_->ctx(), models.data(), hits.data(), capacity_);
    std::vector<Cylinder> map;
    for (int i = 0; i < n && i < capacity_; ++i)
      if (hits[i] > 2) map.push_back(from_abi(models[i]));
    return map;
  }
  int size() { return sloam_b200_map_dump_host(rt_->ctx(), nullptr, nullptr, 0); }

  static Cylinder from_abi(const sloam_cylinder &m) {
    Cylinder cyl;
    for (int a = 0; a < 3; ++a) { cyl.model.root[a] = m.root[a]; cyl.model.ray[a] = m.ray[a]; }
    cyl.model.radius = m.radius;
    cyl.isValid = true;
    return cyl;
  }

 private:
  std::shared_ptr<sloam_b200::Runtime> rt_;
  int capacity_ = 0;
};

// ------------------------------------------------------------------ SLOAMNode::run
// The per-keyframe call sequence of sloamNode.cpp:186-282 without ROS: getSubmap -> segmentation
// mask (supplied by the caller) -> maskCloud x 2 -> computeGraph -> RunSloam -> updateMap, as ONE
// device call (sloam_b200_sequence_step_host).  firstScan_, prevGPlanes_ and the map stay on the
// device between calls.
class SLOAMNodeCore {
 public:
  SLOAMNodeCore(const FeatureModelParams &fm, const sloam_b200::HostConfig &hc, int map_capacity = 1 << 16)
      : rt_(new sloam_b200::Runtime(fm, hc)) {
    rt_->check(sloam_b200_map_init(rt_->ctx(), map_capacity));
    const size_t T = (size_t)rt_->params().max_trees;
    matches_.assign(T, -1); tm_.resize(T); tm_id_.assign(T, 0);
  }
  ~SLOAMNodeCore() { sloam_b200_map_free(rt_->ctx()); }
  SLOAMNodeCore(const SLOAMNodeCore &) = delete;
  SLOAMNodeCore &operator=(const SLOAMNodeCore &) = delete;

  // bool SLOAMNode::run(const SE3 initialGuess, const SE3 prevKeyPose, CloudT::Ptr cloud, stamp, SE3 &outPose);
  // rMask is what segmentator_->run(cloud, rMask) returns (sloamNode.cpp:208-209).
  bool run(const SE3 &initialGuess, const SE3 &prevKeyPose, const CloudT::Ptr &cloud, const Mask &rMask, SE3 &outPose) {
    const sloam_params &p = rt_->params();
    const size_t N = (size_t)p.img_h * p.img_w;
    if (cloud->points.size() != N || (size_t)rMask.rows * rMask.cols != N)
      throw std::runtime_error("SLOAMNodeCore::run: cloud and mask must hold H*W entries");
    const SE3 poseEstimate = compose(prevKeyPose, initialGuess);  // :192
    rt_->check(sloam_b200_sequence_step_host(rt_->ctx(), reinterpret_cast<const sloam_point *>(cloud->points.data()),
                                             rMask.data.data(), &poseEstimate.abi(), &last_, matches_.data(), tm_.data(),
                                             tm_id_.data()));
    if (last_.success) outPose = SE3(last_.T_Map_Curr);  // :238-244
    return last_.success != 0;
  }
  const sloam_kf_result &lastResult() const { return last_; }
  // landmarks of the last keyframe in the map frame and their map matches (sloamOut.tm / .matches)
  std::vector<Cylinder> lastLandmarks() const {
    std::vector<Cylinder> out;
    const int n = SLOAM_KF_CODE(last_.status) == SLOAM_KF_OK || SLOAM_KF_CODE(last_.status) == SLOAM_KF_NOT_CONVERGED ? last_.n_landmarks : 0;
    for (int i = 0; i < n; ++i) { Cylinder c = MapManager::from_abi(tm_[(size_t)i]); c.id = (size_t)tm_id_[(size_t)i]; out.push_back(c); }
    return out;
  }
  std::vector<int> lastMatches() const { return std::vector<int>(matches_.begin(), matches_.begin() + std::max(0, (int)lastLandmarks().size())); }
  int mapSize() { return sloam_b200_map_dump_host(rt_->ctx(), nullptr, nullptr, 0); }

  // Sophus: prevKeyPose * initialGuess
  static SE3 compose(const SE3 &a, const SE3 &b) {
    const double *qa = a.unit_quaternion(), *qb = b.unit_quaternion();  // x y z w
    SE3 r;
    r.setQuaternion(qa[3] * qb[3] - qa[0] * qb[0] - qa[1] * qb[1] - qa[2] * qb[2],
                    qa[3] * qb[0] + qa[0] * qb[3] + qa[1] * qb[2] - qa[2] * qb[1],
                    qa[3] * qb[1] - qa[0] * qb[2] + qa[1] * qb[3] + qa[2] * qb[0],
                    qa[3] * qb[2] + qa[0] * qb[1] - qa[1] * qb[0] + qa[2] * qb[3]);
    r.translation() = a * b.translation();
    return r;
  }

 private:
  std::unique_ptr<sloam_b200::Runtime> rt_;
  sloam_kf_result last_{};
  std::vector<int32_t> matches_, tm_id_;
  std::vector<sloam_cylinder> tm_;
};

