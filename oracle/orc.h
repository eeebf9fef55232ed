/*
 * oracle/orc.h -- CPU restatement of the reference's per-keyframe hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product
 * (sloam_b200/) never links it and has no CPU fallback.
 *
 * It is a single-threaded C++ restatement of KumarRobotics/sloam
 *   sloam/src/core/sloam.cpp, sloam/src/objects/{plane,cylinder}.cpp,
 *   sloam/src/segmentation/trellis.cpp, inference.cpp:80-165,200-273,
 *   sloam/include/helpers/{definitions,utils}.h
 * plus the slices of PCL 1.10 / Eigen 3.3.7 / Ceres@206061a6 / Sophus those
 * files call (none of which is vendored or installed here; the reference
 * cannot be compiled in this image).  Each function cites the
 * reference file:line it follows.
 *
 * Parity pins (tests/test_oracle_golden.py): the four
 * {still,moving}_tree_{t0,t1}.pcd -> *_landmarks_* fixture pairs of the
 * reference (stage a6/a7) reproduce exactly; the restated gtest assertions
 * of sloam/src/tests (all .cpp files) pass.  The third-party arithmetic that no
 * reference fixture pins (PCL line RANSAC, Eigen JacobiSVD sign, Ceres LM) is
 * restated from the published algorithms: "parity unpinned" for those.
 */
#ifndef ORC_H
#define ORC_H

#include <cmath>
#include <cstdint>
#include <vector>

#include "../include/sloam_b200.h"

namespace orc {

/* PointT = pcl::PointXYZI (definitions.h:42); default-constructed = zeros. */
struct Pt {
  float x = 0.f, y = 0.f, z = 0.f, intensity = 0.f;
};
using Cloud = std::vector<Pt>;

struct V3 {
  double x = 0, y = 0, z = 0;
  double &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
/* Eigen's unrolled 3-term reduction is a0 + (a1 + a2) (Redux.h novec unroller). */
inline double dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct Quat {
  double w = 1, x = 0, y = 0, z = 0;
};
/* Sophus::SE3d (definitions.h:33). */
struct SE3 {
  Quat q;
  V3 t;
};
Quat quat_mul(const Quat &a, const Quat &b);
Quat quat_normalized(const Quat &a);
V3 quat_rotate(const Quat &q, const V3 &v); /* Eigen Quaternion::_transformVector */
void quat_to_matrix(const Quat &q, double R[3][3]);
V3 se3_apply(const SE3 &T, const V3 &p);
SE3 se3_inverse(const SE3 &T);
SE3 se3_mul(const SE3 &a, const SE3 &b);
void se3_matrix(const SE3 &T, double M[4][4]);
SE3 pose_from_abi(const sloam_pose &p);
sloam_pose pose_to_abi(const SE3 &T);

/* TreeVertex (definitions.h:56-66). */
struct TreeVertex {
  int treeId = 0;
  int beam = 0;
  int prevVertexSize = 0;
  double radius = 0;
  bool isValid = false;
  Pt coords;
  Cloud points;
  int row = -1; /* not in the reference: scan line, for the flattened ABI */
};
using Landmarks = std::vector<std::vector<TreeVertex>>;

struct PlaneParameters { /* plane.h:14-19 */
  double plane[4] = {0, 0, 0, 0};
  V3 centroid;
};
struct Plane { /* plane.h:21-32 + semanticObject.h */
  bool isValid = false;
  Cloud features;
  PlaneParameters model;
  int n_cell = 0, n_kept = 0; /* diagnostics for the flattened ABI */
};
struct CylinderParameters { /* cylinder.h:14-23 */
  V3 root, ray;
  std::vector<double> radii;
  double radius = 0;
};
struct Cylinder {
  size_t id = 0;
  bool isValid = false;
  Cloud features;
  CylinderParameters model;
  /* diagnostics */
  int n_inliers = 0, best_hypothesis = -1, n_hypotheses = 0, n_refit_inliers = 0;
  int plane_index = -1;
};

struct Options {
  sloam_params p;
  bool use_libm = false; /* true: glibc atan2f/asinf like the reference;
                            false: include/sloam_b200_detmath.h (what the GPU matches) */
};

/* ---- stage a1/a2: inference.cpp:80-165, 200-273 ---- */
void project(const Options &o, const Pt *pts, int n, int32_t *pix, float *range_image);
void mask_cloud(const Options &o, const Pt *pts, int n, const int32_t *pix,
                const uint8_t *mask, Cloud &tree, Cloud &ground);

/* ---- stage a3-a5: sloam.cpp:330-412, plane.cpp ---- */
void bin_ground_points(const Options &o, const V3 &origin, const Pt *pts, int n,
                       std::vector<Cloud> &cells /* [RB*TB] */, std::vector<int> &n_cell);
Plane make_plane(const Cloud &points, int numGroundFeatures);
double plane_distance_point(const PlaneParameters &m, const Pt &p);    /* plane.cpp:136-151 */
double plane_distance_model(const PlaneParameters &a, const PlaneParameters &b); /* :131-134 */
void plane_project(Plane &p, const SE3 &tf);                             /* :153-176 */
bool plane_accept(const Options &o, const SE3 &poseEstimate, const Plane &p); /* sloam.cpp:401-409 */
/* Eigen 3.3.7 JacobiSVD<MatrixXd>(3 x n, ThinU): third left singular vector. */
bool svd_smallest_left_vector(const std::vector<double> &A /* 3 x n col-major */, int n,
                              double out[3]);

/* ---- stage a6/a7: trellis.cpp ---- */
void find_clusters(const Options &o, const Pt *tree, int H, int W,
                   std::vector<uint32_t> &labels,
                   std::vector<std::vector<int>> &label_indices);
void compute_graph(const Options &o, const Pt *tree, int H, int W, Landmarks &landmarks);

/* ---- stage a8-a11: cylinder.cpp ---- */
Cylinder make_cylinder(const Options &o, const std::vector<TreeVertex> &vertices,
                       const Plane &gplane);
double cylinder_distance_model(const CylinderParameters &a, const CylinderParameters &b);
double cylinder_distance_point(const CylinderParameters &m, const Pt &p);
void cylinder_project(Cylinder &c, const SE3 &tf);
/* PCL sampling stream: draw t of a fresh model over n indices. */
void ransac_draw_table(int n, int n_draws, std::vector<int32_t> &pairs /* 2*n_draws */);

/* ---- stage a12-a19: sloam.cpp ---- */
struct TreeMatch { V3 feature; CylinderParameters object; };
struct PlaneMatch { V3 feature; PlaneParameters object; };
struct LMSummary { int iterations = 0; int termination = -1; double initial_cost = 0, final_cost = 0; };

bool optimize_pose(const Options &o, const SE3 &poseEstimate, const std::vector<TreeMatch> &tm,
                   const std::vector<PlaneMatch> &gm, SE3 &tf, LMSummary *s);
bool two_step_optimize_pose(const Options &o, const SE3 &poseEstimate, bool optimTrees,
                            bool optimGround, const std::vector<TreeMatch> &tm,
                            const std::vector<PlaneMatch> &gm, SE3 &tf, LMSummary s[2]);

struct SloamInput {
  SE3 poseEstimate;
  Cloud groundCloud;
  std::vector<Cylinder> mapModels;
  Landmarks landmarks;
};
struct SloamOutput {
  std::vector<int> matches;
  std::vector<Cylinder> tm;
  SE3 T_Map_Curr, T_Delta;
};

class Sloam { /* sloam::sloam, sloam.h:57-107 */
 public:
  explicit Sloam(const Options &o) : o_(o) {}
  bool RunSloam(SloamInput &in, SloamOutput &out);
  void computeModels(SloamInput &in, std::vector<Cylinder> &landmarks, std::vector<Plane> &planes);
  /* state the reference keeps across calls (sloam.h:99-106), made explicit */
  bool firstScan = true;
  std::vector<Plane> prevGPlanes;
  /* diagnostics of the last call */
  sloam_kf_result last{};
  std::vector<Plane> cellPlanes;      /* all RB*TB cells */
  std::vector<char> cellAccepted;
  std::vector<Cylinder> treeModels;   /* one per input landmark */
  std::vector<TreeMatch> lastTreeMatches;   /* match lists handed to the optimiser */
  std::vector<PlaneMatch> lastPlaneMatches;
 private:
  Options o_;
};

}  // namespace orc
#endif
