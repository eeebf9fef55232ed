/* oracle/orc_ground.cpp -- stages a3, a4, a5 (test infrastructure).
 * Restates sloam::binGroundPoints (sloam/src/core/sloam.cpp:330-386),
 * Plane::Plane / computeModel / distance / project (sloam/src/objects/plane.cpp),
 * computeCentroid / euclideanDist2D / pow_2 (include/helpers/utils.h:7-28),
 * the plane acceptance test of computeModels (sloam.cpp:394-412), and the
 * Eigen 3.3.7 pieces those call: JacobiSVD (ColPivHouseholderQR preconditioner
 * + two-sided Jacobi), Quaternion::FromTwoVectors, Matrix3::eulerAngles(0,1,2).
 * Eigen is not available in this image: "believed-upstream semantics",
 * parity of the SVD sign convention is unpinned by any reference fixture. */
#include <algorithm>
#include <cfloat>
#include <cstring>

#include "../include/sloam_b200_detmath.h"
#include "orc.h"

namespace orc {

namespace {
constexpr double PIDEF = 3.14159265; /* definitions.h:28 */

/* utils.h:7: inline float pow_2(const Scalar &x) { return x * x; } */
inline float pow_2(double x) { return (float)(x * x); }
/* utils.h:9-12 */
inline float euclideanDist2D(const Pt &a, const Pt &b) {
  return std::sqrt(pow_2(a.x - b.x) + pow_2(a.y - b.y));
}
}  // namespace

void bin_ground_points(const Options &o, const V3 &origin_d, const Pt *pts, int n,
                       std::vector<Cloud> &cells, std::vector<int> &n_cell) {
  const sloam_params &fm = o.p;
  const int RB = fm.groundRadiiBins, TB = fm.groundThetaBins;
  cells.assign((size_t)RB * TB, Cloud());
  Pt origin; /* sloam.cpp:333-336 */
  origin.x = (float)origin_d.x; origin.y = (float)origin_d.y; origin.z = (float)origin_d.z;
  for (int i = 0; i < n; ++i) { /* :339-360 */
    const Pt &p = pts[i];
    const double pointRadius = euclideanDist2D(origin, p);
    if (pointRadius < fm.maxGroundLidarDist && pointRadius > fm.minGroundLidarDist) {
      const float dy = p.y - origin.y, dx = p.x - origin.x;
      const double pointTheta =
          o.use_libm ? (double)std::atan2(dy, dx) : (double)sloam_det::det_atan2f(dy, dx);
      int rb = (int)std::floor(pointRadius / (fm.maxGroundLidarDist / (double)RB));
      int tb = (int)std::floor((PIDEF + pointTheta) / (2 * PIDEF / (double)TB));
      rb = std::max(0, std::min(rb, RB - 1));
      tb = std::max(0, std::min(tb, TB - 1));
      cells[(size_t)rb * TB + tb].push_back(p);
    }
  }
  n_cell.resize(cells.size());
  for (size_t c = 0; c < cells.size(); ++c) { /* :362-385 */
    Cloud &cell = cells[c];
    n_cell[c] = (int)cell.size();
    if (cell.empty()) continue;
    const double retainNum = 1 / fm.groundRetainThresh;
    if (retainNum < (double)cell.size()) {
      int bottomIdx = (int)((double)cell.size() / retainNum);
      /* B-5: thresh > 1 erases past end() in the reference; clamp. */
      bottomIdx = std::min(bottomIdx, (int)cell.size());
      /* std::sort like the reference (:377-380): not stable, so with exact z ties
       * libstdc++'s introsort decides which tied points survive and in which order
       * (SURVEY B-3); the device replays the same algorithm (csrc/dev_stdsort.h). */
      std::sort(cell.begin(), cell.end(), [](const Pt &a, const Pt &b) { return a.z < b.z; });
      cell.erase(cell.begin() + bottomIdx, cell.end());
    }
  }
}

/* ----------------------------------------------------------------------------
 * Eigen 3.3.7 JacobiSVD<MatrixXd>(A 3 x n, ComputeThinU|ComputeThinV)
 * (Eigen/src/SVD/JacobiSVD.h), returning matrixU().col(2).
 * -------------------------------------------------------------------------- */
namespace {
struct Rot { double c, s; }; /* Eigen::JacobiRotation */

/* JacobiRotation::makeJacobi(x, y, z) (Eigen/src/Jacobi/Jacobi.h) */
inline Rot make_jacobi(double x, double y, double z) {
  Rot r;
  const double deno = 2.0 * std::fabs(y);
  if (deno < DBL_MIN) { r.c = 1; r.s = 0; return r; }
  const double tau = (x - z) / deno;
  const double w = std::sqrt(tau * tau + 1.0);
  const double t = (tau > 0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
  const double sign_t = t > 0 ? 1.0 : -1.0;
  const double n = 1.0 / std::sqrt(t * t + 1.0);
  r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
  r.c = n;
  return r;
}
/* internal::apply_rotation_in_the_plane on two strided vectors */
inline void rot_apply(double *x, int sx, double *y, int sy, int n, Rot j) {
  if (j.c == 1 && j.s == 0) return;
  for (int i = 0; i < n; ++i) {
    const double xi = x[i * sx], yi = y[i * sy];
    x[i * sx] = j.c * xi + j.s * yi;
    y[i * sy] = -j.s * xi + j.c * yi;
  }
}
}  // namespace

bool svd_smallest_left_vector(const std::vector<double> &A_in, int n, double out[3]) {
  if (n < 3) return false; /* ThinU would have < 3 columns: UB in the reference */
  /* JacobiSVD::compute: scale = max |coeff| */
  double scale = 0;
  for (double v : A_in) scale = std::max(scale, std::fabs(v));
  if (scale == 0) scale = 1;
  double Wm[3][3]; /* work matrix */
  double U[3][3];
  if (n > 3) {
    /* qr_preconditioner_impl<ColPivHouseholderQR, MoreColsThanRows>:
     * QR with column pivoting of the adjoint (n x 3), W = R(0:3,0:3)^T,
     * U = column permutation. */
    std::vector<double> Q((size_t)n * 3); /* column-major n x 3 : Q[c*n + r] */
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < 3; ++c) Q[(size_t)c * n + r] = A_in[(size_t)r * 3 + c] / scale;
    const int rows = n, cols = 3;
    double normsUpd[3], normsDir[3];
    auto colnorm = [&](int c, int from) {
      double s = 0;
      for (int r = from; r < rows; ++r) s += Q[(size_t)c * n + r] * Q[(size_t)c * n + r];
      return std::sqrt(s);
    };
    for (int k = 0; k < cols; ++k) normsUpd[k] = normsDir[k] = colnorm(k, 0);
    const double downdate_thr = std::sqrt(DBL_EPSILON);
    int transp[3];
    for (int k = 0; k < 3; ++k) {
      int big = k;
      for (int j = k + 1; j < cols; ++j)
        if (normsUpd[j] > normsUpd[big]) big = j;
      transp[k] = big;
      if (k != big) {
        for (int r = 0; r < rows; ++r) std::swap(Q[(size_t)k * n + r], Q[(size_t)big * n + r]);
        std::swap(normsUpd[k], normsUpd[big]);
        std::swap(normsDir[k], normsDir[big]);
      }
      /* makeHouseholderInPlace on col k, rows k.. (Householder.h) */
      double *ck = &Q[(size_t)k * n];
      double tailSq = 0;
      for (int r = k + 1; r < rows; ++r) tailSq += ck[r] * ck[r];
      const double c0 = ck[k];
      double tau, beta;
      if (tailSq <= DBL_MIN) {
        tau = 0; beta = c0;
        for (int r = k + 1; r < rows; ++r) ck[r] = 0;
      } else {
        beta = std::sqrt(c0 * c0 + tailSq);
        if (c0 >= 0) beta = -beta;
        for (int r = k + 1; r < rows; ++r) ck[r] /= (c0 - beta);
        tau = (beta - c0) / beta;
      }
      ck[k] = beta;
      /* applyHouseholderOnTheLeft to the remaining columns */
      if (tau != 0) {
        for (int j = k + 1; j < cols; ++j) {
          double *cj = &Q[(size_t)j * n];
          double tmp = 0;
          for (int r = k + 1; r < rows; ++r) tmp += ck[r] * cj[r];
          tmp += cj[k];
          cj[k] -= tau * tmp;
          for (int r = k + 1; r < rows; ++r) cj[r] -= tau * ck[r] * tmp;
        }
      }
      /* LAPACK-style norm downdate (ColPivHouseholderQR.h) */
      for (int j = k + 1; j < cols; ++j) {
        if (normsUpd[j] != 0) {
          double temp = std::fabs(Q[(size_t)j * n + k]) / normsUpd[j];
          temp = (1.0 + temp) * (1.0 - temp);
          temp = temp < 0 ? 0 : temp;
          const double r2 = normsUpd[j] / normsDir[j];
          const double temp2 = temp * r2 * r2;
          if (temp2 <= downdate_thr) {
            normsDir[j] = colnorm(j, k + 1);
            normsUpd[j] = normsDir[j];
          } else {
            normsUpd[j] *= std::sqrt(temp);
          }
        }
      }
    }
    /* colsPermutation = product of the transpositions */
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < 3; ++k) std::swap(perm[k], perm[transp[k]]);
    /* P(i, j) = 1 iff i == perm[j] */
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) U[i][j] = (i == perm[j]) ? 1.0 : 0.0;
    /* W = upper-triangular R (3x3) adjoint */
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Wm[i][j] = (j <= i) ? Q[(size_t)i * n + j] : 0.0;
  } else {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        Wm[i][j] = A_in[(size_t)j * 3 + i] / scale;
        U[i][j] = (i == j) ? 1.0 : 0.0;
      }
  }
  /* step 2: Jacobi sweeps */
  const double considerAsZero = DBL_MIN, precision = 2.0 * DBL_EPSILON;
  double maxDiag = std::max(std::fabs(Wm[0][0]), std::max(std::fabs(Wm[1][1]), std::fabs(Wm[2][2])));
  bool finished = false;
  int guard = 0;
  while (!finished && guard++ < 1000) {
    finished = true;
    for (int p = 1; p < 3; ++p)
      for (int q = 0; q < p; ++q) {
        const double threshold = std::max(considerAsZero, precision * maxDiag);
        if (std::fabs(Wm[p][q]) > threshold || std::fabs(Wm[q][p]) > threshold) {
          finished = false;
          /* internal::real_2x2_jacobi_svd */
          double m00 = Wm[p][p], m01 = Wm[p][q], m10 = Wm[q][p], m11 = Wm[q][q];
          Rot rot1;
          const double t = m00 + m11, d = m10 - m01;
          if (std::fabs(d) < DBL_MIN) { rot1.s = 0; rot1.c = 1; }
          else {
            const double u = t / d;
            const double tmp = std::sqrt(1.0 + u * u);
            rot1.s = 1.0 / tmp; rot1.c = u / tmp;
          }
          /* m.applyOnTheLeft(0,1,rot1) */
          {
            const double a0 = rot1.c * m00 + rot1.s * m10, a1 = rot1.c * m01 + rot1.s * m11;
            const double b0 = -rot1.s * m00 + rot1.c * m10, b1 = -rot1.s * m01 + rot1.c * m11;
            m00 = a0; m01 = a1; m10 = b0; m11 = b1;
          }
          const Rot jr = make_jacobi(m00, m01, m11);
          /* j_left = rot1 * j_right.transpose() */
          const Rot jrt{jr.c, -jr.s};
          const Rot jl{rot1.c * jrt.c - rot1.s * jrt.s, rot1.c * jrt.s + rot1.s * jrt.c};
          /* W.applyOnTheLeft(p,q,j_left): rows p,q */
          rot_apply(&Wm[p][0], 1, &Wm[q][0], 1, 3, jl);
          /* U.applyOnTheRight(p,q,j_left.transpose()): cols p,q with j_left */
          rot_apply(&U[0][p], 3, &U[0][q], 3, 3, jl);
          /* W.applyOnTheRight(p,q,j_right): cols p,q with j_right.transpose() */
          rot_apply(&Wm[0][p], 3, &Wm[0][q], 3, 3, jrt);
          maxDiag = std::max(maxDiag, std::max(std::fabs(Wm[p][p]), std::fabs(Wm[q][q])));
        }
      }
  }
  /* step 3: positive singular values */
  double sv[3];
  for (int i = 0; i < 3; ++i) {
    const double a = Wm[i][i];
    sv[i] = std::fabs(a);
    if (a < 0)
      for (int r = 0; r < 3; ++r) U[r][i] = -U[r][i];
  }
  /* step 4: sort descending by selection with column swaps */
  for (int i = 0; i < 3; ++i) {
    int pos = i;
    for (int j = i + 1; j < 3; ++j)
      if (sv[j] > sv[pos]) pos = j;
    if (sv[pos] == 0) break;
    if (pos != i) {
      std::swap(sv[i], sv[pos]);
      for (int r = 0; r < 3; ++r) std::swap(U[r][i], U[r][pos]);
    }
  }
  out[0] = U[0][2]; out[1] = U[1][2]; out[2] = U[2][2];
  return true;
}

Plane make_plane(const Cloud &points, int numGroundFeatures) {
  Plane pl;
  pl.features = points; /* plane.cpp:6 */
  pl.n_kept = (int)points.size();
  if ((int)pl.features.size() < numGroundFeatures) { /* :7-10 */
    pl.isValid = false;
    return pl;
  }
  /* computeModel, :96-128; computeCentroid utils.h:14-28: float32 sequential
   * sums, divided by (double)n and stored back to float */
  Pt centroid;
  for (const Pt &p : pl.features) {
    centroid.x += p.x; centroid.y += p.y; centroid.z += p.z;
  }
  const double nn = (double)pl.features.size();
  centroid.x = (float)((double)centroid.x / nn);
  centroid.y = (float)((double)centroid.y / nn);
  centroid.z = (float)((double)centroid.z / nn);
  const int n = (int)pl.features.size();
  std::vector<double> A((size_t)3 * n);
  for (int i = 0; i < n; ++i) { /* :104-110: float subtraction, stored to double */
    A[(size_t)i * 3 + 0] = (double)(pl.features[i].x - centroid.x);
    A[(size_t)i * 3 + 1] = (double)(pl.features[i].y - centroid.y);
    A[(size_t)i * 3 + 2] = (double)(pl.features[i].z - centroid.z);
  }
  double nrm[3];
  if (!svd_smallest_left_vector(A, n, nrm)) {
    /* n < 3 with ThinU: out-of-range block in the reference (UB). Deviation:
     * the plane is declared invalid. */
    pl.isValid = false;
    return pl;
  }
  const double d = -(nrm[0] * centroid.x + nrm[1] * centroid.y + nrm[2] * centroid.z); /* :117 */
  pl.model.plane[0] = nrm[0]; pl.model.plane[1] = nrm[1]; pl.model.plane[2] = nrm[2];
  pl.model.plane[3] = d;
  pl.model.centroid = {centroid.x, centroid.y, centroid.z};
  pl.features.resize(numGroundFeatures); /* :14 */
  pl.isValid = true;
  return pl;
}

double plane_distance_point(const PlaneParameters &m, const Pt &p) { /* plane.cpp:136-151 */
  const double num = std::fabs(m.plane[0] * p.x + m.plane[1] * p.y + m.plane[2] * p.z + m.plane[3]);
  const V3 nv{m.plane[0], m.plane[1], m.plane[2]};
  return num / norm(nv);
}

double plane_distance_model(const PlaneParameters &a, const PlaneParameters &b) { /* :131-134 */
  return norm(a.centroid - b.centroid);
}

void plane_project(Plane &p, const SE3 &tf) { /* plane.cpp:153-176 */
  double M[4][4];
  se3_matrix(tf, M);
  for (Pt &f : p.features) { /* :158-165, pcl::transformPoint -> float result */
    if (!std::isfinite(f.x) || !std::isfinite(f.y) || !std::isfinite(f.z)) continue;
    /* pcl::transformPoint(point, Affine3d): double matrix times the float
     * coordinates, result rounded to float */
    const double x = f.x, y = f.y, z = f.z;
    Pt r = f;
    r.x = (float)(M[0][0] * x + M[0][1] * y + M[0][2] * z + M[0][3]);
    r.y = (float)(M[1][0] * x + M[1][1] * y + M[1][2] * z + M[1][3]);
    r.z = (float)(M[2][0] * x + M[2][1] * y + M[2][2] * z + M[2][3]);
    f = r;
  }
  /* :168 plane' = (tfm^-1)^T plane */
  const SE3 inv = se3_inverse(tf);
  double Mi[4][4];
  se3_matrix(inv, Mi);
  double np[4];
  for (int i = 0; i < 4; ++i) {
    np[i] = 0;
    for (int j = 0; j < 4; ++j) np[i] += Mi[j][i] * p.model.plane[j];
  }
  std::memcpy(p.model.plane, np, sizeof np);
  p.model.centroid = se3_apply(tf, p.model.centroid); /* :169 */
}

bool plane_accept(const Options &o, const SE3 &poseEstimate, const Plane &ground) {
  /* sloam.cpp:394-409 */
  const double tol = o.p.ground_angle_tol;
  /* tfm = poseEstimate.inverse().matrix().transpose(); normal = tfm * plane */
  double Mi[4][4];
  se3_matrix(se3_inverse(poseEstimate), Mi);
  double nrm4[4];
  for (int i = 0; i < 4; ++i) {
    nrm4[i] = 0;
    for (int j = 0; j < 4; ++j) nrm4[i] += Mi[j][i] * ground.model.plane[j];
  }
  /* Quat::FromTwoVectors((0,0,1), normal.head3) (Eigen Quaternion.h) */
  const V3 b{nrm4[0], nrm4[1], nrm4[2]};
  const V3 v0{0, 0, 1};
  const double bn = norm(b);
  const V3 v1{b.x / bn, b.y / bn, b.z / bn};
  const double c = dot(v1, v0);
  bool angleCheck;
  if (!(c >= -1.0 + 1e-12)) {
    /* antiparallel (or NaN) branch: a rotation by pi about an axis in the xy
     * plane; eulerAngles gives (-0, pi, r2): neither clause of :405-406 can
     * hold, so the test is false for every such input. */
    angleCheck = false;
  } else {
    const V3 axis = cross(v0, v1);
    const double s = std::sqrt((1.0 + c) * 2.0);
    const double invs = 1.0 / s;
    Quat q{s * 0.5, axis.x * invs, axis.y * invs, axis.z * invs};
    double m[3][3];
    quat_to_matrix(q, m);
    /* Matrix3d::eulerAngles(0,1,2), Eigen 3.3.7 EulerAngles.h: i=0 j=1 k=2, even */
    double r0 = std::atan2(m[1][2], m[2][2]);
    const double c2 = std::sqrt(m[0][0] * m[0][0] + m[0][1] * m[0][1]);
    double r1;
    if (r0 > 0) {
      r0 -= M_PI;
      r1 = std::atan2(-m[0][2], -c2);
    } else {
      r1 = std::atan2(-m[0][2], c2);
    }
    const double s1 = std::sin(r0), c1 = std::cos(r0);
    const double r2 = std::atan2(s1 * m[2][0] - c1 * m[1][0], c1 * m[1][1] - s1 * m[2][1]);
    const double a0 = -r0, a1 = -r1, a2 = -r2;
    /* :405-406 (no abs in the first clause: B-8) */
    angleCheck = (a0 < tol && a1 < tol && a2 < tol) ||
                 (M_PI - std::fabs(a0) < tol && M_PI - std::fabs(a1) < tol &&
                  M_PI - std::fabs(a2) < tol);
  }
  /* :408 ground should be under the robot */
  const bool heightCheck = se3_apply(poseEstimate, ground.model.centroid).z < poseEstimate.t.z;
  return ground.isValid && angleCheck && heightCheck;
}

}  // namespace orc
