// dev_plane.h -- scalar pieces of the ground-plane fit and acceptance test.
//   Plane::computeModel            sloam/src/objects/plane.cpp:96-128
//   acceptance in computeModels    sloam/src/core/sloam.cpp:394-409
// with the Eigen 3.3.7 algorithms behind them (JacobiSVD two-sided Jacobi on the
// QR-preconditioned 3x3, Quaternion::FromTwoVectors, Matrix3::eulerAngles(0,1,2)).
#pragma once

#include <float.h>

#include <math.h>

#include "dev_geom.h"

namespace sb {

struct JRot { double c, s; };

// Eigen JacobiRotation::makeJacobi(x, y, z)
SLOAM_HD_FN JRot make_jacobi(double x, double y, double z) {
  JRot r;
  const double deno = 2.0 * fabs(y);
  if (deno < DBL_MIN) { r.c = 1.0; r.s = 0.0; return r; }
  const double tau = (x - z) / deno;
  const double w = sqrt(tau * tau + 1.0);
  const double t = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
  const double sign_t = t > 0.0 ? 1.0 : -1.0;
  const double n = 1.0 / sqrt(t * t + 1.0);
  r.s = -sign_t * (y / fabs(y)) * fabs(t) * n;
  r.c = n;
  return r;
}

// x' = c x + s y ; y' = -s x + c y on n strided entries
SLOAM_HD_FN void rot_pair(double *x, int sx, double *y, int sy, int n, JRot j) {
  if (j.c == 1.0 && j.s == 0.0) return;
  for (int i = 0; i < n; ++i) {
    const double xi = x[i * sx], yi = y[i * sy];
    x[i * sx] = j.c * xi + j.s * yi;
    y[i * sy] = -j.s * xi + j.c * yi;
  }
}

// Steps 2-4 of Eigen's JacobiSVD::compute on a 3x3 work matrix Wm (row-major)
// with U initialised by the preconditioner; returns U.col(2).
SLOAM_HD_FN void jacobi_svd3_last_u(double Wm[9], double U[9], double out[3]) {
  const double considerAsZero = DBL_MIN, precision = 2.0 * DBL_EPSILON;
  double maxDiag = fmax(fabs(Wm[0]), fmax(fabs(Wm[4]), fabs(Wm[8])));
  bool finished = false;
  for (int sweep = 0; !finished && sweep < 1000; ++sweep) {
    finished = true;
    for (int p = 1; p < 3; ++p)
      for (int q = 0; q < p; ++q) {
        const double thr = fmax(considerAsZero, precision * maxDiag);
        if (fabs(Wm[p * 3 + q]) > thr || fabs(Wm[q * 3 + p]) > thr) {
          finished = false;
          double m00 = Wm[p * 3 + p], m01 = Wm[p * 3 + q], m10 = Wm[q * 3 + p], m11 = Wm[q * 3 + q];
          JRot r1;
          const double t = m00 + m11, d = m10 - m01;
          if (fabs(d) < DBL_MIN) { r1.s = 0.0; r1.c = 1.0; }
          else {
            const double u = t / d;
            const double tmp = sqrt(1.0 + u * u);
            r1.s = 1.0 / tmp; r1.c = u / tmp;
          }
          const double a0 = r1.c * m00 + r1.s * m10, a1 = r1.c * m01 + r1.s * m11;
          const double b1 = -r1.s * m01 + r1.c * m11;
          const JRot jr = make_jacobi(a0, a1, b1);
          const JRot jrt = {jr.c, -jr.s};
          const JRot jl = {r1.c * jrt.c - r1.s * jrt.s, r1.c * jrt.s + r1.s * jrt.c};
          rot_pair(&Wm[p * 3], 1, &Wm[q * 3], 1, 3, jl);   // W.applyOnTheLeft(p,q,j_left)
          rot_pair(&U[p], 3, &U[q], 3, 3, jl);             // U.applyOnTheRight(p,q,j_left^T)
          rot_pair(&Wm[p], 3, &Wm[q], 3, 3, jrt);          // W.applyOnTheRight(p,q,j_right)
          maxDiag = fmax(maxDiag, fmax(fabs(Wm[p * 3 + p]), fabs(Wm[q * 3 + q])));
        }
      }
  }
  double sv[3];
  for (int i = 0; i < 3; ++i) {
    const double a = Wm[i * 3 + i];
    sv[i] = fabs(a);
    if (a < 0.0) { U[i] = -U[i]; U[3 + i] = -U[3 + i]; U[6 + i] = -U[6 + i]; }
  }
  for (int i = 0; i < 3; ++i) {  // descending order by selection, swapping columns of U
    int pos = i;
    for (int j = i + 1; j < 3; ++j) if (sv[j] > sv[pos]) pos = j;
    if (sv[pos] == 0.0) break;
    if (pos != i) {
      const double ts = sv[i]; sv[i] = sv[pos]; sv[pos] = ts;
      for (int r = 0; r < 3; ++r) { const double tu = U[r * 3 + i]; U[r * 3 + i] = U[r * 3 + pos]; U[r * 3 + pos] = tu; }
    }
  }
  out[0] = U[2]; out[1] = U[5]; out[2] = U[8];
}

// angleCheck && heightCheck of sloam.cpp:402-408 for a valid plane.
SLOAM_HD_FN bool plane_accept(const sloam_pose &pose, const double plane[4], const double centroid[3],
                              double tol) {
  // normal = (poseEstimate^-1).matrix()^T * plane : rotate the normal into the map frame
  const double n0[3] = {plane[0], plane[1], plane[2]};
  double b[3];
  q_rotate(pose.q, n0, b);
  // Quat::FromTwoVectors((0,0,1), b)
  const double bn = sqrt(b[0] * b[0] + (b[1] * b[1] + b[2] * b[2]));
  const double v1[3] = {b[0] / bn, b[1] / bn, b[2] / bn};
  const double c = v1[2];
  bool angle_ok = false;
  if (c >= -1.0 + 1e-12) {
    // axis = (0,0,1) x v1
    const double ax = -v1[1], ay = v1[0], az = 0.0;
    const double s = sqrt((1.0 + c) * 2.0), invs = 1.0 / s;
    const double q[4] = {ax * invs, ay * invs, az * invs, s * 0.5};
    double m[9];
    q_to_matrix(q, m);
    // Matrix3d::eulerAngles(0,1,2)
    double r0 = atan2(m[5], m[8]);
    const double c2 = sqrt(m[0] * m[0] + m[1] * m[1]);
    double r1;
    if (r0 > 0.0) { r0 -= 3.14159265358979323846; r1 = atan2(-m[2], -c2); }
    else r1 = atan2(-m[2], c2);
    const double s1 = sin(r0), c1 = cos(r0);
    const double r2 = atan2(s1 * m[6] - c1 * m[3], c1 * m[4] - s1 * m[7]);
    const double a0 = -r0, a1 = -r1, a2 = -r2;
    const double PI = 3.14159265358979323846;
    angle_ok = (a0 < tol && a1 < tol && a2 < tol) ||
               (PI - fabs(a0) < tol && PI - fabs(a1) < tol && PI - fabs(a2) < tol);
  }
  double cm[3];
  pose_apply(pose, centroid, cm);
  const bool height_ok = cm[2] < pose.t[2];
  return angle_ok && height_ok;
}

}  // namespace sb
