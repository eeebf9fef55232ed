/* oracle/orc_math.cpp -- Sophus/Eigen pose helpers restated (test infrastructure).
 * Double-precision geometry: operation order inside these is not bit-pinned
 * (tolerances on everything that flows through them are 1e-5 or looser). */
#include "orc.h"

namespace orc {

Quat quat_mul(const Quat &a, const Quat &b) {
  /* Eigen quat product (Quaternion.h, internal::quat_product) */
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
          a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
          a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

Quat quat_normalized(const Quat &a) {
  const double n = std::sqrt(a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z);
  return {a.w / n, a.x / n, a.y / n, a.z / n};
}

V3 quat_rotate(const Quat &q, const V3 &v) {
  /* Eigen QuaternionBase::_transformVector: uv = 2 (q.vec x v);
   * v + w uv + q.vec x uv.  Used by Sophus SO3 * point and by the Ceres cost
   * functors CylinderCost/PlaneCost (cylinder.h:115, plane.h:92). */
  const V3 u{q.x, q.y, q.z};
  V3 uv = cross(u, v);
  uv = uv + uv;
  return v + q.w * uv + cross(u, uv);
}

void quat_to_matrix(const Quat &q, double R[3][3]) {
  /* Eigen QuaternionBase::toRotationMatrix */
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz;       R[0][2] = txz + twy;
  R[1][0] = txy + twz;       R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy;       R[2][1] = tyz + twx;       R[2][2] = 1 - (txx + tyy);
}

V3 se3_apply(const SE3 &T, const V3 &p) { return quat_rotate(T.q, p) + T.t; }

SE3 se3_inverse(const SE3 &T) {
  SE3 r;
  r.q = {T.q.w, -T.q.x, -T.q.y, -T.q.z};
  r.t = quat_rotate(r.q, {-T.t.x, -T.t.y, -T.t.z});
  return r;
}

SE3 se3_mul(const SE3 &a, const SE3 &b) {
  SE3 r;
  r.q = quat_normalized(quat_mul(a.q, b.q));
  r.t = quat_rotate(a.q, b.t) + a.t;
  return r;
}

void se3_matrix(const SE3 &T, double M[4][4]) {
  double R[3][3];
  quat_to_matrix(T.q, R);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) M[i][j] = R[i][j];
    M[i][3] = T.t[i];
    M[3][i] = 0;
  }
  M[3][3] = 1;
}

SE3 pose_from_abi(const sloam_pose &p) {
  SE3 T;
  T.t = {p.t[0], p.t[1], p.t[2]};
  T.q = {p.q[3], p.q[0], p.q[1], p.q[2]};
  return T;
}

sloam_pose pose_to_abi(const SE3 &T) {
  sloam_pose p;
  p.t[0] = T.t.x; p.t[1] = T.t.y; p.t[2] = T.t.z;
  p.q[0] = T.q.x; p.q[1] = T.q.y; p.q[2] = T.q.z; p.q[3] = T.q.w;
  return p;
}

}  // namespace orc
