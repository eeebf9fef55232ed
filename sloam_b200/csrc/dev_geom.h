// dev_geom.h -- quaternion / pose / plane helpers in double, host/device.
// The same operation order as the Eigen / Sophus expressions of the reference (cited per
// function).  Compiles for the host too (tests/hd_geom_check.cpp checks the callers in
// dev_plane.h against the oracle without a GPU); with nvcc the kernels include it through
// common.cuh.
#pragma once

#include <math.h>

#include "../../include/sloam_b200.h"

#ifndef SLOAM_HD_FN
#if defined(__CUDACC__)
#define SLOAM_HD_FN __host__ __device__ __forceinline__
#else
#define SLOAM_HD_FN inline
#endif
#endif

namespace sb {

SLOAM_HD_FN double dot3d(const double *a, const double *b) {
  return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]);
}

// ---- quaternion / pose (double) -------------------------------------------
SLOAM_HD_FN void q_rotate(const double q[4] /*x y z w*/, const double v[3], double out[3]) {
  // Eigen QuaternionBase::_transformVector
  double uvx = q[1] * v[2] - q[2] * v[1];
  double uvy = q[2] * v[0] - q[0] * v[2];
  double uvz = q[0] * v[1] - q[1] * v[0];
  uvx += uvx; uvy += uvy; uvz += uvz;
  out[0] = v[0] + q[3] * uvx + (q[1] * uvz - q[2] * uvy);
  out[1] = v[1] + q[3] * uvy + (q[2] * uvx - q[0] * uvz);
  out[2] = v[2] + q[3] * uvz + (q[0] * uvy - q[1] * uvx);
}
SLOAM_HD_FN void pose_apply(const sloam_pose &T, const double v[3], double out[3]) {
  q_rotate(T.q, v, out);
  out[0] += T.t[0]; out[1] += T.t[1]; out[2] += T.t[2];
}
SLOAM_HD_FN void q_to_matrix(const double q[4], double R[9]) {
  const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
  const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const double txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// plane' = (T^-1)^T plane : n' = R n, d' = d - n'.t   (Plane::project, plane.cpp:168)
SLOAM_HD_FN void plane_transform(const sloam_pose &T, const double pl[4], double out[4]) {
  double n[3] = {pl[0], pl[1], pl[2]}, rn[3];
  q_rotate(T.q, n, rn);
  out[0] = rn[0]; out[1] = rn[1]; out[2] = rn[2];
  out[3] = pl[3] - (rn[0] * T.t[0] + rn[1] * T.t[1] + rn[2] * T.t[2]);
}

}  // namespace sb
