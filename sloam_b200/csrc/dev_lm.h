// dev_lm.h -- pose optimisation core: residuals, analytic Jacobians, Huber
// robustification and the Levenberg-Marquardt trust-region loop.
//
// Replaces the Ceres problems of sloam::OptimizePose / OptimizeXYYaw /
// OptimizeZRollPitch (sloam/src/core/sloam.cpp:55-255) and the cost functors
// CylinderCost / XYYawCylinderCost (sloam/include/objects/cylinder.h:42-133),
// PlaneCost / ZRollPitchPlaneCost (sloam/include/objects/plane.h:34-117).
//
// Same algorithm as Ceres' TrustRegionMinimizer + LevenbergMarquardtStrategy
// with the configuration the reference uses: Huber(0.1)
// corrector (sqrt(rho') scaling), Jacobi column scaling fixed at iteration 0,
// D^2 = clamp(diag(J^T J), 1e-6, 1e32) / radius, step-quality radius update,
// parameter / function tolerance tests on the candidate (which is then
// discarded), gradient tolerance on the accepted point, <= 50 iterations.
// Differences, by design: the damped linear system is solved from the 6x6 (or
// 3x3) normal equations accumulated by a fused residual/Jacobian kernel
// (Cholesky) instead of a dense QR of the stacked Jacobian, and Jacobians are
// closed-form instead of jets.  Both agree with the QR/jet formulation to
// rounding; the parity bar is 1e-5 m / 1e-5 rad on the pose.
//
// Pure scalar code: compiled for the device (k6_pose.cu) and, for the
// host-side unit test tests/hd_check.cpp, for the host.
#pragma once

#include <float.h>
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

enum LMMode { LM_JOINT = 0, LM_XYYAW = 1, LM_ZROLLPITCH = 2 };

SLOAM_HD_FN void cross3(const double a[3], const double b[3], double o[3]) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}

// lp = q * cp + t (Eigen quaternion-vector product, q not assumed unit) and the
// 3x6 derivative of lp in the tangent space [dq(3), t(3)] of
// EigenQuaternionParameterization (x = [qx qy qz qw | tx ty tz]).
SLOAM_HD_FN void transform_joint(const double x[7], const double cp[3], double lp[3], double D[3][6]) {
  const double u[3] = {x[0], x[1], x[2]}, w = x[3];
  double a[3], ua[3];
  cross3(u, cp, a);
  cross3(u, a, ua);
  for (int i = 0; i < 3; ++i) lp[i] = cp[i] + 2.0 * w * a[i] + 2.0 * ua[i] + x[4 + i];
  if (!D) return;
  // d lp / d q (3x4): columns x,y,z then w
  double dq[3][4];
  for (int j = 0; j < 3; ++j) {
    double e[3] = {0, 0, 0}, ev[3], ea[3], uev[3];
    e[j] = 1.0;
    cross3(e, cp, ev);   // d a / d u_j
    cross3(e, a, ea);
    cross3(u, ev, uev);
    for (int i = 0; i < 3; ++i) dq[i][j] = 2.0 * w * ev[i] + 2.0 * ea[i] + 2.0 * uev[i];
  }
  for (int i = 0; i < 3; ++i) dq[i][3] = 2.0 * a[i];
  // Plus jacobian (4x3): rows x,y,z,w
  const double P[4][3] = {{x[3], x[2], -x[1]}, {-x[2], x[3], x[0]}, {x[1], -x[0], x[3]}, {-x[0], -x[1], -x[2]}};
  for (int i = 0; i < 3; ++i) {
    for (int c = 0; c < 3; ++c)
      D[i][c] = dq[i][0] * P[0][c] + dq[i][1] * P[1][c] + dq[i][2] * P[2][c] + dq[i][3] * P[3][c];
    for (int c = 0; c < 3; ++c) D[i][3 + c] = (i == c) ? 1.0 : 0.0;
  }
}

// lp = AngleAxisRotatePoint(aa, cp) + t (ceres/rotation.h) with x = [t | aa];
// D = d lp / d x (3x6), exact derivative of whichever branch is taken.
// Everything that depends on the pose only (theta, sin, cos, the unit axis and its
// derivatives) is prepared once per evaluation (aa_prepare), not once per residual row;
// `jmask` selects the angle-axis columns that the parameter subset actually uses.
struct AaPre {
  bool big;          // theta^2 > eps branch of AngleAxisRotatePoint
  double ct, st, ti;
  double w[3];       // unit axis
  double dw[3][3];   // dw[j][i] = d w_i / d aa_j
};

SLOAM_HD_FN void aa_prepare(const double x[6], AaPre &P) {
  const double *aa = x + 3;
  const double theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  P.big = theta2 > DBL_EPSILON;
  P.ct = 1.0; P.st = 0.0; P.ti = 0.0;
  for (int i = 0; i < 3; ++i) { P.w[i] = 0.0; for (int j = 0; j < 3; ++j) P.dw[j][i] = 0.0; }
  if (P.big) {
    const double theta = sqrt(theta2);
    P.ct = cos(theta); P.st = sin(theta); P.ti = 1.0 / theta;
    for (int i = 0; i < 3; ++i) P.w[i] = aa[i] * P.ti;
    //   d theta / d aa_j = w_j ;  d w_i / d aa_j = (delta_ij - w_i w_j) / theta
    for (int j = 0; j < 3; ++j)
      for (int i = 0; i < 3; ++i) P.dw[j][i] = ((i == j ? 1.0 : 0.0) - P.w[i] * P.w[j]) * P.ti;
  }
}

SLOAM_HD_FN void transform_aa(const double x[6], const AaPre &P, const double cp[3], double lp[3],
                              double D[3][6], unsigned jmask) {
  const double *aa = x + 3;
  if (D)
    for (int i = 0; i < 3; ++i)
      for (int c = 0; c < 3; ++c) D[i][c] = (i == c) ? 1.0 : 0.0;
  if (P.big) {
    const double ct = P.ct, st = P.st;
    const double *w = P.w;
    double wxp[3];
    cross3(w, cp, wxp);
    const double wp = w[0] * cp[0] + w[1] * cp[1] + w[2] * cp[2];
    const double tmp = wp * (1.0 - ct);
    for (int i = 0; i < 3; ++i) lp[i] = cp[i] * ct + wxp[i] * st + w[i] * tmp + x[i];
    if (!D) return;
    // chain rule through theta(aa) and w(aa)
    for (int j = 0; j < 3; ++j) {
      if (!((jmask >> j) & 1u)) { D[0][3 + j] = D[1][3 + j] = D[2][3 + j] = 0.0; continue; }
      const double dth = w[j];
      const double *dw = P.dw[j];
      double dwxp[3];
      cross3(dw, cp, dwxp);
      const double dwp = dw[0] * cp[0] + dw[1] * cp[1] + dw[2] * cp[2];
      const double dtmp = dwp * (1.0 - ct) + wp * st * dth;
      for (int i = 0; i < 3; ++i)
        D[i][3 + j] = -cp[i] * st * dth + dwxp[i] * st + wxp[i] * ct * dth + dw[i] * tmp + w[i] * dtmp;
    }
  } else {
    double axp[3];
    cross3(aa, cp, axp);
    for (int i = 0; i < 3; ++i) lp[i] = cp[i] + axp[i] + x[i];
    if (!D) return;
    for (int j = 0; j < 3; ++j) {
      double e[3] = {0, 0, 0}, ev[3];
      e[j] = 1.0;
      cross3(e, cp, ev);
      for (int i = 0; i < 3; ++i) D[i][3 + j] = ev[i];
    }
  }
}

// cylinder residual ||pp - lp|| - radius and its gradient w.r.t. lp
SLOAM_HD_FN double cylinder_res(const double lp[3], const sloam_cylinder &m, double g[3]) {
  const double d[3] = {lp[0] - m.root[0], lp[1] - m.root[1], lp[2] - m.root[2]};
  const double aa = m.ray[0] * m.ray[0] + (m.ray[1] * m.ray[1] + m.ray[2] * m.ray[2]);
  const double s = (d[0] * m.ray[0] + (d[1] * m.ray[1] + d[2] * m.ray[2])) / aa;
  const double e[3] = {d[0] - s * m.ray[0], d[1] - s * m.ray[1], d[2] - s * m.ray[2]};
  const double ne = sqrt(e[0] * e[0] + (e[1] * e[1] + e[2] * e[2]));
  if (g) { const double inv = 1.0 / ne; g[0] = e[0] * inv; g[1] = e[1] * inv; g[2] = e[2] * inv; }
  return ne - m.radius;
}

// plane residual |n.lp + d| / ||n|| and its gradient (d|v|/dv = +1 at v = 0, like ceres::abs(Jet))
SLOAM_HD_FN double plane_res(const double lp[3], const sloam_plane &m, double g[3]) {
  const double nn = sqrt(m.plane[0] * m.plane[0] + (m.plane[1] * m.plane[1] + m.plane[2] * m.plane[2]));
  const double v = m.plane[0] * lp[0] + m.plane[1] * lp[1] + m.plane[2] * lp[2] + m.plane[3];
  const double sg = v < 0.0 ? -1.0 : 1.0;
  const double inv = 1.0 / nn;
  if (g) { g[0] = sg * m.plane[0] * inv; g[1] = sg * m.plane[1] * inv; g[2] = sg * m.plane[2] * inv; }
  return fabs(v) * inv;
}

// One residual: raw value and tangent-space Jacobian row J[n] (n = 6 joint, 3 subset).
// `pre` = aa_prepare(x) for the angle-axis modes (unused for LM_JOINT).
SLOAM_HD_FN double residual_row(int mode, const double *x, const AaPre &pre, const double feat[3],
                                const sloam_cylinder *cyl, const sloam_plane *pl, double *J) {
  double lp[3], D[3][6], g[3];
  if (mode == LM_JOINT) transform_joint(x, feat, lp, J ? D : nullptr);
  else transform_aa(x, pre, feat, lp, J ? D : nullptr, mode == LM_XYYAW ? 4u : 3u);  // aa_z | aa_x, aa_y
  const double r = cyl ? cylinder_res(lp, *cyl, J ? g : nullptr) : plane_res(lp, *pl, J ? g : nullptr);
  if (J) {
    if (mode == LM_JOINT) {
      for (int c = 0; c < 6; ++c) J[c] = g[0] * D[0][c] + g[1] * D[1][c] + g[2] * D[2][c];
    } else {
      const int f0 = mode == LM_XYYAW ? 0 : 2, f1 = mode == LM_XYYAW ? 1 : 3, f2 = mode == LM_XYYAW ? 5 : 4;
      J[0] = g[0] * D[0][f0] + g[1] * D[1][f0] + g[2] * D[2][f0];
      J[1] = g[0] * D[0][f1] + g[1] * D[1][f1] + g[2] * D[2][f1];
      J[2] = g[0] * D[0][f2] + g[1] * D[1][f2] + g[2] * D[2][f2];
    }
  }
  return r;
}

// HuberLoss(a) + Corrector for rho'' <= 0: returns 0.5 * rho(s) and the sqrt(rho') scale
SLOAM_HD_FN double huber(double r, double a, double *scale) {
  const double s = r * r, b = a * a;
  if (s > b) {
    const double rr = sqrt(s);
    const double rho1 = fmax(DBL_MIN, a / rr);
    *scale = sqrt(rho1);
    return 0.5 * (2.0 * a * rr - b);
  }
  *scale = 1.0;
  return 0.5 * s;
}

// x (+) delta in the tangent space of the mode
SLOAM_HD_FN void lm_plus(int mode, const double *x, const double *d, double *out) {
  if (mode == LM_JOINT) {
    const double nd = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (nd > 0.0) {  // EigenQuaternionParameterization::Plus: q_delta * q
      const double s = sin(nd) / nd, cw = cos(nd);
      const double ax = s * d[0], ay = s * d[1], az = s * d[2];
      const double bx = x[0], by = x[1], bz = x[2], bw = x[3];
      out[3] = cw * bw - ax * bx - ay * by - az * bz;
      out[0] = cw * bx + ax * bw + ay * bz - az * by;
      out[1] = cw * by + ay * bw + az * bx - ax * bz;
      out[2] = cw * bz + az * bw + ax * by - ay * bx;
    } else {
      for (int i = 0; i < 4; ++i) out[i] = x[i];
    }
    for (int i = 0; i < 3; ++i) out[4 + i] = x[4 + i] + d[3 + i];
  } else {  // SubsetParameterization
    for (int i = 0; i < 6; ++i) out[i] = x[i];
    const int f0 = mode == LM_XYYAW ? 0 : 2, f1 = mode == LM_XYYAW ? 1 : 3, f2 = mode == LM_XYYAW ? 5 : 4;
    out[f0] += d[0]; out[f1] += d[1]; out[f2] += d[2];
  }
}

// Work area of one trust-region solve.  Every array of the state machine lives in ONE
// object on purpose: as separate locals nvcc 12.9 gave the normal matrix and the solver's
// scratch the same local-memory slot (its stack colouring starts a slot's lifetime at the
// first use and got the loop wrong), which corrupted the matrix whenever a step was
// rejected and the matrix had to be reused.  On the device the object sits in shared
// memory, one per warp (k6_pose.cu); on the host it is a local of the caller.
struct LMWork {
  double xc[7], cand[7], tmp[7];
  double A[21], g[6];    // linearisation at xc (A Jacobi-scaled in place)
  double A2[21], g2[6];  // linearisation at the candidate (evaluated speculatively)
  double scale[6], diag[6], D2[6], rhs[6], step[6], delta[6], neg[6];
  double L[6][6], inv[6], z[6];
};

// Solve (A + diag(D2)) y = b for symmetric positive definite A (n <= 6, packed upper
// triangle row-major) by a square-root-free Cholesky (L D L^T): n divisions in total, the
// substitutions multiply by the stored reciprocals.  Returns false on breakdown.
SLOAM_HD_FN bool ldl_solve(int n, const double *A, const double *D2, const double *b, double *y, LMWork &w) {
  double(*L)[6] = w.L;
  double *inv = w.inv, *z = w.z;
  for (int j = 0; j < n; ++j) {
    const int dj = j * n - j * (j - 1) / 2;  // packed index of (j, j)
    // below the diagonal L holds l_jk * d_k (the products the later columns need)
    double d = A[dj] + D2[j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * inv[k];
    if (!(d > 0.0)) return false;
    inv[j] = 1.0 / d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[dj + (i - j)];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k] * inv[k];
      L[i][j] = s;  // = l_ij * d_j
    }
  }
  for (int i = 0; i < n; ++i) {  // L z = b, with l_ik = L[i][k] * inv[k]
    double s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i][k] * inv[k] * z[k];
    z[i] = s;
  }
  for (int i = n - 1; i >= 0; --i) {  // D L^T y = z
    double s = z[i] * inv[i];
    for (int k = i + 1; k < n; ++k) s -= L[k][i] * inv[i] * y[k];
    y[i] = s;
  }
  for (int i = 0; i < n; ++i)
    if (!isfinite(y[i])) return false;
  return true;
}

struct LMOut { int iterations; int termination; double initial_cost, final_cost; };

// The trust-region loop.  `ev(x, &cost, JtJ, g)` evaluates the robustified problem at x:
// cost = sum 0.5 rho(r^2), the packed upper triangle of J~^T J~ (n(n+1)/2) and
// g = J~^T r~.  x holds the start point and receives the last accepted point.
//
// Ceres evaluates a candidate cost-only and re-evaluates it with Jacobians once the step is
// accepted.  Here the candidate is linearised speculatively (into A2/g2) in the same pass,
// so an accepted step costs one evaluation round instead of two; the values are the ones
// the second evaluation would produce, so the iterates are unchanged.
// On the device every lane of a warp runs this loop with identical values (ev reduces over
// the lanes and broadcasts).
template <class Eval>
SLOAM_HD_FN LMOut lm_minimize(Eval &ev, LMWork &w, int mode, int n_res, int max_iterations, double *x) {
  LMOut out;
  out.iterations = 0; out.termination = -1; out.initial_cost = 0.0; out.final_cost = 0.0;
  const int n = mode == LM_JOINT ? 6 : 3, na = mode == LM_JOINT ? 7 : 6, np = n * (n + 1) / 2;
  if (n_res == 0) { out.termination = 0; return out; }  // reduced program empty: CONVERGENCE
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double min_relative_decrease = 1e-3, min_radius = 1e-32, max_radius = 1e16;
  const double min_diag = 1e-6, max_diag = 1e32;
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int invalid_steps = 0;
  double *xc = w.xc, *cand = w.cand, *tmp = w.tmp, *A = w.A, *g = w.g, *scale = w.scale, *diag = w.diag,
         *D2 = w.D2, *rhs = w.rhs, *step = w.step, *delta = w.delta, *neg = w.neg;
  for (int i = 0; i < na; ++i) xc[i] = x[i];
  for (int i = 0; i < 6; ++i) { scale[i] = 1.0; diag[i] = 0.0; }
  double cost = 0.0, gmax = 0.0, x_norm = 0.0;
  auto norm_of = [&](const double *v) { double s = 0.0; for (int i = 0; i < na; ++i) s += v[i] * v[i]; return sqrt(s); };
  auto diag_idx = [&](int i) { return i * n - i * (i - 1) / 2; };
  auto after_eval = [&](bool first) {
    if (first)
      for (int c = 0; c < n; ++c) scale[c] = 1.0 / (1.0 + sqrt(A[diag_idx(c)]));
    // Jacobi scaling: J~ <- J~ diag(scale)  =>  A <- S A S ; the gradient stays unscaled
    for (int i = 0; i < n; ++i)
      for (int j = i; j < n; ++j) A[diag_idx(i) + (j - i)] *= scale[i] * scale[j];
    for (int c = 0; c < n; ++c) neg[c] = -g[c];
    lm_plus(mode, xc, neg, tmp);
    gmax = 0.0;
    for (int i = 0; i < na; ++i) gmax = fmax(gmax, fabs(xc[i] - tmp[i]));
  };
  x_norm = norm_of(xc);
  ev(xc, &cost, A, g);
  after_eval(true);
  out.initial_cost = cost;
  int iteration = 0;
  bool last_successful = false;
  for (;;) {
    if (last_successful) for (int i = 0; i < na; ++i) x[i] = xc[i];
    if (iteration >= max_iterations) { out.termination = 1; break; }
    if (gmax <= gradient_tolerance) { out.termination = 0; break; }
    if (radius < min_radius) { out.termination = 0; break; }
    ++iteration;
    last_successful = false;
    if (!reuse_diagonal)
      for (int c = 0; c < n; ++c) diag[c] = fmin(fmax(A[diag_idx(c)], min_diag), max_diag);
    for (int c = 0; c < n; ++c) {
      const double lmd = sqrt(diag[c] / radius);  // lm_diagonal_
      D2[c] = lmd * lmd;
      rhs[c] = g[c] * scale[c];                   // J_s^T r
    }
    reuse_diagonal = true;
    const bool solved = ldl_solve(n, A, D2, rhs, step, w);
    bool step_valid = false;
    double model_cost_change = 0.0;
    if (solved) {
      for (int c = 0; c < n; ++c) step[c] = -step[c];
      // -(J step).(r + J step / 2) = -step.(J^T r) - step^T (J^T J) step / 2
      double lin = 0.0, quad = 0.0;
      for (int i = 0; i < n; ++i) {
        lin += step[i] * rhs[i];
        for (int j = 0; j < n; ++j) {
          const int a = i < j ? i : j, b = i < j ? j : i;
          quad += step[i] * A[diag_idx(a) + (b - a)] * step[j];
        }
      }
      model_cost_change = -lin - 0.5 * quad;
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) {
      if (++invalid_steps >= 5) { out.termination = 2; break; }
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      continue;
    }
    invalid_steps = 0;
    for (int c = 0; c < n; ++c) delta[c] = step[c] * scale[c];
    lm_plus(mode, xc, delta, cand);
    double cand_cost = 0.0;
    ev(cand, &cand_cost, w.A2, w.g2);
    {
      double s = 0.0;
      for (int i = 0; i < na; ++i) s += (xc[i] - cand[i]) * (xc[i] - cand[i]);
      if (sqrt(s) <= parameter_tolerance * (x_norm + parameter_tolerance)) { out.termination = 0; break; }
    }
    if (fabs(cost - cand_cost) <= function_tolerance * cost) { out.termination = 0; break; }
    const double relative_decrease = (cost - cand_cost) / model_cost_change;
    if (relative_decrease > min_relative_decrease) {
      for (int i = 0; i < na; ++i) xc[i] = cand[i];
      x_norm = norm_of(xc);
      cost = cand_cost;
      for (int i = 0; i < np; ++i) A[i] = w.A2[i];
      for (int i = 0; i < n; ++i) g[i] = w.g2[i];
      after_eval(false);
      last_successful = true;
      const double t = 2.0 * relative_decrease - 1.0;
      radius = radius / fmax(1.0 / 3.0, 1.0 - t * t * t);
      radius = fmin(max_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else {
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  out.iterations = iteration;
  out.final_cost = cost;
  return out;
}

// ceres::QuaternionToAngleAxis / AngleAxisToQuaternion, q = (w, x, y, z)
SLOAM_HD_FN void quat_to_angle_axis(const double q[4], double aa[3]) {
  const double s2 = q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  if (s2 > 0.0) {
    const double s = sqrt(s2), c = q[0];
    const double two_theta = 2.0 * ((c < 0.0) ? atan2(-s, -c) : atan2(s, c));
    const double k = two_theta / s;
    aa[0] = q[1] * k; aa[1] = q[2] * k; aa[2] = q[3] * k;
  } else {
    aa[0] = q[1] * 2.0; aa[1] = q[2] * 2.0; aa[2] = q[3] * 2.0;
  }
}
SLOAM_HD_FN void angle_axis_to_quat(const double aa[3], double q[4]) {
  const double t2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (t2 > 0.0) {
    const double t = sqrt(t2), h = t * 0.5, k = sin(h) / t;
    q[0] = cos(h); q[1] = aa[0] * k; q[2] = aa[1] * k; q[3] = aa[2] * k;
  } else {
    q[0] = 1.0; q[1] = aa[0] * 0.5; q[2] = aa[1] * 0.5; q[3] = aa[2] * 0.5;
  }
}

}  // namespace sb
