/* oracle/orc_ceres.cpp -- stages a14..a17 (test infrastructure).
 * Restates the four Ceres cost functors (sloam/include/objects/cylinder.h:42-133,
 * plane.h:34-117), sloam::OptimizePose / TwoStepOptimizePose / OptimizeXYYaw /
 * OptimizeZRollPitch (sloam/src/core/sloam.cpp:33-255) and the parts of
 * Ceres @206061a6 they configure: AutoDiffCostFunction (forward-mode jets),
 * HuberLoss + Corrector, EigenQuaternionParameterization,
 * SubsetParameterization, rotation.h helpers, and the trust-region
 * Levenberg-Marquardt minimizer with DENSE_QR, Jacobi scaling and the
 * default tolerances.  Ceres is not available in
 * this image: "believed-upstream semantics", parity unpinned beyond
 * SLOAMTest.PoseOptimization (|t| < 0.1 m). */
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include "orc.h"

namespace orc {
namespace {

/* ---------------------------------------------------------------- jets ---- */
template <int N>
struct Jet {
  double a = 0;
  double v[N];
  Jet() { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; }
  static Jet var(double x, int k) { Jet j(x); j.v[k] = 1; return j; }
};
template <int N> Jet<N> operator+(const Jet<N> &f, const Jet<N> &g) {
  Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> Jet<N> operator-(const Jet<N> &f, const Jet<N> &g) {
  Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> Jet<N> operator-(const Jet<N> &f) {
  Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> Jet<N> operator*(const Jet<N> &f, const Jet<N> &g) {
  Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> Jet<N> operator/(const Jet<N> &f, const Jet<N> &g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int N> Jet<N> jsqrt(const Jet<N> &f) {
  Jet<N> h; h.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * h.a);
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * t; return h; }
template <int N> Jet<N> jsin(const Jet<N> &f) {
  Jet<N> h; h.a = std::sin(f.a); const double c = std::cos(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> Jet<N> jcos(const Jet<N> &f) {
  Jet<N> h; h.a = std::cos(f.a); const double s = -std::sin(f.a);
  for (int i = 0; i < N; ++i) h.v[i] = s * f.v[i]; return h; }
template <int N> Jet<N> jabs(const Jet<N> &f) { return f.a < 0 ? -f : f; } /* ceres::abs(Jet) */
inline double jsqrt(double x) { return std::sqrt(x); }
inline double jsin(double x) { return std::sin(x); }
inline double jcos(double x) { return std::cos(x); }
inline double jabs(double x) { return std::fabs(x); }
template <int N> inline double scalar(const Jet<N> &f) { return f.a; }
inline double scalar(double x) { return x; }

template <typename T> struct Vec3 { T x, y, z; };
template <typename T> Vec3<T> vadd(const Vec3<T> &a, const Vec3<T> &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> Vec3<T> vsub(const Vec3<T> &a, const Vec3<T> &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> Vec3<T> vscale(const T &s, const Vec3<T> &a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename T> T vdot(const Vec3<T> &a, const Vec3<T> &b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
template <typename T> Vec3<T> vcross(const Vec3<T> &a, const Vec3<T> &b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
template <typename T> T vnorm(const Vec3<T> &a) { return jsqrt(vdot(a, a)); }

/* Eigen Quaternion<T> * Vector3<T> (q = x,y,z,w), not normalised */
template <typename T>
Vec3<T> quat_rotate_t(const T q[4], const Vec3<T> &v) {
  const Vec3<T> u{q[0], q[1], q[2]};
  Vec3<T> uv = vcross(u, v);
  uv = vadd(uv, uv);
  return vadd(vadd(v, vscale(q[3], uv)), vcross(u, uv));
}

/* ceres::AngleAxisRotatePoint (rotation.h) */
template <typename T>
Vec3<T> angle_axis_rotate(const T aa[3], const Vec3<T> &pt) {
  const T theta2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (scalar(theta2) > DBL_EPSILON) {
    const T theta = jsqrt(theta2);
    const T costheta = jcos(theta), sintheta = jsin(theta);
    const T theta_inverse = T(1.0) / theta;
    const Vec3<T> w{aa[0] * theta_inverse, aa[1] * theta_inverse, aa[2] * theta_inverse};
    const Vec3<T> wxp = vcross(w, pt);
    const T tmp = (w.x * pt.x + w.y * pt.y + w.z * pt.z) * (T(1.0) - costheta);
    return {pt.x * costheta + wxp.x * sintheta + w.x * tmp,
            pt.y * costheta + wxp.y * sintheta + w.y * tmp,
            pt.z * costheta + wxp.z * sintheta + w.z * tmp};
  }
  const Vec3<T> a{aa[0], aa[1], aa[2]};
  const Vec3<T> wxp = vcross(a, pt);
  return vadd(pt, wxp);
}

/* residual of CylinderCost / XYYawCylinderCost after the point transform
 * (cylinder.h:66-86, :118-124) */
template <typename T>
T cylinder_residual(const Vec3<T> &lp, const CylinderParameters &m) {
  const Vec3<T> root{T(m.root.x), T(m.root.y), T(m.root.z)};
  const Vec3<T> ray{T(m.ray.x), T(m.ray.y), T(m.ray.z)};
  const Vec3<T> pp = vadd(root, vscale(vdot(vsub(lp, root), ray) / vdot(ray, ray), ray));
  return vnorm(vsub(pp, lp)) - T(m.radius);
}
/* residual of PlaneCost / ZRollPitchPlaneCost (plane.h:60-71, :99-110) */
template <typename T>
T plane_residual(const Vec3<T> &lp, const PlaneParameters &m) {
  const Vec3<T> n{T(m.plane[0]), T(m.plane[1]), T(m.plane[2])};
  const T denominator = vnorm(n);
  const T numerator = jabs(n.x * lp.x + n.y * lp.y + n.z * lp.z + T(m.plane[3]));
  return numerator / denominator;
}

/* ----------------------------------------------------------- problems ---- */
struct Problem {
  int n_ambient = 0, n_tangent = 0;
  virtual ~Problem() {}
  virtual int num_residuals() const = 0;
  /* raw residuals r [m] and tangent-space jacobian J [m x n_tangent] (row major, may be null) */
  virtual void evaluate(const double *x, double *r, double *J) const = 0;
  virtual void plus(const double *x, const double *delta, double *out) const = 0;
};

/* OptimizePose (sloam.cpp:174-255): blocks para_q[4] (x,y,z,w;
 * EigenQuaternionParameterization) then para_t[3]; x = [q, t]. */
struct JointProblem : Problem {
  const std::vector<TreeMatch> &tm;
  const std::vector<PlaneMatch> &gm;
  JointProblem(const std::vector<TreeMatch> &t, const std::vector<PlaneMatch> &g) : tm(t), gm(g) {
    n_ambient = 7; n_tangent = 6;
  }
  int num_residuals() const override { return (int)(tm.size() + gm.size()); }
  void evaluate(const double *x, double *r, double *J) const override {
    using J7 = Jet<7>;
    J7 q[4], t[3];
    for (int i = 0; i < 4; ++i) q[i] = J7::var(x[i], i);
    for (int i = 0; i < 3; ++i) t[i] = J7::var(x[4 + i], 4 + i);
    /* EigenQuaternionParameterization::ComputeJacobian, 4x3 row-major */
    const double P[4][3] = {{x[3], x[2], -x[1]}, {-x[2], x[3], x[0]}, {x[1], -x[0], x[3]},
                            {-x[0], -x[1], -x[2]}};
    int row = 0;
    auto emit = [&](const J7 &res) {
      r[row] = res.a;
      if (J) {
        double *Jr = J + (size_t)row * 6;
        for (int c = 0; c < 3; ++c)
          Jr[c] = res.v[0] * P[0][c] + res.v[1] * P[1][c] + res.v[2] * P[2][c] + res.v[3] * P[3][c];
        for (int c = 0; c < 3; ++c) Jr[3 + c] = res.v[4 + c];
      }
      ++row;
    };
    for (const TreeMatch &m : tm) {
      const Vec3<J7> cp{J7(m.feature.x), J7(m.feature.y), J7(m.feature.z)};
      const Vec3<J7> lp = vadd(quat_rotate_t(q, cp), Vec3<J7>{t[0], t[1], t[2]});
      emit(cylinder_residual(lp, m.object)); /* weight 1 */
    }
    for (const PlaneMatch &m : gm) {
      const Vec3<J7> cp{J7(m.feature.x), J7(m.feature.y), J7(m.feature.z)};
      const Vec3<J7> lp = vadd(quat_rotate_t(q, cp), Vec3<J7>{t[0], t[1], t[2]});
      emit(plane_residual(lp, m.object));
    }
  }
  void plus(const double *x, const double *d, double *out) const override {
    /* EigenQuaternionParameterization::Plus: q_delta (x) q, then t + dt */
    const double nd = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (nd > 0.0) {
      const double s = std::sin(nd) / nd;
      const Quat qd{std::cos(nd), s * d[0], s * d[1], s * d[2]};
      const Quat qx{x[3], x[0], x[1], x[2]};
      const Quat r = quat_mul(qd, qx);
      out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    } else {
      for (int i = 0; i < 4; ++i) out[i] = x[i];
    }
    for (int i = 0; i < 3; ++i) out[4 + i] = x[4 + i] + d[3 + i];
  }
};

/* OptimizeXYYaw / OptimizeZRollPitch (sloam.cpp:55-172): one 6-vector
 * [t, angle-axis] with SubsetParameterization holding `constant` fixed. */
struct SubsetProblem : Problem {
  const std::vector<TreeMatch> *tm = nullptr;
  const std::vector<PlaneMatch> *gm = nullptr;
  int free_idx[3];
  SubsetProblem(const std::vector<TreeMatch> *t, const std::vector<PlaneMatch> *g, int f0, int f1,
                int f2) : tm(t), gm(g) {
    n_ambient = 6; n_tangent = 3;
    free_idx[0] = f0; free_idx[1] = f1; free_idx[2] = f2;
  }
  int num_residuals() const override { return (int)(tm ? tm->size() : gm->size()); }
  void evaluate(const double *x, double *r, double *J) const override {
    using J6 = Jet<6>;
    J6 p[6];
    for (int i = 0; i < 6; ++i) p[i] = J6::var(x[i], i);
    const int m = num_residuals();
    for (int row = 0; row < m; ++row) {
      J6 res;
      if (tm) {
        const TreeMatch &mt = (*tm)[row];
        const Vec3<J6> cp{J6(mt.feature.x), J6(mt.feature.y), J6(mt.feature.z)};
        Vec3<J6> lp = angle_axis_rotate(p + 3, cp);
        lp = vadd(lp, Vec3<J6>{p[0], p[1], p[2]});
        res = cylinder_residual(lp, mt.object);
      } else {
        const PlaneMatch &mp = (*gm)[row];
        const Vec3<J6> cp{J6(mp.feature.x), J6(mp.feature.y), J6(mp.feature.z)};
        Vec3<J6> lp = angle_axis_rotate(p + 3, cp);
        lp = vadd(lp, Vec3<J6>{p[0], p[1], p[2]});
        res = plane_residual(lp, mp.object);
      }
      r[row] = res.a;
      if (J) for (int c = 0; c < 3; ++c) J[(size_t)row * 3 + c] = res.v[free_idx[c]];
    }
  }
  void plus(const double *x, const double *d, double *out) const override {
    for (int i = 0; i < 6; ++i) out[i] = x[i];
    for (int c = 0; c < 3; ++c) out[free_idx[c]] = x[free_idx[c]] + d[c];
  }
};

/* ------------------------------------------------------- LM minimizer ---- */
struct Eval {
  double cost = 0;
  std::vector<double> r, J, g;
};

/* ResidualBlock::Evaluate with HuberLoss(a) + Corrector (rho'' <= 0 branch) */
void evaluate_robust(const Problem &pb, double huber_a, const double *x, Eval &e, bool want_jac) {
  const int m = pb.num_residuals(), n = pb.n_tangent;
  e.r.assign(m, 0.0);
  if (want_jac) e.J.assign((size_t)m * n, 0.0);
  pb.evaluate(x, e.r.data(), want_jac ? e.J.data() : nullptr);
  const double b = huber_a * huber_a;
  double cost = 0;
  for (int i = 0; i < m; ++i) {
    const double s = e.r[i] * e.r[i];
    double rho0, rho1;
    if (s > b) {
      const double rr = std::sqrt(s);
      rho0 = 2.0 * huber_a * rr - b;
      rho1 = std::max(DBL_MIN, huber_a / rr);
    } else {
      rho0 = s; rho1 = 1.0;
    }
    cost += 0.5 * rho0;
    const double sr = std::sqrt(rho1);
    if (want_jac) for (int c = 0; c < n; ++c) e.J[(size_t)i * n + c] *= sr;
    e.r[i] *= sr;
  }
  e.cost = cost;
  if (want_jac) {
    e.g.assign(n, 0.0);
    for (int i = 0; i < m; ++i)
      for (int c = 0; c < n; ++c) e.g[c] += e.J[(size_t)i * n + c] * e.r[i];
  }
}

/* Eigen HouseholderQR least squares: min || A y - b ||, A (rows x n) row-major, overwritten */
bool qr_solve(std::vector<double> &A, std::vector<double> &b, int rows, int n, double *y) {
  for (int k = 0; k < n; ++k) {
    double tailSq = 0;
    for (int r = k + 1; r < rows; ++r) tailSq += A[(size_t)r * n + k] * A[(size_t)r * n + k];
    const double c0 = A[(size_t)k * n + k];
    double tau, beta;
    if (tailSq <= DBL_MIN) { tau = 0; beta = c0; }
    else {
      beta = std::sqrt(c0 * c0 + tailSq);
      if (c0 >= 0) beta = -beta;
      for (int r = k + 1; r < rows; ++r) A[(size_t)r * n + k] /= (c0 - beta);
      tau = (beta - c0) / beta;
    }
    A[(size_t)k * n + k] = beta;
    if (tau != 0) {
      for (int j = k + 1; j <= n; ++j) { /* j == n: the right-hand side */
        auto at = [&](int r) -> double & { return j < n ? A[(size_t)r * n + j] : b[r]; };
        double tmp = at(k);
        for (int r = k + 1; r < rows; ++r) tmp += A[(size_t)r * n + k] * at(r);
        at(k) -= tau * tmp;
        for (int r = k + 1; r < rows; ++r) at(r) -= tau * A[(size_t)r * n + k] * tmp;
      }
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[(size_t)k * n + j] * y[j];
    y[k] = s / A[(size_t)k * n + k];
  }
  for (int k = 0; k < n; ++k) if (!std::isfinite(y[k])) return false;
  return true;
}

/* TrustRegionMinimizer::Minimize + LevenbergMarquardtStrategy + DenseQRSolver
 * with the defaults listed in SURVEY A.7.  x is updated in place to the last
 * accepted point (Ceres leaves the user state at the best accepted iterate). */
LMSummary lm_solve(const Problem &pb, double huber_a, int max_iterations, double *x) {
  LMSummary sum;
  const int n = pb.n_tangent, na = pb.n_ambient, m = pb.num_residuals();
  if (m == 0) { /* reduced program empty: CONVERGENCE, parameters untouched */
    sum.termination = 0;
    return sum;
  }
  const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  const double min_relative_decrease = 1e-3, min_radius = 1e-32, max_radius = 1e16;
  const double min_diag = 1e-6, max_diag = 1e32;
  double radius = 1e4, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  int invalid_steps = 0;

  std::vector<double> xc(x, x + na), cand(na), tmp(na);
  auto xnorm = [&](const std::vector<double> &v) {
    double s = 0; for (double e : v) s += e * e; return std::sqrt(s); };
  Eval ev;
  std::vector<double> scale(n), diag(n), lmdiag(n), step(n), delta(n);
  double gmax = 0;
  auto eval_grad = [&](bool first) {
    evaluate_robust(pb, huber_a, xc.data(), ev, true);
    if (first)
      for (int c = 0; c < n; ++c) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += ev.J[(size_t)i * n + c] * ev.J[(size_t)i * n + c];
        scale[c] = 1.0 / (1.0 + std::sqrt(s));
      }
    for (int i = 0; i < m; ++i)
      for (int c = 0; c < n; ++c) ev.J[(size_t)i * n + c] *= scale[c];
    std::vector<double> neg(n);
    for (int c = 0; c < n; ++c) neg[c] = -ev.g[c];
    pb.plus(xc.data(), neg.data(), tmp.data());
    gmax = 0;
    for (int i = 0; i < na; ++i) gmax = std::max(gmax, std::fabs(xc[i] - tmp[i]));
  };
  /* IterationZero */
  double x_norm = xnorm(xc);
  eval_grad(true);
  double x_cost = ev.cost;
  sum.initial_cost = x_cost;
  int iteration = 0;
  bool last_successful = false;
  for (;;) {
    /* FinalizeIterationAndCheckIfMinimizerCanContinue */
    if (last_successful) for (int i = 0; i < na; ++i) x[i] = xc[i];
    if (iteration >= max_iterations) { sum.termination = 1; break; }
    if (gmax <= gradient_tolerance) { sum.termination = 0; break; }
    if (radius < min_radius) { sum.termination = 0; break; }
    ++iteration;
    last_successful = false;
    /* LevenbergMarquardtStrategy::ComputeStep */
    if (!reuse_diagonal)
      for (int c = 0; c < n; ++c) {
        double s = 0;
        for (int i = 0; i < m; ++i) s += ev.J[(size_t)i * n + c] * ev.J[(size_t)i * n + c];
        diag[c] = std::min(std::max(s, min_diag), max_diag);
      }
    for (int c = 0; c < n; ++c) lmdiag[c] = std::sqrt(diag[c] / radius);
    /* DenseQRSolver: [J; D] y = [r; 0], step = -y */
    std::vector<double> A((size_t)(m + n) * n, 0.0), rhs(m + n, 0.0);
    std::copy(ev.J.begin(), ev.J.end(), A.begin());
    for (int c = 0; c < n; ++c) A[(size_t)(m + c) * n + c] = lmdiag[c];
    std::copy(ev.r.begin(), ev.r.end(), rhs.begin());
    const bool solved = qr_solve(A, rhs, m + n, n, step.data());
    reuse_diagonal = true;
    bool step_valid = false;
    double model_cost_change = 0;
    if (solved) {
      for (int c = 0; c < n; ++c) step[c] = -step[c];
      /* model_cost_change = -(J step).(r + J step / 2) */
      for (int i = 0; i < m; ++i) {
        double mr = 0;
        for (int c = 0; c < n; ++c) mr += ev.J[(size_t)i * n + c] * step[c];
        model_cost_change += -mr * (ev.r[i] + mr / 2.0);
      }
      step_valid = model_cost_change > 0.0;
    }
    if (!step_valid) { /* HandleInvalidStep */
      if (++invalid_steps >= 5) { sum.termination = 2; break; }
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      continue;
    }
    invalid_steps = 0;
    for (int c = 0; c < n; ++c) delta[c] = step[c] * scale[c];
    pb.plus(xc.data(), delta.data(), cand.data());
    Eval ec;
    evaluate_robust(pb, huber_a, cand.data(), ec, false);
    const double cand_cost = ec.cost;
    /* ParameterToleranceReached */
    {
      double s = 0;
      for (int i = 0; i < na; ++i) s += (xc[i] - cand[i]) * (xc[i] - cand[i]);
      if (std::sqrt(s) <= parameter_tolerance * (x_norm + parameter_tolerance)) { sum.termination = 0; break; }
    }
    /* FunctionToleranceReached */
    if (std::fabs(x_cost - cand_cost) <= function_tolerance * x_cost) { sum.termination = 0; break; }
    const double relative_decrease = (x_cost - cand_cost) / model_cost_change;
    if (relative_decrease > min_relative_decrease) { /* HandleSuccessfulStep */
      xc = cand;
      x_norm = xnorm(xc);
      eval_grad(false);
      x_cost = ev.cost;
      last_successful = true;
      radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
      radius = std::min(max_radius, radius);
      decrease_factor = 2.0;
      reuse_diagonal = false;
    } else { /* StepRejected */
      radius /= decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
    }
  }
  sum.iterations = iteration;
  sum.final_cost = x_cost;
  return sum;
}

/* ceres::QuaternionToAngleAxis / AngleAxisToQuaternion (q = w,x,y,z) */
void quat_to_angle_axis(const double q[4], double aa[3]) {
  const double q1 = q[1], q2 = q[2], q3 = q[3];
  const double s2 = q1 * q1 + q2 * q2 + q3 * q3;
  if (s2 > 0.0) {
    const double s = std::sqrt(s2), c = q[0];
    const double two_theta = 2.0 * ((c < 0.0) ? std::atan2(-s, -c) : std::atan2(s, c));
    const double k = two_theta / s;
    aa[0] = q1 * k; aa[1] = q2 * k; aa[2] = q3 * k;
  } else {
    aa[0] = q1 * 2.0; aa[1] = q2 * 2.0; aa[2] = q3 * 2.0;
  }
}
void angle_axis_to_quat(const double aa[3], double q[4]) {
  const double t2 = aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2];
  if (t2 > 0.0) {
    const double t = std::sqrt(t2), h = t * 0.5, k = std::sin(h) / t;
    q[0] = std::cos(h); q[1] = aa[0] * k; q[2] = aa[1] * k; q[3] = aa[2] * k;
  } else {
    q[0] = 1.0; q[1] = aa[0] * 0.5; q[2] = aa[1] * 0.5; q[3] = aa[2] * 0.5;
  }
}
}  // namespace

bool optimize_pose(const Options &o, const SE3 &poseEstimate, const std::vector<TreeMatch> &tm,
                   const std::vector<PlaneMatch> &gm, SE3 &tf, LMSummary *s) {
  /* sloam.cpp:185-193 */
  double x[7] = {poseEstimate.q.x, poseEstimate.q.y, poseEstimate.q.z, poseEstimate.q.w,
                 poseEstimate.t.x, poseEstimate.t.y, poseEstimate.t.z};
  JointProblem pb(tm, gm);
  const LMSummary sum = lm_solve(pb, o.p.huber_delta, o.p.lm_max_iterations, x);
  if (s) *s = sum;
  const bool success = sum.termination == 0; /* :241 */
  if (success) { /* :242-246 setQuaternion normalises */
    tf.q = quat_normalized({x[3], x[0], x[1], x[2]});
    tf.t = {x[4], x[5], x[6]};
  }
  return success;
}

bool two_step_optimize_pose(const Options &o, const SE3 &poseEstimate, bool optimTrees,
                            bool optimGround, const std::vector<TreeMatch> &tm,
                            const std::vector<PlaneMatch> &gm, SE3 &tf, LMSummary s[2]) {
  const double quat[4] = {poseEstimate.q.w, poseEstimate.q.x, poseEstimate.q.y, poseEstimate.q.z};
  double rpy[3];
  quat_to_angle_axis(quat, rpy); /* sloam.cpp:58-63, :119-124 */
  const V3 t = poseEstimate.t;
  double treeOut[3], groundOut[3];
  LMSummary s0, s1;
  { /* OptimizeXYYaw, :55-114: z, roll, pitch constant -> free {0,1,5} */
    double params[6] = {t.x, t.y, t.z, rpy[0], rpy[1], rpy[2]};
    bool success = true;
    if (optimTrees) {
      SubsetProblem pb(&tm, nullptr, 0, 1, 5);
      s0 = lm_solve(pb, o.p.huber_delta, o.p.lm_max_iterations, params);
      success = s0.termination == 0;
    }
    if (optimTrees && success) { treeOut[0] = params[0]; treeOut[1] = params[1]; treeOut[2] = params[5]; }
    else { treeOut[0] = t.x; treeOut[1] = t.y; treeOut[2] = rpy[2]; }
  }
  { /* OptimizeZRollPitch, :116-172: x, y, yaw constant -> free {2,3,4} */
    double params[6] = {t.x, t.y, t.z, rpy[0], rpy[1], rpy[2]};
    bool success = true;
    if (optimGround) {
      SubsetProblem pb(nullptr, &gm, 2, 3, 4);
      s1 = lm_solve(pb, o.p.huber_delta, o.p.lm_max_iterations, params);
      success = s1.termination == 0;
    }
    if (optimGround && success) { groundOut[0] = params[2]; groundOut[1] = params[3]; groundOut[2] = params[4]; }
    else { groundOut[0] = t.z; groundOut[1] = rpy[0]; groundOut[2] = rpy[1]; }
  }
  if (s) { s[0] = s0; s[1] = s1; }
  /* :43-52 */
  const double aa[3] = {groundOut[1], groundOut[2], treeOut[2]};
  double q[4];
  angle_axis_to_quat(aa, q);
  tf.q = quat_normalized({q[0], q[1], q[2], q[3]});
  tf.t = {treeOut[0], treeOut[1], groundOut[0]};
  return true;
}

}  // namespace orc
