// hd_check.cpp -- host build of the scalar device code of the pose optimiser
// (sloam_b200/csrc/dev_lm.h) with a serial evaluator, so that the Jacobians and the
// trust-region state machine can be checked against the oracle without a GPU
// (tests/test_hd_check.py).  Test infrastructure: the product runs this code only inside
// lm_kernel on the device.
//   g++ -O2 -ffp-contract=off -shared -fPIC tests/hd_check.cpp -o hd_check.so
#include "../sloam_b200/csrc/dev_lm.h"

using namespace sb;

struct SerialEval {
  int mode;
  const double *tf; const sloam_cylinder *to; int nt;
  const double *pf; const sloam_plane *po; int np_;
  double huber_a;
  void operator()(const double *x, double *cost, double *A, double *g) {
    const bool want_jac = true;
    const int n = mode == LM_JOINT ? 6 : 3, npk = n * (n + 1) / 2;
    double acc[28];
    for (double &a : acc) a = 0.0;
    AaPre pre;
    if (mode != LM_JOINT) aa_prepare(x, pre);
    for (int r = 0; r < nt + np_; ++r) {
      double J[6], res;
      if (r < nt) res = residual_row(mode, x, pre, tf + 3 * r, to + r, nullptr, want_jac ? J : nullptr);
      else res = residual_row(mode, x, pre, pf + 3 * (r - nt), nullptr, po + (r - nt), want_jac ? J : nullptr);
      double sc;
      acc[27] += huber(res, huber_a, &sc);
      if (want_jac) {
        const double rr = res * sc;
        int p = 0;
        for (int i = 0; i < n; ++i) {
          const double ji = J[i] * sc;
          for (int j = i; j < n; ++j) acc[p++] += ji * (J[j] * sc);
          acc[21 + i] += ji * rr;
        }
      }
    }
    *cost = acc[27];
    if (want_jac) {
      for (int i = 0; i < npk; ++i) A[i] = acc[i];
      for (int i = 0; i < n; ++i) g[i] = acc[21 + i];
    }
  }
};

extern "C" {
// mode: 0 joint (x = qx qy qz qw tx ty tz), 1 XYYaw, 2 ZRollPitch (x = t, angle-axis)
int hd_lm_solve(int mode, double *x, const double *tf, const sloam_cylinder *to, int nt, const double *pf,
                const sloam_plane *po, int np_, double huber_a, int max_it, int *iterations, double *costs) {
  SerialEval ev{mode, tf, to, nt, pf, po, np_, huber_a};
  const bool use_t = mode != LM_ZROLLPITCH, use_p = mode != LM_XYYAW;
  if (!use_t) ev.nt = 0;
  if (!use_p) ev.np_ = 0;
  LMWork w = {};
  const LMOut o = lm_minimize(ev, w, mode, ev.nt + ev.np_, max_it, x);
  *iterations = o.iterations;
  costs[0] = o.initial_cost; costs[1] = o.final_cost;
  return o.termination;
}
// one residual row: value + tangent-space Jacobian (for finite-difference checks)
double hd_residual_row(int mode, const double *x, const double *feat, const sloam_cylinder *cyl,
                       const sloam_plane *pl, double *J) {
  AaPre pre;
  if (mode != LM_JOINT) aa_prepare(x, pre);
  return residual_row(mode, x, pre, feat, cyl, pl, J);
}
void hd_plus(int mode, const double *x, const double *d, double *out) { lm_plus(mode, x, d, out); }
}
