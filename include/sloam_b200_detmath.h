/*
 * sloam_b200_detmath.h -- host/device bit-identical atan2f / asinf.
 *
 * Why: range-image indices (inference.cpp:107-127) and ground-cell indices
 * (sloam.cpp:346-357) are integers derived from atan2f/asinf.  glibc's and
 * CUDA's float transcendentals differ in last-place rounding, so indices can
 * flip at bin edges.  These versions use only IEEE-754 exactly-rounded
 * operations (+ - * / sqrt on double, compiled without FMA contraction:
 * -ffp-contract=off on the host, --fmad=false on the device), so every
 * compiler on every target produces the same bits.  They evaluate in double
 * and round once to float, i.e. they are the correctly-rounded float result
 * except in ~1e-9 of cases.  The deviation from glibc's (not correctly
 * rounded) atan2f/asinf is measured in tests/test_oracle_projection.py and
 * recorded in DESIGN.md; the oracle can run in either mode.
 *
 * Algorithm: the classic Sun fdlibm double atan (argument reduction to
 * [0, 7/16] around 0.5, 1, 1.5, inf + an odd degree-23 polynomial), atan2 by
 * quadrant, asin(v) = atan2(v, sqrt((1-v)(1+v))).
 */
#ifndef SLOAM_B200_DETMATH_H
#define SLOAM_B200_DETMATH_H

#include <math.h>

#if defined(__CUDACC__)
#define SLOAM_HD __host__ __device__ __forceinline__
#else
#define SLOAM_HD inline
#endif

namespace sloam_det {

SLOAM_HD double det_atan_pos(double x) {
  /* x >= 0, finite or +inf */
  const double hi0 = 4.63647609000806093515e-01, lo0 = 2.26987774529616870924e-17;
  const double hi1 = 7.85398163397448278999e-01, lo1 = 3.06161699786838301793e-17;
  const double hi2 = 9.82793723247329054082e-01, lo2 = 1.39033110312309984516e-17;
  const double hi3 = 1.57079632679489655800e+00, lo3 = 6.12323399573676603587e-17;
  const double a0 = 3.33333333333329318027e-01, a1 = -1.99999999998764832476e-01,
               a2 = 1.42857142725034663711e-01, a3 = -1.11111104054623557880e-01,
               a4 = 9.09088713343650656196e-02, a5 = -7.69187620504482999495e-02,
               a6 = 6.66107313738753120669e-02, a7 = -5.83357013379057348645e-02,
               a8 = 4.97687799461593236017e-02, a9 = -3.65315727442169155270e-02,
               a10 = 1.62858201153657823623e-02;
  if (x >= 7.3786976294838206464e19) return hi3 + lo3; /* 2^66 */
  double hi, lo, t;
  int reduced = 1;
  if (x < 0.4375) {
    if (x < 1.862645149230957e-09) return x; /* 2^-29 */
    reduced = 0; hi = 0.0; lo = 0.0; t = x;
  } else if (x < 0.6875) {
    hi = hi0; lo = lo0; t = (2.0 * x - 1.0) / (2.0 + x);
  } else if (x < 1.1875) {
    hi = hi1; lo = lo1; t = (x - 1.0) / (x + 1.0);
  } else if (x < 2.4375) {
    hi = hi2; lo = lo2; t = (x - 1.5) / (1.0 + 1.5 * x);
  } else {
    hi = hi3; lo = lo3; t = -1.0 / x;
  }
  const double z = t * t;
  const double w = z * z;
  const double s1 = z * (a0 + w * (a2 + w * (a4 + w * (a6 + w * (a8 + w * a10)))));
  const double s2 = w * (a1 + w * (a3 + w * (a5 + w * (a7 + w * a9))));
  if (!reduced) return t - t * (s1 + s2);
  return hi - ((t * (s1 + s2) - lo) - t);
}

/* double atan2 for arguments that came from floats (no overflow of y/x). */
SLOAM_HD double det_atan2(double y, double x) {
  const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16;
  const double pio2 = 1.5707963267948965580e+00;
  if (x != x || y != y) return x + y; /* NaN */
  const bool xneg = signbit(x), yneg = signbit(y);
  if (y == 0.0) {
    if (!xneg) return y;             /* +-0 */
    return yneg ? -pi : pi;
  }
  if (x == 0.0) return yneg ? -pio2 : pio2;
  const double ax = fabs(x), ay = fabs(y);
  double z;
  if (isinf(ax)) {
    if (isinf(ay)) z = xneg ? 3.0 * (pi / 4.0) : pi / 4.0;
    else z = xneg ? pi : 0.0;
    return yneg ? -z : z;
  }
  if (isinf(ay)) return yneg ? -pio2 : pio2;
  z = det_atan_pos(ay / ax);
  if (xneg) z = pi - (z - pi_lo);
  return yneg ? -z : z;
}

/* Stand-in for atan2f(float, float). */
SLOAM_HD float det_atan2f(float y, float x) {
  return (float)det_atan2((double)y, (double)x);
}

/* Stand-in for asinf(float): NaN outside [-1, 1] like libm. */
SLOAM_HD float det_asinf(float v) {
  const double d = (double)v;
  if (!(d >= -1.0 && d <= 1.0)) return (float)NAN;
  const double c = sqrt((1.0 - d) * (1.0 + d));
  return (float)det_atan2(d, c);
}

}  // namespace sloam_det
#endif
