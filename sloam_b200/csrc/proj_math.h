// proj_math.h -- per-point spherical projection and polar ground binning.
// Host/device: the same float/double operation sequence as the reference,
//   Segmentation::_doProjection        sloam/src/segmentation/inference.cpp:99-127
//   sloam::binGroundPoints             sloam/src/core/sloam.cpp:339-358
//   euclideanDist2D / pow_2            sloam/include/helpers/utils.h:7-12
// with atan2f/asinf replaced by the bit-reproducible versions of
// include/sloam_b200_detmath.h.  Compiled with --fmad=false.
#pragma once

#include <cmath>
#include <cstring>

#include "../../include/sloam_b200_detmath.h"

namespace sb {

struct ProjGeom {
  float fov_down_abs;  // |fov_down| in radians (float member of the reference)
  float fov;           // |fov_down| + |fov_up|
  float Wf, Hf;        // image size as float
};

// Returns proj_y * W + proj_x; *range gets sqrt(x^2+y^2+z^2) (inference.cpp:103).
SLOAM_HD int project_pixel(const ProjGeom &g, float x, float y, float z, float *range) {
  const float r = sqrtf(x * x + y * y + z * z);
  *range = r;
  const float yaw = -sloam_det::det_atan2f(y, x);
  const float pitch = sloam_det::det_asinf(z / r);
  // :111-112 double expressions rounded to float
  float px = (float)(0.5 * ((double)yaw / 3.14159265358979323846 + 1.0));
  float py = (float)(1.0 - (double)((pitch + g.fov_down_abs) / g.fov));
  px *= g.Wf;
  py *= g.Hf;
  px = floorf(px);
  // std::min(W-1, px) = (px < W-1) ? px : W-1 ; std::max(0, px) = (0 < px) ? px : 0
  // (a NaN px therefore clamps to W-1)
  px = (px < g.Wf - 1.0f) ? px : g.Wf - 1.0f;
  px = (0.0f < px) ? px : 0.0f;
  py = floorf(py);
  py = (py < g.Hf - 1.0f) ? py : g.Hf - 1.0f;
  py = (0.0f < py) ? py : 0.0f;
  return (int)(py * g.Wf + px);
}

struct GroundGeom {
  double max_dist, min_dist;  // maxGroundLidarDist, minGroundLidarDist
  double radial_step;         // maxGroundLidarDist / groundRadiiBins
  double theta_step;          // 2 * PIDEF / groundThetaBins
  int RB, TB;
  float inv_radial_step_f, inv_theta_step_f;  // for the fp32 estimate in k1_project.cu
  // The radius tests and the radial bin as EXACT thresholds on r2 = x*x + y*y (fp32, the argument
  // of the reference's sqrtf, utils.h:9-12): sqrtf is monotone, so every decision that
  // ground_cell_of() takes on (double)sqrtf(r2) is one comparison of r2 with the smallest float
  // for which the decision flips (found by bisection over the float bit patterns, ctx.cu).
  float r2_in_lo;   // min{s : sqrtf(s) > min_dist}
  float r2_in_hi;   // min{s : !(sqrtf(s) < max_dist)}
  float r2_bin[3];  // r2_bin[j-1] = min{s : floor(sqrtf(s) / radial_step) >= j}, j = 1 .. RB-1; +inf beyond
  int r2_bins;      // RB - 1 when RB <= 4, else -1: radial bin by estimate + exact fallback
};

// Polar cell (rb * TB + tb) of a ground point seen from the origin, or -1 when
// the point is outside (min, max) radius.  sloam.cpp:342-358 with origin 0.
SLOAM_HD int ground_cell_of(const GroundGeom &g, float x, float y) {
  const double PIDEF = 3.14159265;  // definitions.h:28
  // pow_2(double) returns float; euclideanDist2D adds two floats and takes sqrtf
  const float dx = 0.0f - x, dy = 0.0f - y;          // vecA - vecB with vecA = origin
  const float sx = (float)((double)dx * (double)dx);
  const float sy = (float)((double)dy * (double)dy);
  const double radius = (double)sqrtf(sx + sy);
  if (!(radius < g.max_dist && radius > g.min_dist)) return -1;
  const double theta = (double)sloam_det::det_atan2f(y - 0.0f, x - 0.0f);
  int rb = (int)floor(radius / g.radial_step);
  int tb = (int)floor((PIDEF + theta) / g.theta_step);
  rb = rb < g.RB - 1 ? rb : g.RB - 1;
  rb = rb > 0 ? rb : 0;
  tb = tb < g.TB - 1 ? tb : g.TB - 1;
  tb = tb > 0 ? tb : 0;
  return rb * g.TB + tb;
}

// Radial part of ground_cell_of() from r2 = x*x + y*y by the thresholds of GroundGeom:
// -1 outside (min_dist, max_dist), else the clamped radial bin.  Only valid when g.r2_bins >= 0.
SLOAM_HD int ground_radial_bin_of_r2(const GroundGeom &g, float r2) {
  if (!(r2 >= g.r2_in_lo && r2 < g.r2_in_hi)) return -1;
  const int rb = ((r2 >= g.r2_bin[0]) ? 1 : 0) + ((r2 >= g.r2_bin[1]) ? 1 : 0) + ((r2 >= g.r2_bin[2]) ? 1 : 0);
  return rb < g.RB - 1 ? rb : g.RB - 1;
}

// Host only: fills the r2 thresholds from max_dist / min_dist / radial_step / RB.  Each is the
// smallest non-negative float (+inf included) for which a monotone predicate on sqrtf(s) holds,
// found by bisection over the bit patterns (non-negative floats order like their bits).
template <class Pred>
inline float sloam_first_float(Pred pred) {
  unsigned lo = 0u, hi = 0x7f800000u;  // pred(+inf) holds for all uses below
  while (lo < hi) {
    const unsigned mid = lo + (hi - lo) / 2;
    float f;
    std::memcpy(&f, &mid, 4);
    if (pred(f)) hi = mid; else lo = mid + 1;
  }
  float f;
  std::memcpy(&f, &lo, 4);
  return f;
}
inline void ground_geom_thresholds(GroundGeom &g) {
  const double mn = g.min_dist, mx = g.max_dist, step = g.radial_step;
  g.r2_in_lo = sloam_first_float([&](float s) { return (double)sqrtf(s) > mn; });
  g.r2_in_hi = sloam_first_float([&](float s) { return !((double)sqrtf(s) < mx); });
  g.r2_bins = (g.RB <= 4 && step > 0.0) ? g.RB - 1 : -1;
  for (int j = 1; j <= 3; ++j)
    g.r2_bin[j - 1] = (g.r2_bins >= j)
        ? sloam_first_float([&](float s) { return floor((double)sqrtf(s) / step) >= (double)j; })
        : INFINITY;
}

}  // namespace sb
