// synth.h -- seeded synthetic forest scans with ground-truth labels
// (SURVEY.md 8(d)): organized H x W lidar returns of a sloped ground plane and
// n vertical-ish cylinders, one beam per thread.  Test/bench infrastructure
// compiled for host (sloam_synth_generate_host) and device (..._dev).
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/sloam_b200.h"
#include "common.cuh"
#include "proj_math.h"

namespace sb {

SLOAM_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
// uniform in (0,1)
SLOAM_HD double hash_uniform(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  uint64_t h = splitmix64(seed ^ splitmix64(a ^ splitmix64(b ^ splitmix64(c))));
  return ((double)(h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
SLOAM_HD double hash_normal(uint64_t seed, uint64_t a, uint64_t b, uint64_t c) {
  const double u1 = hash_uniform(seed, a, b, 2 * c), u2 = hash_uniform(seed, a, b, 2 * c + 1);
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}

struct SynthTree {  // map frame
  float cx, cy, cz;  // axis point on the ground
  float ax, ay, az;  // unit axis
  float radius, height;
};

struct SynthScene {
  // ground plane n.x + d = 0 in the map frame (n unit, nz > 0)
  float gn[3], gd;
  int n_trees;
};

// Ground-truth sensor pose of keyframe k: a circle of radius 1.5 m around the
// scene origin walked in `step` metre arcs, heading along the tangent.
SLOAM_HD void synth_gt_pose(const sloam_synth_config &c, int64_t k, sloam_pose *T) {
  const double R = 1.5;
  const double ang = (double)k * (double)c.step_per_keyframe / R;
  const double yaw = ang + 1.5707963267948966;
  T->t[0] = R * cos(ang);
  T->t[1] = R * sin(ang);
  T->t[2] = (double)c.sensor_height;
  T->q[0] = 0; T->q[1] = 0; T->q[2] = sin(0.5 * yaw); T->q[3] = cos(0.5 * yaw);
}

// One beam.  Returns the label (0 none, 1 ground, 255 tree) and the point in
// the sensor frame.
SLOAM_HD int synth_beam(const sloam_synth_config &c, const SynthScene &sc, const SynthTree *trees,
                        const sloam_pose &T, int64_t k, int row, int col, sloam_point *out) {
  const double fov_up = (double)c.fov_up_deg * 0.017453292519943295;
  const double fov_down = (double)c.fov_down_deg * 0.017453292519943295;
  const double fov = fabs(fov_up) + fabs(fov_down);
  const double yaw = 3.141592653589793 * (2.0 * ((double)col + 0.5 + (double)c.azimuth_offset_cols) / (double)c.img_w - 1.0);
  const double pitch = fov_up - ((double)row + 0.5) * fov / (double)c.img_h;
  // sensor-frame direction consistent with yaw = -atan2(y, x), pitch = asin(z / r)
  const double ds[3] = {cos(pitch) * cos(yaw), -cos(pitch) * sin(yaw), sin(pitch)};
  double d[3];
  q_rotate(T.q, ds, d);
  const double o[3] = {T.t[0], T.t[1], T.t[2]};
  double best_t = (double)c.max_range;
  int label = 0;
  // ground
  {
    const double denom = sc.gn[0] * d[0] + sc.gn[1] * d[1] + sc.gn[2] * d[2];
    if (denom < -1e-9) {
      const double t = -(sc.gn[0] * o[0] + sc.gn[1] * o[1] + sc.gn[2] * o[2] + sc.gd) / denom;
      if (t > 0.3 && t < best_t) { best_t = t; label = 1; }
    }
  }
  // trunks: |(o + t d - c) - ((o + t d - c).a) a| = r
  for (int i = 0; i < sc.n_trees; ++i) {
    const SynthTree &tr = trees[i];
    const double w[3] = {o[0] - tr.cx, o[1] - tr.cy, o[2] - tr.cz};
    const double a[3] = {tr.ax, tr.ay, tr.az};
    const double da = d[0] * a[0] + d[1] * a[1] + d[2] * a[2];
    const double wa = w[0] * a[0] + w[1] * a[1] + w[2] * a[2];
    const double dp[3] = {d[0] - da * a[0], d[1] - da * a[1], d[2] - da * a[2]};
    const double wp[3] = {w[0] - wa * a[0], w[1] - wa * a[1], w[2] - wa * a[2]};
    const double A = dp[0] * dp[0] + dp[1] * dp[1] + dp[2] * dp[2];
    const double Bq = dp[0] * wp[0] + dp[1] * wp[1] + dp[2] * wp[2];
    const double Cq = wp[0] * wp[0] + wp[1] * wp[1] + wp[2] * wp[2] - (double)tr.radius * tr.radius;
    const double disc = Bq * Bq - A * Cq;
    if (A < 1e-12 || disc <= 0) continue;
    const double t = (-Bq - sqrt(disc)) / A;
    if (t <= 0.3 || t >= best_t) continue;
    const double h = wa + t * da;  // height along the axis
    if (h < 0 || h > (double)tr.height) continue;
    best_t = t; label = 255;
  }
  if (label == 0) {
    const float v = c.nan_no_return ? NAN : 0.0f;
    out->x = v; out->y = v; out->z = v; out->intensity = 0.0f;
    return 0;
  }
  const uint64_t pixel = (uint64_t)row * (uint64_t)c.img_w + (uint64_t)col;
  const double sigma = label == 1 ? (double)c.ground_noise : (double)c.range_noise;
  const double tn = best_t + sigma * hash_normal(c.seed, (uint64_t)k, pixel, 1);
  out->x = (float)(tn * ds[0]);
  out->y = (float)(tn * ds[1]);
  out->z = (float)(tn * ds[2]);
  out->intensity = (float)(100.0 * hash_uniform(c.seed, (uint64_t)k, pixel, 7));
  return label;
}

}  // namespace sb
