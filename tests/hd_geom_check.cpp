// hd_geom_check.cpp -- host build of the bit-exact per-point arithmetic the kernels fall back
// to (sloam_b200/csrc/proj_math.h: spherical projection, polar ground cell) and of the ground
// plane acceptance test (sloam_b200/csrc/dev_plane.h), so that they can be compared with the
// oracle without a GPU (tests/test_hd_geom.py).  Test infrastructure: the
// product runs this code only inside project_split_kernel / ground_tag_kernel on the device.
//   g++ -O2 -ffp-contract=off -shared -fPIC tests/hd_geom_check.cpp -o hd_geom_check.so
#include <cmath>

#include "../include/sloam_b200.h"
#include "../sloam_b200/csrc/dev_plane.h"
#include "../sloam_b200/csrc/proj_math.h"

using namespace sb;

extern "C" {
// geometry derived like ctx.cu:derive() (inference.cpp:7-9: double expressions stored to floats)
void hd_project(const sloam_params *p, const sloam_point *pts, int n, int32_t *pix, float *range) {
  const float fov_up = (float)((double)p->fov_up_deg / 180.0 * 3.14159265358979323846);
  const float fov_down = (float)((double)p->fov_down_deg / 180.0 * 3.14159265358979323846);
  ProjGeom g;
  g.fov_down_abs = std::fabs(fov_down);
  g.fov = std::fabs(fov_down) + std::fabs(fov_up);
  g.Wf = (float)p->img_w;
  g.Hf = (float)p->img_h;
  for (int i = 0; i < n; ++i) pix[i] = project_pixel(g, pts[i].x, pts[i].y, pts[i].z, &range[i]);
}
void hd_ground_cells(const sloam_params *p, const sloam_point *pts, int n, int32_t *cell) {
  GroundGeom g;
  g.max_dist = p->maxGroundLidarDist;
  g.min_dist = p->minGroundLidarDist;
  g.radial_step = p->maxGroundLidarDist / (double)p->groundRadiiBins;
  g.theta_step = 2 * 3.14159265 / (double)p->groundThetaBins;
  g.RB = p->groundRadiiBins;
  g.TB = p->groundThetaBins;
  g.inv_radial_step_f = (float)(1.0 / g.radial_step);
  g.inv_theta_step_f = (float)(1.0 / g.theta_step);
  for (int i = 0; i < n; ++i) cell[i] = ground_cell_of(g, pts[i].x, pts[i].y);
}
// the radial decisions of the same points from the r^2 thresholds (what project_split_kernel's
// fast path uses): -1 outside the radius range, else the radial bin; -2 when RB > 4 (no thresholds)
void hd_ground_radial_by_threshold(const sloam_params *p, const sloam_point *pts, int n, int32_t *rb) {
  GroundGeom g;
  g.max_dist = p->maxGroundLidarDist;
  g.min_dist = p->minGroundLidarDist;
  g.radial_step = p->maxGroundLidarDist / (double)p->groundRadiiBins;
  g.theta_step = 2 * 3.14159265 / (double)p->groundThetaBins;
  g.RB = p->groundRadiiBins;
  g.TB = p->groundThetaBins;
  g.inv_radial_step_f = g.inv_theta_step_f = 0.f;
  ground_geom_thresholds(g);
  for (int i = 0; i < n; ++i)
    rb[i] = g.r2_bins < 0 ? -2 : ground_radial_bin_of_r2(g, pts[i].x * pts[i].x + pts[i].y * pts[i].y);
}
// angleCheck && heightCheck (sloam.cpp:401-409) as plane_finish_kernel evaluates it
int hd_plane_accept(const sloam_pose *pose, const double *plane, const double *centroid, double tol) {
  return plane_accept(*pose, plane, centroid, tol) ? 1 : 0;
}
}
