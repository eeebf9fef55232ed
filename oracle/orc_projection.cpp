/* oracle/orc_projection.cpp -- stage a1/a2 (test infrastructure).
 * Restates Segmentation::_doProjection (sloam/src/segmentation/inference.cpp:80-165),
 * sort_indexes (include/segmentation/inference.h:232-243) and
 * Segmentation::maskCloud (inference.cpp:230-273). */
#include <algorithm>
#include <numeric>

#include "../include/sloam_b200_detmath.h"
#include "orc.h"

namespace orc {

namespace {
inline float ref_atan2f(const Options &o, float y, float x) {
  return o.use_libm ? std::atan2(y, x) : sloam_det::det_atan2f(y, x);
}
inline float ref_asinf(const Options &o, float v) {
  return o.use_libm ? std::asin(v) : sloam_det::det_asinf(v);
}
/* std::min(a,b) = (b<a)?b:a ; std::max(a,b) = (a<b)?b:a -- NaN in b is dropped */
inline float std_min(float a, float b) { return (b < a) ? b : a; }
inline float std_max(float a, float b) { return (a < b) ? b : a; }
}  // namespace

void project(const Options &o, const Pt *pts, int n, int32_t *pix, float *range_image) {
  const int W = o.p.img_w, H = o.p.img_h;
  /* constructor, inference.cpp:7-9: double arithmetic stored to float members */
  const float fov_up = (float)(o.p.fov_up_deg / 180.0 * M_PI);
  const float fov_down = (float)(o.p.fov_down_deg / 180.0 * M_PI);
  const float fov = std::abs(fov_down) + std::abs(fov_up);

  std::vector<float> ranges(n);
  for (int i = 0; i < n; ++i) {
    const float x = pts[i].x, y = pts[i].y, z = pts[i].z;
    const float range = std::sqrt(x * x + y * y + z * z); /* :103 */
    ranges[i] = range;
    const float yaw = -ref_atan2f(o, y, x);      /* :107 */
    const float pitch = ref_asinf(o, z / range); /* :108 */
    /* :111-112: double expressions assigned to float */
    float proj_x = (float)(0.5 * ((double)yaw / M_PI + 1.0));
    float proj_y = (float)(1.0 - (double)((pitch + std::abs(fov_down)) / fov));
    proj_x *= (float)W; /* :115-116 */
    proj_y *= (float)H;
    proj_x = std::floor(proj_x); /* :119-122 */
    proj_x = std_min((float)W - 1.0f, proj_x);
    proj_x = std_max(0.0f, proj_x);
    proj_y = std::floor(proj_y); /* :124-127 */
    proj_y = std_min((float)H - 1.0f, proj_y);
    proj_y = std_max(0.0f, proj_y);
    /* :131-132 keeps proj_xs/proj_ys in input order; maskCloud indexes with
     * proj_ys[i] * W + proj_xs[i] (:242) */
    pix[i] = (int32_t)(proj_y * (float)W + proj_x);
  }
  if (!range_image) return;
  /* :135 order by decreasing range, :160-162 overwrite so the closest point
   * wins.  Points with a NaN range are excluded (std::sort with NaN keys is
   * undefined behaviour in the reference). */
  std::vector<size_t> order;
  order.reserve(n);
  for (int i = 0; i < n; ++i)
    if (ranges[i] == ranges[i]) order.push_back((size_t)i);
  std::sort(order.begin(), order.end(),
            [&ranges](size_t a, size_t b) { return ranges[a] > ranges[b]; });
  std::fill(range_image, range_image + (size_t)W * H, 0.0f); /* :150-158 */
  for (size_t idx : order) range_image[pix[idx]] = ranges[idx];
}

/* Segmentation::_destaggerCloud, inference.cpp:200-228, statement for statement.  `out`
 * starts as a copy of `in` (pcl::copyPointCloud, :255).  The bound test `im_col > W` lets
 * col + 32 == W through: the write lands on the first pixel of the next row (SURVEY B-12),
 * which the next row's own pass overwrites; for an odd H the last such write is past the end
 * of the cloud (undefined behaviour in the reference) and is dropped here. */
static void destagger_cloud(const Cloud &in, Cloud &out, int H, int W) {
  bool col_valid = true;
  for (int irow = 0; irow < H; irow++) {
    for (int icol = 0; icol < W; icol++) {
      int im_col = icol;
      if (irow % 2 == 0) {
        im_col += 32;
        if (im_col < 0 || im_col > W) {
          col_valid = false;
          im_col = im_col % W;
        }
      }
      if (col_valid) {
        const Pt &pt = in[(size_t)irow * W + icol];
        const size_t dst = (size_t)irow * W + im_col;
        if (dst < out.size()) {
          out[dst].x = pt.x;
          out[dst].y = pt.y;
          out[dst].z = pt.z;
        }
      }
      col_valid = true;
    }
  }
}

void mask_cloud(const Options &o, const Pt *pts, int n, const int32_t *pix,
                const uint8_t *mask, Cloud &tree, Cloud &ground) {
  /* sloamNode.cpp:212: maskCloud(cloud, mask, ground, 1)          -> sparse
   * sloamNode.cpp:215: maskCloud(cloud, mask, tree, 255, dense)  -> organized
   * inference.cpp:241-253 */
  tree.assign(n, Pt());
  ground.clear();
  const float qnan = std::numeric_limits<float>::quiet_NaN();
  for (int i = 0; i < n; ++i) {
    const uint8_t m = mask[pix[i]];
    if (m == 1) ground.push_back(pts[i]);
    if (m == 255) {
      tree[i] = pts[i];
    } else {
      Pt p; /* default-constructed PointXYZI: intensity 0 */
      p.x = p.y = p.z = qnan;
      tree[i] = p;
    }
  }
  if (o.p.do_destagger && n == o.p.img_h * o.p.img_w) { /* inference.cpp:256-259: the dense cloud only */
    Cloud out = tree;
    destagger_cloud(tree, out, o.p.img_h, o.p.img_w);
    tree.swap(out);
  }
}

}  // namespace orc
