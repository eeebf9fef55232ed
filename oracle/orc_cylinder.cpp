/* oracle/orc_cylinder.cpp -- stages a8..a11 (test infrastructure).
 * Restates Cylinder::Cylinder / computeModel / groundBasedRoot / filter /
 * distance / project (sloam/src/objects/cylinder.cpp:3-211),
 * rayPlaneIntersection (include/helpers/utils.h:41-52) and the PCL 1.10 code
 * behind pcl::SACSegmentation<PointT> with SACMODEL_LINE / SAC_RANSAC
 * (cylinder.cpp:116-124): RandomSampleConsensus::computeModel,
 * SampleConsensusModel::drawIndexSample, SampleConsensusModelLine::
 * {isSampleGood, computeModelCoefficients, countWithinDistance,
 * selectWithinDistance, optimizeModelCoefficients}, pcl::eigen33,
 * pcl::computeCorrespondingEigenVector.
 * PCL is not available in this image: "believed-upstream semantics"; the only
 * reference pin is CylinderTest.DistanceToFeature (+-0.1 m). */
#include <algorithm>
#include <cfloat>
#include <climits>
#include <random>

#include "orc.h"

namespace orc {

namespace {
constexpr double PIDEF = 3.14159265;

inline float pow_2(double x) { return (float)(x * x); }
inline float euclideanDist2D(const Pt &a, const Pt &b) { /* utils.h:9-12 */
  return std::sqrt(pow_2(a.x - b.x) + pow_2(a.y - b.y));
}

struct F3 { float x, y, z; };

/* SampleConsensusModelLine::countWithinDistance / selectWithinDistance:
 * Eigen::Vector4f arithmetic (SSE2 packet reductions: (a0+a2)+(a1+a3)),
 * cross3, squared distance promoted to double and compared with thr^2. */
struct LineScorer {
  float px, py, pz, dx, dy, dz;
  double sqr_thr;
  LineScorer(const float coef[6], double thr) {
    px = coef[0]; py = coef[1]; pz = coef[2];
    float x = coef[3], y = coef[4], z = coef[5];
    /* line_dir.normalize(): Vector4f (x,y,z,0) */
    const float sq = (x * x + z * z) + (y * y + 0.0f * 0.0f);
    if (sq > 0) { const float n = std::sqrt(sq); x /= n; y /= n; z /= n; }
    dx = x; dy = y; dz = z;
    sqr_thr = thr * thr;
  }
  bool inlier(const F3 &p) const {
    const float ax = px - p.x, ay = py - p.y, az = pz - p.z;
    const float cx = ay * dz - az * dy;
    const float cy = az * dx - ax * dz;
    const float cz = ax * dy - ay * dx;
    const double sqr = (double)((cx * cx + cz * cz) + (cy * cy + 0.0f));
    return sqr < sqr_thr;
  }
};

/* SampleConsensusModelLine::computeModelCoefficients */
bool line_from_samples(const std::vector<F3> &pts, int i0, int i1, float coef[6]) {
  const F3 &a = pts[i0], &b = pts[i1];
  if (std::fabs(a.x - b.x) <= FLT_EPSILON && std::fabs(a.y - b.y) <= FLT_EPSILON &&
      std::fabs(a.z - b.z) <= FLT_EPSILON)
    return false;
  coef[0] = a.x; coef[1] = a.y; coef[2] = a.z;
  float x = b.x - a.x, y = b.y - a.y, z = b.z - a.z;
  /* model_coefficients.tail<3>().normalize(): x^2 + (y^2 + z^2) */
  const float sq = x * x + (y * y + z * z);
  if (sq > 0) { const float n = std::sqrt(sq); x /= n; y /= n; z /= n; }
  coef[3] = x; coef[4] = y; coef[5] = z;
  return true;
}

inline bool sample_good(const std::vector<F3> &pts, int i0, int i1) {
  /* PCL 1.10 SampleConsensusModelLine::isSampleGood: all three differ (&&) */
  return pts[i0].x != pts[i1].x && pts[i0].y != pts[i1].y && pts[i0].z != pts[i1].z;
}

/* SampleConsensusModel sampling state: boost::mt19937 seeded 12345u,
 * uniform_int<>(0, INT_MAX) == mt() >> 1; shuffled_indices_ persists. */
struct Sampler {
  std::mt19937 rng{12345u};
  std::vector<int> shuf;
  explicit Sampler(int n) : shuf(n) { for (int i = 0; i < n; ++i) shuf[i] = i; }
  void draw(int &s0, int &s1) {
    const size_t n = shuf.size();
    for (unsigned i = 0; i < 2; ++i) {
      const unsigned r = (unsigned)(rng() >> 1);
      std::swap(shuf[i], shuf[i + (r % (n - i))]);
    }
    s0 = shuf[0]; s1 = shuf[1];
  }
};

/* pcl::computeRoots (closed-form eigenvalues of a symmetric 3x3, float) */
void compute_roots2(float b, float c, float roots[3]) {
  roots[0] = 0.f;
  /* Scalar d = Scalar(b * b - 4.0 * c): the subtraction is done in double */
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  const float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}
void compute_roots(const float m[3][3], float roots[3]) {
  const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] -
                   m[0][0] * m[1][2] * m[1][2] - m[1][1] * m[0][2] * m[0][2] -
                   m[2][2] * m[0][1] * m[0][1];
  const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] -
                   m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2];
  const float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < FLT_EPSILON) {
    compute_roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = 1.0f / 3.0f;
  const float s_sqrt3 = std::sqrt(3.0f);
  const float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  const float rho = std::sqrt(-a_over_3);
  const float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
  const float cos_theta = std::cos(theta), sin_theta = std::sin(theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0) compute_roots2(c2, c1, roots);
}

/* SampleConsensusModelLine::optimizeModelCoefficients (float32 PCA refit) */
void refit_line(const std::vector<F3> &pts, const std::vector<int> &inl, const float in[6],
                float out[6]) {
  for (int i = 0; i < 6; ++i) out[i] = in[i];
  if (inl.size() <= 2) return;
  /* compute3DCentroid: float sequential sums */
  float cx = 0, cy = 0, cz = 0;
  for (int i : inl) { cx += pts[i].x; cy += pts[i].y; cz += pts[i].z; }
  const float nf = (float)inl.size();
  cx /= nf; cy /= nf; cz /= nf;
  /* computeCovarianceMatrix (un-normalised) */
  float C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i : inl) {
    float x = pts[i].x - cx, y = pts[i].y - cy, z = pts[i].z - cz;
    C[1][1] += y * y; C[1][2] += y * z; C[2][2] += z * z;
    const float s = x;
    x *= s; y *= s; z *= s;
    C[0][0] += x; C[0][1] += y; C[0][2] += z;
  }
  C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
  out[0] = cx; out[1] = cy; out[2] = cz;
  /* pcl::eigen33(mat, evals): scale, computeRoots, unscale */
  float scale = 0;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(C[i][j]));
  if (scale <= FLT_MIN) scale = 1.0f;
  float S[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[i][j] = C[i][j] / scale;
  float ev[3];
  compute_roots(S, ev);
  for (float &e : ev) e *= scale;
  /* pcl::computeCorrespondingEigenVector(mat, ev[2], v) */
  const float lam = ev[2] / scale;
  S[0][0] -= lam; S[1][1] -= lam; S[2][2] -= lam;
  auto crossf = [](const float a[3], const float b[3], float r[3]) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = a[2] * b[0] - a[0] * b[2];
    r[2] = a[0] * b[1] - a[1] * b[0];
  };
  float v1[3], v2[3], v3[3];
  crossf(S[0], S[1], v1); crossf(S[0], S[2], v2); crossf(S[1], S[2], v3);
  auto sq = [](const float v[3]) { return v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]); };
  const float l1 = sq(v1), l2 = sq(v2), l3 = sq(v3);
  const float *v; float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  else { v = v3; l = l3; }
  const float nl = std::sqrt(l);
  out[3] = v[0] / nl; out[4] = v[1] / nl; out[5] = v[2] / nl;
}
}  // namespace

void ransac_draw_table(int n, int n_draws, std::vector<int32_t> &pairs) {
  Sampler s(n);
  pairs.resize((size_t)2 * n_draws);
  for (int t = 0; t < n_draws; ++t) {
    int a, b;
    s.draw(a, b);
    pairs[2 * t] = a; pairs[2 * t + 1] = b;
  }
}

/* pcl::SACSegmentation::segment with SACMODEL_LINE/SAC_RANSAC + optimize.
 * Returns false when no model was found (inliers empty). */
static bool line_ransac(const Options &o, const std::vector<F3> &pts, float coef_out[6],
                        Cylinder &diag) {
  const int n = (int)pts.size();
  const double thr = o.p.ransac_threshold;
  const int max_iterations = o.p.ransac_max_iterations;
  const int fixed = o.p.ransac_fixed_hypotheses;
  if (n < 2) return false; /* getSamples: indices < sample size */
  Sampler sampler(n);
  int iterations = 0, best = -INT_MAX, skipped = 0;
  double k = 1.0;
  const double log_probability = std::log(1.0 - o.p.ransac_probability);
  const double one_over_indices = 1.0 / (double)n;
  const unsigned max_skip = (unsigned)max_iterations * 10u;
  float best_coef[6];
  bool have_model = false;
  int best_hyp = -1;
  while (true) {
    if (fixed > 0) {
      if (iterations >= fixed) break; /* stress mode: exactly `fixed` hypotheses */
    } else {
      if (!((double)iterations < k && (unsigned)skipped < max_skip)) break;
    }
    /* getSamples: up to 1000 draws until isSampleGood */
    int s0 = -1, s1 = -1;
    bool got = false;
    for (int it = 0; it < 1000; ++it) {
      sampler.draw(s0, s1);
      if (sample_good(pts, s0, s1)) { got = true; break; }
    }
    if (!got) break;
    float coef[6];
    if (!line_from_samples(pts, s0, s1, coef)) {
      ++skipped;
      if (fixed > 0 && (unsigned)skipped >= max_skip) break;
      continue;
    }
    LineScorer sc(coef, thr);
    int cnt = 0;
    for (const F3 &p : pts) cnt += sc.inlier(p) ? 1 : 0;
    if (cnt > best) { /* strict: first maximum wins */
      best = cnt;
      for (int i = 0; i < 6; ++i) best_coef[i] = coef[i];
      have_model = true;
      best_hyp = iterations;
      const double w = (double)best * one_over_indices;
      double p_no_outliers = 1.0 - std::pow(w, 2.0);
      p_no_outliers = std::max(DBL_EPSILON, p_no_outliers);
      p_no_outliers = std::min(1.0 - DBL_EPSILON, p_no_outliers);
      k = log_probability / std::log(p_no_outliers);
    }
    ++iterations;
    if (fixed <= 0 && iterations > max_iterations) break;
  }
  diag.n_hypotheses = iterations;
  diag.best_hypothesis = best_hyp;
  if (!have_model) return false;
  diag.n_inliers = best;
  std::vector<int> inliers;
  {
    LineScorer sc(best_coef, thr);
    for (int i = 0; i < n; ++i) if (sc.inlier(pts[i])) inliers.push_back(i);
  }
  float refined[6];
  refit_line(pts, inliers, best_coef, refined);
  {
    LineScorer sc(refined, thr);
    inliers.clear();
    for (int i = 0; i < n; ++i) if (sc.inlier(pts[i])) inliers.push_back(i);
  }
  diag.n_refit_inliers = (int)inliers.size();
  for (int i = 0; i < 6; ++i) coef_out[i] = refined[i];
  return !inliers.empty();
}

/* cylinder.cpp:75-173 */
static void compute_model(const Options &o, Cylinder &c, const std::vector<TreeVertex> &vtxs) {
  const TreeVertex &firstVtx = vtxs[2];
  std::vector<float> validRadii;
  std::vector<F3> tree;
  for (const TreeVertex &vtx : vtxs) { /* :82-98 */
    tree.push_back({vtx.coords.x, vtx.coords.y, vtx.coords.z});
    for (Pt p : vtx.points) {
      p.intensity = (float)firstVtx.treeId;
      c.features.push_back(p);
    }
    c.model.radii.push_back(vtx.radius);
    if (vtx.points.size() > 3) validRadii.push_back((float)vtx.radius);
  }
  c.id = (size_t)firstVtx.treeId;
  const Pt bottom = vtxs[1].coords, top = vtxs[vtxs.size() - 2].coords;
  c.model.root = {bottom.x, bottom.y, bottom.z};
  /* :109 pcl::geometry::squaredDistance: float (x^2 + (y^2 + z^2)) */
  {
    const float dx = top.x - bottom.x, dy = top.y - bottom.y, dz = top.z - bottom.z;
    const float sq = dx * dx + (dy * dy + dz * dz);
    if ((double)sq < o.p.min_tree_height_sq || tree.empty()) {
      c.model.radius = -1;
      return;
    }
  }
  float coef[6];
  if (!line_ransac(o, tree, coef, c)) { /* :126-131 */
    c.model.radius = -1;
    return;
  }
  c.model.ray = {coef[3], coef[4], coef[5]}; /* :137-139 */
  if (!validRadii.empty()) { /* :141-159 */
    std::sort(validRadii.begin(), validRadii.end());
    int d = 0;
    while (d < (int)validRadii.size() && !(validRadii[d] > 0)) ++d;
    const int middle = std::min((int)(validRadii.size() - 1),
                                (d + 1) + (int)((validRadii.size() - (size_t)d) / 2));
    c.model.radius = validRadii[middle];
    if (c.model.radius == 0) c.model.radius = -1;
    else if (c.model.radius < o.p.defaultTreeRadius) c.model.radius = o.p.defaultTreeRadius;
  } else {
    c.model.radius = -1;
  }
  c.features.resize(o.p.featuresPerTree); /* :172 zero-pads (B-6) */
}

Cylinder make_cylinder(const Options &o, const std::vector<TreeVertex> &vertices,
                       const Plane &gplane) {
  Cylinder c; /* cylinder.cpp:3-28 */
  Pt centroid;
  centroid.x = (float)gplane.model.centroid.x;
  centroid.y = (float)gplane.model.centroid.y;
  centroid.z = (float)gplane.model.centroid.z;
  const TreeVertex &firstVtx = vertices[1];
  const bool withinMaxDist = euclideanDist2D(firstVtx.coords, centroid) < o.p.maxLidarDist;
  c.isValid = false;
  /* id and ray are uninitialised in the reference on the early-return paths; the
   * oracle defines them as vertices[2].treeId and zero (the cylinder is invalid there). */
  c.id = (size_t)vertices[2].treeId;
  if (!withinMaxDist) return c;
  compute_model(o, c, vertices);
  /* groundBasedRoot, :41-55 */
  const double *g = gplane.model.plane;
  bool validZ = false;
  {
    const V3 nv{g[0], g[1], g[2]};
    const float dist =
        (float)(std::fabs(g[0] * c.model.root.x + g[1] * c.model.root.y + g[2] * c.model.root.z + g[3]) /
                norm(nv));
    if ((double)dist < o.p.root_plane_max_dist) {
      /* rayPlaneIntersection, utils.h:41-52 (float denom and t) */
      const float denom = (float)dot(nv, c.model.ray);
      if (std::fabs(denom) > 0.001f) {
        const float t = (float)(dot(gplane.model.centroid - c.model.root, nv) / (double)denom);
        if (t >= 0.001f) c.model.root = c.model.root + (double)t * c.model.ray;
      }
      validZ = true;
    }
  }
  const bool validNorm = norm(c.model.root) > 0.01; /* :22 */
  const bool validRadius = c.model.radius > 0.0;    /* :23 */
  /* filter, :57-73 */
  bool validTree = false;
  if (c.model.radius != -1) {
    const V3 up{g[0], g[1], g[2]};
    const double theta =
        (180 / PIDEF) * std::acos(dot(c.model.ray, up) / (norm(c.model.ray) * norm(up)));
    if (c.model.radius < o.p.maxTreeRadius &&
        (theta <= o.p.maxAxisTheta || theta >= 180 - o.p.maxAxisTheta))
      validTree = true;
  }
  c.isValid = validZ && validRadius && validTree && validNorm; /* :26 */
  return c;
}

double cylinder_distance_model(const CylinderParameters &m, const CylinderParameters &tgt) {
  /* cylinder.cpp:175-194 */
  const double heights[3] = {0.0, 3.0, 6.0};
  double distance = 0.0;
  for (double h : heights) {
    const double src_t = (h - m.root.z) / m.ray.z;
    const V3 a = m.root + src_t * m.ray;
    const double tgt_t = (h - tgt.root.z) / tgt.ray.z;
    const V3 b = tgt.root + tgt_t * tgt.ray;
    distance += norm(a - b);
  }
  return distance / 3.0;
}

double cylinder_distance_point(const CylinderParameters &m, const Pt &p) { /* :196-203 */
  const V3 e{p.x, p.y, p.z};
  const V3 proj = m.root + (dot(e - m.root, m.ray) / dot(m.ray, m.ray)) * m.ray;
  return norm(e - proj) - m.radius;
}

void cylinder_project(Cylinder &c, const SE3 &tf) { /* :205-211 */
  V3 other = c.model.root + c.model.ray;
  c.model.root = se3_apply(tf, c.model.root);
  other = se3_apply(tf, other);
  c.model.ray = other - c.model.root;
}

}  // namespace orc
