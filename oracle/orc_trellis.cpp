/* oracle/orc_trellis.cpp -- stages a6, a7 (test infrastructure).
 * Restates Instance::findClusters / findTrees / computeTreeVertex /
 * computeVertexProperties (sloam/src/segmentation/trellis.cpp:15-140) and the
 * PCL 1.10 OrganizedConnectedComponentSegmentation::segment +
 * EuclideanClusterComparator::compare it calls.
 * Pinned by the reference's four *_tree_*.pcd -> *_landmarks_* fixture pairs
 * (tests/test_oracle_golden.py). */
#include <algorithm>
#include <limits>

#include "orc.h"

namespace orc {

namespace {
/* Eigen Vector3f::norm(): sqrt(x^2 + (y^2 + z^2)) in float (Redux.h unroller) */
inline float dist3f(const Pt &a, const Pt &b) {
  const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
  return std::sqrt(dx * dx + (dy * dy + dz * dz));
}
inline unsigned find_root(const std::vector<unsigned> &runs, unsigned i) {
  while (runs[i] != i) i = runs[i];
  return i;
}
}  // namespace

void find_clusters(const Options &o, const Pt *pc, int H, int W, std::vector<uint32_t> &labels,
                   std::vector<std::vector<int>> &label_indices) {
  const uint32_t invalid = std::numeric_limits<uint32_t>::max();
  const size_t N = (size_t)H * W;
  labels.assign(N, invalid);
  label_indices.clear();
  if (N == 0) return; /* trellis.cpp:17 */
  const float thr = o.p.cluster_dist_thresh; /* trellis.cpp:23 */
  /* EuclideanClusterComparator::compare, depth_dependent = false */
  auto compare = [&](size_t a, size_t b) { return dist3f(pc[a], pc[b]) < thr; };
  std::vector<unsigned> run_ids;
  unsigned clust_id = 0;
  if (std::isfinite(pc[0].x)) {
    labels[0] = clust_id++;
    run_ids.push_back(labels[0]);
  }
  for (int c = 1; c < W; ++c) { /* first row */
    if (!std::isfinite(pc[c].x)) continue;
    if (compare(c, c - 1)) {
      labels[c] = labels[c - 1];
    } else {
      labels[c] = clust_id++;
      run_ids.push_back(labels[c]);
    }
  }
  for (int r = 1; r < H; ++r) {
    const size_t cur = (size_t)r * W, prev = cur - W;
    if (std::isfinite(pc[cur].x)) {
      if (compare(cur, prev)) {
        labels[cur] = labels[prev];
      } else {
        labels[cur] = clust_id++;
        run_ids.push_back(labels[cur]);
      }
    }
    for (int c = 1; c < W; ++c) {
      const size_t i = cur + c;
      if (!std::isfinite(pc[i].x)) continue;
      if (compare(i, i - 1)) labels[i] = labels[i - 1];
      if (compare(i, prev + c)) {
        if (labels[i] == invalid) {
          labels[i] = labels[prev + c];
        } else if (labels[prev + c] != invalid) {
          const unsigned r1 = find_root(run_ids, labels[i]);
          const unsigned r2 = find_root(run_ids, labels[prev + c]);
          if (r1 < r2) run_ids[r2] = r1; else run_ids[r1] = r2;
        }
      }
      if (labels[i] == invalid) {
        labels[i] = clust_id++;
        run_ids.push_back(labels[i]);
      }
    }
  }
  std::vector<unsigned> map(clust_id);
  unsigned max_id = 0;
  for (unsigned k = 0; k < run_ids.size(); ++k) {
    if (run_ids[k] == k) map[k] = max_id++;
    else map[k] = map[find_root(run_ids, k)];
  }
  label_indices.resize(max_id + 1); /* one empty tail entry, as in PCL */
  for (size_t i = 0; i < N; ++i)
    if (labels[i] != invalid) {
      labels[i] = map[labels[i]];
      label_indices[labels[i]].push_back((int)i);
    }
}

/* trellis.cpp:63-102 */
static bool vertex_properties(const Options &o, Cloud &pc, Cloud &filtered, Pt &median,
                              double &radius) {
  const int num_points = (int)pc.size();
  const int middle = (int)(num_points / 2.0);
  /* Three std::sort calls exactly like the reference (:71-82); libstdc++'s
   * introsort is what breaks ties (SURVEY B-3). */
  std::sort(pc.begin(), pc.end(), [](const Pt &a, const Pt &b) { return a.x < b.x; });
  const double mx = pc[middle].x;
  std::sort(pc.begin(), pc.end(), [](const Pt &a, const Pt &b) { return a.y < b.y; });
  const double my = pc[middle].y;
  std::sort(pc.begin(), pc.end(), [](const Pt &a, const Pt &b) { return a.z < b.z; });
  const double mz = pc[middle].z;
  median.x = (float)mx; median.y = (float)my; median.z = (float)mz;
  for (const Pt &p : pc) /* :89-93 pcl::euclideanDistance: float */
    if (dist3f(p, median) < o.p.max_dist_to_centroid) filtered.push_back(p);
  if (filtered.size() > 1) { /* :95-100 */
    radius = dist3f(filtered.front(), filtered.back());
    return true;
  }
  return false;
}

void compute_graph(const Options &o, const Pt *pc, int H, int W, Landmarks &landmarks) {
  std::vector<uint32_t> labels;
  std::vector<std::vector<int>> label_indices;
  find_clusters(o, pc, H, W, labels, label_indices); /* trellis.cpp:138 */
  /* findTrees, trellis.cpp:104-132 */
  for (size_t i = 0; i < label_indices.size(); ++i) {
    if ((int)label_indices[i].size() > o.p.min_cluster_points) { /* :109 */
      std::vector<TreeVertex> tree;
      for (int row = H - 1; row >= 0; --row) { /* :111 */
        Cloud beam;
        for (int col = 0; col < W; ++col)
          if (labels[(size_t)row * W + col] == i) beam.push_back(pc[(size_t)row * W + col]);
        if ((int)beam.size() > o.p.min_vertex_points) { /* :119 */
          TreeVertex v; /* computeTreeVertex :45-61 */
          double radius = 0;
          v.isValid = vertex_properties(o, beam, v.points, v.coords, radius);
          v.treeId = (int)i;
          v.radius = radius;
          v.row = row;
          if (v.isValid) tree.push_back(v);
        }
      }
      if ((int)tree.size() > o.p.min_tree_vertices) { /* :124-128 */
        if ((int)tree.size() > o.p.max_tree_vertices) tree.resize(o.p.max_tree_vertices);
        landmarks.push_back(tree);
      }
    }
  }
}

}  // namespace orc
