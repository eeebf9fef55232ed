/* oracle/orc_map.cpp -- MapManager (test infrastructure), SURVEY 8(f)-1.
 * Restates sloam/src/core/mapManager.cpp:8-71 (updateMap, getMap, getSubmap) with the
 * pcl::KdTreeFLANN<PointT>::nearestKSearch call (:44-55) replaced by an exact brute-force
 * k-nearest search: FLANN's single kd-tree search with default parameters is exact, returns
 * the neighbours sorted by squared L2 distance over (x, y, z) in float
 * (L2_Simple: ((dx^2) + dy^2) + dz^2); ties are broken here by the lower index. */
#include <algorithm>
#include <map>

#include "orc.h"

namespace orc {

struct MapManager {
  std::vector<Pt> landmarks;          /* landmarks_ : tree roots */
  std::vector<sloam_cylinder> models; /* treeModels_ (the part association reads) */
  std::vector<int> ids;
  std::vector<size_t> hits;           /* treeHits_ */
  std::map<int, int> matchesMap;
  int knn = 100, recent = 200;        /* mapManager.cpp:54,60 */

  void getSubmap(const SE3 &pose, std::vector<sloam_cylinder> &submap, std::vector<int> &map_index) {
    if (landmarks.empty()) return; /* :42 */
    Pt q; q.x = (float)pose.t.x; q.y = (float)pose.t.y; q.z = 1.0f; /* :51-53 */
    std::vector<std::pair<float, int>> d(landmarks.size());
    for (size_t i = 0; i < landmarks.size(); ++i) {
      const float dx = landmarks[i].x - q.x, dy = landmarks[i].y - q.y, dz = landmarks[i].z - q.z;
      d[i] = {(dx * dx + dy * dy) + dz * dz, (int)i};
    }
    const size_t kk = std::min<size_t>(knn, d.size());
    std::partial_sort(d.begin(), d.begin() + kk, d.end());
    int idx_count = 0;
    const size_t map_size = models.size();
    for (size_t j = 0; j < kk; ++j) { /* :57-66 */
      const int map_idx = d[j].second;
      if (map_size - (size_t)map_idx < (size_t)recent) {
        matchesMap.insert({idx_count, map_idx});
        submap.push_back(models[map_idx]);
        map_index.push_back(map_idx);
        idx_count++;
      }
    }
  }

  void updateMap(const std::vector<sloam_cylinder> &obs, const std::vector<int> &obs_ids,
                 const std::vector<int> &matches) { /* :8-28 */
    for (size_t i = 0; i < obs.size(); ++i) {
      Pt pt; pt.x = (float)obs[i].root[0]; pt.y = (float)obs[i].root[1]; pt.z = (float)obs[i].root[2];
      if (matches[i] == -1) {
        landmarks.push_back(pt); models.push_back(obs[i]); ids.push_back(obs_ids[i]); hits.push_back(1);
      } else {
        const int matchIdx = matchesMap[matches[i]]; /* operator[]: 0 when the key is missing */
        landmarks[matchIdx] = pt; models[matchIdx] = obs[i]; ids[matchIdx] = obs_ids[i]; hits[matchIdx] += 1;
      }
    }
    matchesMap.clear();
  }
};

}  // namespace orc

using orc::MapManager;

extern "C" {
void *orc_map_create() { return new MapManager(); }
void orc_map_destroy(void *m) { delete static_cast<MapManager *>(m); }
int orc_map_size(void *m) { return (int)static_cast<MapManager *>(m)->models.size(); }
int orc_map_get_submap(void *m, const sloam_pose *pose, sloam_cylinder *submap, int32_t *map_index, int cap) {
  std::vector<sloam_cylinder> s; std::vector<int> mi;
  static_cast<MapManager *>(m)->getSubmap(orc::pose_from_abi(*pose), s, mi);
  const int n = std::min<int>((int)s.size(), cap);
  for (int i = 0; i < n; ++i) { submap[i] = s[i]; map_index[i] = mi[i]; }
  return n;
}
void orc_map_update(void *m, const sloam_cylinder *obs, const int32_t *ids, const int32_t *matches, int n) {
  std::vector<sloam_cylinder> o(obs, obs + n);
  std::vector<int> id(ids, ids + n), mt(matches, matches + n);
  static_cast<MapManager *>(m)->updateMap(o, id, mt);
}
/* getMap (:30-39): models with more than 2 hits; also dumps the whole state for tests */
int orc_map_dump(void *m, sloam_cylinder *models, int32_t *hits, int cap) {
  MapManager *mm = static_cast<MapManager *>(m);
  const int n = std::min<int>((int)mm->models.size(), cap);
  for (int i = 0; i < n; ++i) { models[i] = mm->models[i]; hits[i] = (int32_t)mm->hits[i]; }
  return n;
}
}
