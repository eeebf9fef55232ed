/* oracle/orc_sloam.cpp -- stages a8, a12, a13, a18, a19 (test infrastructure).
 * Restates sloam::computeModels / matchFeatures / addFeatureMatches /
 * matchModels / projectModels / RunSloam (sloam/src/core/sloam.cpp:257-532). */
#include "orc.h"

namespace orc {

void Sloam::computeModels(SloamInput &in, std::vector<Cylinder> &landmarks,
                          std::vector<Plane> &planes) {
  /* sloam.cpp:388-436 */
  const sloam_params &fm = o_.p;
  std::vector<Cloud> cells;
  std::vector<int> n_cell;
  bin_ground_points(o_, V3{0, 0, 0} /* SE3() :392 */, in.groundCloud.data(),
                    (int)in.groundCloud.size(), cells, n_cell);
  cellPlanes.clear();
  cellAccepted.clear();
  for (int r = 0; r < fm.groundRadiiBins; ++r)
    for (int t = 0; t < fm.groundThetaBins; ++t) {
      const size_t c = (size_t)r * fm.groundThetaBins + t;
      Plane ground = make_plane(cells[c], fm.numGroundFeatures); /* :400 */
      ground.n_cell = n_cell[c];
      /* In the reference angleCheck/heightCheck are evaluated on an
       * uninitialised model when !isValid; only the conjunction matters. */
      const bool ok = ground.isValid && plane_accept(o_, in.poseEstimate, ground);
      cellPlanes.push_back(ground);
      cellAccepted.push_back(ok ? 1 : 0);
      if (ok) planes.push_back(ground); /* :409-410 */
    }
  treeModels.clear();
  if (planes.empty()) return; /* :414 */
  for (const std::vector<TreeVertex> &t : in.landmarks) { /* :418-435 */
    const Pt approxTreePos = t[1].coords;
    double bestDist = 100000;
    int bestPlane = 0;
    for (size_t g = 0; g < planes.size(); ++g) {
      const double d = plane_distance_point(planes[g].model, approxTreePos);
      if (d < bestDist) { bestDist = d; bestPlane = (int)g; }
    }
    Cylinder c = make_cylinder(o_, t, planes[bestPlane]);
    c.plane_index = bestPlane;
    treeModels.push_back(c);
    if (c.isValid) landmarks.push_back(c);
  }
}

namespace {
/* matchFeatures<Cylinder>, sloam.cpp:257-296 */
std::vector<TreeMatch> match_tree_features(const SE3 &tf, const std::vector<Cylinder> &curr,
                                           const std::vector<Cylinder> &map, double thresh) {
  std::vector<TreeMatch> matches;
  if (map.empty()) return matches; /* B-15 */
  for (const Cylinder &co : curr) {
    Cylinder proj = co;
    cylinder_project(proj, tf);
    double bestDist = thresh + 100;
    size_t best = 0;
    for (size_t k = 0; k < map.size(); ++k) {
      const double d = cylinder_distance_model(map[k].model, proj.model);
      if (d < bestDist) { bestDist = d; best = k; }
    }
    if (bestDist < thresh)
      for (const Pt &f : co.features) /* addFeatureMatches :288-296, ObjectMatch sloam.h:15-22 */
        matches.push_back({V3{f.x, f.y, f.z}, map[best].model});
  }
  return matches;
}
/* matchFeatures<Plane> */
std::vector<PlaneMatch> match_plane_features(const SE3 &tf, const std::vector<Plane> &curr,
                                             const std::vector<Plane> &map, double thresh) {
  std::vector<PlaneMatch> matches;
  if (map.empty()) return matches; /* B-15 */
  for (const Plane &co : curr) {
    Plane proj = co;
    plane_project(proj, tf);
    double bestDist = thresh + 100;
    size_t best = 0;
    for (size_t k = 0; k < map.size(); ++k) {
      const double d = plane_distance_model(map[k].model, proj.model);
      if (d < bestDist) { bestDist = d; best = k; }
    }
    if (bestDist < thresh)
      for (const Pt &f : co.features) matches.push_back({V3{f.x, f.y, f.z}, map[best].model});
  }
  return matches;
}
}  // namespace

bool Sloam::RunSloam(SloamInput &in, SloamOutput &out) {
  /* sloam.cpp:453-532 */
  const sloam_params &fm = o_.p;
  last = sloam_kf_result{};
  last.lm_termination[0] = last.lm_termination[1] = -1;
  std::vector<Plane> planes;
  std::vector<Cylinder> landmarks;
  computeModels(in, landmarks, planes);
  last.n_ground = (int)in.groundCloud.size();
  last.n_planes = (int)planes.size();
  last.n_trees = (int)in.landmarks.size();
  last.n_landmarks = (int)landmarks.size();
  std::vector<int> matchIndices(landmarks.size(), -1);
  bool success = true;
  if (firstScan) { /* :463-473 */
    for (Plane &p : planes) plane_project(p, in.poseEstimate);
    for (Cylinder &l : landmarks) cylinder_project(l, in.poseEstimate);
    out.T_Map_Curr = in.poseEstimate;
    out.matches = matchIndices;
    out.tm = landmarks;
    prevGPlanes = planes;
    firstScan = false;
  } else {
    if (in.mapModels.empty()) { /* :476-480 */
      last.status = SLOAM_KF_EMPTY_MAP;
      last.success = 0;
      return false;
    }
    if (planes.empty() || landmarks.empty()) { /* :482-486 */
      last.status = SLOAM_KF_NO_MODELS;
      last.success = 0;
      return false;
    }
    const std::vector<TreeMatch> treeMatches =
        match_tree_features(in.poseEstimate, landmarks, in.mapModels, fm.treeMatchThresh);
    const std::vector<PlaneMatch> planeMatches =
        match_plane_features(in.poseEstimate, planes, prevGPlanes, fm.plane_match_thresh);
    lastTreeMatches = treeMatches;
    lastPlaneMatches = planeMatches;
    last.n_tree_matches = (int)treeMatches.size();
    last.n_plane_matches = (int)planeMatches.size();
    SE3 T_Delta;
    SE3 currPose = in.poseEstimate;
    /* B-1: minPlanes_ is computed from the parameters actually set */
    const double minPlanes = (fm.groundRadiiBins * fm.groundThetaBins) * 0.1;
    const bool treeCheck = (double)landmarks.size() > fm.minTreeModels &&
                           (double)treeMatches.size() > 5.0 * fm.featuresPerTree; /* :499 */
    const bool groundCheck = (double)planeMatches.size() > fm.minGroundModels &&
                             (double)planes.size() > minPlanes; /* :500 */
    LMSummary s[2];
    if (fm.twoStepOptim) { /* :501-504 */
      success = two_step_optimize_pose(o_, in.poseEstimate, treeCheck, groundCheck, treeMatches,
                                       planeMatches, currPose, s);
      last.lm_iterations[0] = s[0].iterations; last.lm_termination[0] = s[0].termination;
      last.lm_iterations[1] = s[1].iterations; last.lm_termination[1] = s[1].termination;
    } else if (treeCheck && groundCheck) { /* :505-510 */
      success = optimize_pose(o_, in.poseEstimate, treeMatches, planeMatches, T_Delta, &s[0]);
      last.lm_iterations[0] = s[0].iterations; last.lm_termination[0] = s[0].termination;
      if (success) currPose = T_Delta;
    }
    for (Plane &p : planes) plane_project(p, currPose); /* :520 */
    for (Cylinder &l : landmarks) cylinder_project(l, currPose);
    /* matchModels, :298-328 */
    for (size_t i = 0; i < landmarks.size(); ++i) {
      double bestDist = fm.treeMatchThresh + 100;
      size_t bestKey = 0;
      for (size_t k = 0; k < in.mapModels.size(); ++k) {
        const double d = cylinder_distance_model(in.mapModels[k].model, landmarks[i].model);
        if (d < bestDist) { bestDist = d; bestKey = k; }
      }
      if (bestDist < fm.AddNewTreeThreshDist) matchIndices[i] = (int)bestKey;
    }
    prevGPlanes = planes; /* :525-529 */
    out.matches = matchIndices;
    out.T_Map_Curr = currPose;
    out.T_Delta = T_Delta;
    out.tm = landmarks;
  }
  last.status = success ? SLOAM_KF_OK : SLOAM_KF_NOT_CONVERGED;
  last.success = success ? 1 : 0;
  last.T_Map_Curr = pose_to_abi(out.T_Map_Curr);
  last.T_Delta = pose_to_abi(out.T_Delta);
  return success;
}

}  // namespace orc
