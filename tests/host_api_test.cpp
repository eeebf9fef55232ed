// host_api_test.cpp -- the reference's gtest cases (sloam/src/tests/{plane,cylinder,core}_test.cpp)
// written against sloam_b200/host/sloam_host.h: same classes, same calls, same assertions,
// executed on the GPU through the C ABI.  Fixtures (the reference's still_* files converted
// by scripts/make_golden.py) arrive as one binary blob written by tests/test_host_api.py.
//
//   g++ -std=c++17 -O1 tests/host_api_test.cpp -Lsloam_b200/lib -lsloam_b200 -o host_api_test
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "../sloam_b200/host/sloam_host.h"

static int g_fail = 0, g_run = 0;
#define EXPECT_TRUE(c) do { ++g_run; if (!(c)) { ++g_fail; std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #c); } } while (0)
#define EXPECT_NEAR(a, b, t) EXPECT_TRUE(std::fabs((double)(a) - (double)(b)) <= (t))
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))

struct Fixture {
  VectorType ground;
  std::vector<std::vector<TreeVertex>> landmarks;
};

static bool read_fixture(std::ifstream &f, Fixture &fx) {
  int32_t ng = 0, nt = 0;
  f.read((char *)&ng, 4);
  fx.ground.resize(ng);
  f.read((char *)fx.ground.data(), sizeof(PointT) * (size_t)ng);
  f.read((char *)&nt, 4);
  for (int t = 0; t < nt; ++t) {
    int32_t nv = 0;
    f.read((char *)&nv, 4);
    std::vector<TreeVertex> tree(nv);
    for (auto &v : tree) {
      int32_t id = 0, np = 0; float radius = 0, c[3];
      f.read((char *)&id, 4); f.read((char *)&radius, 4); f.read((char *)c, 12); f.read((char *)&np, 4);
      v.treeId = id; v.radius = radius; v.isValid = true;
      v.coords.x = c[0]; v.coords.y = c[1]; v.coords.z = c[2];
      v.points.resize(np);
      f.read((char *)v.points.data(), sizeof(PointT) * (size_t)np);
    }
    fx.landmarks.push_back(tree);
  }
  return (bool)f;
}

static PointT pt(float x, float y, float z) { PointT p; p.x = x; p.y = y; p.z = z; return p; }

// ---------------------------------------------------------------- plane_test.cpp
static FeatureModelParams plane_params() {  // plane_test.cpp:28-52
  FeatureModelParams p;
  p.maxLidarDist = 15.0; p.maxGroundLidarDist = 30.0; p.minGroundLidarDist = 0.0;
  p.groundRadiiBins = 1; p.groundThetaBins = 18; p.groundRetainThresh = 0.1; p.numGroundFeatures = 1;
  p.treeMatchThresh = 1.0; p.AddNewTreeThreshDist = 2.0; p.featuresPerTree = 2; p.defaultTreeRadius = 0.1;
  return p;
}
static void plane_tests() {
  const FeatureModelParams params = plane_params();
  VectorType g{pt(0, 0, 0), pt(0, 1, 0), pt(1, 0, 0)};
  {  // PlaneTest.Initalizes :57-66
    Plane plane(g, params);
    EXPECT_TRUE(plane.isValid);
  }
  {  // PlaneTest.DistanceToFeature :68-80
    Plane plane(g, params);
    float d = plane.distance(g[0]);
    EXPECT_NEAR(0.0, d, 0.1);
    EXPECT_NEAR(1.0, std::fabs(plane.model.plane[2]), 1e-12);
  }
  {  // PlaneTest.TranslateModel :82-97
    Plane plane(g, params);
    SE3 tf;
    tf.translation()[2] = 1;
    auto centroid = plane.model.centroid;
    plane.project(tf);
    EXPECT_EQ(centroid[2] + 1, plane.model.centroid[2]);
  }
}

// ------------------------------------------------------------- cylinder_test.cpp
static void cylinder_tests(const Fixture &t0) {
  FeatureModelParams params;  // cylinder_test.cpp:38-62
  params.maxLidarDist = 30.0; params.maxGroundLidarDist = 30.0; params.minGroundLidarDist = 0.0;
  params.groundRadiiBins = 1; params.groundThetaBins = 1; params.groundRetainThresh = 0.1;
  params.treeMatchThresh = 1.0; params.AddNewTreeThreshDist = 2.0;
  params.featuresPerTree = 2; params.numGroundFeatures = 3; params.defaultTreeRadius = 0.1;
  VectorType gf{pt(0, 0, 0), pt(0, 1, 0), pt(1, 0, 0), pt(1, 1, 0)};
  Plane plane(gf, params);
  EXPECT_TRUE(plane.isValid);
  std::vector<Cylinder> cylinders;
  for (auto l : t0.landmarks) {  // Initalizes :69-77
    auto c = Cylinder(l, plane, params);
    if (c.isValid) cylinders.push_back(c);
  }
  EXPECT_TRUE(cylinders.size() > 0);
  if (cylinders.empty()) return;
  {  // DistanceToModel :79-93
    float d = cylinders[0].distance(cylinders[0].model);
    EXPECT_EQ(0.0, d);
  }
  {  // DistanceToFeature :95-108
    float d = cylinders[0].distance(cylinders[0].features[0]);
    EXPECT_NEAR(0.0, d, 0.1);
  }
  {  // TranslateModel :110-127
    SE3 tf;
    tf.translation()[0] = 1;
    auto root = cylinders[0].model.root;
    cylinders[0].project(tf);
    EXPECT_EQ(root[0] + 1, cylinders[0].model.root[0]);
  }
}

// ----------------------------------------------------------------- core_test.cpp
static FeatureModelParams core_params(bool two_step) {  // core_test.cpp:94-119 + YAML for the unset ones
  FeatureModelParams p;
  p.maxLidarDist = 15.0; p.maxGroundLidarDist = 30.0; p.minGroundLidarDist = 0.0;
  p.groundRadiiBins = 1; p.groundThetaBins = 18; p.groundMatchThresh = 2.0; p.groundRetainThresh = 0.05;
  p.maxTreeRadius = 0.3; p.maxAxisTheta = 10; p.roughTreeMatchThresh = 3.0; p.treeMatchThresh = 1.0;
  p.AddNewTreeThreshDist = 2.0; p.featuresPerTree = 2; p.numGroundFeatures = 60; p.defaultTreeRadius = 0.1;
  p.minTreeModels = 5; p.minGroundModels = 36; p.twoStepOptim = two_step;
  return p;
}
static SloamInput make_input(const Fixture &fx) {  // readInputData :72-92
  SloamInput in;
  in.landmarks = fx.landmarks;
  for (auto &tree : in.landmarks)
    for (auto &vtx : tree) vtx.points.resize(5);
  in.groundCloud->points = fx.ground;
  in.poseEstimate = SE3();
  return in;
}
static void core_tests(const Fixture &t0, const Fixture &t1, bool two_step) {
  const FeatureModelParams params = core_params(two_step);
  {  // FirstScan / SecondScan :128-146
    for (const Fixture *fx : {&t0, &t1}) {
      sloam::sloam s;
      SloamOutput out;
      s.setFmParams(params);
      SloamInput in = make_input(*fx);
      s.RunSloam(in, out);
      EXPECT_TRUE(out.tm.size() > 0);
      EXPECT_TRUE(s.getPrevGroundModel().size() > 0);
    }
  }
  {  // SLOAMSucess, PoseOptimization, ObjectAssociation :148-194
    sloam::sloam s;
    SloamOutput out;
    s.setFmParams(params);
    SloamInput in0 = make_input(t0), in1 = make_input(t1);
    s.RunSloam(in0, out);
    in1.mapModels = out.tm;  // first scan output is the initial map
    bool success = s.RunSloam(in1, out);
    EXPECT_TRUE(success);
    float translation = out.T_Delta.translation().norm();
    std::printf("  two_step=%d  |T_Delta.t| = %.6f  |T_Map_Curr.t| = %.6f  lm iters %d/%d\n", (int)two_step,
                translation, out.T_Map_Curr.translation().norm(), s.lastResult().lm_iterations[0],
                s.lastResult().lm_iterations[1]);
    EXPECT_NEAR(0.0, translation, two_step ? 0.1 : 0.2);  // see tests/test_oracle_reference_tests.py
    bool matched = false;
    for (auto m : out.matches) matched = matched || m != -1;
    EXPECT_TRUE(matched);
  }
  {  // not the first scan but the map is empty -> false (sloam.cpp:476-480)
    sloam::sloam s;
    SloamOutput out;
    s.setFmParams(params);
    SloamInput in0 = make_input(t0), in1 = make_input(t1);
    s.RunSloam(in0, out);
    SloamOutput out2;
    EXPECT_TRUE(!s.RunSloam(in1, out2));
    EXPECT_TRUE(out2.tm.empty());
  }
}

// The public per-stage methods of sloam::sloam (sloam.h:71-96): RunSloam's own body
// (sloam.cpp:453-532) re-assembled from them must give what RunSloam gives.
#include <chrono>
static void stage_method_tests(const Fixture &t0, const Fixture &t1, bool two_step) {
  const FeatureModelParams params = core_params(two_step);
  // ---- the fused entry
  sloam::sloam ref;
  ref.setFmParams(params);
  SloamOutput r0, r1;
  SloamInput a0 = make_input(t0), a1 = make_input(t1);
  ref.RunSloam(a0, r0);
  a1.mapModels = r0.tm;
  const bool ok_ref = ref.RunSloam(a1, r1);
  // ---- the same from the stage methods
  sloam::sloam s;
  s.setFmParams(params);
  SloamInput in0 = make_input(t0), in1 = make_input(t1);
  std::vector<Cylinder> lm0, lm1;
  std::vector<Plane> pl0, pl1;
  s.computeModels(in0, lm0, pl0);                      // :487 (first scan: :463-473)
  EXPECT_EQ(lm0.size(), r0.tm.size());
  EXPECT_TRUE(pl0.size() > 0 && pl0[0].features.size() == (size_t)params.numGroundFeatures);
  s.projectModels(in0.poseEstimate, lm0, pl0);
  in1.mapModels = lm0;
  s.computeModels(in1, lm1, pl1);
  EXPECT_EQ(lm1.size(), r1.tm.size());
  const SE3 currPose = in1.poseEstimate;
  auto treeMatches = s.matchFeatures(currPose, lm1, in1.mapModels, params.treeMatchThresh);   // :489
  auto planeMatches = s.matchFeatures(currPose, pl1, pl0, 1.0);                               // :490
  EXPECT_EQ((int)treeMatches.size(), ref.lastResult().n_tree_matches);
  EXPECT_EQ((int)planeMatches.size(), ref.lastResult().n_plane_matches);
  const double minPlanes = params.groundRadiiBins * params.groundThetaBins * 0.1;
  const bool treeCheck = lm1.size() > params.minTreeModels && treeMatches.size() > 5.0 * params.featuresPerTree;   // :499
  const bool groundCheck = planeMatches.size() > params.minGroundModels && pl1.size() > minPlanes;                 // :500
  SE3 tf = currPose;
  bool success = true;
  if (two_step) success = s.TwoStepOptimizePose(currPose, treeCheck, groundCheck, treeMatches, planeMatches, tf);
  else if (treeCheck && groundCheck) success = s.OptimizePose(currPose, treeMatches, planeMatches, tf);
  EXPECT_EQ(success, ok_ref);
  for (int a = 0; a < 3; ++a) EXPECT_NEAR(tf.translation()[a], r1.T_Map_Curr.translation()[a], 1e-9);
  for (int a = 0; a < 4; ++a) EXPECT_NEAR(tf.unit_quaternion()[a], r1.T_Map_Curr.unit_quaternion()[a], 1e-9);
  s.projectModels(tf, lm1, pl1);                        // :511
  std::vector<int> matches(lm1.size(), -1);
  s.matchModels(lm1, in1.mapModels, matches);           // :516
  EXPECT_TRUE(matches == r1.matches);
  for (size_t i = 0; i < lm1.size() && i < r1.tm.size(); ++i) {
    for (int a = 0; a < 3; ++a) EXPECT_NEAR(lm1[i].model.root[a], r1.tm[i].model.root[a], 1e-9);
    EXPECT_EQ(lm1[i].id, r1.tm[i].id);
  }
  // the two half problems on their own: the pieces TwoStepOptimizePose composes (:33-53)
  if (two_step) {
    double treeOut[3], groundOut[3];
    s.OptimizeXYYaw(currPose, treeCheck, treeMatches, treeOut);
    s.OptimizeZRollPitch(currPose, groundCheck, planeMatches, groundOut);
    EXPECT_NEAR(treeOut[0], tf.translation()[0], 1e-9);
    EXPECT_NEAR(treeOut[1], tf.translation()[1], 1e-9);
    EXPECT_NEAR(groundOut[0], tf.translation()[2], 1e-9);
  }
  // getPrevGroundFeatures: numGroundFeatures points per remembered plane, in the map frame
  EXPECT_EQ(ref.getPrevGroundFeatures().points.size(), ref.getPrevGroundModel().size() * (size_t)params.numGroundFeatures);
  // binGroundPoints: the retained lists are the lowest points of their cells, z ascending
  GroundGrid scgf;
  s.binGroundPoints(SE3(), in1.groundCloud->points, scgf);
  EXPECT_EQ((int)scgf.size(), params.groundRadiiBins);
  size_t kept = 0;
  bool sorted = true;
  for (auto &row : scgf)
    for (auto &cell : row) {
      kept += cell.size();
      for (size_t i = 1; i < cell.size(); ++i) sorted = sorted && !(cell[i].z < cell[i - 1].z);
    }
  EXPECT_TRUE(sorted && kept > 100 && kept < in1.groundCloud->points.size() / 10);
  {  // about another origin: shifting cloud and pose together selects the same number of points
    SE3 shifted;
    shifted.translation()[0] = 2.0; shifted.translation()[1] = -1.0;
    VectorType moved = in1.groundCloud->points;
    for (auto &q : moved) { q.x += 2.0f; q.y -= 1.0f; }
    GroundGrid g2;
    s.binGroundPoints(shifted, moved, g2);
    size_t kept2 = 0;
    for (auto &row : g2) for (auto &cell : row) kept2 += cell.size();
    EXPECT_TRUE(kept2 + 20 > kept && kept2 < kept + 20);
  }
  // cost of one RunSloam through the mirror once its staging arena is warm
  sloam::sloam timed;
  timed.setFmParams(params);
  SloamOutput o;
  SloamInput w0 = make_input(t0);
  timed.RunSloam(w0, o);
  double best = 1e9;
  for (int rep = 0; rep < 5; ++rep) {
    SloamInput w1 = make_input(t1);
    w1.mapModels = r0.tm;
    const auto c0 = std::chrono::steady_clock::now();
    timed.RunSloam(w1, o);
    best = std::min(best, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - c0).count());
  }
  std::printf("  two_step=%d  RunSloam through the C++ mirror: %.3f ms (best of 5, %zu ground points, %zu landmarks)\n",
              (int)two_step, best, t1.ground.size(), t1.landmarks.size());
}

// MapManager (mapManager.cpp) through the mirror.
static void map_manager_tests() {
  sloam_b200::HostConfig hc;
  hc.img_h = 16; hc.img_w = 64;
  std::shared_ptr<sloam_b200::Runtime> rt(new sloam_b200::Runtime(FeatureModelParams(), hc));
  MapManager mm(rt, 1024);
  EXPECT_TRUE(mm.size() == 0);
  std::vector<Cylinder> obs(5);
  for (size_t i = 0; i < obs.size(); ++i) {
    obs[i].model.root[0] = 2.0 * (double)i; obs[i].model.root[1] = 1.0; obs[i].model.root[2] = 1.0;
    obs[i].model.ray[2] = 1.0; obs[i].model.radius = 0.1 + 0.01 * (double)i; obs[i].isValid = true;
  }
  mm.updateMap(obs, std::vector<int>(5, -1));  // all new (:21-27)
  EXPECT_TRUE(mm.size() == 5);
  EXPECT_TRUE(mm.getMap().empty());            // hits == 1, getMap wants > 2 (:33)
  for (int rep = 0; rep < 2; ++rep) {
    std::vector<Cylinder> sub;
    mm.getSubmap(SE3(), sub);
    EXPECT_TRUE(sub.size() == 5);
    // match every observation to the submap entry with the same root
    std::vector<int> matches(5, -1);
    for (size_t i = 0; i < obs.size(); ++i)
      for (size_t j = 0; j < sub.size(); ++j)
        if (std::fabs(sub[j].model.root[0] - obs[i].model.root[0]) < 1e-9) matches[i] = (int)j;
    obs[0].model.radius += 0.05;
    mm.updateMap(obs, matches);
    EXPECT_TRUE(mm.size() == 5);
  }
  const std::vector<Cylinder> map = mm.getMap();
  EXPECT_TRUE(map.size() == 5);                // 1 + 2 hits
  if (map.size() == 5) EXPECT_NEAR(map[0].model.radius, 0.2, 1e-6);  // matched landmarks are overwritten (:14-16)
}

// SLOAMNode::run on a synthetic sequence; the lines are compared with the oracle
// by tests/test_host_api.py.
static void node_sequence(int n_keyframes) {
  sloam_b200::HostConfig hc;
  hc.img_h = 64; hc.img_w = 1024;
  sloam_synth_config cfg;
  sloam_synth_default_config(&cfg, hc.img_h, hc.img_w, 20);
  SLOAMNodeCore node(FeatureModelParams(), hc, 4096);
  const size_t N = (size_t)hc.img_h * hc.img_w;
  CloudT::Ptr cloud(new CloudT());
  cloud->width = hc.img_w; cloud->height = hc.img_h;
  cloud->points.resize(N);
  Mask mask(hc.img_h, hc.img_w);
  for (int k = 0; k < n_keyframes; ++k) {
    sloam_pose gt, guess;
    sloam_synth_pose(&cfg, k, &gt, &guess);
    EXPECT_TRUE(sloam_synth_generate_host(&cfg, k, 1, reinterpret_cast<sloam_point *>(cloud->points.data()),
                                          mask.data.data()) == 0);
    SE3 out;
    const bool ok = node.run(SE3(guess), SE3(), cloud, mask, out);
    const sloam_kf_result &r = node.lastResult();
    EXPECT_TRUE(ok == (r.success != 0));
    if (k == 0) EXPECT_TRUE(ok);
    std::printf("SEQ %d %d %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", k, (int)r.status, (int)ok, (int)r.n_landmarks,
                node.mapSize(), out.translation()[0], out.translation()[1], out.translation()[2],
                out.unit_quaternion()[0], out.unit_quaternion()[1], out.unit_quaternion()[2], out.unit_quaternion()[3]);
  }
}

int main(int argc, char **argv) {
  std::setvbuf(stdout, nullptr, _IOLBF, 0);
  if (argc < 2) { std::printf("usage: %s fixtures.bin\n", argv[0]); return 2; }
  std::ifstream f(argv[1], std::ios::binary);
  Fixture t0, t1;
  if (!read_fixture(f, t0) || !read_fixture(f, t1)) { std::printf("cannot read %s\n", argv[1]); return 2; }
  std::printf("fixtures: t0 %zu ground / %zu trees, t1 %zu ground / %zu trees\n", t0.ground.size(),
              t0.landmarks.size(), t1.ground.size(), t1.landmarks.size());
  try {
    plane_tests();
    cylinder_tests(t0);
    core_tests(t0, t1, false);
    core_tests(t0, t1, true);
    stage_method_tests(t0, t1, false);
    stage_method_tests(t0, t1, true);
    {  // Instance::computeGraph on an empty organized cloud: no landmarks, no crash
      Instance inst;
      CloudT::Ptr cloud(new CloudT());
      cloud->width = 64; cloud->height = 16;
      PointT nanp; nanp.x = nanp.y = nanp.z = std::numeric_limits<float>::quiet_NaN();
      cloud->points.assign(64 * 16, nanp);
      std::vector<std::vector<TreeVertex>> lm;
      inst.computeGraph(cloud, cloud, lm);
      EXPECT_TRUE(lm.empty());
    }
    map_manager_tests();
    node_sequence(argc > 2 ? std::atoi(argv[2]) : 6);
  } catch (const std::exception &e) {
    std::printf("EXCEPTION %s\n", e.what());
    return 3;
  }
  std::printf("%d checks, %d failed\n", g_run, g_fail);
  return g_fail ? 1 : 0;
}
