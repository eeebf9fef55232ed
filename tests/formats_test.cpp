// formats_test.cpp -- sloam_b200/host/formats.h.
//   g++ -std=c++17 -O1 -Iinclude tests/formats_test.cpp -o formats_test && ./formats_test [aux_dir]
// With the reference's fixture directory (sloam/src/tests/aux) as argument the fixtures are
// read and re-written: the re-serialised bytes must equal the files.
#include <cassert>
#include <cstdio>
#include <limits>
#include "../sloam_b200/host/formats.h"

using namespace sloam_formats;

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

static bool same_bits(float a, float b) { return std::memcmp(&a, &b, 4) == 0 || (std::isnan(a) && std::isnan(b)); }

int main(int argc, char **argv) {
  // PCD round trips
  CloudT c;
  c.width = 4; c.height = 2;
  const float nan = std::numeric_limits<float>::quiet_NaN();
  c.points = {{1.5f, -2.25f, 3.0f, 7.f}, {nan, nan, nan, 0.f}, {-2.7191279f, 2.4954553f, 1.4520754f, 0.f},
              {1e-7f, 123456.79f, -0.1f, 255.f}, {0.f, 0.f, 0.f, 0.f}, {3.4e38f, -1.2e-38f, 0.33333334f, 1.f},
              {10.f, 20.f, 30.f, 40.f}, {nan, 1.f, 2.f, 3.f}};
  for (int binary = 0; binary < 2; ++binary) {
    const std::string bytes = binary ? pcd_to_string_binary(c) : pcd_to_string_ascii(c);
    const CloudT r = pcd_from_string(bytes);
    CHECK(r.width == 4 && r.height == 2 && r.points.size() == c.points.size() && !r.is_dense);
    for (size_t i = 0; i < c.points.size(); ++i)
      CHECK(same_bits(r.points[i].x, c.points[i].x) && same_bits(r.points[i].y, c.points[i].y) &&
            same_bits(r.points[i].z, c.points[i].z) && same_bits(r.points[i].intensity, c.points[i].intensity));
  }
  CHECK(pcd_to_string_ascii(c).find("\nnan nan nan 0\n") != std::string::npos);
  CHECK(pcd_to_string_ascii(c).find("\n-2.7191279 2.4954553 1.4520754 0\n") != std::string::npos);
  // landmarks archive round trip
  Landmarks lm(2);
  lm[0].resize(2); lm[1].resize(1);
  lm[0][0].treeId = 31; lm[0][0].beam = 3; lm[0][0].radius = 0.017828299f; lm[0][0].isValid = true;
  lm[0][0].coords = {-1.8650216f, 10.382545f, 0.19929208f, 0.f};
  lm[0][0].points = {{-1.9263206f, 10.358125f, 0.19909006f, 0.f}, {1.f, 2.f, 3.f, 0.f}};
  lm[0][1].treeId = 31; lm[0][1].beam = 4; lm[0][1].prevVertexSize = 2; lm[0][1].radius = 0.5; lm[0][1].isValid = false;
  lm[1][0].treeId = 7;
  const std::string ar = landmarks_to_string(lm);
  CHECK(ar.rfind("22 serialization::archive 17 0 0 2 0 0 0 2 0 0 0 31 3 0 1.782829873e-02 1 0 0 -1.865021586e+00", 0) == 0);
  const Landmarks back = landmarks_from_string(ar);
  CHECK(back.size() == 2 && back[0].size() == 2 && back[1].size() == 1);
  CHECK(back[0][0].treeId == 31 && back[0][0].beam == 3 && back[0][0].isValid && back[0][0].points.size() == 2);
  CHECK(same_bits(back[0][0].points[0].y, 10.358125f) && back[0][1].prevVertexSize == 2 && !back[0][1].isValid);
  CHECK(landmarks_to_string(back) == ar);
  // ROS wire
  std::vector<ROSCylinder> cy(2);
  cy[0].root[0] = 1.f; cy[0].ray[2] = 1.f; cy[0].radii = {0.1, 0.2, 0.3}; cy[0].radius = 0.2f; cy[0].id = 42;
  cy[1].id = -7;
  const std::string wire = ros_encode_cylinders(cy);
  CHECK(wire.size() == 4 + (24 + 4 + 24 + 4 + 8) + (24 + 4 + 0 + 4 + 8));
  const std::vector<ROSCylinder> cb = ros_decode_cylinders(wire);
  CHECK(cb.size() == 2 && cb[0].radii.size() == 3 && cb[0].radii[2] == 0.3 && cb[0].id == 42 && cb[1].id == -7 &&
        cb[0].ray[2] == 1.f && cb[0].radius == 0.2f);
  bool threw = false;
  try { ros_decode_cylinders(wire.substr(0, wire.size() - 3)); } catch (const std::runtime_error &) { threw = true; }
  CHECK(threw);
  {  // ROSGround / ROSObservation: known bytes of the fixed part, round trip of the whole
    ROSHeader h; h.seq = 7; h.secs = 1630612694u; h.nsecs = 758621199u; h.frame_id = "map";
    std::string hb; ros_encode(hb, h);
    const unsigned char want[] = {7, 0, 0, 0, 0xd6, 0x2c, 0x31, 0x61, 0x0f, 0xa4, 0x37, 0x2d, 3, 0, 0, 0, 'm', 'a', 'p'};
    CHECK(hb.size() == sizeof want && std::memcmp(hb.data(), want, sizeof want) == 0);
    sloam_kf_result res{};
    res.success = 1; res.n_landmarks = 2;
    res.T_Map_Curr.t[0] = 0.625; res.T_Map_Curr.q[3] = 1.0;
    sloam_cylinder tm[2] = {};
    tm[0].root[0] = 1.5; tm[0].ray[2] = 1.0; tm[0].radius = 0.2; tm[1].root[1] = -2.0; tm[1].ray[2] = 1.0; tm[1].radius = 0.1;
    const int32_t ids[2] = {4, 9}, matches[2] = {3, -1};
    sloam_pose guess{}; guess.q[3] = 1.0; guess.t[0] = 0.6;
    sloam_plane gp{}; gp.plane[2] = 1.0; gp.plane[3] = 3.4;
    ROSObservation m = ros_observation(res, tm, ids, matches, guess, &gp, "map");
    std::vector<sloam_point> feat = {{1.f, 2.f, 3.f, 4.f}, {5.f, 6.f, 7.f, 8.f}};
    m.ground.features = ros_cloud_xyzi(feat, "map");
    m.ground.id = 11;
    const std::string wire = ros_encode(m);
    const ROSObservation b = ros_decode_observation(wire);
    CHECK(ros_encode(b) == wire);
    CHECK(b.treeModels.size() == 2 && b.treeModels[1].id == 9 && b.matches.size() == 2 && b.matches[1] == -1 && b.success == 1);
    CHECK(b.pose.position[0] == 0.625 && b.initialGuess.position[0] == 0.6 && b.ground.coefs[3] == 3.4f && b.ground.id == 11);
    CHECK(b.ground.features.width == 2 && b.ground.features.point_step == 32 && b.ground.features.data.size() == 64 &&
          b.ground.features.fields.size() == 4 && b.ground.features.fields[3].name == "intensity" && b.ground.features.fields[3].offset == 16);
    float back = 0.f;
    std::memcpy(&back, &b.ground.features.data[32 + 16], 4);
    CHECK(back == 8.f);
    bool threw2 = false;
    try { ros_decode_observation(wire.substr(0, wire.size() - 1)); } catch (const std::runtime_error &) { threw2 = true; }
    CHECK(threw2);
  }
  // the reference's own fixtures, byte for byte
  if (argc > 1) {
    const std::string dir = argv[1];
    int n_files = 0;
    for (const char *scene : {"still", "moving"})
      for (const char *t : {"t0", "t1"}) {
        const std::string lpath = dir + "/" + scene + "_landmarks_" + t;
        const std::string lbytes = read_file(lpath);
        const Landmarks l = landmarks_from_string(lbytes);
        CHECK(!l.empty());
        const std::string again = landmarks_to_string(l);
        if (again != lbytes) {
          size_t d = 0;
          while (d < again.size() && d < lbytes.size() && again[d] == lbytes[d]) ++d;
          std::printf("landmarks %s differ at byte %zu of %zu/%zu: '%s' vs '%s'\n", lpath.c_str(), d, again.size(),
                      lbytes.size(), again.substr(d, 30).c_str(), lbytes.substr(d, 30).c_str());
          ++fails;
        }
        for (const char *kind : {"tree", "ground"}) {
          const std::string ppath = dir + "/" + scene + "_" + kind + "_" + t + ".pcd";
          const std::string pbytes = read_file(ppath);
          const CloudT pc = pcd_from_string(pbytes);
          CHECK(pc.points.size() == (size_t)pc.width * pc.height && !pc.points.empty());
          const std::string pagain = pcd_to_string_ascii(pc);
          if (pagain != pbytes) {
            size_t d = 0;
            while (d < pagain.size() && d < pbytes.size() && pagain[d] == pbytes[d]) ++d;
            std::printf("pcd %s differs at byte %zu of %zu/%zu: '%s' vs '%s'\n", ppath.c_str(), d, pagain.size(),
                        pbytes.size(), pagain.substr(d, 40).c_str(), pbytes.substr(d, 40).c_str());
            ++fails;
          }
          ++n_files;
        }
      }
    std::printf("fixtures: %d pcd + 4 archives re-serialised\n", n_files);
  }
  std::printf(fails ? "FAILED %d\n" : "OK\n", fails);
  return fails ? 1 : 0;
}
