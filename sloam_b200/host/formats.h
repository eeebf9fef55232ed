// formats.h -- SURVEY 8(f)-3: the fixture / wire formats either side of the hot path, so
// that dumps from a real SLOAM install can be read and written without PCL, Boost or ROS.
//
//  * PCD v0.7 of pcl::PointXYZI clouds, DATA ascii and DATA binary
//    (what pcl::io::loadPCDFile / savePCDFileASCII / savePCDFileBinary exchange; the
//    reference fixtures sloam/src/tests/aux/*_{tree,ground}_t{0,1}.pcd are ascii).
//  * Boost text archive version 17 of std::vector<std::vector<TreeVertex>> as declared in
//    sloam/include/helpers/serialization.h:13-33 (the *_landmarks_t{0,1} fixtures that
//    sloam/src/tests/core_test.cpp:79 loads).
//  * ROS1 wire encoding of sloam_msgs/ROSCylinder and of a ROSCylinder[] field
//    (sloam_msgs/msg/ROSCylinder.msg; filled from a Cylinder in sloamNode.cpp:140-146), of
//    sloam_msgs/ROSGround and of sloam_msgs/ROSObservation (the message SLOAMNode advertises,
//    sloamNode.cpp:45) with the std_msgs / geometry_msgs / sensor_msgs messages they embed.
//
// Header-only, host-only; uses the stand-in types of sloam_host.h.
#ifndef SLOAM_B200_FORMATS_H
#define SLOAM_B200_FORMATS_H

#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "sloam_host.h"

namespace sloam_formats {

using Landmarks = std::vector<std::vector<TreeVertex>>;

// ----------------------------------------------------------------- PCD v0.7 --
// pcl::PCDWriter::writeASCII prints every field through an ostream with precision 8
// (general format, "nan" for NaN); the header below is PCL's for PointXYZI.
inline std::string pcd_header(const CloudT &c, const char *data) {
  std::ostringstream h;
  const size_t n = c.points.size();
  const uint32_t w = c.width ? c.width : (uint32_t)n, ht = c.width ? (c.height ? c.height : 1) : 1;
  h << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\n"
       "SIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH " << w << "\nHEIGHT " << ht
    << "\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA " << data << "\n";
  return h.str();
}

inline void pcd_value(std::string &out, float v) {
  if (std::isnan(v)) { out += "nan"; return; }
  char buf[32];
  std::snprintf(buf, sizeof buf, "%.8g", (double)v);
  out += buf;
}

inline std::string pcd_to_string_ascii(const CloudT &c) {
  std::string out = pcd_header(c, "ascii");
  out.reserve(out.size() + c.points.size() * 48);
  for (const PointT &p : c.points) {
    pcd_value(out, p.x); out += ' ';
    pcd_value(out, p.y); out += ' ';
    pcd_value(out, p.z); out += ' ';
    pcd_value(out, p.intensity); out += '\n';
  }
  return out;
}

inline std::string pcd_to_string_binary(const CloudT &c) {
  std::string out = pcd_header(c, "binary");
  const size_t off = out.size();
  out.resize(off + c.points.size() * sizeof(PointT));
  if (!c.points.empty()) std::memcpy(&out[off], c.points.data(), c.points.size() * sizeof(PointT));
  return out;
}

inline void write_file(const std::string &path, const std::string &bytes) {
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  f.write(bytes.data(), (std::streamsize)bytes.size());
}
inline std::string read_file(const std::string &path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::ostringstream s;
  s << f.rdbuf();
  return s.str();
}

// Reads x y z [intensity] clouds with 4-byte float fields, ascii or binary (not
// binary_compressed).  Fields other than x, y, z, intensity are skipped.
inline CloudT pcd_from_string(const std::string &s) {
  CloudT c;
  size_t pos = 0, points = 0;
  std::vector<std::string> fields;
  std::vector<int> sizes, counts;
  std::string data;
  while (pos < s.size()) {
    const size_t eol = s.find('\n', pos);
    if (eol == std::string::npos) throw std::runtime_error("pcd: truncated header");
    std::istringstream line(s.substr(pos, eol - pos));
    pos = eol + 1;
    std::string key;
    line >> key;
    if (key.empty() || key[0] == '#') continue;
    if (key == "FIELDS") { std::string f; while (line >> f) fields.push_back(f); }
    else if (key == "SIZE") { int v; while (line >> v) sizes.push_back(v); }
    else if (key == "COUNT") { int v; while (line >> v) counts.push_back(v); }
    else if (key == "WIDTH") line >> c.width;
    else if (key == "HEIGHT") line >> c.height;
    else if (key == "POINTS") line >> points;
    else if (key == "DATA") { line >> data; break; }
  }
  if (fields.empty() || data.empty()) throw std::runtime_error("pcd: missing FIELDS / DATA");
  if (points == 0) points = (size_t)c.width * c.height;
  if (counts.empty()) counts.assign(fields.size(), 1);
  if (sizes.empty()) sizes.assign(fields.size(), 4);
  int ix = -1, iy = -1, iz = -1, ii = -1, slot = 0;
  std::vector<int> first(fields.size());
  for (size_t f = 0; f < fields.size(); ++f) {
    first[f] = slot;
    if (fields[f] == "x") ix = slot; else if (fields[f] == "y") iy = slot;
    else if (fields[f] == "z") iz = slot; else if (fields[f] == "intensity") ii = slot;
    if (sizes[f] != 4) throw std::runtime_error("pcd: only 4-byte fields are supported");
    slot += counts[f];
  }
  if (ix < 0 || iy < 0 || iz < 0) throw std::runtime_error("pcd: no x y z fields");
  c.points.resize(points);
  c.is_dense = true;
  if (data == "ascii") {
    const char *p = s.c_str() + pos;
    std::vector<float> row(slot);
    for (size_t n = 0; n < points; ++n) {
      for (int q = 0; q < slot; ++q) {
        char *end = nullptr;
        row[q] = std::strtof(p, &end);  // accepts "nan"
        if (end == p) throw std::runtime_error("pcd: truncated ascii data");
        p = end;
      }
      PointT &o = c.points[n];
      o.x = row[ix]; o.y = row[iy]; o.z = row[iz]; o.intensity = ii >= 0 ? row[ii] : 0.f;
      if (std::isnan(o.x) || std::isnan(o.y) || std::isnan(o.z)) c.is_dense = false;
    }
  } else if (data == "binary") {
    const size_t stride = (size_t)slot * 4;
    if (s.size() - pos < points * stride) throw std::runtime_error("pcd: truncated binary data");
    for (size_t n = 0; n < points; ++n) {
      float row[64];
      if (slot > 64) throw std::runtime_error("pcd: too many fields");
      std::memcpy(row, s.data() + pos + n * stride, stride);
      PointT &o = c.points[n];
      o.x = row[ix]; o.y = row[iy]; o.z = row[iz]; o.intensity = ii >= 0 ? row[ii] : 0.f;
      if (std::isnan(o.x) || std::isnan(o.y) || std::isnan(o.z)) c.is_dense = false;
    }
  } else {
    throw std::runtime_error("pcd: unsupported DATA " + data);
  }
  return c;
}

// ------------------------------------------- Boost text archive, version 17 --
// Grammar (tokens separated by one space): the signature "22 serialization::archive 17";
// every class writes "tracking version" (0 0) the FIRST time an object of it appears;
// a vector writes "count item_version".  Floats print as %.9e (max_digits10 = 9).  The
// fixtures print TreeVertex::radius with 9 digits as well (it was a float when they were
// recorded; definitions.h:63 declares it Scalar = double now), so the writer does too.
inline void ar_float(std::string &o, double v) {
  char buf[40];
  std::snprintf(buf, sizeof buf, " %.9e", v);
  o += buf;
}

inline std::string landmarks_to_string(const Landmarks &lm) {
  std::string o = "22 serialization::archive 17";
  bool first_inner = true, first_vertex = true, first_coords = true, first_points = true;
  o += " 0 0 " + std::to_string(lm.size()) + " 0";
  for (const std::vector<TreeVertex> &tree : lm) {
    if (first_inner) { o += " 0 0"; first_inner = false; }
    o += " " + std::to_string(tree.size()) + " 0";
    for (const TreeVertex &v : tree) {
      if (first_vertex) { o += " 0 0"; first_vertex = false; }
      o += " " + std::to_string(v.treeId) + " " + std::to_string(v.beam) + " " + std::to_string(v.prevVertexSize);
      ar_float(o, v.radius);
      o += v.isValid ? " 1" : " 0";
      if (first_coords) { o += " 0 0"; first_coords = false; }
      ar_float(o, v.coords.x); ar_float(o, v.coords.y); ar_float(o, v.coords.z);
      if (first_points) { o += " 0 0"; first_points = false; }
      o += " " + std::to_string(v.points.size()) + " 0";
      for (const PointT &p : v.points) { ar_float(o, p.x); ar_float(o, p.y); ar_float(o, p.z); }
    }
  }
  o += "\n";
  return o;
}

inline Landmarks landmarks_from_string(const std::string &s) {
  std::istringstream in(s);
  auto tok = [&]() { std::string t; if (!(in >> t)) throw std::runtime_error("archive: truncated"); return t; };
  auto skip = [&](int n) { for (int i = 0; i < n; ++i) tok(); };
  if (tok() != "22" || tok() != "serialization::archive") throw std::runtime_error("archive: bad signature");
  if (tok() != "17") throw std::runtime_error("archive: version 17 expected");
  skip(2);
  Landmarks lm((size_t)std::stoul(tok()));
  skip(1);
  bool first_inner = true, first_vertex = true, first_coords = true, first_points = true;
  for (std::vector<TreeVertex> &tree : lm) {
    if (first_inner) { skip(2); first_inner = false; }
    tree.resize((size_t)std::stoul(tok()));
    skip(1);
    for (TreeVertex &v : tree) {
      if (first_vertex) { skip(2); first_vertex = false; }
      v.treeId = std::stoi(tok()); v.beam = std::stoi(tok()); v.prevVertexSize = std::stoi(tok());
      v.radius = std::stod(tok());
      v.isValid = tok() != "0";
      if (first_coords) { skip(2); first_coords = false; }
      v.coords.x = std::stof(tok()); v.coords.y = std::stof(tok()); v.coords.z = std::stof(tok());
      if (first_points) { skip(2); first_points = false; }
      v.points.resize((size_t)std::stoul(tok()));
      skip(1);
      for (PointT &p : v.points) { p.x = std::stof(tok()); p.y = std::stof(tok()); p.z = std::stof(tok()); }
    }
  }
  return lm;
}

// ------------------------------------------------ ROS1 wire: ROSCylinder ----
// float32[3] root, float32[3] ray, float64[] radii, float32 radius, int64 id: little-endian,
// fixed arrays inline, variable arrays behind a uint32 element count.
struct ROSCylinder {
  float root[3] = {0, 0, 0}, ray[3] = {0, 0, 0};
  std::vector<double> radii;
  float radius = 0;
  int64_t id = 0;
};

template <class T> inline void put(std::string &o, const T &v) { o.append(reinterpret_cast<const char *>(&v), sizeof(T)); }
template <class T> inline T get(const std::string &s, size_t &pos) {
  if (pos + sizeof(T) > s.size()) throw std::runtime_error("ros: truncated message");
  T v;
  std::memcpy(&v, s.data() + pos, sizeof(T));
  pos += sizeof(T);
  return v;
}

inline void ros_encode(std::string &o, const ROSCylinder &c) {
  for (float v : c.root) put(o, v);
  for (float v : c.ray) put(o, v);
  put(o, (uint32_t)c.radii.size());
  for (double v : c.radii) put(o, v);
  put(o, c.radius);
  put(o, c.id);
}
inline ROSCylinder ros_decode_cylinder(const std::string &s, size_t &pos) {
  ROSCylinder c;
  for (float &v : c.root) v = get<float>(s, pos);
  for (float &v : c.ray) v = get<float>(s, pos);
  c.radii.resize(get<uint32_t>(s, pos));
  for (double &v : c.radii) v = get<double>(s, pos);
  c.radius = get<float>(s, pos);
  c.id = get<int64_t>(s, pos);
  return c;
}
// ROSCylinder[] (e.g. ROSObservation.treeModels)
inline std::string ros_encode_cylinders(const std::vector<ROSCylinder> &v) {
  std::string o;
  put(o, (uint32_t)v.size());
  for (const ROSCylinder &c : v) ros_encode(o, c);
  return o;
}
inline std::vector<ROSCylinder> ros_decode_cylinders(const std::string &s) {
  size_t pos = 0;
  std::vector<ROSCylinder> v(get<uint32_t>(s, pos));
  for (ROSCylinder &c : v) c = ros_decode_cylinder(s, pos);
  if (pos != s.size()) throw std::runtime_error("ros: trailing bytes");
  return v;
}
// the message the node fills from a landmark (sloamNode.cpp:140-146)
inline ROSCylinder ros_from_cylinder(const sloam_cylinder &m, int64_t id, const std::vector<double> &radii = {}) {
  ROSCylinder c;
  for (int i = 0; i < 3; ++i) { c.root[i] = (float)m.root[i]; c.ray[i] = (float)m.ray[i]; }
  c.radius = (float)m.radius;
  c.radii = radii;
  c.id = id;
  return c;
}

// ------------------------------------- ROS1 wire: ROSGround, ROSObservation ----
// sloam_msgs/msg/ROSGround.msg and ROSObservation.msg (the message SLOAMNode advertises on
// "observation", sloamNode.cpp:45,132) with the standard messages they embed: std_msgs/Header,
// geometry_msgs/PoseStamped, sensor_msgs/PointCloud2.  ROS1 serialisation: little-endian,
// fields in declaration order, strings and variable arrays behind a uint32 length.
struct ROSHeader {  // std_msgs/Header
  uint32_t seq = 0, secs = 0, nsecs = 0;
  std::string frame_id;
};
struct ROSPoseStamped {  // geometry_msgs/PoseStamped
  ROSHeader header;
  double position[3] = {0, 0, 0}, orientation[4] = {0, 0, 0, 1};  // x y z w
};
struct ROSPointField {  // sensor_msgs/PointField
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 7;  // FLOAT32
  uint32_t count = 1;
};
struct ROSPointCloud2 {  // sensor_msgs/PointCloud2
  ROSHeader header;
  uint32_t height = 1, width = 0;
  std::vector<ROSPointField> fields;
  uint8_t is_bigendian = 0;
  uint32_t point_step = 0, row_step = 0;
  std::string data;
  uint8_t is_dense = 0;
};
struct ROSGround {  // float32[4] coefs, PointCloud2 features, int64 id
  float coefs[4] = {0, 0, 0, 0};
  ROSPointCloud2 features;
  int64_t id = 0;
};
struct ROSObservation {
  ROSHeader header;
  ROSPoseStamped pose, initialGuess;
  ROSPointCloud2 pc;
  std::vector<ROSCylinder> treeModels;
  ROSGround ground;
  std::vector<int32_t> matches;
  uint8_t success = 0;
};

inline void put_str(std::string &o, const std::string &v) { put(o, (uint32_t)v.size()); o.append(v); }
inline std::string get_str(const std::string &s, size_t &pos) {
  const uint32_t n = get<uint32_t>(s, pos);
  if (pos + n > s.size()) throw std::runtime_error("ros: truncated message");
  std::string v = s.substr(pos, n);
  pos += n;
  return v;
}
inline void ros_encode(std::string &o, const ROSHeader &h) { put(o, h.seq); put(o, h.secs); put(o, h.nsecs); put_str(o, h.frame_id); }
inline void ros_decode(const std::string &s, size_t &pos, ROSHeader &h) {
  h.seq = get<uint32_t>(s, pos); h.secs = get<uint32_t>(s, pos); h.nsecs = get<uint32_t>(s, pos); h.frame_id = get_str(s, pos);
}
inline void ros_encode(std::string &o, const ROSPoseStamped &p) {
  ros_encode(o, p.header);
  for (double v : p.position) put(o, v);
  for (double v : p.orientation) put(o, v);
}
inline void ros_decode(const std::string &s, size_t &pos, ROSPoseStamped &p) {
  ros_decode(s, pos, p.header);
  for (double &v : p.position) v = get<double>(s, pos);
  for (double &v : p.orientation) v = get<double>(s, pos);
}
inline void ros_encode(std::string &o, const ROSPointCloud2 &c) {
  ros_encode(o, c.header);
  put(o, c.height); put(o, c.width);
  put(o, (uint32_t)c.fields.size());
  for (const ROSPointField &f : c.fields) { put_str(o, f.name); put(o, f.offset); put(o, f.datatype); put(o, f.count); }
  put(o, c.is_bigendian); put(o, c.point_step); put(o, c.row_step);
  put_str(o, c.data);
  put(o, c.is_dense);
}
inline void ros_decode(const std::string &s, size_t &pos, ROSPointCloud2 &c) {
  ros_decode(s, pos, c.header);
  c.height = get<uint32_t>(s, pos); c.width = get<uint32_t>(s, pos);
  c.fields.resize(get<uint32_t>(s, pos));
  for (ROSPointField &f : c.fields) { f.name = get_str(s, pos); f.offset = get<uint32_t>(s, pos); f.datatype = get<uint8_t>(s, pos); f.count = get<uint32_t>(s, pos); }
  c.is_bigendian = get<uint8_t>(s, pos); c.point_step = get<uint32_t>(s, pos); c.row_step = get<uint32_t>(s, pos);
  c.data = get_str(s, pos);
  c.is_dense = get<uint8_t>(s, pos);
}
inline void ros_encode(std::string &o, const ROSGround &g) {
  for (float v : g.coefs) put(o, v);
  ros_encode(o, g.features);
  put(o, g.id);
}
inline void ros_decode(const std::string &s, size_t &pos, ROSGround &g) {
  for (float &v : g.coefs) v = get<float>(s, pos);
  ros_decode(s, pos, g.features);
  g.id = get<int64_t>(s, pos);
}
inline std::string ros_encode(const ROSObservation &m) {
  std::string o;
  ros_encode(o, m.header);
  ros_encode(o, m.pose);
  ros_encode(o, m.initialGuess);
  ros_encode(o, m.pc);
  put(o, (uint32_t)m.treeModels.size());
  for (const ROSCylinder &c : m.treeModels) ros_encode(o, c);
  ros_encode(o, m.ground);
  put(o, (uint32_t)m.matches.size());
  for (int32_t v : m.matches) put(o, v);
  put(o, m.success);
  return o;
}
inline ROSObservation ros_decode_observation(const std::string &s) {
  ROSObservation m;
  size_t pos = 0;
  ros_decode(s, pos, m.header);
  ros_decode(s, pos, m.pose);
  ros_decode(s, pos, m.initialGuess);
  ros_decode(s, pos, m.pc);
  m.treeModels.resize(get<uint32_t>(s, pos));
  for (ROSCylinder &c : m.treeModels) c = ros_decode_cylinder(s, pos);
  ros_decode(s, pos, m.ground);
  m.matches.resize(get<uint32_t>(s, pos));
  for (int32_t &v : m.matches) v = get<int32_t>(s, pos);
  m.success = get<uint8_t>(s, pos);
  if (pos != s.size()) throw std::runtime_error("ros: trailing bytes");
  return m;
}
// pcl::toROSMsg of a PointXYZI cloud as the node publishes its feature clouds: fields x, y, z,
// intensity (FLOAT32) at offsets 0, 4, 8, 16 of PCL's 32-byte point
inline ROSPointCloud2 ros_cloud_xyzi(const std::vector<sloam_point> &pts, const std::string &frame) {
  ROSPointCloud2 c;
  c.header.frame_id = frame;
  c.width = (uint32_t)pts.size();
  const char *names[4] = {"x", "y", "z", "intensity"};
  const uint32_t offs[4] = {0, 4, 8, 16};
  for (int i = 0; i < 4; ++i) { ROSPointField f; f.name = names[i]; f.offset = offs[i]; c.fields.push_back(f); }
  c.point_step = 32;
  c.row_step = 32 * c.width;
  c.data.assign((size_t)32 * pts.size(), '\0');
  for (size_t i = 0; i < pts.size(); ++i) {
    const float one = 1.0f;
    std::memcpy(&c.data[32 * i], &pts[i].x, 12);
    std::memcpy(&c.data[32 * i + 12], &one, 4);          // PCL's padding lane holds 1.0f
    std::memcpy(&c.data[32 * i + 16], &pts[i].intensity, 4);
  }
  return c;
}
// SloamOutput + the accepted ground plane of a keyframe -> the observation message
inline ROSObservation ros_observation(const sloam_kf_result &res, const sloam_cylinder *tm, const int32_t *tm_id,
                                      const int32_t *matches, const sloam_pose &guess, const sloam_plane *ground,
                                      const std::string &frame) {
  ROSObservation m;
  m.header.frame_id = m.pose.header.frame_id = m.initialGuess.header.frame_id = frame;
  for (int i = 0; i < 3; ++i) { m.pose.position[i] = res.T_Map_Curr.t[i]; m.initialGuess.position[i] = guess.t[i]; }
  for (int i = 0; i < 4; ++i) { m.pose.orientation[i] = res.T_Map_Curr.q[i]; m.initialGuess.orientation[i] = guess.q[i]; }
  for (int i = 0; i < res.n_landmarks; ++i) { m.treeModels.push_back(ros_from_cylinder(tm[i], tm_id[i])); m.matches.push_back(matches[i]); }
  if (ground) for (int i = 0; i < 4; ++i) m.ground.coefs[i] = (float)ground->plane[i];
  m.pc.header.frame_id = m.ground.features.header.frame_id = frame;
  m.success = res.success ? 1 : 0;
  return m;
}

}  // namespace sloam_formats
#endif
