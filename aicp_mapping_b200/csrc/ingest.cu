// ingest.cu -- the data formats and the sweep accumulation in front of the path.
//
// replaces VelodyneAccumulatorROS::processLidar (aicp_ros/src/velodyne_accumulator.cpp:31-73): every sweep is cropped to +-30 m
//   around the sensor (getPointsInOrientedBox with the identity pose, :59-60), moved to the inertial frame with
//   pcl::transformPointCloud(cloud, out, body_pose.translation().cast<float>(), Quaternionf(body_pose.rotation().cast<float>()))
//   (:62-63) and appended to the accumulated cloud (:66).  Here the accumulated cloud lives on the device: k_crop_box
//   (crop.cu) compacts the sweep, k_acc_transform writes the transformed points behind the points already accumulated, and the
//   result (aicp_b200_get_accumulated) feeds the pre-filter without ever visiting the host.
// and the readers of the replay format (App::processFromFile, aicp_core/src/registration/app.cpp:250-279):
//   pcl::io::loadPCDFile<pcl::PointXYZ>  ->  read_pcd  (PCD v0.7 "ascii" and "binary" DATA, float32 x y z fields at any offset;
//                                            "binary_compressed" is rejected with a message)
//   PoseFileReader::readPoseFile (aicp_core/include/aicp_utils/poseFileReader.hpp:46-78)  ->  read_pose_file
//   pcl::io::loadPLYFile<pcl::PointXYZ> of the prior map (aicp_ros/src/app_ros.cpp:301)     ->  read_ply
// The readers are plain host code (file parsing is not GPU work); they are part of the library so that a replay needs nothing
// else, and they are exposed without a handle so the CPU tests cover them.
#include <math.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "handle.cuh"

namespace aicp {

struct Xform16 { float T[16]; };

__global__ void __launch_bounds__(256) k_acc_transform(const float4* __restrict__ in, long long n, Xform16 X, float4* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(&in[i]);
  const float3 o = xform_f(X.T, p.x, p.y, p.z);          // ((m00 x + m01 y) + m02 z) + m03, float32, no FMA
  out[i] = make_float4(o.x, o.y, o.z, p.w);
}

// Translation3f(t) * Quaternionf(R): Eigen's matrix -> quaternion (Shoemake) -> matrix round trip in float32, column-major out
void pose_to_float_transform(const double* pose, float* T) {
  float m[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[r][c] = (float)pose[c * 4 + r];
  float q[4];
  float t = (m[0][0] + m[1][1]) + m[2][2];
  if (t > 0.f) {
    t = sqrtf(t + 1.0f);
    q[3] = 0.5f * t;
    t = 0.5f / t;
    q[0] = (m[2][1] - m[1][2]) * t;
    q[1] = (m[0][2] - m[2][0]) * t;
    q[2] = (m[1][0] - m[0][1]) * t;
  } else {
    int i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrtf(((m[i][i] - m[j][j]) - m[k][k]) + 1.0f);
    q[i] = 0.5f * t;
    t = 0.5f / t;
    q[3] = (m[k][j] - m[j][k]) * t;
    q[j] = (m[j][i] + m[i][j]) * t;
    q[k] = (m[k][i] + m[i][k]) * t;
  }
  const float tx = 2.0f * q[0], ty = 2.0f * q[1], tz = 2.0f * q[2];
  const float twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
  const float txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
  const float tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
  float R[3][3];
  R[0][0] = 1.0f - (tyy + tzz); R[0][1] = txy - twz;          R[0][2] = txz + twy;
  R[1][0] = txy + twz;          R[1][1] = 1.0f - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy;          R[2][1] = tyz + twx;          R[2][2] = 1.0f - (txx + tyy);
  for (int c = 0; c < 3; ++c) { for (int r = 0; r < 3; ++r) T[c * 4 + r] = R[r][c]; T[c * 4 + 3] = 0.f; }
  T[12] = (float)pose[12]; T[13] = (float)pose[13]; T[14] = (float)pose[14]; T[15] = 1.f;
}

int run_accumulate_sweep(Handle* h, const float4* sweep, int64_t n, float half, const double* body_pose, int clear_first, int64_t* n_added) {
  if (clear_first) h->acc_n = 0;
  *n_added = 0;
  if (n == 0) return AICP_B200_OK;
  cudaStream_t s = h->stream;
  const float zero[3] = {0.f, 0.f, 0.f};
  CUDA_TRY(h->acc_tmp.reserve((size_t)n));
  int64_t kept = 0;
  int rc = run_crop_box(h, sweep, n, -half, half, zero, zero, h->acc_tmp.p, &kept);
  if (rc) return rc;
  if (kept == 0) return AICP_B200_OK;
  const int64_t total = h->acc_n + kept;
  if (total > (1ll << 28)) return fail(h, AICP_B200_ERR_BAD_ARG, "accumulate_sweep: more than 2^28 accumulated points");
  if ((size_t)total > h->acc.cap) {               // grow and carry over (DevBuf::reserve alone drops the content)
    DevBuf<float4> bigger;
    CUDA_TRY(bigger.reserve((size_t)total * 2));
    if (h->acc_n > 0) CUDA_TRY(cudaMemcpyAsync(bigger.p, h->acc.p, sizeof(float4) * (size_t)h->acc_n, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    h->acc.release();
    h->acc = bigger;
  }
  Xform16 X;
  pose_to_float_transform(body_pose, X.T);
  k_acc_transform<<<(unsigned)((kept + 255) / 256), 256, 0, s>>>(h->acc_tmp.p, (long long)kept, X, h->acc.p + h->acc_n);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  h->acc_n = total;
  *n_added = kept;
  return AICP_B200_OK;
}

// ---- PCD ---------------------------------------------------------------------------------------------------------------
// Reads the x, y, z fields (float32) of a PCD v0.7 file.  out: capacity cap records of 4 floats (nullable: count only).
int read_pcd(const char* path, float* out, int64_t cap, int64_t* n_out, std::string* err) {
  *n_out = 0;
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) { *err = std::string("cannot open ") + path; return AICP_B200_ERR_CONFIG; }
  std::vector<std::string> fields;
  std::vector<int> sizes, counts;
  std::vector<char> types;
  long long points = -1, width = -1, height = 1;
  std::string data_kind, line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "FIELDS" || key == "COLUMNS") { std::string v; while (ls >> v) fields.push_back(v); }
    else if (key == "SIZE") { int v; while (ls >> v) sizes.push_back(v); }
    else if (key == "TYPE") { char v; while (ls >> v) types.push_back(v); }
    else if (key == "COUNT") { int v; while (ls >> v) counts.push_back(v); }
    else if (key == "WIDTH") ls >> width;
    else if (key == "HEIGHT") ls >> height;
    else if (key == "POINTS") ls >> points;
    else if (key == "DATA") { ls >> data_kind; break; }
  }
  if (data_kind.empty()) { *err = "PCD header without a DATA line"; return AICP_B200_ERR_CONFIG; }
  if (points < 0) points = width > 0 ? width * height : -1;
  if (points < 0 || fields.empty()) { *err = "PCD header without POINTS / FIELDS"; return AICP_B200_ERR_CONFIG; }
  if (counts.empty()) counts.assign(fields.size(), 1);
  if (sizes.size() != fields.size() || types.size() != fields.size() || counts.size() != fields.size()) { *err = "PCD header: FIELDS / SIZE / TYPE / COUNT disagree"; return AICP_B200_ERR_CONFIG; }
  int off[3] = {-1, -1, -1}, col[3] = {-1, -1, -1}, stride = 0, cols = 0;
  for (size_t i = 0; i < fields.size(); ++i) {
    // header values are untrusted: a negative or huge SIZE / COUNT must not reach the record buffer or the offsets
    if ((sizes[i] != 1 && sizes[i] != 2 && sizes[i] != 4 && sizes[i] != 8) || counts[i] < 0 || counts[i] > 4096 ||
        stride > (1 << 20)) { *err = "PCD header: SIZE must be 1, 2, 4 or 8 and COUNT in [0, 4096]"; return AICP_B200_ERR_CONFIG; }
    for (int d = 0; d < 3; ++d)
      if (fields[i] == (d == 0 ? "x" : d == 1 ? "y" : "z")) {
        if (sizes[i] != 4 || types[i] != 'F' || counts[i] != 1) { *err = "PCD: x y z must be float32 fields"; return AICP_B200_ERR_CONFIG; }
        off[d] = stride; col[d] = cols;
      }
    stride += sizes[i] * counts[i];
    cols += counts[i];
  }
  if (off[0] < 0 || off[1] < 0 || off[2] < 0) { *err = "PCD: no x y z fields"; return AICP_B200_ERR_CONFIG; }
  *n_out = points;
  if (!out) return AICP_B200_OK;
  if (cap < points) { *err = "PCD: output buffer too small"; return AICP_B200_ERR_BAD_ARG; }
  if (data_kind == "binary") {
    std::vector<char> rec((size_t)stride);
    for (long long i = 0; i < points; ++i) {
      f.read(rec.data(), stride);
      if (f.gcount() != stride) { *err = "PCD: truncated binary data"; return AICP_B200_ERR_CONFIG; }
      for (int d = 0; d < 3; ++d) memcpy(&out[4 * i + d], rec.data() + off[d], 4);
      out[4 * i + 3] = 1.0f;
    }
  } else if (data_kind == "ascii") {
    for (long long i = 0; i < points; ++i) {
      if (!std::getline(f, line)) { *err = "PCD: truncated ascii data"; return AICP_B200_ERR_CONFIG; }
      const char* p = line.c_str();
      char* end = nullptr;
      int c = 0;
      float v[3] = {0.f, 0.f, 0.f};
      int got = 0;
      while (c < cols) {
        const float x = strtof(p, &end);         // accepts "nan" like PCL's reader
        if (end == p) break;
        for (int d = 0; d < 3; ++d) if (col[d] == c) { v[d] = x; ++got; }
        p = end; ++c;
      }
      if (got != 3) { *err = "PCD: malformed ascii record"; return AICP_B200_ERR_CONFIG; }
      out[4 * i] = v[0]; out[4 * i + 1] = v[1]; out[4 * i + 2] = v[2]; out[4 * i + 3] = 1.0f;
    }
  } else {
    *err = "PCD: DATA " + data_kind + " is not supported (ascii, binary)";
    return AICP_B200_ERR_CONFIG;
  }
  return AICP_B200_OK;
}

// pcl::PCDWriter::writeBinary for a pcl::PointXYZ cloud: FIELDS x y z, float32, unorganised
int write_pcd_binary(const char* path, const float* xyzw, int64_t n, std::string* err) {
  FILE* f = fopen(path, "wb");
  if (!f) { *err = std::string("cannot create ") + path; return AICP_B200_ERR_CONFIG; }
  fprintf(f, "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH %lld\nHEIGHT 1\n"
             "VIEWPOINT 0 0 0 1 0 0 0\nPOINTS %lld\nDATA binary\n", (long long)n, (long long)n);
  for (int64_t i = 0; i < n; ++i) if (fwrite(xyzw + 4 * i, 4, 3, f) != 3) { fclose(f); *err = "short write"; return AICP_B200_ERR_CONFIG; }
  fclose(f);
  return AICP_B200_OK;
}

// pcl::io::loadPLYFile<pcl::PointXYZ> as AppROS loads the prior map (aicp_ros/src/app_ros.cpp:301): the x, y, z properties of the
// vertex element, "ascii" or "binary_little_endian"; the vertex element must come first (it does in every PLY PCL or CloudCompare
// writes); elements after it (faces) are ignored.  float32 or float64 coordinates (converted to float32).
int read_ply(const char* path, float* out, int64_t cap, int64_t* n_out, std::string* err) {
  *n_out = 0;
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) { *err = std::string("cannot open ") + path; return AICP_B200_ERR_CONFIG; }
  std::string line, format;
  long long n_vertex = -1;
  bool in_vertex = false, seen_other_first = false, header_done = false;
  struct Prop { std::string name; int size; char kind; };     // kind: f float, d double, i other scalar
  std::vector<Prop> props;
  auto type_info = [](const std::string& t, int* size, char* kind) -> bool {
    if (t == "float" || t == "float32") { *size = 4; *kind = 'f'; return true; }
    if (t == "double" || t == "float64") { *size = 8; *kind = 'd'; return true; }
    if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") { *size = 1; *kind = 'i'; return true; }
    if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") { *size = 2; *kind = 'i'; return true; }
    if (t == "int" || t == "uint" || t == "int32" || t == "uint32") { *size = 4; *kind = 'i'; return true; }
    return false;
  };
  if (!std::getline(f, line)) { *err = "PLY: empty file"; return AICP_B200_ERR_CONFIG; }
  if (!line.empty() && line.back() == '\r') line.pop_back();
  if (line != "ply") { *err = "PLY: missing magic line"; return AICP_B200_ERR_CONFIG; }
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "format") ls >> format;
    else if (key == "element") {
      std::string name; long long cnt = 0;
      ls >> name >> cnt;
      if (name == "vertex") { in_vertex = true; n_vertex = cnt; }
      else { if (n_vertex < 0) seen_other_first = true; in_vertex = false; }
    } else if (key == "property" && in_vertex) {
      std::string t, name;
      ls >> t;
      if (t == "list") { *err = "PLY: list property in the vertex element"; return AICP_B200_ERR_CONFIG; }
      ls >> name;
      Prop p; p.name = name;
      if (!type_info(t, &p.size, &p.kind)) { *err = "PLY: unknown property type " + t; return AICP_B200_ERR_CONFIG; }
      props.push_back(p);
    } else if (key == "end_header") { header_done = true; break; }
  }
  if (!header_done || n_vertex < 0) { *err = "PLY: header without a vertex element / end_header"; return AICP_B200_ERR_CONFIG; }
  if (seen_other_first) { *err = "PLY: an element precedes the vertex element (not supported)"; return AICP_B200_ERR_CONFIG; }
  int idx[3] = {-1, -1, -1}, off[3] = {0, 0, 0}, stride = 0;
  for (size_t i = 0; i < props.size(); ++i) {
    for (int d = 0; d < 3; ++d)
      if (props[i].name == (d == 0 ? "x" : d == 1 ? "y" : "z")) {
        if (props[i].kind == 'i') { *err = "PLY: x y z must be float or double properties"; return AICP_B200_ERR_CONFIG; }
        idx[d] = (int)i; off[d] = stride;
      }
    stride += props[i].size;
  }
  if (idx[0] < 0 || idx[1] < 0 || idx[2] < 0) { *err = "PLY: no x y z properties"; return AICP_B200_ERR_CONFIG; }
  *n_out = n_vertex;
  if (!out) return AICP_B200_OK;
  if (cap < n_vertex) { *err = "PLY: output buffer too small"; return AICP_B200_ERR_BAD_ARG; }
  if (format == "binary_little_endian") {
    std::vector<char> rec((size_t)stride);
    for (long long i = 0; i < n_vertex; ++i) {
      f.read(rec.data(), stride);
      if (f.gcount() != stride) { *err = "PLY: truncated binary data"; return AICP_B200_ERR_CONFIG; }
      for (int d = 0; d < 3; ++d) {
        if (props[(size_t)idx[d]].kind == 'f') memcpy(&out[4 * i + d], rec.data() + off[d], 4);
        else { double v; memcpy(&v, rec.data() + off[d], 8); out[4 * i + d] = (float)v; }
      }
      out[4 * i + 3] = 1.0f;
    }
  } else if (format == "ascii") {
    for (long long i = 0; i < n_vertex; ++i) {
      if (!std::getline(f, line)) { *err = "PLY: truncated ascii data"; return AICP_B200_ERR_CONFIG; }
      const char* p = line.c_str();
      char* end = nullptr;
      int got = 0;
      for (size_t c = 0; c < props.size(); ++c) {
        const double v = strtod(p, &end);
        if (end == p) break;
        for (int d = 0; d < 3; ++d) if (idx[d] == (int)c) { out[4 * i + d] = props[c].kind == 'f' ? strtof(p, nullptr) : (float)v; ++got; }
        p = end;
      }
      if (got != 3) { *err = "PLY: malformed ascii vertex"; return AICP_B200_ERR_CONFIG; }
      out[4 * i + 3] = 1.0f;
    }
  } else {
    *err = "PLY: format " + format + " is not supported (ascii, binary_little_endian)";
    return AICP_B200_ERR_CONFIG;
  }
  return AICP_B200_OK;
}

// PoseFileReader::readPoseFile: rows "counter, sec, nsec, x, y, z, qx, qy, qz, qw", '#' comment lines skipped.
// rows_out: n x 3 int64 (counter, sec, nsec); poses_out: n x 16 doubles column-major (translation; Quaterniond(w,x,y,z) matrix)
int read_pose_file(const char* path, int64_t* rows_out, double* poses_out, int64_t cap, int64_t* n_out, std::string* err) {
  *n_out = 0;
  std::ifstream f(path);
  if (!f.good()) { *err = std::string("cannot open ") + path; return AICP_B200_ERR_CONFIG; }
  std::string line;
  int64_t n = 0;
  while (std::getline(f, line)) {
    if (line.empty()) { *err = "pose file: empty line (the reference's reader throws on it, poseFileReader.hpp:54)"; return AICP_B200_ERR_CONFIG; }
    if (line[0] == '#') continue;
    std::vector<double> row;
    std::stringstream sep(line);
    std::string field;
    while (std::getline(sep, field, ',')) {
      char* end = nullptr;
      const double v = strtod(field.c_str(), &end);
      if (end == field.c_str()) { *err = "pose file: not a number: " + field; return AICP_B200_ERR_CONFIG; }
      row.push_back(v);
    }
    if (row.size() < 10) { *err = "pose file: a row has fewer than 10 fields"; return AICP_B200_ERR_CONFIG; }
    if (rows_out && poses_out) {
      if (n >= cap) { *err = "pose file: output buffer too small"; return AICP_B200_ERR_BAD_ARG; }
      rows_out[3 * n] = (int64_t)(int)row[0]; rows_out[3 * n + 1] = (int64_t)(int)row[1]; rows_out[3 * n + 2] = (int64_t)(int)row[2];
      // Isometry3d::Identity(); translation() << x, y, z; rotate(Quaterniond(w, x, y, z)): Eigen's toRotationMatrix in float64
      const double x = row[6], y = row[7], z = row[8], w = row[9];
      const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
      const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
      double* P = poses_out + 16 * n;
      P[0] = 1.0 - (tyy + tzz); P[4] = txy - twz;         P[8] = txz + twy;          P[12] = row[3];
      P[1] = txy + twz;         P[5] = 1.0 - (txx + tzz); P[9] = tyz - twx;          P[13] = row[4];
      P[2] = txz - twy;         P[6] = tyz + twx;         P[10] = 1.0 - (txx + tyy); P[14] = row[5];
      P[3] = 0.0; P[7] = 0.0; P[11] = 0.0; P[15] = 1.0;
    }
    ++n;
  }
  *n_out = n;
  return AICP_B200_OK;
}

}  // namespace aicp
