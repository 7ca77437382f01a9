// Minimal stand-ins for the PCL / Eigen types that aicp_core's plug-in headers mention, so that the adapter
// (include/aicp_b200_adapter.hpp) can be compile-tested in an image without PCL, Eigen or octomap.
// Layout-compatible where it matters: pcl::PointXYZ is a 16-byte (x, y, z, pad) record, Eigen::Matrix4f is 16
// column-major floats.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace Eigen {
struct Matrix4f {
  float m[16];
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  float operator()(int r, int c) const { return m[c * 4 + r]; }
  float* data() { return m; }
  const float* data() const { return m; }
  static Matrix4f Identity() { Matrix4f I; for (int i = 0; i < 16; ++i) I.m[i] = (i % 5 == 0) ? 1.f : 0.f; return I; }
};
struct Vector3d {
  double v[3];
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
};
struct Matrix4d {
  double m[16];
  const double* data() const { return m; }
};
struct Isometry3d {
  double m[16];                                           // column-major 4x4, what Isometry3d::matrix().data() points at
  Vector3d translation() const { return Vector3d{{m[12], m[13], m[14]}}; }
  Matrix4d matrix() const { Matrix4d r; for (int i = 0; i < 16; ++i) r.m[i] = m[i]; return r; }
  static Isometry3d Identity() { Isometry3d I; for (int i = 0; i < 16; ++i) I.m[i] = (i % 5 == 0) ? 1.0 : 0.0; return I; }
};
}  // namespace Eigen

namespace pcl {
struct alignas(16) PointXYZ { float x, y, z, pad; };
struct alignas(16) PointXYZRGB { float x, y, z, pad; float rgb; float pad2[3]; };
struct alignas(16) PointXYZRGBNormal { float x, y, z, pad; float normal_x, normal_y, normal_z, pad1; float rgb, curvature, pad2[2]; };
template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT> > Ptr;      // boost::shared_ptr in PCL 1.8; only -> and * are used
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  size_t size() const { return points.size(); }
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ is a 16-byte record");
static_assert(sizeof(PointXYZRGB) == 32, "pcl::PointXYZRGB is a 32-byte record");
}  // namespace pcl
