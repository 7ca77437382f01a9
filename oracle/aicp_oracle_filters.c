/* aicp_oracle_filters.c -- CPU restatement of the map-handling filter next to the registration path.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product.
 *
 * getPointsInOrientedBox (aicp_core/src/utils/filteringUtils.cpp:621-637): pcl::CropBox with min = (m,m,m), max = (M,M,M),
 * rotation = origin.block<3,3>(0,0).eulerAngles(0,1,2), translation = origin.col(3); the input cloud is replaced by the
 * points inside the box, input order preserved.  Called by App on the prior / built map before every registration
 * (app.cpp:41-69) with +-crop_map_around_base (15 m, aicp_ros/launch/aicp.launch:56).
 *
 * [UPSTREAM, recalled] pcl::CropBox<PointT>::applyFilter: transform = pcl::getTransformation(0,0,0, roll, pitch, yaw) (float),
 * local = inverse(transform) * (p - translation); a point is removed when any local coordinate is < min or > max.
 * Decisions where the upstream arithmetic is not reproducible bit for bit (identical in oracle and CUDA):
 *   - the inverse of the rotation is its transpose (Eigen computes a general 3x3 inverse of the same matrix);
 *   - local_k = (m_k0 * dx + m_k1 * dy) + m_k2 * dz in float32, no FMA;
 *   - non-finite points are dropped (CropBox does this when !is_dense). */
#include <math.h>
#include <stdint.h>

#include "aicp_oracle.h"

/* pcl::getTransformation(0,0,0,roll,pitch,yaw): R = Rz(yaw) Ry(pitch) Rx(roll), float, row-major out[9] */
void orc_rpy_to_matrix(const float* rpy, float* R) {
  float A = cosf(rpy[2]), B = sinf(rpy[2]), C = cosf(rpy[1]), D = sinf(rpy[1]), E = cosf(rpy[0]), F = sinf(rpy[0]);
  float DE = D * E, DF = D * F;
  R[0] = A * C;  R[1] = A * DF - B * E;  R[2] = B * F + A * DE;
  R[3] = B * C;  R[4] = A * E + B * DF;  R[5] = B * DE - A * F;
  R[6] = -D;     R[7] = C * F;           R[8] = C * E;
}

int64_t orc_crop_box(const float* xyzw, int64_t n, float bmin, float bmax, const float* rpy, const float* t, float* out) {
  float R[9];
  orc_rpy_to_matrix(rpy, R);
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = xyzw + 4 * i;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    float dx = p[0] - t[0], dy = p[1] - t[1], dz = p[2] - t[2];
    /* inverse rotation = transpose: row k of R^T is column k of R */
    float lx = (R[0] * dx + R[3] * dy) + R[6] * dz;
    float ly = (R[1] * dx + R[4] * dy) + R[7] * dz;
    float lz = (R[2] * dx + R[5] * dy) + R[8] * dz;
    if (lx < bmin || ly < bmin || lz < bmin || lx > bmax || ly > bmax || lz > bmax) continue;
    if (out) { out[4 * m] = p[0]; out[4 * m + 1] = p[1]; out[4 * m + 2] = p[2]; out[4 * m + 3] = p[3]; }
    ++m;
  }
  return m;
}
