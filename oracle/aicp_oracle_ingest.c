/* aicp_oracle_ingest.c -- CPU restatement of the sweep accumulation in front of the path.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, never by the product.
 *
 * VelodyneAccumulatorROS::processLidar (aicp_ros/src/velodyne_accumulator.cpp:31-73): every incoming sweep is cropped to
 * +-30 m around the sensor (getPointsInOrientedBox with the identity pose, :59-60), moved to the inertial frame with
 * pcl::transformPointCloud(cloud, out, body_pose.translation().cast<float>(), Quaternionf(body_pose.rotation().cast<float>()))
 * (:62-63) and appended to the accumulated cloud (:66) until batch_size sweeps are in (:70-72).
 *
 * [UPSTREAM, recalled] Eigen 3.3 Quaternionf(Matrix3f) (Shoemake), Quaternionf::toRotationMatrix, and PCL 1.8's
 * transformPointCloud loop:  out.x = (float)(m00*x + m01*y + m02*z + m03), left to right, float32.  No FMA. */
#include <math.h>
#include <stdint.h>

#include "aicp_oracle.h"

/* pose: 16 doubles column-major.  T: 16 floats column-major = Translation3f(t) * Quaternionf(R) as a matrix */
void orc_pose_to_float_transform(const double* pose, float* T) {
  float m[3][3];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) m[r][c] = (float)pose[c * 4 + r];
  float q[4];                                   /* x, y, z, w */
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

/* one processLidar step: crop to +-half around the sensor, transform, append at out + 4 * n_acc.  Returns the points appended. */
int64_t orc_accumulate_sweep(const float* sweep, int64_t n, float half, const double* body_pose, float* out) {
  static const float zero[3] = {0.f, 0.f, 0.f};
  float T[16];
  orc_pose_to_float_transform(body_pose, T);
  int64_t m = orc_crop_box(sweep, n, -half, half, zero, zero, out);
  for (int64_t i = 0; i < m; ++i) {
    float* p = out + 4 * i;
    const float x = p[0], y = p[1], z = p[2];
    p[0] = ((T[0] * x + T[4] * y) + T[8] * z) + T[12];
    p[1] = ((T[1] * x + T[5] * y) + T[9] * z) + T[13];
    p[2] = ((T[2] * x + T[6] * y) + T[10] * z) + T[14];
  }
  return m;
}
