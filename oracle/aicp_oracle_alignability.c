/* aicp_oracle_alignability.c -- CPU restatement of the FOV overlap filter and the alignability filter.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product.
 *
 * overlapFilter        aicp_core/src/utils/filteringUtils.cpp:111-193   (App::computeAlignmentRisk, app.cpp:153-156)
 * alignabilityFilter   filteringUtils.cpp:196-430 with computeNormalsCentroid :432-445, getOrientedBoundingBox :448-478,
 *                      getPointsInOrientedBox :481-505, overlapBoxFilter :507-576   (app.cpp:164-166)
 * Both call PCL (VoxelGrid / NormalEstimation / RegionGrowing through the pre-filter, MomentOfInertiaEstimation::getOBB,
 * CropBox, PCA) and Eigen; neither is installed or vendored, so the PCL parts are [UPSTREAM, recalled] -- PARITY UNPINNED.
 *
 * Decisions where the upstream arithmetic is not reproducible bit for bit (identical here and in the CUDA path):
 *   - FOV test: the reference computes theta = atan2f(y, x) * 180 / pi and keeps |theta| < thresh; restated without libm as
 *     x / sqrt(x^2 + y^2) > cos(thresh * pi / 180) in float64 (theta = 0 for x = y = 0);
 *   - every per-cluster sum (points, normals, n n^T, (p - mean)(p - mean)^T) is exact fixed point (2^-20 m for coordinates,
 *     2^-30 for products), so cluster statistics do not depend on summation order;
 *   - eigen-decompositions (MomentOfInertiaEstimation::computeEigenVectors, pcl::PCA) by cyclic Jacobi in float64; axes in
 *     canonical sign (largest |component| positive) before PCL's right-handedness fix (major axis flipped when det <= 0);
 *   - Eigen's eulerAngles(0, 1, 2) restated with float libm calls on the HOST (the product does the same on its host side).
 * Kept as in the reference, bugs included: the OBB rotation goes through eulerAngles(0,1,2) (R = Rx Ry Rz) into
 * pcl::CropBox::setRotation, which rebuilds it as Rz Ry Rx (getPointsInOrientedBox, :497-502). */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "aicp_oracle.h"

/* pose: 16 doubles column-major (Eigen::Isometry3d::matrix().data()) */
static void iso_inverse(const double* P, double* Q) {
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Q[c * 4 + r] = P[r * 4 + c];
  for (int r = 0; r < 3; ++r) Q[12 + r] = -((Q[0 * 4 + r] * P[12] + Q[1 * 4 + r] * P[13]) + Q[2 * 4 + r] * P[14]);
  Q[3] = Q[7] = Q[11] = 0.0; Q[15] = 1.0;
}

/* one direction of overlapFilter: points of `cloud` seen from `pose_other`; returns the number accepted */
static int64_t fov_one(const float* cloud, int64_t n, const double* pose_other, float range, double cos_thr, int thr_positive, float* out) {
  double Pi[16];
  iso_inverse(pose_other, Pi);
  float Rf[9], tf[3];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) Rf[3 * r + c] = (float)pose_other[c * 4 + r]; tf[r] = (float)pose_other[12 + r]; }
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    const double x = cloud[4 * i], y = cloud[4 * i + 1], z = cloud[4 * i + 2];
    const float lx = (float)(((Pi[0] * x + Pi[4] * y) + Pi[8] * z) + Pi[12]);
    const float ly = (float)(((Pi[1] * x + Pi[5] * y) + Pi[9] * z) + Pi[13]);
    const float lz = (float)(((Pi[2] * x + Pi[6] * y) + Pi[10] * z) + Pi[14]);
    const double dx = lx, dy = ly, dz = lz;
    const float r = (float)sqrt((dx * dx + dy * dy) + dz * dz);
    const double h = sqrt(dx * dx + dy * dy);
    const int in_fov = h > 0.0 ? (dx / h > cos_thr) : thr_positive;
    if (!(in_fov && r < range)) continue;
    if (out) {
      out[4 * m + 0] = ((Rf[0] * lx + Rf[1] * ly) + Rf[2] * lz) + tf[0];
      out[4 * m + 1] = ((Rf[3] * lx + Rf[4] * ly) + Rf[5] * lz) + tf[1];
      out[4 * m + 2] = ((Rf[6] * lx + Rf[7] * ly) + Rf[8] * lz) + tf[2];
      out[4 * m + 3] = 1.0f;
    }
    ++m;
  }
  return m;
}

/* overlapFilter.  outA / outB: capacity nA / nB x 4 (nullable).  counts = {accepted A, accepted B}.  Returns the percentage. */
float orc_fov_overlap(const float* A, int64_t nA, const float* B, int64_t nB, const double* poseA, const double* poseB, float range,
                      float angular_view, float* outA, float* outB, int64_t* counts) {
  const float thresh = (float)(180.0 - ((360.0 - (double)angular_view) / 2));
  const double cos_thr = cos((double)thresh * M_PI / 180.0);
  counts[0] = fov_one(A, nA, poseB, range, cos_thr, thresh > 0.f, outA);
  counts[1] = fov_one(B, nB, poseA, range, cos_thr, thresh > 0.f, outB);
  const float pa = (float)counts[0] / (float)nA, pb = (float)counts[1] / (float)nB;
  const float overlap = pa * pb;
  return (float)((double)overlap * 100.0);
}

/* ---- per-cluster statistics -------------------------------------------------------------------------------------- */
typedef struct {
  int64_t n;
  float ncen[3];          /* computeNormalsCentroid */
  float mean[3];          /* MomentOfInertiaEstimation::computeMeanValue */
  float axis[9];          /* major, middle, minor as COLUMNS of the row-major 3x3 obb_rotational_matrix */
  float bmin[3], bmax[3], pos[3];
  int64_t snn[6];         /* sum n n^T, 2^-30 */
} cluster_t;

static void jacobi3d(double a[3][3], double v[3][3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 2; ++p) for (int q = p + 1; q < 3; ++q) off = off + a[p][q] * a[p][q];
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        double apq = a[p][q];
        if (apq == 0.0) continue;
        double app = a[p][p], aqq = a[q][q];
        double theta = (aqq - app) / (2.0 * apq);
        double t;
        if (theta >= 0.0) t = 1.0 / (theta + sqrt(theta * theta + 1.0));
        else t = -1.0 / (-theta + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0);
        double s = t * c;
        a[p][p] = app - t * apq;
        a[q][q] = aqq + t * apq;
        a[p][q] = 0.0; a[q][p] = 0.0;
        for (int r = 0; r < 3; ++r) {
          if (r == p || r == q) continue;
          double arp = a[r][p], arq = a[r][q];
          double nrp = c * arp - s * arq, nrq = s * arp + c * arq;
          a[r][p] = nrp; a[p][r] = nrp; a[r][q] = nrq; a[q][r] = nrq;
        }
        for (int r = 0; r < 3; ++r) {
          double vrp = v[r][p], vrq = v[r][q];
          v[r][p] = c * vrp - s * vrq;
          v[r][q] = s * vrp + c * vrq;
        }
      }
    }
  }
}

/* unit column `col` of v in canonical sign, as floats */
static void canonical_axis(double v[3][3], int col, float* out) {
  double x = v[0][col], y = v[1][col], z = v[2][col];
  double nn = sqrt((x * x + y * y) + z * z);
  x = x / nn; y = y / nn; z = z / nn;
  double lead = x, al = fabs(x);
  if (fabs(y) > al) { lead = y; al = fabs(y); }
  if (fabs(z) > al) { lead = z; al = fabs(z); }
  if (lead < 0.0) { x = -x; y = -y; z = -z; }
  out[0] = (float)x; out[1] = (float)y; out[2] = (float)z;
}

/* Eigen 3.3 Matrix3f::eulerAngles(0, 1, 2) [UPSTREAM, recalled]; R row-major */
void orc_euler_angles_012(const float* R, float* rpy) {
  float r0 = atan2f(R[3 * 1 + 2], R[3 * 2 + 2]);
  const float c2 = sqrtf(R[0] * R[0] + R[1] * R[1]);
  float r1;
  if (r0 > 0.f) { r0 = r0 - (float)M_PI; r1 = atan2f(-R[2], -c2); }
  else r1 = atan2f(-R[2], c2);
  const float s1 = sinf(r0), c1 = cosf(r0);
  const float r2 = atan2f(s1 * R[3 * 2 + 0] - c1 * R[3 * 1 + 0], c1 * R[3 * 1 + 1] - s1 * R[3 * 2 + 1]);
  rpy[0] = -r0; rpy[1] = -r1; rpy[2] = -r2;
}

static void cluster_stats(const float* pts, const float* nrm, const int32_t* labels, int64_t m, int64_t n_clusters, cluster_t* cl) {
  int64_t (*sp)[3] = calloc((size_t)n_clusters, sizeof(*sp));
  int64_t (*sn)[3] = calloc((size_t)n_clusters, sizeof(*sn));
  int64_t (*sc)[6] = calloc((size_t)n_clusters, sizeof(*sc));
  for (int64_t c = 0; c < n_clusters; ++c) { cl[c].n = 0; memset(cl[c].snn, 0, sizeof(cl[c].snn)); }
  for (int64_t i = 0; i < m; ++i) {
    const int32_t c = labels[i];
    if (c < 0) continue;
    const float* p = pts + 4 * i; const float* q = nrm + 4 * i;
    cl[c].n++;
    for (int d = 0; d < 3; ++d) { sp[c][d] += llrint((double)p[d] * 1048576.0); sn[c][d] += llrint((double)q[d] * 1073741824.0); }
    int t = 0;
    for (int a = 0; a < 3; ++a) for (int b = a; b < 3; ++b) cl[c].snn[t++] += llrint(((double)q[a] * (double)q[b]) * 1073741824.0);
  }
  for (int64_t c = 0; c < n_clusters; ++c) {
    const double cnt = (double)cl[c].n;
    for (int d = 0; d < 3; ++d) {
      cl[c].mean[d] = (float)(((double)sp[c][d] / cnt) * (1.0 / 1048576.0));
      cl[c].ncen[d] = (float)(((double)sn[c][d] / cnt) * (1.0 / 1073741824.0));
    }
  }
  for (int64_t i = 0; i < m; ++i) {
    const int32_t c = labels[i];
    if (c < 0) continue;
    const float* p = pts + 4 * i;
    const float d[3] = {p[0] - cl[c].mean[0], p[1] - cl[c].mean[1], p[2] - cl[c].mean[2]};
    int t = 0;
    for (int a = 0; a < 3; ++a) for (int b = a; b < 3; ++b) sc[c][t++] += llrint(((double)d[a] * (double)d[b]) * 1073741824.0);
  }
  for (int64_t c = 0; c < n_clusters; ++c) {
    const double cnt = (double)cl[c].n, s = 1.0 / 1073741824.0;
    double a[3][3], v[3][3];
    a[0][0] = (double)sc[c][0] * s / cnt; a[0][1] = a[1][0] = (double)sc[c][1] * s / cnt; a[0][2] = a[2][0] = (double)sc[c][2] * s / cnt;
    a[1][1] = (double)sc[c][3] * s / cnt; a[1][2] = a[2][1] = (double)sc[c][4] * s / cnt; a[2][2] = (double)sc[c][5] * s / cnt;
    jacobi3d(a, v);
    /* MomentOfInertiaEstimation::computeEigenVectors: three compare-swaps on the indices */
    int major = 0, middle = 1, minor = 2, tmp;
    if (a[major][major] < a[middle][middle]) { tmp = major; major = middle; middle = tmp; }
    if (a[major][major] < a[minor][minor]) { tmp = major; major = minor; minor = tmp; }
    if (a[middle][middle] < a[minor][minor]) { tmp = minor; minor = middle; middle = tmp; }
    float ax[3][3];
    canonical_axis(v, major, ax[0]); canonical_axis(v, middle, ax[1]); canonical_axis(v, minor, ax[2]);
    const float cx = ax[1][1] * ax[2][2] - ax[1][2] * ax[2][1], cy = ax[1][2] * ax[2][0] - ax[1][0] * ax[2][2],
                cz = ax[1][0] * ax[2][1] - ax[1][1] * ax[2][0];
    const float det = (ax[0][0] * cx + ax[0][1] * cy) + ax[0][2] * cz;
    if (det <= 0.f) { ax[0][0] = -ax[0][0]; ax[0][1] = -ax[0][1]; ax[0][2] = -ax[0][2]; }
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) cl[c].axis[3 * r + k] = ax[k][r];      /* axes as columns */
    for (int k = 0; k < 3; ++k) { cl[c].bmin[k] = FLT_MAX; cl[c].bmax[k] = -FLT_MAX; }
  }
  for (int64_t i = 0; i < m; ++i) {
    const int32_t c = labels[i];
    if (c < 0) continue;
    const float* p = pts + 4 * i;
    const float d[3] = {p[0] - cl[c].mean[0], p[1] - cl[c].mean[1], p[2] - cl[c].mean[2]};
    for (int k = 0; k < 3; ++k) {
      const float v = (d[0] * cl[c].axis[0 + k] + d[1] * cl[c].axis[3 + k]) + d[2] * cl[c].axis[6 + k];
      if (v < cl[c].bmin[k]) cl[c].bmin[k] = v;
      if (v > cl[c].bmax[k]) cl[c].bmax[k] = v;
    }
  }
  for (int64_t c = 0; c < n_clusters; ++c) {      /* computeOBB: centre the box */
    float shift[3];
    for (int k = 0; k < 3; ++k) { shift[k] = (cl[c].bmax[k] + cl[c].bmin[k]) / 2.0f; cl[c].bmin[k] -= shift[k]; cl[c].bmax[k] -= shift[k]; }
    for (int r = 0; r < 3; ++r)
      cl[c].pos[r] = cl[c].mean[r] + ((cl[c].axis[3 * r] * shift[0] + cl[c].axis[3 * r + 1] * shift[1]) + cl[c].axis[3 * r + 2] * shift[2]);
  }
  free(sp); free(sn); free(sc);
}

/* the CropBox that overlapBoxFilter builds around one cluster (:515-531 / :543-559): M = inverse rotation (row-major) */
typedef struct { float M[9], t[3], bmin[3], bmax[3]; } crop_t;

static void cluster_box(const cluster_t* c, crop_t* b) {
  float rpy[3], R[9];
  orc_euler_angles_012(c->axis, rpy);            /* rotational_matrix_OBB.eulerAngles(0, 1, 2), :492 */
  orc_rpy_to_matrix(rpy, R);                     /* CropBox::setRotation -> pcl::getTransformation */
  for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) b->M[3 * r + k] = R[3 * k + r];
  for (int k = 0; k < 3; ++k) { b->t[k] = c->pos[k]; b->bmin[k] = c->bmin[k]; b->bmax[k] = c->bmax[k]; }
  b->bmin[2] = (float)(3.0 * (double)c->bmin[2]);  /* "direction perpendicular to plane", :522-523 */
  b->bmax[2] = (float)(3.0 * (double)c->bmax[2]);
}

static inline int in_box(const crop_t* b, const float* p) {
  const float dx = p[0] - b->t[0], dy = p[1] - b->t[1], dz = p[2] - b->t[2];
  const float lx = (b->M[0] * dx + b->M[1] * dy) + b->M[2] * dz;
  const float ly = (b->M[3] * dx + b->M[4] * dy) + b->M[5] * dz;
  const float lz = (b->M[6] * dx + b->M[7] * dy) + b->M[8] * dz;
  return !(lx < b->bmin[0] || ly < b->bmin[1] || lz < b->bmin[2] || lx > b->bmax[0] || ly > b->bmax[1] || lz > b->bmax[2]);
}

/* counts[i * nY + j] = points of Y's cluster j inside the box of X's cluster i */
static void pair_counts(const crop_t* boxes, int64_t nX, const float* ptsY, const int32_t* labY, int64_t mY, int64_t nY, int64_t* counts) {
  memset(counts, 0, sizeof(int64_t) * (size_t)(nX * nY));
  for (int64_t p = 0; p < mY; ++p) {
    const int32_t j = labY[p];
    if (j < 0) continue;
    for (int64_t i = 0; i < nX; ++i) if (in_box(&boxes[i], ptsY + 4 * p)) counts[i * nY + j]++;
  }
}

/* alignabilityFilter.  matching (nullable, capacity = number of B clusters): matching_indeces of :229-282.
 * info = {clusters A, clusters B, matched}.  Returns the alignability in percent (0 when nothing matches, :331-335). */
int orc_alignability(const float* A, int64_t nA, const float* B, int64_t nB, const double* poseA, const double* poseB,
                     const orc_prefilter_config* cfg, int threads, float* out_alignability, int32_t* matching, int64_t* info) {
  *out_alignability = 0.f;
  info[0] = info[1] = info[2] = 0;
  const float* in[2] = {A, B};
  const int64_t nn[2] = {nA, nB};
  const double* pose[2] = {poseA, poseB};
  float* smp[2] = {NULL, NULL}; float* nrm[2] = {NULL, NULL}; int32_t* lab[2] = {NULL, NULL};
  int64_t cnt[2][3];
  cluster_t* cl[2] = {NULL, NULL};
  int rc = ORC_OK;
  for (int s = 0; s < 2 && rc == ORC_OK; ++s) {
    const size_t cap = (size_t)(nn[s] > 0 ? nn[s] : 1);
    smp[s] = malloc(sizeof(float) * 4 * cap); nrm[s] = malloc(sizeof(float) * 4 * cap); lab[s] = malloc(sizeof(int32_t) * cap);
    const float vp[3] = {(float)pose[s][12], (float)pose[s][13], (float)pose[s][14]};
    rc = orc_prefilter(in[s], nn[s], cfg, vp, threads, smp[s], nrm[s], lab[s], NULL, cnt[s]);
    if (rc == ORC_OK) {
      cl[s] = malloc(sizeof(cluster_t) * (size_t)(cnt[s][1] > 0 ? cnt[s][1] : 1));
      cluster_stats(smp[s], nrm[s], lab[s], cnt[s][0], cnt[s][1], cl[s]);
    }
  }
  if (rc == ORC_OK) {
    const int64_t kA = cnt[0][1], kB = cnt[1][1];
    info[0] = kA; info[1] = kB;
    crop_t* bA = malloc(sizeof(crop_t) * (size_t)(kA > 0 ? kA : 1));
    crop_t* bB = malloc(sizeof(crop_t) * (size_t)(kB > 0 ? kB : 1));
    for (int64_t i = 0; i < kA; ++i) cluster_box(&cl[0][i], &bA[i]);
    for (int64_t j = 0; j < kB; ++j) cluster_box(&cl[1][j], &bB[j]);
    int64_t* b_in_a = malloc(sizeof(int64_t) * (size_t)(kA * kB + 1));   /* [i * kB + j] */
    int64_t* a_in_b = malloc(sizeof(int64_t) * (size_t)(kA * kB + 1));   /* [j * kA + i] */
    pair_counts(bA, kA, smp[1], lab[1], cnt[1][0], kB, b_in_a);
    pair_counts(bB, kB, smp[0], lab[0], cnt[0][0], kA, a_in_b);
    int32_t* mi = malloc(sizeof(int32_t) * (size_t)(kB + 1));
    float* mo = malloc(sizeof(float) * (size_t)(kB + 1));
    for (int64_t j = 0; j < kB; ++j) { mi[j] = -1; mo[j] = -1.f; }
    for (int64_t i = 0; i < kA; ++i) {
      float max_overlap = 0.f;
      int64_t best = -1;
      const float* ca = cl[0][i].ncen;
      for (int64_t j = 0; j < kB; ++j) {
        const float* cb = cl[1][j].ncen;
        const float dot = (ca[0] * cb[0] + ca[1] * cb[1]) + ca[2] * cb[2];
        const float na = sqrtf((ca[0] * ca[0] + ca[1] * ca[1]) + ca[2] * ca[2]), nb = sqrtf((cb[0] * cb[0] + cb[1] * cb[1]) + cb[2] * cb[2]);
        const float dist = (float)((double)acosf(dot / (na * nb)) * 180.0 / M_PI);
        const float perc_a = (float)a_in_b[j * kA + i] / (float)cl[0][i].n;
        const float perc_b = (float)b_in_a[i * kB + j] / (float)cl[1][j].n;
        const float ov = perc_a * perc_b;
        const float current = (float)((double)ov * 100.0);
        if (current > max_overlap && dist < 20) { best = j; max_overlap = current; }
      }
      if (max_overlap > 0.f) {
        if (mi[best] == -1 || max_overlap > mo[best]) { mi[best] = (int32_t)i; mo[best] = max_overlap; }
      }
    }
    int64_t S[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t j = 0; j < kB; ++j) {
      if (matching) matching[j] = mi[j];
      if (mi[j] < 0) continue;
      info[2]++;
      for (int t = 0; t < 6; ++t) S[t] += cl[0][mi[j]].snn[t];
    }
    if (info[2] > 0) {
      /* pcl::PCA on the normals and their mirror images: mean 0, covariance proportional to sum n n^T */
      double a[3][3], v[3][3];
      a[0][0] = (double)S[0]; a[0][1] = a[1][0] = (double)S[1]; a[0][2] = a[2][0] = (double)S[2];
      a[1][1] = (double)S[3]; a[1][2] = a[2][1] = (double)S[4]; a[2][2] = (double)S[5];
      jacobi3d(a, v);
      double l[3] = {a[0][0], a[1][1], a[2][2]}, t;
      if (l[0] < l[1]) { t = l[0]; l[0] = l[1]; l[1] = t; }
      if (l[0] < l[2]) { t = l[0]; l[0] = l[2]; l[2] = t; }
      if (l[1] < l[2]) { t = l[1]; l[1] = l[2]; l[2] = t; }
      const double sum = (l[0] + l[1]) + l[2];
      const float lambda0 = (float)(l[0] / sum), lambda2 = (float)(l[2] / sum);
      const float scattering = lambda2 / lambda0;
      *out_alignability = (float)((double)scattering * 100.0);
    }
    free(bA); free(bB); free(b_in_a); free(a_in_b); free(mi); free(mo);
  }
  for (int s = 0; s < 2; ++s) { free(smp[s]); free(nrm[s]); free(lab[s]); free(cl[s]); }
  return rc;
}
