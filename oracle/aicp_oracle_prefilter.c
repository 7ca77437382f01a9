/* aicp_oracle_prefilter.c -- CPU restatement of AICP's cloud pre-filter.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's CPU legs, never by the product.
 *
 * regionGrowingUniformPlaneSegmentationFilter (aicp_core/src/utils/filteringUtils.cpp:5-45 and :51-104), which App runs on
 * every reading (app.cpp:77-110), on the first cloud (app.cpp:295) and periodically on the merged map (app.cpp:486-493):
 *   pcl::VoxelGrid{leaf 0.08}  ->  pcl::NormalEstimation{KdTree, k = 30}  ->  pcl::RegionGrowing{min 50, max 1e6, 15
 *   neighbours, smoothness 3 deg, curvature 1.0}  ->  concatenation of the clusters.
 * PCL is not installed in this image and is not vendored under /root/reference (find_package(PCL), aicp_core/CMakeLists.txt),
 * so the stages below are [UPSTREAM, recalled] restatements of PCL 1.8 (Ubuntu 18.04 / ROS Melodic, README.md:57-65):
 *
 * VoxelGrid<PointXYZ>::applyFilter   inverse_leaf = 1/leaf (float); min/max over the finite points; "leaf size too small"
 *   (dx*dy*dz > INT32_MAX) returns the input unchanged; min_b = floor(min * inverse_leaf), div_b = max_b - min_b + 1;
 *   ijk = (int)(floor(p * inverse_leaf) - (float)min_b); idx = ijk . (1, div_x, div_x*div_y); points sorted by idx; one
 *   centroid per occupied voxel, emitted in ascending idx order.
 * NormalEstimation::computeFeature   k nearest neighbours of every point (itself included), computeMeanAndCovarianceMatrix
 *   = single-pass float32 accumulators (sum xx, xy, xz, yy, yz, zz, x, y, z) / k over coordinates SHIFTED by the first
 *   neighbour, cov = E[ab] - E[a]E[b].  (The shift is PCL >= 1.10's; PCL 1.8 accumulates the raw coordinates, which loses
 *   the covariance of a 0.5 m neighbourhood a few tens of metres from the origin -- at 200 m every normal is noise.  The
 *   shifted form is restated so that the result does not depend on where the world origin is.)
 *   solvePlaneParameters: eigenvector of the smallest eigenvalue, curvature = |lambda_min / trace|;
 *   flipNormalTowardsViewpoint: n -> -n when (vp - p) . n < 0.
 * RegionGrowing::extract             points sorted by curvature; the lowest-curvature unlabelled point seeds a region that
 *   grows breadth-first over the DIRECTED k-NN graph (15 neighbours): an unlabelled neighbour joins when
 *   |n_neighbour . n_current| >= cos(smoothness); it is expanded in turn when its curvature <= the curvature threshold;
 *   clusters with size in [min, max] are kept, in seed order, point indices ascending inside a cluster (assembleRegions).
 *
 * Decisions where the upstream arithmetic is not reproducible bit for bit (identical here and in CUDA):
 *   - voxel centroid: PCL sums floats in std::sort's (unstable) order; here the sum is EXACT in fixed point
 *     (llrint(x * 2^20), int64) and centroid = (float)((double)S / (double)count * 2^-20): order independent;
 *     |coordinate| must stay below 32768 m (ORC_ERR_EXTENT);
 *   - k-NN ties: neighbours ordered by (d2, index) (FLANN's order among equal distances is tree dependent);
 *   - eigen solve: PCL's eigen33 (closed-form roots with float cos/atan2) is libm dependent; here a cyclic Jacobi in float64
 *     on the SAME float32 covariance matrix, first strict minimum, unit normal, canonical sign (largest |component|
 *     positive) before the viewpoint flip; curvature = fabsf((float)lambda_min / ((c00 + c11) + c22));
 *   - seed order: (curvature, index) ascending (PCL's std::sort compares curvature only);
 *   - dot products ((a0*b0) + (a1*b1)) + (a2*b2) in float32 without FMA; cos threshold = cosf((float)smoothness).
 * The region growing below is the SEQUENTIAL algorithm as PCL runs it (seed queue); the CUDA path computes the same
 * partition as a min-label fixed point, and tests/ check that equivalence. */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "aicp_oracle.h"

typedef struct { uint32_t idx; int32_t pt; } vox_pair;

static int cmp_vox(const void* a, const void* b) {
  const vox_pair* x = (const vox_pair*)a; const vox_pair* y = (const vox_pair*)b;
  if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
  return x->pt < y->pt ? -1 : (x->pt > y->pt ? 1 : 0);
}

/* returns the number of output points (<= n), or -(error code).  out: capacity n x 4 floats (nullable: count only). */
int64_t orc_voxel_grid(const float* xyzw, int64_t n, float leaf, float* out) {
  if (n < 0 || !(leaf > 0.f)) return -ORC_ERR_BAD_ARG;
  if (n == 0) return 0;
  const float inv = 1.0f / leaf;
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  int64_t n_fin = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = xyzw + 4 * i;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    for (int d = 0; d < 3; ++d) {
      if (fabsf(p[d]) >= 32768.f) return -ORC_ERR_EXTENT;
      if (p[d] < mn[d]) mn[d] = p[d];
      if (p[d] > mx[d]) mx[d] = p[d];
    }
    ++n_fin;
  }
  if (n_fin == 0) return 0;
  int64_t dd[3];
  for (int d = 0; d < 3; ++d) dd[d] = (int64_t)((mx[d] - mn[d]) * inv) + 1;
  if (dd[0] * dd[1] * dd[2] > (int64_t)INT32_MAX) {              /* "Leaf size is too small for the input dataset" */
    if (out) memcpy(out, xyzw, sizeof(float) * 4 * (size_t)n);
    return n;
  }
  int32_t min_b[3], div_b[3];
  for (int d = 0; d < 3; ++d) {
    min_b[d] = (int32_t)floorf(mn[d] * inv);
    int32_t max_b = (int32_t)floorf(mx[d] * inv);
    div_b[d] = max_b - min_b[d] + 1;
  }
  const int32_t mul1 = div_b[0], mul2 = div_b[0] * div_b[1];
  vox_pair* v = (vox_pair*)malloc(sizeof(vox_pair) * (size_t)n_fin);
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = xyzw + 4 * i;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    int32_t i0 = (int32_t)(floorf(p[0] * inv) - (float)min_b[0]);
    int32_t i1 = (int32_t)(floorf(p[1] * inv) - (float)min_b[1]);
    int32_t i2 = (int32_t)(floorf(p[2] * inv) - (float)min_b[2]);
    v[m].idx = (uint32_t)(i0 + i1 * mul1 + i2 * mul2);
    v[m].pt = (int32_t)i;
    ++m;
  }
  qsort(v, (size_t)m, sizeof(vox_pair), cmp_vox);
  int64_t n_out = 0;
  for (int64_t a = 0; a < m;) {
    int64_t b = a;
    int64_t s[3] = {0, 0, 0};
    while (b < m && v[b].idx == v[a].idx) {
      const float* p = xyzw + 4 * (int64_t)v[b].pt;
      for (int d = 0; d < 3; ++d) s[d] += llrint((double)p[d] * 1048576.0);
      ++b;
    }
    if (out) {
      const double cnt = (double)(b - a);
      for (int d = 0; d < 3; ++d) out[4 * n_out + d] = (float)(((double)s[d] / cnt) * (1.0 / 1048576.0));
      out[4 * n_out + 3] = 1.0f;
    }
    ++n_out;
    a = b;
  }
  free(v);
  return n_out;
}

/* cyclic Jacobi, 3x3 symmetric, float64 -- the same sequence of operations as orc_jacobi(3, ...) in aicp_oracle.c */
static void jacobi3(double a[3][3], double v[3][3]) {
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

/* pcl::computePointNormal + flipNormalTowardsViewpoint for one point.  nb: k neighbour ids in (d2, id) order. */
void orc_pcl_point_normal(const float* pts, const int32_t* nb, int32_t k, const float* query, const float* viewpoint, float* out4) {
  float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  const float* K = pts + 4 * (int64_t)nb[0];          /* shift by the first neighbour (the query point itself) */
  for (int32_t j = 0; j < k; ++j) {
    const float* p = pts + 4 * (int64_t)nb[j];
    const float x = p[0] - K[0], y = p[1] - K[1], z = p[2] - K[2];
    acc[0] = acc[0] + x * x; acc[1] = acc[1] + x * y; acc[2] = acc[2] + x * z;
    acc[3] = acc[3] + y * y; acc[4] = acc[4] + y * z; acc[5] = acc[5] + z * z;
    acc[6] = acc[6] + x; acc[7] = acc[7] + y; acc[8] = acc[8] + z;
  }
  const float kf = (float)k;
  for (int i = 0; i < 9; ++i) acc[i] = acc[i] / kf;
  const float c00 = acc[0] - acc[6] * acc[6], c01 = acc[1] - acc[6] * acc[7], c02 = acc[2] - acc[6] * acc[8];
  const float c11 = acc[3] - acc[7] * acc[7], c12 = acc[4] - acc[7] * acc[8], c22 = acc[5] - acc[8] * acc[8];
  double a[3][3] = {{c00, c01, c02}, {c01, c11, c12}, {c02, c12, c22}}, v[3][3];
  jacobi3(a, v);
  int smallest = 0; double sv = a[0][0];
  if (a[1][1] < sv) { smallest = 1; sv = a[1][1]; }
  if (a[2][2] < sv) { smallest = 2; sv = a[2][2]; }
  double nx = v[0][smallest], ny = v[1][smallest], nz = v[2][smallest];
  double nn = sqrt((nx * nx + ny * ny) + nz * nz);
  nx = nx / nn; ny = ny / nn; nz = nz / nn;
  double lead = nx, al = fabs(nx);
  if (fabs(ny) > al) { lead = ny; al = fabs(ny); }
  if (fabs(nz) > al) { lead = nz; al = fabs(nz); }
  if (lead < 0.0) { nx = -nx; ny = -ny; nz = -nz; }
  float fx = (float)nx, fy = (float)ny, fz = (float)nz;
  const float vx = viewpoint[0] - query[0], vy = viewpoint[1] - query[1], vz = viewpoint[2] - query[2];
  const float cos_theta = (vx * fx + vy * fy) + vz * fz;
  if (cos_theta < 0.f) { fx = -fx; fy = -fy; fz = -fz; }
  const float eig_sum = (c00 + c11) + c22;
  out4[0] = fx; out4[1] = fy; out4[2] = fz;
  out4[3] = eig_sum != 0.f ? fabsf((float)sv / eig_sum) : 0.f;
}

typedef struct { float c; int32_t i; } curv_pair;
static int cmp_curv(const void* a, const void* b) {
  const curv_pair* x = (const curv_pair*)a; const curv_pair* y = (const curv_pair*)b;
  if (x->c < y->c) return -1;
  if (x->c > y->c) return 1;
  return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}

/* pcl::RegionGrowing::extract as PCL runs it (applySmoothRegionGrowingAlgorithm + growRegion + assembleRegions).
 * normals: m x 4 (nx, ny, nz, curvature); knn: m x k_stride neighbour ids of which the first n_nb are used.
 * labels (m): cluster ordinal among the KEPT clusters (seed order) or -1.  Returns the number of kept clusters. */
int64_t orc_region_growing(const float* normals, const int32_t* knn, int64_t m, int32_t k_stride, int32_t n_nb,
                           int32_t min_size, int32_t max_size, float cos_thr, float curv_thr, int32_t* labels) {
  curv_pair* order = (curv_pair*)malloc(sizeof(curv_pair) * (size_t)m);
  int32_t* seg = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
  int32_t* queue = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
  int64_t* seg_size = (int64_t*)malloc(sizeof(int64_t) * (size_t)(m + 1));
  for (int64_t i = 0; i < m; ++i) { order[i].c = normals[4 * i + 3]; order[i].i = (int32_t)i; seg[i] = -1; }
  qsort(order, (size_t)m, sizeof(curv_pair), cmp_curv);
  int64_t n_seg = 0, next = 0;
  while (next < m) {
    const int32_t seed = order[next].i;
    if (seg[seed] != -1) { ++next; continue; }
    int64_t head = 0, tail = 0, count = 1;
    queue[tail++] = seed;
    seg[seed] = (int32_t)n_seg;
    while (head < tail) {
      const int32_t cur = queue[head++];
      const float* nc = normals + 4 * (int64_t)cur;
      for (int32_t j = 0; j < n_nb; ++j) {
        const int32_t nb = knn[(int64_t)cur * k_stride + j];
        if (seg[nb] != -1) continue;
        const float* nn = normals + 4 * (int64_t)nb;
        const float dot = fabsf((nn[0] * nc[0] + nn[1] * nc[1]) + nn[2] * nc[2]);
        if (dot < cos_thr) continue;
        seg[nb] = (int32_t)n_seg;
        ++count;
        if (!(nn[3] > curv_thr)) queue[tail++] = nb;
      }
    }
    seg_size[n_seg++] = count;
    ++next;
  }
  int32_t* ordinal = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n_seg + 1));
  int64_t kept = 0;
  for (int64_t s = 0; s < n_seg; ++s) ordinal[s] = (seg_size[s] >= min_size && seg_size[s] <= max_size) ? (int32_t)kept++ : -1;
  for (int64_t i = 0; i < m; ++i) labels[i] = ordinal[seg[i]];
  free(order); free(seg); free(queue); free(seg_size); free(ordinal);
  return kept;
}

/* The whole pre-filter.  sampled (nullable, cap n x 4), normals (nullable, cap n x 4: nx, ny, nz, curvature), labels
 * (nullable, cap n), out (nullable, cap n x 4).  counts = {n_sampled, n_clusters, n_out}.  Returns 0 or an error code. */
int orc_prefilter(const float* xyzw, int64_t n, const orc_prefilter_config* cfg, const float* viewpoint, int threads,
                  float* sampled, float* normals, int32_t* labels, float* out, int64_t* counts) {
  static const float origin[3] = {0.f, 0.f, 0.f};
  if (!viewpoint) viewpoint = origin;
  counts[0] = counts[1] = counts[2] = 0;
  if (n == 0) return ORC_OK;
  float* s = (float*)malloc(sizeof(float) * 4 * (size_t)n);
  int64_t m = orc_voxel_grid(xyzw, n, cfg->leaf_size, s);
  if (m < 0) { free(s); return (int)-m; }
  counts[0] = m;
  if (sampled) memcpy(sampled, s, sizeof(float) * 4 * (size_t)m);
  if (m <= cfg->knn_normals) {
    /* fewer points than neighbours: every region is smaller than knn_normals + 1 <= min_cluster_size -> nothing kept */
    int rc = m < cfg->min_cluster_size ? ORC_OK : ORC_ERR_KNN_TOO_LARGE;
    if (labels) for (int64_t i = 0; i < m; ++i) labels[i] = -1;
    if (normals) memset(normals, 0, sizeof(float) * 4 * (size_t)m);
    free(s);
    return rc;
  }
  const int32_t k = cfg->knn_normals;
  int32_t* knn = (int32_t*)malloc(sizeof(int32_t) * (size_t)m * k);
  int rc = orc_surface_normals(s, m, k, 1, threads, NULL, knn);
  if (rc) { free(s); free(knn); return rc; }
  float* nrm = (float*)malloc(sizeof(float) * 4 * (size_t)m);
  for (int64_t i = 0; i < m; ++i) orc_pcl_point_normal(s, knn + i * k, k, s + 4 * i, viewpoint, nrm + 4 * i);
  int32_t* lab = (int32_t*)malloc(sizeof(int32_t) * (size_t)m);
  const float cos_thr = cosf(cfg->smoothness_threshold);
  counts[1] = orc_region_growing(nrm, knn, m, k, cfg->n_neighbours, cfg->min_cluster_size, cfg->max_cluster_size, cos_thr,
                                 cfg->curvature_threshold, lab);
  /* "*cloud_out = *cloud_out + cloud_cluster" per cluster: cluster order, ascending point index inside */
  int64_t* offs = (int64_t*)calloc((size_t)counts[1] + 1, sizeof(int64_t));
  for (int64_t i = 0; i < m; ++i) if (lab[i] >= 0) offs[lab[i] + 1]++;
  for (int64_t c = 0; c < counts[1]; ++c) offs[c + 1] += offs[c];
  counts[2] = offs[counts[1]];
  if (out) {
    for (int64_t i = 0; i < m; ++i) {
      if (lab[i] < 0) continue;
      const int64_t o = offs[lab[i]]++;
      memcpy(out + 4 * o, s + 4 * i, sizeof(float) * 4);
    }
  }
  if (normals) memcpy(normals, nrm, sizeof(float) * 4 * (size_t)m);
  if (labels) memcpy(labels, lab, sizeof(int32_t) * (size_t)m);
  free(offs); free(lab); free(nrm); free(knn); free(s);
  return ORC_OK;
}

void orc_prefilter_default_config(orc_prefilter_config* cfg) {
  cfg->leaf_size = 0.08f;                                   /* filteringUtils.cpp:12 */
  cfg->knn_normals = 30;                                    /* :22 */
  cfg->n_neighbours = 15;                                   /* :30 */
  cfg->min_cluster_size = 50;                               /* :27 */
  cfg->max_cluster_size = 1000000;                          /* :28 */
  cfg->smoothness_threshold = (float)(3.0 / 180.0 * M_PI);  /* :33 (setSmoothnessThreshold takes a float) */
  cfg->curvature_threshold = 1.0f;                          /* :34 */
}
