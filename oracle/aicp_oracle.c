/*
 * aicp_oracle.c -- CPU ORACLE (test infrastructure only; PARITY UNPINNED, see aicp_oracle.h).
 *
 * Restates, in plain C, the libpointmatcher point-to-plane chain that
 * aicp_core/src/registration/pointmatcher_registration.cpp:92-151 runs through PM::ICP::operator(), configured by
 * aicp_core/config/icp/icp_autotuned.yaml:9-58.  Section numbers "A.n" refer to SURVEY.md Appendix A
 * ([UPSTREAM] libpointmatcher / libnabo behaviour recalled from their public sources).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).  -ffp-contract=off is REQUIRED: the float
 * operation order below is the parity contract.
 */
#include "aicp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef __int128 i128;

/* ------------------------------------------------------------------------------------------------
 * Deterministic libm subset: only + - * / sqrt and floor, so that the CUDA restatement is bit-identical.
 * ---------------------------------------------------------------------------------------------- */

/* sin/cos by Cody-Waite reduction to [-pi/4, pi/4] and forward Taylor summation. */
void orc_sincos(double x, double* s, double* c) {
  const double two_over_pi = 0.63661977236758138;      /* 0x3FE45F306DC9C883 */
  const double pio2_hi = 1.5707963267341256;           /* fdlibm pio2_1: first 33 bits of pi/2 */
  const double pio2_lo = 6.0771005065061922e-11;       /* fdlibm pio2_1t */
  double kd = floor(x * two_over_pi + 0.5);
  double r = (x - kd * pio2_hi) - kd * pio2_lo;
  double r2 = r * r;
  /* sin r: forward Taylor sum, factorial ratios applied as reciprocal multiplications; stops once a term no longer
   * changes the sum (at most 11 terms) */
  double term = r, ss = r;
  for (int n = 1; n <= 11; ++n) {
    double inv = 1.0 / (double)((2 * n) * (2 * n + 1));
    term = (term * r2) * inv;
    term = -term;
    double ns = ss + term;
    if (ns == ss) break;
    ss = ns;
  }
  /* cos r */
  double cterm = 1.0, cc = 1.0;
  for (int n = 1; n <= 11; ++n) {
    double inv = 1.0 / (double)((2 * n - 1) * (2 * n));
    cterm = (cterm * r2) * inv;
    cterm = -cterm;
    double nc = cc + cterm;
    if (nc == cc) break;
    cc = nc;
  }
  long long k = (long long)kd;
  int quad = (int)(((k % 4) + 4) % 4);
  switch (quad) {
    case 0: *s = ss;  *c = cc;  break;
    case 1: *s = cc;  *c = -ss; break;
    case 2: *s = -ss; *c = -cc; break;
    default: *s = -cc; *c = ss; break;
  }
}

/* atan(z) for z in [0,1]: three half-angle reductions, then an 11-term alternating series. */
static double orc_atan01(double z) {
  double u = z;
  for (int h = 0; h < 3; ++h) u = u / (1.0 + sqrt(1.0 + u * u));
  double u2 = u * u, p = u, sum = u;
  for (int n = 1; n <= 11; ++n) {
    double inv = 1.0 / (double)(2 * n + 1);
    p = p * u2;
    p = -p;
    double ns = sum + p * inv;
    if (ns == sum) break;
    sum = ns;
  }
  return 8.0 * sum;
}

/* atan2 for y >= 0, x >= 0 */
double orc_atan2_pos(double y, double x) {
  const double pio2 = 1.5707963267948966;
  if (y == 0.0 && x == 0.0) return 0.0;
  if (y <= x) return orc_atan01(y / x);
  return pio2 - orc_atan01(x / y);
}

/* Cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 6), float64.
 * a is overwritten (diagonal = eigenvalues), v receives eigenvectors in columns. */
static void orc_jacobi(int n, double a[6][6], double v[6][6]) {
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) off = off + a[p][q] * a[p][q];
    if (off == 0.0) break;
    for (int p = 0; p < n - 1; ++p) {
      for (int q = p + 1; q < n; ++q) {
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
        a[p][q] = 0.0;
        a[q][p] = 0.0;
        for (int r = 0; r < n; ++r) {
          if (r == p || r == q) continue;
          double arp = a[r][p], arq = a[r][q];
          double nrp = c * arp - s * arq;
          double nrq = s * arp + c * arq;
          a[r][p] = nrp; a[p][r] = nrp;
          a[r][q] = nrq; a[q][r] = nrq;
        }
        for (int r = 0; r < n; ++r) {
          double vrp = v[r][p], vrq = v[r][q];
          v[r][p] = c * vrp - s * vrq;
          v[r][q] = s * vrp + c * vrq;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Geometry primitives (float32, fixed order)
 * ---------------------------------------------------------------------------------------------- */

/* libnabo leaf scan: dist = 0; for i: diff = q[i]-p[i]; dist += diff*diff  (A.3) */
static inline float d2_f(const float* q, const float* p) {
  float dx = q[0] - p[0], dy = q[1] - p[1], dz = q[2] - p[2];
  float d = dx * dx;
  d = d + dy * dy;
  d = d + dz * dz;
  return d;
}

static inline float box_d2_f(const float* lo, const float* hi, const float* q) {
  float dx = 0.f, dy = 0.f, dz = 0.f;
  if (q[0] < lo[0]) dx = lo[0] - q[0]; else if (q[0] > hi[0]) dx = q[0] - hi[0];
  if (q[1] < lo[1]) dy = lo[1] - q[1]; else if (q[1] > hi[1]) dy = q[1] - hi[1];
  if (q[2] < lo[2]) dz = lo[2] - q[2]; else if (q[2] > hi[2]) dz = q[2] - hi[2];
  float d = dx * dx;
  d = d + dy * dy;
  d = d + dz * dz;
  return d;
}

/* RigidTransformation::compute, A.6: features <- T * features; float, k ascending, no FMA. */
static inline void xform_f(const float* T, const float* in, float* out) {
  float x = in[0], y = in[1], z = in[2];
  for (int r = 0; r < 3; ++r) {
    float acc = T[0 * 4 + r] * x;
    acc = acc + T[1 * 4 + r] * y;
    acc = acc + T[2 * 4 + r] * z;
    acc = acc + T[3 * 4 + r];
    out[r] = acc;
  }
}

void orc_transform_points(const float* T, const float* in, int64_t n, float* out) {
  for (int64_t i = 0; i < n; ++i) {
    float o[3];
    xform_f(T, in + 4 * i, o);
    out[4 * i + 0] = o[0]; out[4 * i + 1] = o[1]; out[4 * i + 2] = o[2]; out[4 * i + 3] = in[4 * i + 3];
  }
}

/* C = A * B for rigid 4x4 (column-major), bottom row fixed to (0,0,0,1); float, k ascending. */
static void mat4_mul_f(const float* A, const float* B, float* C) {
  float R[16];
  for (int c = 0; c < 4; ++c) {
    for (int r = 0; r < 3; ++r) {
      float acc = A[0 * 4 + r] * B[c * 4 + 0];
      acc = acc + A[1 * 4 + r] * B[c * 4 + 1];
      acc = acc + A[2 * 4 + r] * B[c * 4 + 2];
      if (c == 3) acc = acc + A[3 * 4 + r];
      R[c * 4 + r] = acc;
    }
    R[c * 4 + 3] = (c == 3) ? 1.f : 0.f;
  }
  memcpy(C, R, sizeof(R));
}

/* ------------------------------------------------------------------------------------------------
 * kd-tree restating libnabo's KDTREE_LINEAR_HEAP tree shape (sliding-midpoint split on the longest side,
 * bucket size 8, points in leaves; A.3) but with explicit tight node boxes so that pruning is provably
 * exact in float arithmetic: box_d2_f(node) <= d2_f(any point in node), by monotonicity of rounding.
 * ---------------------------------------------------------------------------------------------- */
#define KD_BUCKET 8

typedef struct {
  float lo[3], hi[3];
  int32_t left, right;     /* children, or -1 for leaf */
  int32_t start, count;    /* range in sorted point array */
} kd_node;

typedef struct {
  kd_node* nodes;
  int32_t n_nodes, cap_nodes;
  int32_t* perm;           /* sorted position -> original index */
  float* spts;             /* 4 floats per sorted point: x,y,z,(unused) */
  const float* pts;
  int64_t n;
} kd_tree;

static int32_t kd_new_node(kd_tree* t) {
  if (t->n_nodes == t->cap_nodes) {
    t->cap_nodes = t->cap_nodes ? t->cap_nodes * 2 : 1024;
    t->nodes = (kd_node*)realloc(t->nodes, sizeof(kd_node) * (size_t)t->cap_nodes);
  }
  return t->n_nodes++;
}

static int32_t kd_build_rec(kd_tree* t, int32_t start, int32_t count) {
  int32_t id = kd_new_node(t);
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int32_t i = start; i < start + count; ++i) {
    const float* p = t->pts + 4 * (int64_t)t->perm[i];
    for (int d = 0; d < 3; ++d) { if (p[d] < lo[d]) lo[d] = p[d]; if (p[d] > hi[d]) hi[d] = p[d]; }
  }
  for (int d = 0; d < 3; ++d) { t->nodes[id].lo[d] = lo[d]; t->nodes[id].hi[d] = hi[d]; }
  t->nodes[id].start = start; t->nodes[id].count = count;
  t->nodes[id].left = t->nodes[id].right = -1;
  int dim = 0;
  float ext = hi[0] - lo[0];
  if (hi[1] - lo[1] > ext) { ext = hi[1] - lo[1]; dim = 1; }
  if (hi[2] - lo[2] > ext) { ext = hi[2] - lo[2]; dim = 2; }
  if (count <= KD_BUCKET || !(ext > 0.f)) return id;
  float split = 0.5f * (lo[dim] + hi[dim]);
  /* partition: < split to the left */
  int32_t i = start, j = start + count - 1;
  while (i <= j) {
    if (t->pts[4 * (int64_t)t->perm[i] + dim] < split) ++i;
    else { int32_t tmp = t->perm[i]; t->perm[i] = t->perm[j]; t->perm[j] = tmp; --j; }
  }
  int32_t nl = i - start;
  if (nl == 0 || nl == count) {
    /* sliding midpoint cannot happen with a tight box and ext>0 except through rounding of the midpoint */
    nl = count / 2;
    /* fall back to a median-ish split by simple selection on dim */
    for (int32_t a = start; a < start + nl; ++a) {
      int32_t m = a;
      for (int32_t b = a + 1; b < start + count; ++b)
        if (t->pts[4 * (int64_t)t->perm[b] + dim] < t->pts[4 * (int64_t)t->perm[m] + dim]) m = b;
      int32_t tmp = t->perm[a]; t->perm[a] = t->perm[m]; t->perm[m] = tmp;
    }
  }
  int32_t l = kd_build_rec(t, start, nl);
  int32_t r = kd_build_rec(t, start + nl, count - nl);
  t->nodes[id].left = l; t->nodes[id].right = r;
  return id;
}

static kd_tree* kd_build(const float* pts, int64_t n) {
  kd_tree* t = (kd_tree*)calloc(1, sizeof(kd_tree));
  t->pts = pts; t->n = n;
  t->perm = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) t->perm[i] = (int32_t)i;
  if (n > 0) kd_build_rec(t, 0, (int32_t)n);
  t->spts = (float*)malloc(sizeof(float) * 4 * (size_t)(n > 0 ? n : 1));
  for (int64_t i = 0; i < n; ++i) memcpy(t->spts + 4 * i, pts + 4 * (int64_t)t->perm[i], 16);
  return t;
}

static void kd_free(kd_tree* t) {
  if (!t) return;
  free(t->nodes); free(t->perm); free(t->spts); free(t);
}

typedef struct { float d2; int32_t id; } cand;

static inline int cand_less(float d2a, int32_t ia, float d2b, int32_t ib) {
  return (d2a < d2b) || (d2a == d2b && ia < ib);
}

/* insert (d2,id) into ascending list of at most k entries */
static inline void knn_insert(cand* list, int32_t* cnt, int32_t k, float d2, int32_t id) {
  int32_t n = *cnt;
  if (n == k) {
    if (!cand_less(d2, id, list[k - 1].d2, list[k - 1].id)) return;
    n = k - 1;
  }
  int32_t pos = n;
  while (pos > 0 && cand_less(d2, id, list[pos - 1].d2, list[pos - 1].id)) { list[pos] = list[pos - 1]; --pos; }
  list[pos].d2 = d2; list[pos].id = id;
  *cnt = n + 1;
}

static void kd_knn_rec(const kd_tree* t, int32_t node, const float* q, int32_t k, cand* list, int32_t* cnt) {
  const kd_node* nd = &t->nodes[node];
  if (nd->left < 0) {
    for (int32_t i = nd->start; i < nd->start + nd->count; ++i)
      knn_insert(list, cnt, k, d2_f(q, t->spts + 4 * (int64_t)i), t->perm[i]);
    return;
  }
  const kd_node* L = &t->nodes[nd->left];
  const kd_node* R = &t->nodes[nd->right];
  float dl = box_d2_f(L->lo, L->hi, q), dr = box_d2_f(R->lo, R->hi, q);
  int32_t first = nd->left, second = nd->right;
  float df = dl, ds = dr;
  if (dr < dl) { first = nd->right; second = nd->left; df = dr; ds = dl; }
  /* visit when the box could still hold a pair (d2,id) smaller than the current worst: d2box <= worst.d2 */
  if (*cnt < k || df <= list[k - 1].d2) kd_knn_rec(t, first, q, k, list, cnt);
  if (*cnt < k || ds <= list[k - 1].d2) kd_knn_rec(t, second, q, k, list, cnt);
}

static void brute_knn(const float* pts, int64_t n, const float* q, int32_t k, cand* list, int32_t* cnt) {
  for (int64_t j = 0; j < n; ++j) knn_insert(list, cnt, k, d2_f(q, pts + 4 * j), (int32_t)j);
}

/* ------------------------------------------------------------------------------------------------
 * A.3  KDTreeMatcher{knn=1, epsilon=0}
 * ---------------------------------------------------------------------------------------------- */
int orc_match(const float* ref, int64_t n_ref, const float* qry, int64_t n_qry, int use_kdtree, int threads,
              int32_t* out_idx, float* out_d2) {
  if (!ref || !qry || n_ref < 1 || n_qry < 0) return ORC_ERR_BAD_ARG;
  kd_tree* t = use_kdtree ? kd_build(ref, n_ref) : NULL;
  if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 512) num_threads(threads)
  for (int64_t i = 0; i < n_qry; ++i) {
    cand best; int32_t cnt = 0;
    if (t) kd_knn_rec(t, 0, qry + 4 * i, 1, &best, &cnt);
    else brute_knn(ref, n_ref, qry + 4 * i, 1, &best, &cnt);
    out_idx[i] = best.id;
    out_d2[i] = best.d2;
  }
  kd_free(t);
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * A.2  SurfaceNormalDataPointsFilter{knn, keepNormals, keepDensities}
 * ---------------------------------------------------------------------------------------------- */

/* neighbours (sorted by (d2,id)) -> unit normal + density.  float64, sequential in list order. */
static void normal_from_neighbours(const float* pts, const cand* list, int32_t k, float* out4) {
  double mx = 0, my = 0, mz = 0;
  for (int32_t j = 0; j < k; ++j) {
    const float* p = pts + 4 * (int64_t)list[j].id;
    mx = mx + (double)p[0]; my = my + (double)p[1]; mz = mz + (double)p[2];
  }
  double kd = (double)k;
  mx = mx / kd; my = my / kd; mz = mz / kd;
  double cxx = 0, cxy = 0, cxz = 0, cyy = 0, cyz = 0, czz = 0, r2max = 0;
  for (int32_t j = 0; j < k; ++j) {
    const float* p = pts + 4 * (int64_t)list[j].id;
    double dx = (double)p[0] - mx, dy = (double)p[1] - my, dz = (double)p[2] - mz;
    cxx = cxx + dx * dx; cxy = cxy + dx * dy; cxz = cxz + dx * dz;
    cyy = cyy + dy * dy; cyz = cyz + dy * dz; czz = czz + dz * dz;
    double r2 = dx * dx + dy * dy;
    r2 = r2 + dz * dz;
    if (r2 > r2max) r2max = r2;
  }
  double a[6][6], v[6][6];
  memset(a, 0, sizeof(a));
  a[0][0] = cxx; a[0][1] = cxy; a[0][2] = cxz;
  a[1][0] = cxy; a[1][1] = cyy; a[1][2] = cyz;
  a[2][0] = cxz; a[2][1] = cyz; a[2][2] = czz;
  orc_jacobi(3, a, v);
  double l0 = a[0][0], l1 = a[1][1], l2 = a[2][2];
  /* A.2: "first strict minimum when scanning j = 0..2" */
  int smallest = 0; double sv = l0;
  if (l1 < sv) { smallest = 1; sv = l1; }
  if (l2 < sv) { smallest = 2; sv = l2; }
  double lmax = l0; if (l1 > lmax) lmax = l1; if (l2 > lmax) lmax = l2;
  /* middle eigenvalue by selection (median of three), never by arithmetic */
  double mn01 = l0 < l1 ? l0 : l1, mx01 = l0 < l1 ? l1 : l0;
  double t2 = mx01 < l2 ? mx01 : l2;
  double lmid = mn01 > t2 ? mn01 : t2;
  double nx, ny, nz;
  /* A.2 rank test (fullPivHouseholderQr(C).rank()+1 >= 3 in float): restated as lambda_mid > 3*eps_f*lambda_max.
   * Rank <= 1 (collinear / coincident neighbourhood) -> eigenvalues (1,0,0), eigenvectors I -> normal (0,1,0). */
  const double rank_tol = 3.0 * (double)FLT_EPSILON;
  if (!(lmid > rank_tol * lmax)) {
    nx = 0.0; ny = 1.0; nz = 0.0;
  } else {
    nx = v[0][smallest]; ny = v[1][smallest]; nz = v[2][smallest];
    double nn = sqrt((nx * nx + ny * ny) + nz * nz);
    nx = nx / nn; ny = ny / nn; nz = nz / nn;
    /* sign is arbitrary upstream (no orientation step); canonical here: largest-|component| positive */
    double ax = fabs(nx), ay = fabs(ny), az = fabs(nz);
    double lead = nx; double al = ax;
    if (ay > al) { lead = ny; al = ay; }
    if (az > al) { lead = nz; al = az; }
    if (lead < 0.0) { nx = -nx; ny = -ny; nz = -nz; }
  }
  out4[0] = (float)nx; out4[1] = (float)ny; out4[2] = (float)nz;
  /* density = k / (4/3 pi r^3), r = max_j ||NN_j|| */
  const double four_thirds_pi = 4.1887902047863905;
  double r = sqrt(r2max);
  double vol = four_thirds_pi * ((r * r) * r);
  out4[3] = (float)(kd / vol);
}

int orc_surface_normals(const float* pts, int64_t n, int32_t k, int use_kdtree, int threads,
                        float* out_normals, int32_t* out_knn) {
  if (!pts || n < 1 || k < 1) return ORC_ERR_BAD_ARG;
  if (k >= n) return ORC_ERR_KNN_TOO_LARGE;     /* A.2: requires knn < N */
  kd_tree* t = use_kdtree ? kd_build(pts, n) : NULL;
  if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
  {
    cand* list = (cand*)malloc(sizeof(cand) * (size_t)k);
#pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < n; ++i) {
      int32_t cnt = 0;
      if (t) kd_knn_rec(t, 0, pts + 4 * i, k, list, &cnt);
      else brute_knn(pts, n, pts + 4 * i, k, list, &cnt);
      if (out_knn) for (int32_t j = 0; j < k; ++j) out_knn[i * k + j] = list[j].id;
      if (out_normals) normal_from_neighbours(pts, list, k, out_normals + 4 * i);
    }
    free(list);
  }
  kd_free(t);
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * A.4  TrimmedDistOutlierFilter
 * ---------------------------------------------------------------------------------------------- */
static int cmp_float(const void* a, const void* b) {
  float x = *(const float*)a, y = *(const float*)b;
  return (x > y) - (x < y);
}

int orc_trim_threshold(const float* d2, int64_t n, float ratio, float* out_limit, int64_t* out_n_valid) {
  if (!(ratio > 0.f) || ratio > 1.f) return ORC_ERR_BAD_ARG;
  float* vals = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  int64_t m = 0;
  for (int64_t i = 0; i < n; ++i)
    if (d2[i] != INFINITY && d2[i] > 0.f) vals[m++] = d2[i];   /* Matches::getDistsQuantile filter */
  if (out_n_valid) *out_n_valid = m;
  if (m == 0) { free(vals); return ORC_ERR_NO_VALID_MATCH; }
  qsort(vals, (size_t)m, sizeof(float), cmp_float);
  int64_t idx;
  if (ratio == 1.0f) idx = m - 1;
  else {
    float fi = (float)m * ratio;               /* size_t * float evaluated in float32, then truncated */
    idx = (int64_t)fi;
    if (idx > m - 1) idx = m - 1;              /* guard (upstream would index past the end) */
  }
  *out_limit = vals[idx];
  free(vals);
  return ORC_OK;
}

/* ------------------------------------------------------------------------------------------------
 * A.5  PointToPlaneErrorMinimizer: exact fixed-point accumulation of A = sum F F^T, g = sum F (delta.n)
 * ---------------------------------------------------------------------------------------------- */
static inline i128 fixed_term(float a, float b) {
  /* product of two floats is exact in double; scaling by 2^30 is exact; llrint rounds half to even */
  double prod = (double)a * (double)b;
  return (i128)llrint(prod * 1073741824.0);
}

static void accumulate_pair(const float* p, const float* q, const float* nrm, i128* sums) {
  float F[6];
  float t0, t1;
  t0 = p[1] * nrm[2]; t1 = p[2] * nrm[1]; F[0] = t0 - t1;      /* c = p x n */
  t0 = p[2] * nrm[0]; t1 = p[0] * nrm[2]; F[1] = t0 - t1;
  t0 = p[0] * nrm[1]; t1 = p[1] * nrm[0]; F[2] = t0 - t1;
  F[3] = nrm[0]; F[4] = nrm[1]; F[5] = nrm[2];
  float ddx = p[0] - q[0], ddy = p[1] - q[1], ddz = p[2] - q[2];
  float r = ddx * nrm[0];
  r = r + ddy * nrm[1];
  r = r + ddz * nrm[2];
  int s = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) sums[s++] += fixed_term(F[i], F[j]);
  for (int i = 0; i < 6; ++i) sums[s++] += fixed_term(F[i], r);
}

int orc_normal_equations(const float* p, int64_t n, const float* ref, const float* normals, const int32_t* idx,
                         const float* d2, float limit, int64_t* sums_hi, uint64_t* sums_lo, int64_t* out_n_used) {
  i128 sums[27];
  memset(sums, 0, sizeof(sums));
  int64_t used = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (!(d2[i] <= limit)) continue;             /* weight = (d2 <= limit) ? 1 : 0 */
    accumulate_pair(p + 4 * i, ref + 4 * (int64_t)idx[i], normals + 4 * (int64_t)idx[i], sums);
    ++used;
  }
  for (int s = 0; s < 27; ++s) {
    sums_hi[s] = (int64_t)(sums[s] >> 64);
    sums_lo[s] = (uint64_t)sums[s];
  }
  if (out_n_used) *out_n_used = used;
  return ORC_OK;
}

static double fixed128_to_double(int64_t hi, uint64_t lo) {
  /* sign-magnitude so that the common case (|v| < 2^64) converts with a single rounding */
  int neg = hi < 0;
  uint64_t mh = (uint64_t)hi, ml = lo;
  if (neg) { ml = ~ml + 1u; mh = ~mh + (ml == 0 ? 1u : 0u); }
  double v = (double)mh * 18446744073709551616.0 + (double)ml;
  v = v * (1.0 / 1073741824.0);
  return neg ? -v : v;
}

int orc_solve6(const int64_t* sums_hi, const uint64_t* sums_lo, double* x) {
  double A[6][6], b[6];
  int s = 0;
  for (int i = 0; i < 6; ++i)
    for (int j = i; j < 6; ++j) { double v = fixed128_to_double(sums_hi[s], sums_lo[s]); A[i][j] = v; A[j][i] = v; ++s; }
  for (int i = 0; i < 6; ++i) { b[i] = -fixed128_to_double(sums_hi[s], sums_lo[s]); ++s; }

  /* LLT (A.5: x = A.llt().solve(b) when A is invertible) */
  const double rtol = 6.0 * (double)FLT_EPSILON;   /* stands in for fullPivHouseholderQr(A).isInvertible() in float */
  double L[6][6];
  memset(L, 0, sizeof(L));
  int ok = 1;
  for (int j = 0; j < 6 && ok; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d = d - L[j][k] * L[j][k];
    if (!(d > rtol * A[j][j]) || !(d > 0.0)) { ok = 0; break; }
    double ljj = sqrt(d);
    L[j][j] = ljj;
    for (int i = j + 1; i < 6; ++i) {
      double v = A[i][j];
      for (int k = 0; k < j; ++k) v = v - L[i][k] * L[j][k];
      L[i][j] = v / ljj;
    }
  }
  if (ok) {
    double y[6];
    for (int i = 0; i < 6; ++i) {
      double v = b[i];
      for (int k = 0; k < i; ++k) v = v - L[i][k] * y[k];
      y[i] = v / L[i][i];
    }
    for (int i = 5; i >= 0; --i) {
      double v = y[i];
      for (int k = i + 1; k < 6; ++k) v = v - L[k][i] * x[k];
      x[i] = v / L[i][i];
    }
    return 1;
  }
  /* rank-revealing fallback: minimal-norm solution through the symmetric eigen-decomposition */
  double a[6][6], v[6][6];
  memcpy(a, A, sizeof(a));
  orc_jacobi(6, a, v);
  double lmax = 0.0;
  for (int i = 0; i < 6; ++i) { double l = fabs(a[i][i]); if (l > lmax) lmax = l; }
  for (int i = 0; i < 6; ++i) x[i] = 0.0;
  for (int e = 0; e < 6; ++e) {
    double l = a[e][e];
    if (!(l > rtol * lmax)) continue;
    double proj = 0.0;
    for (int i = 0; i < 6; ++i) proj = proj + v[i][e] * b[i];
    double coef = proj / l;
    for (int i = 0; i < 6; ++i) x[i] = x[i] + coef * v[i][e];
  }
  return 2;
}

/* A.5 pose increment: R = AngleAxis(|w|, w/|w|), t = x[3:6]; zero rotation vector -> identity rotation. */
void orc_pose_increment(const double* x, float* dT) {
  double wx = x[0], wy = x[1], wz = x[2];
  double th2 = (wx * wx + wy * wy) + wz * wz;
  double th = sqrt(th2);
  double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  if (th > 0.0) {
    double ux = wx / th, uy = wy / th, uz = wz / th;
    double s, c;
    orc_sincos(th, &s, &c);
    double omc = 1.0 - c;
    R[0][0] = c + (ux * ux) * omc;        R[0][1] = (ux * uy) * omc - uz * s;   R[0][2] = (ux * uz) * omc + uy * s;
    R[1][0] = (uy * ux) * omc + uz * s;   R[1][1] = c + (uy * uy) * omc;        R[1][2] = (uy * uz) * omc - ux * s;
    R[2][0] = (uz * ux) * omc - uy * s;   R[2][1] = (uz * uy) * omc + ux * s;   R[2][2] = c + (uz * uz) * omc;
  }
  for (int c4 = 0; c4 < 3; ++c4) {
    for (int r = 0; r < 3; ++r) dT[c4 * 4 + r] = (float)R[r][c4];
    dT[c4 * 4 + 3] = 0.f;
  }
  dT[12] = (float)x[3]; dT[13] = (float)x[4]; dT[14] = (float)x[5]; dT[15] = 1.f;
}

/* quaternion (w,x,y,z) of the rotation block of a column-major float 4x4, float64 (Eigen's branch structure) */
static void quat_from_T(const float* T, double* q) {
#define M_(r, c) ((double)T[(c) * 4 + (r)])
  double tr = (M_(0, 0) + M_(1, 1)) + M_(2, 2);
  if (tr > 0.0) {
    double t = sqrt(tr + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (M_(2, 1) - M_(1, 2)) * t;
    q[2] = (M_(0, 2) - M_(2, 0)) * t;
    q[3] = (M_(1, 0) - M_(0, 1)) * t;
  } else {
    int i = 0;
    if (M_(1, 1) > M_(0, 0)) i = 1;
    if (M_(2, 2) > M_(i, i)) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    double t = sqrt(((M_(i, i) - M_(j, j)) - M_(k, k)) + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (M_(k, j) - M_(j, k)) * t;
    q[1 + j] = (M_(j, i) + M_(i, j)) * t;
    q[1 + k] = (M_(k, i) + M_(i, k)) * t;
  }
#undef M_
}

/* |angularDistance(a,b)| = 2 atan2(||vec(a * conj(b))||, |w(a * conj(b))|) */
static double quat_angular_distance(const double* a, const double* b) {
  double w = ((a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]) + a[3] * b[3];
  double vx = ((a[1] * b[0] - a[0] * b[1]) - a[2] * b[3]) + a[3] * b[2];
  double vy = ((a[2] * b[0] - a[0] * b[2]) - a[3] * b[1]) + a[1] * b[3];
  double vz = ((a[3] * b[0] - a[0] * b[3]) - a[1] * b[2]) + a[2] * b[1];
  double vn = sqrt((vx * vx + vy * vy) + vz * vz);
  return 2.0 * orc_atan2_pos(vn, fabs(w));
}

/* ------------------------------------------------------------------------------------------------
 * A.1  PM::ICP::operator()(reading, reference, T_init)   + AICP wrapper pointmatcher_registration.cpp:92-151
 * ---------------------------------------------------------------------------------------------- */
static int check_finite(const float* pts, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d)
      if (!isfinite(pts[4 * i + d])) return 0;
  return 1;
}

int orc_icp(const float* ref, int64_t n_ref, const float* read, int64_t n_read, const float* init_T,
            const orc_icp_config* cfg, float* out_T, float* out_reading, float* out_normals,
            int32_t* trace_idx, orc_icp_result* res) {
  if (!ref || !read || !cfg || !out_T || !res || n_ref < 1 || n_read < 1) return ORC_ERR_BAD_ARG;
  if (cfg->max_iterations < 1 || cfg->max_iterations > ORC_MAX_ITERS || cfg->smooth_length < 1) return ORC_ERR_BAD_ARG;
  if (!(cfg->ratio > 0.f) || cfg->ratio > 1.f) return ORC_ERR_BAD_ARG;
  if (cfg->knn_normals >= n_ref) return ORC_ERR_KNN_TOO_LARGE;
  if (cfg->reading_normals && cfg->knn_normals >= n_read) return ORC_ERR_KNN_TOO_LARGE;
  if (!check_finite(ref, n_ref) || !check_finite(read, n_read)) return ORC_ERR_NONFINITE_INPUT;
  memset(res, 0, sizeof(*res));
  int rc = ORC_OK;

  float* normals = (float*)malloc(sizeof(float) * 4 * (size_t)n_ref);
  float* refc = (float*)malloc(sizeof(float) * 4 * (size_t)n_ref);
  float* read0 = (float*)malloc(sizeof(float) * 4 * (size_t)n_read);
  float* step = (float*)malloc(sizeof(float) * 4 * (size_t)n_read);
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_read);
  float* d2 = (float*)malloc(sizeof(float) * (size_t)n_read);
  double (*quat_hist)[4] = (double (*)[4])malloc(sizeof(double) * 4 * (size_t)(cfg->max_iterations + 1));
  double (*tr_hist)[3] = (double (*)[3])malloc(sizeof(double) * 3 * (size_t)(cfg->max_iterations + 1));
  kd_tree* tree = NULL;

  /* step 1: referenceDataPointsFilters (icp_autotuned.yaml:18-23) on the un-centred reference */
  rc = orc_surface_normals(ref, n_ref, cfg->knn_normals, cfg->use_kdtree, cfg->threads, normals, NULL);
  if (rc) goto done;
  /* step 4 (dead work for PointToPlane, kept optional for timing fidelity): readingDataPointsFilters */
  if (cfg->reading_normals) {
    float* rn = (float*)malloc(sizeof(float) * 4 * (size_t)n_read);
    rc = orc_surface_normals(read, n_read, cfg->knn_normals, cfg->use_kdtree, cfg->threads, rn, NULL);
    free(rn);
    if (rc) goto done;
  }
  /* step 2: centre the reference on its mean (exact fixed-point sum, order independent) */
  float mu[3];
  for (int d = 0; d < 3; ++d) {
    i128 s = 0;
    for (int64_t i = 0; i < n_ref; ++i) s += (i128)llrint((double)ref[4 * i + d] * 65536.0);
    mu[d] = (float)((double)(int64_t)s / (65536.0 * (double)n_ref));
    res->mean_ref[d] = mu[d];
  }
  for (int64_t i = 0; i < n_ref; ++i) {
    for (int d = 0; d < 3; ++d) refc[4 * i + d] = ref[4 * i + d] - mu[d];
    refc[4 * i + 3] = ref[4 * i + 3];
  }
  /* step 5: T_refMean_dataIn = T_refIn_refMean^-1 * T_init ;  reading' = T_refMean_dataIn * reading */
  float Tinit[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  if (init_T) memcpy(Tinit, init_T, sizeof(Tinit));
  float M0[16];
  memcpy(M0, Tinit, sizeof(M0));
  for (int d = 0; d < 3; ++d) M0[12 + d] = Tinit[12 + d] - mu[d];
  M0[3] = M0[7] = M0[11] = 0.f; M0[15] = 1.f;
  orc_transform_points(M0, read, n_read, read0);
  for (int64_t i = 0; i < n_ref; ++i)
    for (int d = 0; d < 3; ++d)
      if (!(fabsf(refc[4 * i + d]) <= 1024.f)) { rc = ORC_ERR_EXTENT; goto done; }
  for (int64_t i = 0; i < n_read; ++i)
    for (int d = 0; d < 3; ++d)
      if (!(fabsf(read0[4 * i + d]) <= 1024.f)) { rc = ORC_ERR_EXTENT; goto done; }

  /* step 3: matcher.init(reference') */
  if (cfg->use_kdtree) tree = kd_build(refc, n_ref);

  /* step 6: iterate */
  float Titer[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  quat_from_T(Titer, quat_hist[0]);
  tr_hist[0][0] = tr_hist[0][1] = tr_hist[0][2] = 0.0;
  int hist_n = 1;
  int iterate = 1, it = 0;
  int threads = cfg->threads < 1 ? 1 : cfg->threads;
  int64_t n_used = 0;
  while (iterate) {
    orc_transform_points(Titer, read0, n_read, step);
#pragma omp parallel for schedule(dynamic, 512) num_threads(threads)
    for (int64_t i = 0; i < n_read; ++i) {
      cand best; int32_t cnt = 0;
      if (tree) kd_knn_rec(tree, 0, step + 4 * i, 1, &best, &cnt);
      else brute_knn(refc, n_ref, step + 4 * i, 1, &best, &cnt);
      idx[i] = best.id; d2[i] = best.d2;
    }
    if (trace_idx) memcpy(trace_idx + (int64_t)it * n_read, idx, sizeof(int32_t) * (size_t)n_read);
    float limit; int64_t n_valid;
    rc = orc_trim_threshold(d2, n_read, cfg->ratio, &limit, &n_valid);
    if (rc) goto done;
    int64_t hi[27]; uint64_t lo[27];
    orc_normal_equations(step, n_read, refc, normals, idx, d2, limit, hi, lo, &n_used);
    double x[6];
    orc_solve6(hi, lo, x);
    float dT[16];
    orc_pose_increment(x, dT);
    mat4_mul_f(dT, Titer, Titer);
    orc_iter_trace* trc = &res->trace[it];
    memcpy(trc->T_iter, Titer, sizeof(Titer));
    trc->limit_d2 = limit; trc->n_valid = n_valid; trc->n_used = n_used;
    trc->rot_err = NAN; trc->trans_err = NAN;
    ++it;
    /* A.7 checkers in YAML order */
    int nan_found = 0;
    for (int e = 0; e < 16; ++e) if (isnan(Titer[e])) nan_found = 1;
    if (nan_found) { rc = ORC_ERR_NAN; res->iterations = it; goto done; }
    if (it >= cfg->max_iterations) { iterate = 0; res->stop_reason = ORC_STOP_COUNTER; }
    quat_from_T(Titer, quat_hist[hist_n]);
    tr_hist[hist_n][0] = (double)Titer[12]; tr_hist[hist_n][1] = (double)Titer[13]; tr_hist[hist_n][2] = (double)Titer[14];
    ++hist_n;
    if (hist_n > cfg->smooth_length) {
      double re = 0.0, te = 0.0;
      for (int i = hist_n - 1; i >= hist_n - cfg->smooth_length; --i) {
        re = re + quat_angular_distance(quat_hist[i], quat_hist[i - 1]);
        double dx = tr_hist[i][0] - tr_hist[i - 1][0], dy = tr_hist[i][1] - tr_hist[i - 1][1], dz = tr_hist[i][2] - tr_hist[i - 1][2];
        te = te + sqrt((dx * dx + dy * dy) + dz * dz);
      }
      re = re / (double)cfg->smooth_length;
      te = te / (double)cfg->smooth_length;
      trc->rot_err = re; trc->trans_err = te;
      if (re < (double)cfg->min_diff_rot && te < (double)cfg->min_diff_trans) {
        if (iterate) res->stop_reason = ORC_STOP_DIFFERENTIAL;
        iterate = 0;
      }
    }
  }
  res->iterations = it;
  res->weighted_point_used_ratio = (float)n_used / (float)n_read;

  /* step 7: T = T_refIn_refMean * T_iter * T_refMean_dataIn (Eigen evaluates left to right) */
  {
    float Tmu[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, mu[0], mu[1], mu[2], 1};
    float Y[16];
    mat4_mul_f(Tmu, Titer, Y);
    mat4_mul_f(Y, M0, out_T);
  }
  /* pointmatcher_registration.cpp:128-131: out_read_cloud_ = T * reading (unfiltered copy) */
  if (out_reading) orc_transform_points(out_T, read, n_read, out_reading);
  if (out_normals) memcpy(out_normals, normals, sizeof(float) * 4 * (size_t)n_ref);

done:
  kd_free(tree);
  free(normals); free(refc); free(read0); free(step); free(idx); free(d2); free(quat_hist); free(tr_hist);
  return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Text glue: app.cpp:198-202 (clamp) and fileIO.cpp:194-198 (ostream << float, 6 significant digits)
 * ---------------------------------------------------------------------------------------------- */
float orc_autotune_ratio(float overlap_pct, char* text_out) {
  float r = overlap_pct / 100.0;                 /* float / double -> double -> float, as in app.cpp:198 */
  if (r < 0.25) r = 0.25;
  else if (r > 0.70) r = 0.70;
  char buf[64];
  snprintf(buf, sizeof(buf), "%g", (double)r);   /* default ostream precision(6), general notation */
  if (text_out) { strncpy(text_out, buf, 31); text_out[31] = 0; }
  return strtof(buf, NULL);
}
