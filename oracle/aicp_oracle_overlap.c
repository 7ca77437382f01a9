/*
 * aicp_oracle_overlap.c -- CPU ORACLE for the octree overlap parameter (test infrastructure only; PARITY UNPINNED).
 *
 * Restates aicp::OctreesOverlap::computeOverlap (aicp_core/src/overlap/octrees_overlap.cpp:29-72) with
 * createTree (:153-218), convertPointCloudToScanGraph (:220-241) and getOverlappingNodes (:113-151), plus the
 * octomap 1.9.x pieces they call ([UPSTREAM]): coordToKey / keyToCoord (OcTreeBaseImpl.hxx),
 * computeRayKeys (Amanatides-Woo DDA), OccupancyOcTreeBase::computeUpdate / insertPointCloud.
 *
 * Because createTree force-marks every leaf occupied after ray casting (octrees_overlap.cpp:205-215) and both
 * trees are fully expanded before counting (:115-116), the result is a pure set computation on depth-16 keys:
 *   A = keys(ref rays U ref end voxels), B likewise for the reading,
 *   overlap = 100 * min(|A^B|/|A|, |A^B|/|B|)  in float32 (:47-53).
 */
#include "aicp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TREE_MAX_VAL 32768

/* ---- open-addressing set of 48-bit keys ---- */
typedef struct { uint64_t* slot; uint64_t cap, count; } keyset;
#define EMPTY_SLOT 0xFFFFFFFFFFFFFFFFull

static uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
static void ks_init(keyset* s, uint64_t cap) {
  s->cap = cap; s->count = 0;
  s->slot = (uint64_t*)malloc(sizeof(uint64_t) * cap);
  memset(s->slot, 0xFF, sizeof(uint64_t) * cap);
}
static int ks_contains(const keyset* s, uint64_t k) {
  uint64_t h = mix64(k) & (s->cap - 1);
  while (s->slot[h] != EMPTY_SLOT) { if (s->slot[h] == k) return 1; h = (h + 1) & (s->cap - 1); }
  return 0;
}
static void ks_insert(keyset* s, uint64_t k);
static void ks_grow(keyset* s) {
  keyset n; ks_init(&n, s->cap * 2);
  for (uint64_t i = 0; i < s->cap; ++i) if (s->slot[i] != EMPTY_SLOT) ks_insert(&n, s->slot[i]);
  free(s->slot); *s = n;
}
static void ks_insert(keyset* s, uint64_t k) {
  if ((s->count + 1) * 2 > s->cap) ks_grow(s);
  uint64_t h = mix64(k) & (s->cap - 1);
  while (s->slot[h] != EMPTY_SLOT) { if (s->slot[h] == k) return; h = (h + 1) & (s->cap - 1); }
  s->slot[h] = k; s->count++;
}

static inline uint64_t pack_key(const int* k) { return ((uint64_t)k[0] << 32) | ((uint64_t)k[1] << 16) | (uint64_t)k[2]; }

/* octomap coordToKeyChecked: key = int(floor(resolution_factor * coord)) + tree_max_val, multiply by the reciprocal */
static int coord_to_key(float coord, double res_factor, int* key) {
  double v = floor(res_factor * (double)coord);
  if (!(v > -1.0e9 && v < 1.0e9)) return 0;            /* NaN / huge: out of bounds */
  int scaled = (int)v + TREE_MAX_VAL;
  if (scaled >= 0 && (unsigned)scaled < 2u * TREE_MAX_VAL) { *key = scaled; return 1; }
  return 0;
}
static inline double key_to_coord(int key, double res) { return ((double)(key - TREE_MAX_VAL) + 0.5) * res; }

/* computeUpdate for one point: ray keys (origin voxel .. last voxel before the end voxel) + end voxel key */
static void insert_ray(keyset* set, const float* origin, const float* end, double res, double res_factor) {
  int ko[3], ke[3];
  int ok_o = coord_to_key(origin[0], res_factor, &ko[0]) && coord_to_key(origin[1], res_factor, &ko[1]) &&
             coord_to_key(origin[2], res_factor, &ko[2]);
  int ok_e = coord_to_key(end[0], res_factor, &ke[0]) && coord_to_key(end[1], res_factor, &ke[1]) &&
             coord_to_key(end[2], res_factor, &ke[2]);
  if (ok_e) ks_insert(set, pack_key(ke));               /* occupied end point */
  if (!ok_o || !ok_e) return;                           /* computeRayKeys returns false */
  if (ko[0] == ke[0] && ko[1] == ke[1] && ko[2] == ke[2]) return;
  ks_insert(set, pack_key(ko));
  float dir[3] = {end[0] - origin[0], end[1] - origin[1], end[2] - origin[2]};
  float nsq = dir[0] * dir[0] + dir[1] * dir[1];
  nsq = nsq + dir[2] * dir[2];
  float length = (float)sqrt((double)nsq);
  dir[0] = dir[0] / length; dir[1] = dir[1] / length; dir[2] = dir[2] / length;
  int step[3], cur[3] = {ko[0], ko[1], ko[2]};
  double tMax[3], tDelta[3];
  for (int i = 0; i < 3; ++i) {
    if (dir[i] > 0.0f) step[i] = 1; else if (dir[i] < 0.0f) step[i] = -1; else step[i] = 0;
    if (step[i] != 0) {
      double border = key_to_coord(cur[i], res);
      border += (double)(float)((double)step[i] * res * 0.5);
      tMax[i] = (border - (double)origin[i]) / (double)dir[i];
      tDelta[i] = res / fabs((double)dir[i]);
    } else { tMax[i] = DBL_MAX; tDelta[i] = DBL_MAX; }
  }
  for (int guard = 0; guard < 400000; ++guard) {
    int dim;
    if (tMax[0] < tMax[1]) dim = (tMax[0] < tMax[2]) ? 0 : 2;
    else dim = (tMax[1] < tMax[2]) ? 1 : 2;
    cur[dim] += step[dim];
    tMax[dim] += tDelta[dim];
    if (cur[0] == ke[0] && cur[1] == ke[1] && cur[2] == ke[2]) break;
    double dmin = tMax[0] < tMax[1] ? tMax[0] : tMax[1];
    if (tMax[2] < dmin) dmin = tMax[2];
    if (dmin > (double)length) break;
    if ((unsigned)cur[0] >= 65536u || (unsigned)cur[1] >= 65536u || (unsigned)cur[2] >= 65536u) break;
    ks_insert(set, pack_key(cur));
  }
}

static void build_keyset(keyset* set, const float* pts, int64_t n, const double* origin, double res) {
  double res_factor = 1.0 / res;
  /* octomap::pose6d(float x, float y, float z, ...) -- octrees_overlap.cpp:229-230 */
  float o[3] = {(float)origin[0], (float)origin[1], (float)origin[2]};
  ks_init(set, 1u << 16);
  for (int64_t i = 0; i < n; ++i) {
    const float* p = pts + 4 * i;
    if (!isfinite(p[0]) || !isfinite(p[1]) || !isfinite(p[2])) continue;
    insert_ray(set, o, p, res, res_factor);
  }
}

static int cmp_u64(const void* a, const void* b) {
  uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
  return (x > y) - (x < y);
}

int64_t orc_ray_keys(const float* pts, int64_t n, const double* origin, double resolution, uint64_t* keys, int64_t cap) {
  keyset s; build_keyset(&s, pts, n, origin, resolution);
  int64_t cnt = (int64_t)s.count;
  if (keys) {
    int64_t m = 0;
    for (uint64_t i = 0; i < s.cap && m < cap; ++i) if (s.slot[i] != EMPTY_SLOT) keys[m++] = s.slot[i];
    qsort(keys, (size_t)m, sizeof(uint64_t), cmp_u64);
  }
  free(s.slot);
  return cnt;
}

int orc_overlap(const float* ref, int64_t n_ref, const double* ref_origin, const float* read, int64_t n_read,
                const double* read_origin, double resolution, float* out_overlap_pct, int64_t* counts) {
  if (!ref || !read || !ref_origin || !read_origin || !(resolution > 0.0)) return ORC_ERR_BAD_ARG;
  keyset A, B;
  build_keyset(&A, ref, n_ref, ref_origin, resolution);
  build_keyset(&B, read, n_read, read_origin, resolution);
  int64_t inter = 0;
  for (uint64_t i = 0; i < A.cap; ++i)
    if (A.slot[i] != EMPTY_SLOT && ks_contains(&B, A.slot[i])) ++inter;
  /* octrees_overlap.cpp:47-53 */
  float treeA = (float)inter / (float)(int64_t)A.count;
  float treeB = (float)inter / (float)(int64_t)B.count;
  float mn = treeA < treeB ? treeA : treeB;          /* std::min(a,b): returns a unless b < a */
  if (treeB < treeA) mn = treeB; else mn = treeA;
  if (out_overlap_pct) *out_overlap_pct = (float)((double)mn * 100.0);
  if (counts) { counts[0] = inter; counts[1] = (int64_t)A.count; counts[2] = (int64_t)B.count; }
  free(A.slot); free(B.slot);
  return ORC_OK;
}
