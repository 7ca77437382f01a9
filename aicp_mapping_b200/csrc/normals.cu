// normals.cu -- SurfaceNormalDataPointsFilter{knn, keepNormals, keepDensities} on the GPU.
//
// replaces [UPSTREAM] libpointmatcher DataPointsFilters/SurfaceNormal.cpp as configured by
// aicp_core/config/icp/icp_autotuned.yaml:18-23: per point, the knn nearest neighbours (self
// included, ordered by (d2, index)), C = sum (p - mean)(p - mean)^T, normal = eigenvector of the smallest
// eigenvalue, density = knn / (4/3 pi r_max^3).
//
// Two kernels:
//   k_knn_warp          one WARP per query point.  The sorted candidate list (32 entries, the first knn matter) lives
//                       one entry per lane; candidates arrive 32 at a time as a coalesced 512-byte chunk of the
//                       Morton-ordered cloud, each lane evaluates one, and accepted ones are inserted with a ballot +
//                       shuffle-up (warp-shuffle top-k).  The walk over the index is warp-uniform: it descends the
//                       implicit binary tree five levels at a time, the 32 lanes testing the 32 descendant boxes of a
//                       node in one coalesced 1 KiB load, nearest box first, pruning against the current k-th distance.
//   k_normals_from_knn  one THREAD per point: gathers the neighbours in list order and runs covariance + cyclic Jacobi
//                       in float64 registers (sequential in list order, so the result does not depend on the index).
// Algorithmic HBM bytes per point: 16 (query) + 4*knn (neighbour ids written) + 4*knn (read back) + 16 (normal written).
#include "detmath.cuh"
#include "handle.cuh"

namespace aicp {

struct WarpKnn {
  float qx, qy, qz;
  float ld;        // this lane's list entry: squared distance
  int lid;         // original index
  int lpos;        // position in the Morton-ordered array
  float worst_d;   // entry k-1 (uniform)
  int worst_id;
  int k;
  int own_chunk;
  int lane;
};

__device__ __forceinline__ void warp_insert(WarpKnn& w, float cd, int cid, int cpos) {
  bool less = cand_less(cd, cid, w.ld, w.lid);
  unsigned m = __ballot_sync(0xFFFFFFFFu, less);
  if (m == 0) return;
  int p = __ffs(m) - 1;
  float ud = __shfl_up_sync(0xFFFFFFFFu, w.ld, 1);
  int uid = __shfl_up_sync(0xFFFFFFFFu, w.lid, 1);
  int upos = __shfl_up_sync(0xFFFFFFFFu, w.lpos, 1);
  if (w.lane > p) { w.ld = ud; w.lid = uid; w.lpos = upos; }
  else if (w.lane == p) { w.ld = cd; w.lid = cid; w.lpos = cpos; }
  w.worst_d = __shfl_sync(0xFFFFFFFFu, w.ld, w.k - 1);
  w.worst_id = __shfl_sync(0xFFFFFFFFu, w.lid, w.k - 1);
}

// 32 consecutive Morton-ordered points: one candidate per lane
__device__ __forceinline__ void scan_chunk(const IndexView& ix, WarpKnn& w, int chunk) {
  int pos = chunk * 32 + w.lane;
  float4 p = __ldg(&ix.pts[pos]);
  int id = __float_as_int(p.w);
  float d = d2_f(w.qx, w.qy, w.qz, p.x, p.y, p.z);
  unsigned m = __ballot_sync(0xFFFFFFFFu, id != 0x7FFFFFFF && cand_less(d, id, w.worst_d, w.worst_id));
  while (m) {
    int b = __ffs(m) - 1;
    m &= m - 1;
    float cd = __shfl_sync(0xFFFFFFFFu, d, b);
    int cid = __shfl_sync(0xFFFFFFFFu, id, b);
    if (cand_less(cd, cid, w.worst_d, w.worst_id)) warp_insert(w, cd, cid, chunk * 32 + b);
  }
}

// node at depth `depth`; chunk_depth = depth of the nodes that cover exactly 32 points (two levels above the leaves)
template <int LEVEL>
__device__ void knn_visit(const IndexView& ix, WarpKnn& w, int node, int depth, int chunk_depth) {
  if (depth == chunk_depth) {
    int chunk = node - (1 << chunk_depth);
    if (chunk != w.own_chunk) scan_chunk(ix, w, chunk);
    return;
  }
  if constexpr (LEVEL < 6) {
    int step = min(5, chunk_depth - depth);
    int base = node << step;
    float bd = INFINITY;
    if (w.lane < (1 << step)) bd = node_d2(ix, base + w.lane, w.qx, w.qy, w.qz);
    unsigned m = __ballot_sync(0xFFFFFFFFu, bd <= w.worst_d && bd < INFINITY);
    while (m) {
      // nearest remaining box first
      unsigned key = ((m >> w.lane) & 1u) ? __float_as_uint(bd) : 0xFFFFFFFFu;
      unsigned kmin = __reduce_min_sync(0xFFFFFFFFu, key);
      if (__uint_as_float(kmin) > w.worst_d) break;        // every remaining box is farther than the k-th neighbour
      unsigned pick = __ballot_sync(0xFFFFFFFFu, key == kmin) & m;
      int c = __ffs(pick) - 1;
      m &= ~(1u << c);
      knn_visit<LEVEL + 1>(ix, w, base + c, depth + step, chunk_depth);
    }
  }
}

__global__ void __launch_bounds__(256) k_knn_warp(IndexView ix, int k, int chunk_depth, int* __restrict__ knn_pos,
                                                  int* __restrict__ knn_out_orig) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // query = Morton position
  if (i >= ix.n) return;
  float4 q = __ldg(&ix.pts[i]);
  WarpKnn w;
  w.qx = q.x; w.qy = q.y; w.qz = q.z;
  w.ld = INFINITY; w.lid = 0x7FFFFFFF; w.lpos = -1;
  w.worst_d = INFINITY; w.worst_id = 0x7FFFFFFF;
  w.k = k; w.lane = lane;
  w.own_chunk = -1;
  scan_chunk(ix, w, i >> 5);          // the query's own chunk first: a tight bound before the tree is touched
  w.own_chunk = i >> 5;
  knn_visit<0>(ix, w, 1, 0, chunk_depth);
  if (lane < k) {
    knn_pos[(size_t)i * k + lane] = w.lpos;
    if (knn_out_orig) knn_out_orig[(size_t)__float_as_int(q.w) * k + lane] = w.lid;
  }
}

__global__ void __launch_bounds__(128) k_normals_from_knn(IndexView ix, int k, const int* __restrict__ knn_pos,
                                                          float4* __restrict__ normals_morton) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ix.n) return;
  const int* nb = knn_pos + (size_t)i * k;
  // mean and covariance in list order, float64
  double mx = 0, my = 0, mz = 0;
  for (int j = 0; j < k; ++j) {
    float4 p = __ldg(&ix.pts[__ldg(&nb[j])]);
    mx = mx + (double)p.x; my = my + (double)p.y; mz = mz + (double)p.z;
  }
  double kd = (double)k;
  mx = mx / kd; my = my / kd; mz = mz / kd;
  double cxx = 0, cxy = 0, cxz = 0, cyy = 0, cyz = 0, czz = 0, r2max = 0;
  for (int j = 0; j < k; ++j) {
    float4 p = __ldg(&ix.pts[__ldg(&nb[j])]);
    double dx = (double)p.x - mx, dy = (double)p.y - my, dz = (double)p.z - mz;
    cxx = cxx + dx * dx; cxy = cxy + dx * dy; cxz = cxz + dx * dz;
    cyy = cyy + dy * dy; cyz = cyz + dy * dz; czz = czz + dz * dz;
    double r2 = dx * dx + dy * dy;
    r2 = r2 + dz * dz;
    if (r2 > r2max) r2max = r2;
  }
  double a[3][3] = {{cxx, cxy, cxz}, {cxy, cyy, cyz}, {cxz, cyz, czz}};
  double v[3][3];
  det_jacobi<3>(a, v);
  double l0 = a[0][0], l1 = a[1][1], l2 = a[2][2];
  int smallest = 0; double sv = l0;
  if (l1 < sv) { smallest = 1; sv = l1; }
  if (l2 < sv) { smallest = 2; sv = l2; }
  double lmax = l0; if (l1 > lmax) lmax = l1; if (l2 > lmax) lmax = l2;
  double mn01 = l0 < l1 ? l0 : l1, mx01 = l0 < l1 ? l1 : l0;
  double t2 = mx01 < l2 ? mx01 : l2;
  double lmid = mn01 > t2 ? mn01 : t2;
  double nx, ny, nz;
  const double rank_tol = 3.0 * (double)FLT_EPSILON;     // A.2 rank test, see DESIGN.md "Degenerate neighbourhoods"
  if (!(lmid > rank_tol * lmax)) {
    nx = 0.0; ny = 1.0; nz = 0.0;
  } else {
    nx = smallest == 0 ? v[0][0] : (smallest == 1 ? v[0][1] : v[0][2]);
    ny = smallest == 0 ? v[1][0] : (smallest == 1 ? v[1][1] : v[1][2]);
    nz = smallest == 0 ? v[2][0] : (smallest == 1 ? v[2][1] : v[2][2]);
    double nn = sqrt((nx * nx + ny * ny) + nz * nz);
    nx = nx / nn; ny = ny / nn; nz = nz / nn;
    double ax = fabs(nx), ay = fabs(ny), az = fabs(nz);
    double lead = nx, al = ax;
    if (ay > al) { lead = ny; al = ay; }
    if (az > al) { lead = nz; al = az; }
    if (lead < 0.0) { nx = -nx; ny = -ny; nz = -nz; }
  }
  const double four_thirds_pi = 4.1887902047863905;
  double r = sqrt(r2max);
  double vol = four_thirds_pi * ((r * r) * r);
  normals_morton[i] = make_float4((float)nx, (float)ny, (float)nz, (float)(kd / vol));
}

int run_surface_normals(Handle* h, const SpatialIndex& ix, int knn, float4* normals_morton, int* knn_out_orig) {
  if (knn < 1 || knn > 32) return fail(h, AICP_B200_ERR_BAD_ARG, "SurfaceNormalDataPointsFilter: knn %d outside [1,32]", knn);
  if (knn >= ix.n) return fail(h, AICP_B200_ERR_KNN_TOO_LARGE, "SurfaceNormalDataPointsFilter: knn %d >= %d points", knn, ix.n);
  CUDA_TRY(h->knn_pos.reserve((size_t)ix.n * knn));
  int depth = 0;
  while ((1 << depth) < ix.first_leaf) ++depth;
  const int chunk_depth = depth - 2;                   // build_index guarantees at least 4 leaves
  const long long threads = (long long)ix.n * 32;
  k_knn_warp<<<(unsigned)((threads + 255) / 256), 256, 0, h->stream>>>(ix.view(), knn, chunk_depth, h->knn_pos.p, knn_out_orig);
  k_normals_from_knn<<<(ix.n + 127) / 128, 128, 0, h->stream>>>(ix.view(), knn, h->knn_pos.p, normals_morton);
  CUDA_TRY(cudaGetLastError());
  h->launches += 2;
  return AICP_B200_OK;
}

}  // namespace aicp
