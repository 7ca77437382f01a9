// normals.cu -- SurfaceNormalDataPointsFilter{knn, keepNormals, keepDensities} on the GPU.
//
// replaces [UPSTREAM] libpointmatcher DataPointsFilters/SurfaceNormal.cpp as configured by
// aicp_core/config/icp/icp_autotuned.yaml:18-23: per point, the knn nearest neighbours (self
// included, ordered by (d2, index)), C = sum (p - mean)(p - mean)^T, normal = eigenvector of the smallest
// eigenvalue, density = knn / (4/3 pi r_max^3).
//
// One thread per point, Morton order (neighbouring threads walk the same tree nodes, so the traversal is served by
// L1/L2); the candidate list lives in shared memory, strided by thread; covariance and the 3x3 Jacobi eigen-solver run
// in float64 registers in list order, which makes the result independent of the index shape.
#include "detmath.cuh"
#include "handle.cuh"

namespace aicp {

__global__ void __launch_bounds__(128) k_surface_normals(IndexView ix, int k, float4* __restrict__ normals_morton,
                                                         int* __restrict__ knn_out_orig) {
  extern __shared__ unsigned char smem_raw[];
  const int nt = blockDim.x;
  float* s_d2 = reinterpret_cast<float*>(smem_raw);
  int* s_id = reinterpret_cast<int*>(s_d2 + (size_t)k * nt);
  int* s_pos = s_id + (size_t)k * nt;
  int i = blockIdx.x * nt + threadIdx.x;
  if (i >= ix.n) return;
  KnnList L{s_d2 + threadIdx.x, s_id + threadIdx.x, s_pos + threadIdx.x, nt, k};
  float4 q = __ldg(&ix.pts[i]);
  knn_search(ix, q.x, q.y, q.z, L);

  // mean and covariance in list order, float64
  double mx = 0, my = 0, mz = 0;
  for (int j = 0; j < k; ++j) {
    float4 p = __ldg(&ix.pts[L.pos[j * nt]]);
    mx = mx + (double)p.x; my = my + (double)p.y; mz = mz + (double)p.z;
  }
  double kd = (double)k;
  mx = mx / kd; my = my / kd; mz = mz / kd;
  double cxx = 0, cxy = 0, cxz = 0, cyy = 0, cyz = 0, czz = 0, r2max = 0;
  for (int j = 0; j < k; ++j) {
    float4 p = __ldg(&ix.pts[L.pos[j * nt]]);
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
  if (knn_out_orig) {
    int self = __float_as_int(q.w);
    for (int j = 0; j < k; ++j) knn_out_orig[(size_t)self * k + j] = L.id[j * nt];
  }
}

int run_surface_normals(Handle* h, const SpatialIndex& ix, int knn, float4* normals_morton, int* knn_out_orig) {
  if (knn < 1 || knn > AICP_MAX_KNN) return fail(h, AICP_B200_ERR_BAD_ARG, "knn %d outside [1,%d]", knn, AICP_MAX_KNN);
  if (knn >= ix.n) return fail(h, AICP_B200_ERR_KNN_TOO_LARGE, "SurfaceNormalDataPointsFilter: knn %d >= %d points", knn, ix.n);
  int nt = knn <= 32 ? 128 : 64;
  size_t smem = (size_t)knn * nt * 12;
  CUDA_TRY(cudaFuncSetAttribute(k_surface_normals, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  k_surface_normals<<<(ix.n + nt - 1) / nt, nt, smem, h->stream>>>(ix.view(), knn, normals_morton, knn_out_orig);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return AICP_B200_OK;
}

}  // namespace aicp
