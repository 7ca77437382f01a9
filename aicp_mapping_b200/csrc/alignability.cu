// alignability.cu -- the FOV overlap filter and the alignability filter on the GPU.
//
// replaces overlapFilter       (aicp_core/src/utils/filteringUtils.cpp:111-193; App::computeAlignmentRisk, app.cpp:153-156)
//          alignabilityFilter  (filteringUtils.cpp:196-430 with computeNormalsCentroid :432-445, getOrientedBoundingBox :448-478,
//                               getPointsInOrientedBox :481-505, overlapBoxFilter :507-576; app.cpp:164-166)
// The arithmetic contract is the one of oracle/aicp_oracle_alignability.c (decisions listed there).
//
// overlapFilter: one order-preserving compaction per cloud -- k_fov_flags (pose transform in float64 as the reference, range
//   and field-of-view test, accepted point transformed back in float32) -> exclusive scan -> k_fov_scatter.
// alignabilityFilter: both clouds go through the pre-filter (prefilter.cu: sampled cloud, normals flipped towards the sensor,
//   plane clusters).  What the reference then does per PAIR of clusters with PCL objects (two MomentOfInertiaEstimation runs and
//   two CropBox filters per pair, 10^4 pairs) becomes
//   k_al_sums     per-cluster exact fixed-point sums of points, normals and n n^T (integer atomics: order independent)
//   k_al_means    cluster means
//   k_al_cov      per-cluster exact sums of (p - mean)(p - mean)^T
//   [host]        3x3 eigen-decompositions of ~100 clusters (Jacobi, float64) -> OBB axes
//   k_al_extents  OBB extents: min / max of the projections (ordered-int atomics)
//   [host]        OBB centre, enlargement, Euler angles, CropBox matrices (as getPointsInOrientedBox builds them)
//   k_al_pairs    ONE pass over each cloud: every clustered point is tested against all boxes of the other cloud
//                 (boxes in shared memory) -> the full cluster x cluster count matrix
//   [host]        the reference's greedy matching loop (:236-282) on that matrix, sum of n n^T over the matched clusters,
//                 3x3 eigenvalues -> alignability = 100 * lambda_min / lambda_max.
#include <float.h>
#include <math.h>
#include <string.h>

#include <thread>
#include <vector>

#include "handle.cuh"

namespace aicp {

#define AL_NSUM 18              // per cluster: 3 point sums, 3 normal sums, 6 n n^T sums, 6 covariance sums
#define AL_P20 1048576.0
#define AL_P30 1073741824.0

struct FovParams {
  double Pi[12];                // inverse pose, rows 0..2 of the 3x4 matrix, column-major as Pi[c * 3 + r]
  float R[9], t[3];             // pose as floats (row-major R)
  float range;
  double cos_thr;
  int thr_positive;
};

struct CropT { float M[9], t[3], bmin[3], bmax[3]; };

// ---- overlapFilter ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fov_flags(const float4* __restrict__ pts, int n, FovParams P, unsigned int* __restrict__ flag,
                                                   float4* __restrict__ moved) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(&pts[i]);
  const double x = p.x, y = p.y, z = p.z;
  const float lx = (float)(((P.Pi[0] * x + P.Pi[3] * y) + P.Pi[6] * z) + P.Pi[9]);
  const float ly = (float)(((P.Pi[1] * x + P.Pi[4] * y) + P.Pi[7] * z) + P.Pi[10]);
  const float lz = (float)(((P.Pi[2] * x + P.Pi[5] * y) + P.Pi[8] * z) + P.Pi[11]);
  const double dx = lx, dy = ly, dz = lz;
  const float r = (float)sqrt((dx * dx + dy * dy) + dz * dz);
  const double hh = sqrt(dx * dx + dy * dy);
  const bool in_fov = hh > 0.0 ? (dx / hh > P.cos_thr) : (P.thr_positive != 0);
  const bool keep = in_fov && r < P.range;
  flag[i] = keep ? 1u : 0u;
  if (keep) {
    float4 o;
    o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.R[0], lx), __fmul_rn(P.R[1], ly)), __fmul_rn(P.R[2], lz)), P.t[0]);
    o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.R[3], lx), __fmul_rn(P.R[4], ly)), __fmul_rn(P.R[5], lz)), P.t[1]);
    o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(P.R[6], lx), __fmul_rn(P.R[7], ly)), __fmul_rn(P.R[8], lz)), P.t[2]);
    o.w = 1.0f;
    moved[i] = o;
  }
}

__global__ void __launch_bounds__(256) k_fov_scatter(const unsigned int* __restrict__ flag, const unsigned int* __restrict__ slot,
                                                     const float4* __restrict__ moved, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && __ldg(&flag[i])) out[__ldg(&slot[i])] = moved[i];
}

static void fov_params(const double* pose_other, float range, double cos_thr, int thr_positive, FovParams* P) {
  // Isometry3d::inverse(): [R^T | -(R^T t)], float64
  double Q[16];
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Q[c * 4 + r] = pose_other[r * 4 + c];
  for (int r = 0; r < 3; ++r) Q[12 + r] = -((Q[0 * 4 + r] * pose_other[12] + Q[1 * 4 + r] * pose_other[13]) + Q[2 * 4 + r] * pose_other[14]);
  for (int c = 0; c < 4; ++c) for (int r = 0; r < 3; ++r) P->Pi[c * 3 + r] = Q[c * 4 + r];
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) P->R[3 * r + c] = (float)pose_other[c * 4 + r]; P->t[r] = (float)pose_other[12 + r]; }
  P->range = range; P->cos_thr = cos_thr; P->thr_positive = thr_positive;
}

// one direction; result in `out` (capacity n), count in *n_out
static int fov_one(Handle* h, const float4* pts, int64_t n, const double* pose_other, float range, double cos_thr, int thr_positive,
                   DevBuf<float4>& out, int64_t* n_out) {
  *n_out = 0;
  if (n == 0) return AICP_B200_OK;
  if (n > (1ll << 28)) return fail(h, AICP_B200_ERR_BAD_ARG, "fov_overlap: cloud larger than 2^28 points");
  cudaStream_t s = h->stream;
  CUDA_TRY(h->pf_flag.reserve((size_t)n)); CUDA_TRY(h->pf_slot.reserve((size_t)n)); CUDA_TRY(h->pf_tiles.reserve((size_t)n / 1024 + 2));
  CUDA_TRY(h->al_moved.reserve((size_t)n)); CUDA_TRY(out.reserve((size_t)n)); CUDA_TRY(h->al_counts.reserve(64));
  FovParams P;
  fov_params(pose_other, range, cos_thr, thr_positive, &P);
  const int blocks = (int)((n + 255) / 256);
  k_fov_flags<<<blocks, 256, 0, s>>>(pts, (int)n, P, h->pf_flag.p, h->al_moved.p);
  h->launches += 1;
  int rc = exclusive_scan_u32(h, h->pf_flag.p, h->pf_slot.p, (int)n, h->pf_tiles.p, h->al_counts.p);
  if (rc) return rc;
  k_fov_scatter<<<blocks, 256, 0, s>>>(h->pf_flag.p, h->pf_slot.p, h->al_moved.p, (int)n, out.p);
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  unsigned int total = 0;
  CUDA_TRY(cudaMemcpyAsync(&total, h->al_counts.p, sizeof(total), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  *n_out = (int64_t)total;
  return AICP_B200_OK;
}

int run_fov_overlap(Handle* h, const float4* A, int64_t nA, const float4* B, int64_t nB, const double* poseA, const double* poseB,
                    float range, float angular_view, float* overlap_pct, int64_t* counts) {
  const float thresh = (float)(180.0 - ((360.0 - (double)angular_view) / 2));
  const double cos_thr = cos((double)thresh * M_PI / 180.0);       // host libm, as the oracle
  int rc;
  if ((rc = fov_one(h, A, nA, poseB, range, cos_thr, thresh > 0.f, h->al_fov[0], &h->al_fov_n[0]))) return rc;
  if ((rc = fov_one(h, B, nB, poseA, range, cos_thr, thresh > 0.f, h->al_fov[1], &h->al_fov_n[1]))) return rc;
  const float pa = (float)h->al_fov_n[0] / (float)nA, pb = (float)h->al_fov_n[1] / (float)nB;
  const float overlap = pa * pb;
  *overlap_pct = (float)((double)overlap * 100.0);
  if (counts) { counts[0] = h->al_fov_n[0]; counts[1] = h->al_fov_n[1]; }
  return AICP_B200_OK;
}

// ---- alignabilityFilter ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_al_sums(const float4* __restrict__ pts, const float4* __restrict__ nrm, const int* __restrict__ labels,
                                                 int m, unsigned long long* sums, unsigned int* cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int c = __ldg(&labels[i]);
  if (c < 0) return;
  const float4 p = __ldg(&pts[i]), q = __ldg(&nrm[i]);
  unsigned long long* s = sums + (size_t)c * AL_NSUM;
  atomicAdd(&cnt[c], 1u);
  atomicAdd(&s[0], (unsigned long long)__double2ll_rn((double)p.x * AL_P20));
  atomicAdd(&s[1], (unsigned long long)__double2ll_rn((double)p.y * AL_P20));
  atomicAdd(&s[2], (unsigned long long)__double2ll_rn((double)p.z * AL_P20));
  atomicAdd(&s[3], (unsigned long long)__double2ll_rn((double)q.x * AL_P30));
  atomicAdd(&s[4], (unsigned long long)__double2ll_rn((double)q.y * AL_P30));
  atomicAdd(&s[5], (unsigned long long)__double2ll_rn((double)q.z * AL_P30));
  const double n3[3] = {(double)q.x, (double)q.y, (double)q.z};
  int t = 6;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b) atomicAdd(&s[t++], (unsigned long long)__double2ll_rn((n3[a] * n3[b]) * AL_P30));
}

__global__ void k_al_means(const unsigned long long* __restrict__ sums, const unsigned int* __restrict__ cnt, int k, float4* __restrict__ mean) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  const double n = (double)cnt[c];
  const long long* s = reinterpret_cast<const long long*>(sums + (size_t)c * AL_NSUM);
  mean[c] = make_float4((float)(((double)s[0] / n) * (1.0 / AL_P20)), (float)(((double)s[1] / n) * (1.0 / AL_P20)),
                        (float)(((double)s[2] / n) * (1.0 / AL_P20)), 0.f);
}

__global__ void __launch_bounds__(256) k_al_cov(const float4* __restrict__ pts, const int* __restrict__ labels, int m, const float4* __restrict__ mean,
                                                unsigned long long* sums) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int c = __ldg(&labels[i]);
  if (c < 0) return;
  const float4 p = __ldg(&pts[i]), mu = __ldg(&mean[c]);
  const double d[3] = {(double)__fsub_rn(p.x, mu.x), (double)__fsub_rn(p.y, mu.y), (double)__fsub_rn(p.z, mu.z)};
  unsigned long long* s = sums + (size_t)c * AL_NSUM;
  int t = 12;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = a; b < 3; ++b) atomicAdd(&s[t++], (unsigned long long)__double2ll_rn((d[a] * d[b]) * AL_P30));
}

// axes: k x 9 floats, row-major R with the major / middle / minor axis as columns; ext: k x 6 ordered ints (min xyz, max xyz)
__global__ void __launch_bounds__(256) k_al_extents(const float4* __restrict__ pts, const int* __restrict__ labels, int m, const float4* __restrict__ mean,
                                                    const float* __restrict__ axes, int* ext) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int c = __ldg(&labels[i]);
  if (c < 0) return;
  const float4 p = __ldg(&pts[i]), mu = __ldg(&mean[c]);
  const float d0 = __fsub_rn(p.x, mu.x), d1 = __fsub_rn(p.y, mu.y), d2 = __fsub_rn(p.z, mu.z);
  const float* R = axes + (size_t)c * 9;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float v = __fadd_rn(__fadd_rn(__fmul_rn(d0, __ldg(&R[k])), __fmul_rn(d1, __ldg(&R[3 + k]))), __fmul_rn(d2, __ldg(&R[6 + k])));
    const int o = float_to_ordered(v);
    atomicMin(&ext[c * 6 + k], o);
    atomicMax(&ext[c * 6 + 3 + k], o);
  }
}

__global__ void k_al_ext_init(int* ext, int k) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  for (int d = 0; d < 3; ++d) { ext[c * 6 + d] = 0x7FFFFFFF; ext[c * 6 + 3 + d] = (int)0x80000000; }
}

// counts[i * kY + j] += 1 for every point of Y's cluster j inside box i of X
#define AL_BOX_TILE 64
__global__ void __launch_bounds__(256) k_al_pairs(const CropT* __restrict__ boxes, int kX, const float4* __restrict__ ptsY, const int* __restrict__ labY,
                                                  int mY, int kY, unsigned int* counts) {
  __shared__ CropT s_box[AL_BOX_TILE];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = p < mY ? __ldg(&labY[p]) : -1;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j >= 0) q = __ldg(&ptsY[p]);
  for (int base = 0; base < kX; base += AL_BOX_TILE) {
    const int nb = min(AL_BOX_TILE, kX - base);
    __syncthreads();
    for (int w = threadIdx.x; w < nb * (int)(sizeof(CropT) / 4); w += blockDim.x)
      reinterpret_cast<float*>(s_box)[w] = reinterpret_cast<const float*>(boxes + base)[w];
    __syncthreads();
    if (j < 0) continue;
    for (int i = 0; i < nb; ++i) {
      const CropT& b = s_box[i];
      const float dx = __fsub_rn(q.x, b.t[0]), dy = __fsub_rn(q.y, b.t[1]), dz = __fsub_rn(q.z, b.t[2]);
      const float lx = __fadd_rn(__fadd_rn(__fmul_rn(b.M[0], dx), __fmul_rn(b.M[1], dy)), __fmul_rn(b.M[2], dz));
      const float ly = __fadd_rn(__fadd_rn(__fmul_rn(b.M[3], dx), __fmul_rn(b.M[4], dy)), __fmul_rn(b.M[5], dz));
      const float lz = __fadd_rn(__fadd_rn(__fmul_rn(b.M[6], dx), __fmul_rn(b.M[7], dy)), __fmul_rn(b.M[8], dz));
      if (!(lx < b.bmin[0] || ly < b.bmin[1] || lz < b.bmin[2] || lx > b.bmax[0] || ly > b.bmax[1] || lz > b.bmax[2]))
        atomicAdd(&counts[(size_t)(base + i) * kY + j], 1u);
    }
  }
}

// ---- host-side small math: the same operation sequences as oracle/aicp_oracle_alignability.c ---------------------------
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

// Eigen 3.3 Matrix3f::eulerAngles(0, 1, 2) [UPSTREAM, recalled]; R row-major (the adapter's crop path calls Eigen itself)
static void euler_angles_012(const float* R, float* rpy) {
  float r0 = atan2f(R[3 * 1 + 2], R[3 * 2 + 2]);
  const float c2 = sqrtf(R[0] * R[0] + R[1] * R[1]);
  float r1;
  if (r0 > 0.f) { r0 = r0 - (float)M_PI; r1 = atan2f(-R[2], -c2); }
  else r1 = atan2f(-R[2], c2);
  const float s1 = sinf(r0), c1 = cosf(r0);
  const float r2 = atan2f(s1 * R[3 * 2 + 0] - c1 * R[3 * 1 + 0], c1 * R[3 * 1 + 1] - s1 * R[3 * 2 + 1]);
  rpy[0] = -r0; rpy[1] = -r1; rpy[2] = -r2;
}

// pcl::getTransformation(0,0,0,roll,pitch,yaw), row-major (the same expression as crop.cu / the oracle)
static void rpy_to_matrix(const float* rpy, float* R) {
  const float A = cosf(rpy[2]), B = sinf(rpy[2]), C = cosf(rpy[1]), D = sinf(rpy[1]), E = cosf(rpy[0]), F = sinf(rpy[0]);
  const float DE = D * E, DF = D * F;
  R[0] = A * C;  R[1] = A * DF - B * E;  R[2] = B * F + A * DE;
  R[3] = B * C;  R[4] = A * E + B * DF;  R[5] = B * DE - A * F;
  R[6] = -D;     R[7] = C * F;           R[8] = C * E;
}

struct ClusterHost {
  long long n;
  float ncen[3], mean[3], axis[9];
  long long snn[6];
};

// per-cloud device state of one alignability call
struct AlSide {
  int m = 0, k = 0;
  std::vector<ClusterHost> cl;
  std::vector<CropT> boxes;
};

static int al_side(Handle* h, int s, const float4* pts, int64_t n, const aicp_b200_prefilter_config* cfg, const double* pose, AlSide* out) {
  cudaStream_t st = h->stream;
  const float vp[3] = {(float)pose[12], (float)pose[13], (float)pose[14]};
  int rc = run_prefilter(h, pts, n, cfg, vp, nullptr);
  if (rc) return rc;
  const int m = (int)h->pf_n_sampled, k = h->pf_has_segments ? (int)h->pf_n_clusters : 0;
  out->m = m; out->k = k;
  out->cl.assign((size_t)k, ClusterHost());
  out->boxes.assign((size_t)k, CropT());
  if (k == 0) return AICP_B200_OK;
  // keep this cloud's by-products: the second pre-filter run reuses the pf_* buffers
  CUDA_TRY(h->al_pts[s].reserve((size_t)m)); CUDA_TRY(h->al_lab[s].reserve((size_t)m));
  CUDA_TRY(cudaMemcpyAsync(h->al_pts[s].p, h->pf_sampled.p, sizeof(float4) * (size_t)m, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(h->al_lab[s].p, h->pf_labels_out.p, sizeof(int) * (size_t)m, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(h->al_sums.reserve((size_t)k * AL_NSUM)); CUDA_TRY(h->al_cnt.reserve((size_t)k)); CUDA_TRY(h->al_mean.reserve((size_t)k));
  CUDA_TRY(h->al_axes.reserve((size_t)k * 9)); CUDA_TRY(h->al_ext.reserve((size_t)k * 6)); CUDA_TRY(h->al_boxes[s].reserve((size_t)k * sizeof(CropT) / 4));
  CUDA_TRY(cudaMemsetAsync(h->al_sums.p, 0, sizeof(unsigned long long) * (size_t)k * AL_NSUM, st));
  CUDA_TRY(cudaMemsetAsync(h->al_cnt.p, 0, sizeof(unsigned int) * (size_t)k, st));
  const int blocks = (m + 255) / 256, kb = (k + 127) / 128;
  k_al_sums<<<blocks, 256, 0, st>>>(h->al_pts[s].p, h->pf_normals_orig.p, h->al_lab[s].p, m, h->al_sums.p, h->al_cnt.p);
  k_al_means<<<kb, 128, 0, st>>>(h->al_sums.p, h->al_cnt.p, k, h->al_mean.p);
  k_al_cov<<<blocks, 256, 0, st>>>(h->al_pts[s].p, h->al_lab[s].p, m, h->al_mean.p, h->al_sums.p);
  k_al_ext_init<<<kb, 128, 0, st>>>(h->al_ext.p, k);
  h->launches += 4;
  CUDA_TRY(cudaGetLastError());
  std::vector<long long> sums((size_t)k * AL_NSUM);
  std::vector<unsigned int> cnt((size_t)k);
  std::vector<float4> mean((size_t)k);
  CUDA_TRY(cudaMemcpyAsync(sums.data(), h->al_sums.p, sizeof(long long) * sums.size(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(cnt.data(), h->al_cnt.p, sizeof(unsigned int) * cnt.size(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(mean.data(), h->al_mean.p, sizeof(float4) * mean.size(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  std::vector<float> axes((size_t)k * 9);
  for (int c = 0; c < k; ++c) {
    ClusterHost& C = out->cl[(size_t)c];
    const long long* s18 = &sums[(size_t)c * AL_NSUM];
    C.n = (long long)cnt[(size_t)c];
    const double dn = (double)C.n;
    C.mean[0] = mean[(size_t)c].x; C.mean[1] = mean[(size_t)c].y; C.mean[2] = mean[(size_t)c].z;
    for (int d = 0; d < 3; ++d) C.ncen[d] = (float)(((double)s18[3 + d] / dn) * (1.0 / AL_P30));
    for (int t = 0; t < 6; ++t) C.snn[t] = s18[6 + t];
    const double sc = 1.0 / AL_P30;
    double a[3][3], v[3][3];
    a[0][0] = (double)s18[12] * sc / dn; a[0][1] = a[1][0] = (double)s18[13] * sc / dn; a[0][2] = a[2][0] = (double)s18[14] * sc / dn;
    a[1][1] = (double)s18[15] * sc / dn; a[1][2] = a[2][1] = (double)s18[16] * sc / dn; a[2][2] = (double)s18[17] * sc / dn;
    jacobi3d(a, v);
    // MomentOfInertiaEstimation::computeEigenVectors: three compare-swaps on the indices
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
    for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) C.axis[3 * r + q] = ax[q][r];
    memcpy(&axes[(size_t)c * 9], C.axis, sizeof(float) * 9);
  }
  CUDA_TRY(cudaMemcpyAsync(h->al_axes.p, axes.data(), sizeof(float) * axes.size(), cudaMemcpyHostToDevice, st));
  k_al_extents<<<blocks, 256, 0, st>>>(h->al_pts[s].p, h->al_lab[s].p, m, h->al_mean.p, h->al_axes.p, h->al_ext.p);
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  std::vector<int> ext((size_t)k * 6);
  CUDA_TRY(cudaMemcpyAsync(ext.data(), h->al_ext.p, sizeof(int) * ext.size(), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  for (int c = 0; c < k; ++c) {
    const ClusterHost& C = out->cl[(size_t)c];
    float bmin[3], bmax[3], shift[3], pos[3];
    for (int q = 0; q < 3; ++q) { bmin[q] = ordered_to_float(ext[(size_t)c * 6 + q]); bmax[q] = ordered_to_float(ext[(size_t)c * 6 + 3 + q]); }
    for (int q = 0; q < 3; ++q) { shift[q] = (bmax[q] + bmin[q]) / 2.0f; bmin[q] -= shift[q]; bmax[q] -= shift[q]; }      // computeOBB
    for (int r = 0; r < 3; ++r) pos[r] = C.mean[r] + ((C.axis[3 * r] * shift[0] + C.axis[3 * r + 1] * shift[1]) + C.axis[3 * r + 2] * shift[2]);
    CropT& b = out->boxes[(size_t)c];
    float rpy[3], R[9];
    euler_angles_012(C.axis, rpy);                 // rotational_matrix_OBB.eulerAngles(0, 1, 2), filteringUtils.cpp:492
    rpy_to_matrix(rpy, R);                         // CropBox::setRotation -> pcl::getTransformation
    for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) b.M[3 * r + q] = R[3 * q + r];
    for (int q = 0; q < 3; ++q) { b.t[q] = pos[q]; b.bmin[q] = bmin[q]; b.bmax[q] = bmax[q]; }
    b.bmin[2] = (float)(3.0 * (double)bmin[2]);    // "direction perpendicular to plane", :522-523
    b.bmax[2] = (float)(3.0 * (double)bmax[2]);
  }
  CUDA_TRY(cudaMemcpyAsync(h->al_boxes[s].p, out->boxes.data(), sizeof(CropT) * (size_t)k, cudaMemcpyHostToDevice, st));
  return AICP_B200_OK;
}

int run_alignability(Handle* h, const float4* A, int64_t nA, const float4* B, int64_t nB, const double* poseA, const double* poseB,
                     const aicp_b200_prefilter_config* cfg, float* out_alignability, int32_t* matching, int64_t* info) {
  *out_alignability = 0.f;
  if (info) info[0] = info[1] = info[2] = 0;
  AlSide side[2];
  int rc;
  // The two clouds are independent until the box counts, and each side is a chain of small latency-bound kernels with
  // host round trips in between: outside a batch (where the other streams already fill the GPU) the second cloud runs on a
  // child handle -- its own stream and buffers -- driven by a second host thread.
  Handle* hb = h;                    // the handle that holds side 1's device buffers
  if (!h->batch_worker) {
    if (!h->al_child) {
      aicp_b200_handle* c = nullptr;
      if (aicp_b200_create(nullptr, h->device, &c) == AICP_B200_OK) h->al_child = reinterpret_cast<Handle*>(c);
    }
    if (h->al_child) hb = h->al_child;
  }
  if (hb != h) {
    CUDA_TRY(cudaStreamSynchronize(h->stream));      // B may still be on its way in on this handle's stream
    int rc1 = AICP_B200_OK;
    std::thread tb([&]() {
      cudaSetDevice(hb->device);
      rc1 = al_side(hb, 1, B, nB, cfg, poseB, &side[1]);
      if (!rc1 && cudaStreamSynchronize(hb->stream) != cudaSuccess) rc1 = fail(hb, AICP_B200_ERR_CUDA, "alignability: stream synchronisation failed");
    });
    rc = al_side(h, 0, A, nA, cfg, poseA, &side[0]);
    tb.join();
    if (rc) return rc;
    if (rc1) return fail(h, rc1, "%s", hb->last_error.c_str());
  } else {
    if ((rc = al_side(h, 0, A, nA, cfg, poseA, &side[0]))) return rc;
    if ((rc = al_side(h, 1, B, nB, cfg, poseB, &side[1]))) return rc;
  }
  const int kA = side[0].k, kB = side[1].k;
  if (info) { info[0] = kA; info[1] = kB; }
  if (kA == 0 || kB == 0) {
    for (int j = 0; matching && j < kB; ++j) matching[j] = -1;
    return AICP_B200_OK;
  }
  cudaStream_t st = h->stream;
  const size_t cells = (size_t)kA * kB;
  CUDA_TRY(h->al_counts.reserve(2 * cells + 64));
  unsigned int* b_in_a = h->al_counts.p;            // [i * kB + j]
  unsigned int* a_in_b = h->al_counts.p + cells;    // [j * kA + i]
  CUDA_TRY(cudaMemsetAsync(h->al_counts.p, 0, sizeof(unsigned int) * 2 * cells, st));
  k_al_pairs<<<(side[1].m + 255) / 256, 256, 0, st>>>(reinterpret_cast<const CropT*>(h->al_boxes[0].p), kA, hb->al_pts[1].p, hb->al_lab[1].p,
                                                      side[1].m, kB, b_in_a);
  k_al_pairs<<<(side[0].m + 255) / 256, 256, 0, st>>>(reinterpret_cast<const CropT*>(hb->al_boxes[1].p), kB, h->al_pts[0].p, h->al_lab[0].p,
                                                      side[0].m, kA, a_in_b);
  h->launches += 2;
  CUDA_TRY(cudaGetLastError());
  std::vector<unsigned int> counts(2 * cells);
  CUDA_TRY(cudaMemcpyAsync(counts.data(), h->al_counts.p, sizeof(unsigned int) * 2 * cells, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  // filteringUtils.cpp:229-282: greedy matching on the overlap matrix
  std::vector<int32_t> mi((size_t)kB, -1);
  std::vector<float> mo((size_t)kB, -1.f);
  for (int i = 0; i < kA; ++i) {
    float max_overlap = 0.f;
    int best = -1;
    const float* ca = side[0].cl[(size_t)i].ncen;
    for (int j = 0; j < kB; ++j) {
      const float* cb = side[1].cl[(size_t)j].ncen;
      const float dot = (ca[0] * cb[0] + ca[1] * cb[1]) + ca[2] * cb[2];
      const float na = sqrtf((ca[0] * ca[0] + ca[1] * ca[1]) + ca[2] * ca[2]), nb = sqrtf((cb[0] * cb[0] + cb[1] * cb[1]) + cb[2] * cb[2]);
      const float dist = (float)((double)acosf(dot / (na * nb)) * 180.0 / M_PI);
      const float perc_a = (float)counts[cells + (size_t)j * kA + i] / (float)side[0].cl[(size_t)i].n;
      const float perc_b = (float)counts[(size_t)i * kB + j] / (float)side[1].cl[(size_t)j].n;
      const float ov = perc_a * perc_b;
      const float current = (float)((double)ov * 100.0);
      if (current > max_overlap && dist < 20) { best = j; max_overlap = current; }
    }
    if (max_overlap > 0.f) {
      if (mi[(size_t)best] == -1 || max_overlap > mo[(size_t)best]) { mi[(size_t)best] = i; mo[(size_t)best] = max_overlap; }
    }
  }
  long long S[6] = {0, 0, 0, 0, 0, 0};
  int matched = 0;
  for (int j = 0; j < kB; ++j) {
    if (matching) matching[j] = mi[(size_t)j];
    if (mi[(size_t)j] < 0) continue;
    ++matched;
    for (int t = 0; t < 6; ++t) S[t] += side[0].cl[(size_t)mi[(size_t)j]].snn[t];
  }
  if (info) info[2] = matched;
  if (matched > 0) {
    // pcl::PCA on the matched normals and their mirror images (:284-372): mean 0, covariance proportional to sum n n^T
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
  return AICP_B200_OK;
}

}  // namespace aicp
