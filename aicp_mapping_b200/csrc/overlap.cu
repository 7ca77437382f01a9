// overlap.cu -- the octree overlap parameter as a GPU occupancy intersection.
//
// replaces aicp::OctreesOverlap::computeOverlap (aicp_core/src/overlap/octrees_overlap.cpp:29-72): createTree (:153-218)
// ray-casts every point of a cloud from the sensor origin into an octomap::ColorOcTree (insertPointCloud, :184), then
// force-marks every leaf occupied (:205-215); getOverlappingNodes (:113-151) counts depth-16 leaves of both trees and of
// their intersection.  The outcome depends only on the two SETS of voxel keys touched by the rays, so the
// trees are replaced by two dense occupancy bitmaps over the joint key bounding box:
//   k_key_bounds   key bounding box of (origins, end points) of both clouds         (every DDA step stays inside it)
//   k_ray_mark     one thread per ray: octomap's computeRayKeys (Amanatides-Woo 3-D DDA, same float/double mix, same
//                  tie order and early-exit rule), setting one bit per visited voxel (test-before-atomicOr)
//   k_popcount     |A|, |B|, |A & B| by popcount over 32-bit words
// overlap = 100 * min(|A^B|/|A|, |A^B|/|B|) in float32 as in octrees_overlap.cpp:47-53.
#include <float.h>

#include "handle.cuh"

namespace aicp {

#define OCT_MAX_VAL 32768
#define RAY_BATCH 4

struct KeyBounds { int lo[3]; int hi[3]; };

__device__ __forceinline__ bool coord_to_key(float coord, double res_factor, int* key) {
  double v = floor(res_factor * (double)coord);
  if (!(v > -1.0e9 && v < 1.0e9)) return false;
  int scaled = (int)v + OCT_MAX_VAL;
  if (scaled >= 0 && scaled < 2 * OCT_MAX_VAL) { *key = scaled; return true; }
  return false;
}
__device__ __forceinline__ double key_to_coord(int key, double res) { return ((double)(key - OCT_MAX_VAL) + 0.5) * res; }

__global__ void k_bounds_init(KeyBounds* kb, unsigned long long* counts) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int d = 0; d < 3; ++d) { kb->lo[d] = 0x7FFFFFFF; kb->hi[d] = (int)0x80000000; }
    counts[0] = counts[1] = counts[2] = counts[3] = 0;
  }
}

__global__ void __launch_bounds__(256) k_key_bounds(const float4* __restrict__ pts, int n, float ox, float oy, float oz,
                                                    double res_factor, KeyBounds* kb) {
  int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n + 1; i += gridDim.x * blockDim.x) {
    float c[3];
    if (i < n) { float4 p = __ldg(&pts[i]); c[0] = p.x; c[1] = p.y; c[2] = p.z; }
    else { c[0] = ox; c[1] = oy; c[2] = oz; }
    int k[3];
    if (coord_to_key(c[0], res_factor, &k[0]) && coord_to_key(c[1], res_factor, &k[1]) && coord_to_key(c[2], res_factor, &k[2])) {
#pragma unroll
      for (int d = 0; d < 3; ++d) { lo[d] = min(lo[d], k[d]); hi[d] = max(hi[d], k[d]); }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xFFFFFFFFu, lo[d], off));
      hi[d] = max(hi[d], __shfl_xor_sync(0xFFFFFFFFu, hi[d], off));
    }
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int d = 0; d < 3; ++d) { atomicMin(&kb->lo[d], lo[d]); atomicMax(&kb->hi[d], hi[d]); }
}

struct Grid { int lo[3]; int dim[3]; };

// The DDA may step past the end voxel before its length guard fires (octomap's "accumulating discretization errors"
// case), so the bitmap box carries a 2-voxel margin; a key outside even that raises a flag instead of being dropped.
__device__ __forceinline__ void mark(unsigned int* bits, const Grid& g, int kx, int ky, int kz, unsigned long long* oob) {
  if ((unsigned)(kx - g.lo[0]) >= (unsigned)g.dim[0] || (unsigned)(ky - g.lo[1]) >= (unsigned)g.dim[1] ||
      (unsigned)(kz - g.lo[2]) >= (unsigned)g.dim[2]) { atomicAdd(oob, 1ull); return; }
  unsigned long long idx = ((unsigned long long)(kx - g.lo[0]) * (unsigned)g.dim[1] + (unsigned)(ky - g.lo[1])) * (unsigned)g.dim[2] +
                           (unsigned)(kz - g.lo[2]);
  unsigned int* w = bits + (idx >> 5);
  unsigned int m = 1u << (idx & 31);
  if (!(__ldcg(w) & m)) atomicOr(w, m);
}

// OccupancyOcTreeBase::computeUpdate for one point: computeRayKeys(origin, p) -> free cells, key(p) -> occupied cell
__global__ void __launch_bounds__(128) k_ray_mark(const float4* __restrict__ pts, int n, float ox, float oy, float oz, double res,
                                                  double res_factor, Grid g, unsigned int* bits, unsigned long long* oob) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(&pts[i]);
  if (!isfinite(p.x) || !isfinite(p.y) || !isfinite(p.z)) return;
  const float origin[3] = {ox, oy, oz};
  const float end[3] = {p.x, p.y, p.z};
  int ko[3], ke[3];
  bool ok_o = coord_to_key(origin[0], res_factor, &ko[0]) && coord_to_key(origin[1], res_factor, &ko[1]) &&
              coord_to_key(origin[2], res_factor, &ko[2]);
  bool ok_e = coord_to_key(end[0], res_factor, &ke[0]) && coord_to_key(end[1], res_factor, &ke[1]) &&
              coord_to_key(end[2], res_factor, &ke[2]);
  if (ok_e) mark(bits, g, ke[0], ke[1], ke[2], oob);
  if (!ok_o || !ok_e) return;
  if (ko[0] == ke[0] && ko[1] == ke[1] && ko[2] == ke[2]) return;
  mark(bits, g, ko[0], ko[1], ko[2], oob);
  float dir[3] = {__fsub_rn(end[0], origin[0]), __fsub_rn(end[1], origin[1]), __fsub_rn(end[2], origin[2])};
  float nsq = __fadd_rn(__fmul_rn(dir[0], dir[0]), __fmul_rn(dir[1], dir[1]));
  nsq = __fadd_rn(nsq, __fmul_rn(dir[2], dir[2]));
  float length = (float)sqrt((double)nsq);
  dir[0] = __fdiv_rn(dir[0], length); dir[1] = __fdiv_rn(dir[1], length); dir[2] = __fdiv_rn(dir[2], length);
  int step[3], cur[3] = {ko[0], ko[1], ko[2]};
  double tMax[3], tDelta[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    step[a] = dir[a] > 0.0f ? 1 : (dir[a] < 0.0f ? -1 : 0);
    if (step[a] != 0) {
      double border = key_to_coord(cur[a], res);
      border += (double)(float)((double)step[a] * res * 0.5);
      tMax[a] = (border - (double)origin[a]) / (double)dir[a];
      tDelta[a] = res / fabs((double)dir[a]);
    } else { tMax[a] = DBL_MAX; tDelta[a] = DBL_MAX; }
  }
  // The DDA itself is a short dependent chain of register arithmetic, but every visited voxel costs an L2 round trip for
  // the test-before-atomicOr.  Steps are therefore taken in batches of RAY_BATCH: advance (sequentially, exactly as
  // octomap does), issue all the bitmap loads of the batch, then test them -- RAY_BATCH loads in flight per thread
  // instead of one.
  bool done = false;
  for (int guard = 0; guard < 400000 && !done; guard += RAY_BATCH) {
    unsigned int* wp[RAY_BATCH];
    unsigned int wm[RAY_BATCH], wv[RAY_BATCH];
#pragma unroll
    for (int b = 0; b < RAY_BATCH; ++b) {
      wp[b] = nullptr;
      if (done) continue;
      int dim;
      if (tMax[0] < tMax[1]) dim = (tMax[0] < tMax[2]) ? 0 : 2;
      else dim = (tMax[1] < tMax[2]) ? 1 : 2;
      // static indexing keeps tMax / cur in registers
      if (dim == 0) { cur[0] += step[0]; tMax[0] += tDelta[0]; }
      else if (dim == 1) { cur[1] += step[1]; tMax[1] += tDelta[1]; }
      else { cur[2] += step[2]; tMax[2] += tDelta[2]; }
      if (cur[0] == ke[0] && cur[1] == ke[1] && cur[2] == ke[2]) { done = true; continue; }
      double dmin = tMax[0] < tMax[1] ? tMax[0] : tMax[1];
      if (tMax[2] < dmin) dmin = tMax[2];
      if (dmin > (double)length) { done = true; continue; }
      if ((unsigned)cur[0] >= 65536u || (unsigned)cur[1] >= 65536u || (unsigned)cur[2] >= 65536u) { done = true; continue; }
      if ((unsigned)(cur[0] - g.lo[0]) >= (unsigned)g.dim[0] || (unsigned)(cur[1] - g.lo[1]) >= (unsigned)g.dim[1] ||
          (unsigned)(cur[2] - g.lo[2]) >= (unsigned)g.dim[2]) { atomicAdd(oob, 1ull); continue; }
      unsigned long long idx = ((unsigned long long)(cur[0] - g.lo[0]) * (unsigned)g.dim[1] + (unsigned)(cur[1] - g.lo[1])) * (unsigned)g.dim[2] +
                               (unsigned)(cur[2] - g.lo[2]);
      wp[b] = bits + (idx >> 5);
      wm[b] = 1u << (idx & 31);
    }
#pragma unroll
    for (int b = 0; b < RAY_BATCH; ++b) wv[b] = wp[b] ? __ldcg(wp[b]) : 0xFFFFFFFFu;
#pragma unroll
    for (int b = 0; b < RAY_BATCH; ++b) if (wp[b] && !(wv[b] & wm[b])) atomicOr(wp[b], wm[b]);
  }
}

__global__ void __launch_bounds__(256) k_popcount(const unsigned int* __restrict__ a, const unsigned int* __restrict__ b,
                                                  unsigned long long n_words, unsigned long long* counts) {
  unsigned long long ca = 0, cb = 0, ci = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n_words;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned int wa = __ldg(&a[i]), wb = __ldg(&b[i]);
    ca += __popc(wa); cb += __popc(wb); ci += __popc(wa & wb);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    ca += __shfl_xor_sync(0xFFFFFFFFu, ca, off);
    cb += __shfl_xor_sync(0xFFFFFFFFu, cb, off);
    ci += __shfl_xor_sync(0xFFFFFFFFu, ci, off);
  }
  if ((threadIdx.x & 31) == 0) {
    if (ci) atomicAdd(&counts[0], ci);
    if (ca) atomicAdd(&counts[1], ca);
    if (cb) atomicAdd(&counts[2], cb);
  }
}

int run_overlap(Handle* h, const float4* ref, int64_t n_ref, const double* ref_origin, const float4* read, int64_t n_read,
                const double* read_origin, double resolution, float* overlap_pct, int64_t* counts) {
  cudaStream_t s = h->stream;
  if (!(resolution > 0.0)) return fail(h, AICP_B200_ERR_BAD_ARG, "octomapResolution must be positive");
  const double res_factor = 1.0 / resolution;
  // octomap::pose6d(float x, float y, float z, ...): the origin is narrowed to float (octrees_overlap.cpp:229-230)
  const float ro[3] = {(float)ref_origin[0], (float)ref_origin[1], (float)ref_origin[2]};
  const float so[3] = {(float)read_origin[0], (float)read_origin[1], (float)read_origin[2]};
  CUDA_TRY(h->ovl_counts.reserve(16));
  KeyBounds* kb = reinterpret_cast<KeyBounds*>(h->ovl_counts.p + 8);
  unsigned long long* dcounts = h->ovl_counts.p;
  k_bounds_init<<<1, 32, 0, s>>>(kb, dcounts);
  auto nblk = [](int64_t n) { int64_t b = (n + 1 + 255) / 256; return (unsigned)(b < 148 * 4 ? b : 148 * 4); };
  k_key_bounds<<<nblk(n_ref), 256, 0, s>>>(ref, (int)n_ref, ro[0], ro[1], ro[2], res_factor, kb);
  k_key_bounds<<<nblk(n_read), 256, 0, s>>>(read, (int)n_read, so[0], so[1], so[2], res_factor, kb);
  KeyBounds hb;
  CUDA_TRY(cudaMemcpyAsync(&hb, kb, sizeof(hb), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  h->launches += 3;
  unsigned long long hc[4] = {0, 0, 0, 0};
  if (hb.lo[0] <= hb.hi[0]) {
    Grid g;
    unsigned long long n_bits = 1;
    for (int d = 0; d < 3; ++d) { g.lo[d] = hb.lo[d] - 2; g.dim[d] = hb.hi[d] - hb.lo[d] + 5; n_bits *= (unsigned long long)g.dim[d]; }
    if (n_bits > (1ull << 35)) return fail(h, AICP_B200_ERR_EXTENT, "overlap key box %d x %d x %d voxels exceeds the 4 GiB bitmap limit", g.dim[0], g.dim[1], g.dim[2]);
    unsigned long long n_words = (n_bits + 31) / 32;
    CUDA_TRY(h->ovl_bits_a.reserve((size_t)n_words));
    CUDA_TRY(h->ovl_bits_b.reserve((size_t)n_words));
    CUDA_TRY(cudaMemsetAsync(h->ovl_bits_a.p, 0, n_words * 4, s));
    CUDA_TRY(cudaMemsetAsync(h->ovl_bits_b.p, 0, n_words * 4, s));
    // an empty cloud has an empty tree: |A| (or |B|) = 0 and the ratio below is 0/0 = NaN, as in the reference
    if (n_ref > 0) k_ray_mark<<<(unsigned)((n_ref + 127) / 128), 128, 0, s>>>(ref, (int)n_ref, ro[0], ro[1], ro[2], resolution, res_factor, g, h->ovl_bits_a.p, dcounts + 3);
    if (n_read > 0) k_ray_mark<<<(unsigned)((n_read + 127) / 128), 128, 0, s>>>(read, (int)n_read, so[0], so[1], so[2], resolution, res_factor, g, h->ovl_bits_b.p, dcounts + 3);
    unsigned long long pb = (n_words + 255) / 256;
    k_popcount<<<(unsigned)(pb < 148 * 8 ? pb : 148 * 8), 256, 0, s>>>(h->ovl_bits_a.p, h->ovl_bits_b.p, n_words, dcounts);
    CUDA_TRY(cudaMemcpyAsync(hc, dcounts, sizeof(hc), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaGetLastError());
    h->launches += 3;
    if (hc[3]) return fail(h, AICP_B200_ERR_EXTENT, "overlap: %llu ray voxels fell outside the occupancy box", hc[3]);
  }
  // octrees_overlap.cpp:47-53
  float treeA = (float)(long long)hc[0] / (float)(long long)hc[1];
  float treeB = (float)(long long)hc[0] / (float)(long long)hc[2];
  float mn = (treeB < treeA) ? treeB : treeA;
  if (overlap_pct) *overlap_pct = (float)((double)mn * 100.0);
  if (counts) { counts[0] = (int64_t)hc[0]; counts[1] = (int64_t)hc[1]; counts[2] = (int64_t)hc[2]; }
  return AICP_B200_OK;
}

}  // namespace aicp
