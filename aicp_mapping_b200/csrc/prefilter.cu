// prefilter.cu -- AICP's cloud pre-filter on the GPU.
//
// replaces regionGrowingUniformPlaneSegmentationFilter (aicp_core/src/utils/filteringUtils.cpp:5-45 and :51-104), which App
// runs on every reading (app.cpp:77-110), on the first cloud (app.cpp:295) and periodically on the merged map
// (app.cpp:486-493):  pcl::VoxelGrid{0.08} -> pcl::NormalEstimation{k 30} -> pcl::RegionGrowing{50, 1e6, 15 nb, 3 deg, 1.0}
// -> concatenation of the kept clusters.  The arithmetic contract is the one of oracle/aicp_oracle_prefilter.c.
//
// Stage V, VoxelGrid (HBM-bound: 16 B/point read, 8 B/point key+index written and sorted, 16 B/voxel written):
//   k_vg_minmax    bounding box of the finite points (ordered-int atomics after a block reduction), finite count
//   k_vg_params    PCL's min_b / div_b / divb_mul and its "leaf size too small" overflow test, one thread
//   k_vg_keys      voxel index of every point as a 32-bit sort key (non-finite points: 0xFFFFFFFF, sorted to the end)
//   radix sort     (voxel index, point index), stable (sort.cu)
//   k_vg_fused     one pass over the sorted pairs: heads of the runs of equal keys, their output slots (chained scan with
//                  decoupled look-back over tiles of 2048) and, per head, the exact fixed-point sums (llrint(x * 2^20),
//                  int64 -- order independent) -> centroid.  (k_vg_heads / k_scan_* / k_vg_centroids: the first, five-launch
//                  version of the same step; the scan is still used by the cluster numbering and the FOV filter.)
// Stage N, normals: Morton index + exact k-NN (index.cu, normals.cu: the same kernels as the ICP chain's SurfaceNormal filter),
//   k_pf_normals   one thread per point: PCL's single-pass float32 covariance (shifted by the first neighbour) in list order,
//                  float64 Jacobi, curvature,
//                  viewpoint flip
// Stage R, region growing.  PCL grows regions sequentially from seeds in ascending-curvature order over the DIRECTED k-NN
//   graph (edge u -> w when w is one of u's 15 neighbours and |n_u . n_w| >= cos 3 deg).  Its result is a pure function of the
//   graph: region(v) = the lowest-ranked point that reaches v.  (Let u* be that point for v.  Nothing ranked lower reaches
//   u* -- it would reach v -- so u* is unlabelled when its turn comes and becomes a seed; no vertex on the path u* -> v was
//   taken earlier, for the same reason; so v joins u*'s region.)  That is a min-label fixed point, computed data-parallel:
//   k_pf_curv_keys + radix sort   rank of every point in (curvature, index) order
//   k_pf_edges     15-bit mask of the neighbours that pass the smoothness test, and which of those edges are two-way
//   k_pf_cc_*      points joined by two-way edges reach each other, so they share their ancestors and their final label:
//                  components of the two-way graph by lock-free union-find (atomicCAS hooking, path halving) in ONE pass over
//                  the edges -- a wall or the ground collapses into one component without any wave crossing it
//   k_pf_propagate min-label propagation between components along the remaining one-way edges (atomicMin), plus the shortcut
//                  label[c] <- label[component of seed(label[c])] (the seed's ancestors are c's ancestors); passes are enqueued
//                  in batches until one makes no change (first version without the components: 72 passes, 1.9 ms on an
//                  HDL-64 sweep)
//   k_pf_count / k_pf_seed_flags / scan / k_pf_final / radix sort / k_pf_gather   cluster sizes, the [min, max] size filter,
//                  cluster ordinals in seed order, output = clusters in seed order, ascending point index inside
//   The curvature threshold only matters for points whose curvature exceeds it (they join a region but are not expanded);
//   with the reference's 1.0 that needs |lambda_min| > |trace|, i.e. a numerically broken covariance.  Such a cloud is
//   rejected (AICP_B200_ERR_EXTENT) instead of being processed with different semantics.
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "detmath.cuh"
#include "handle.cuh"

namespace aicp {

#define PF_FIXED 1048576.0          // 2^20: centroid fixed point (oracle contract)
#define PF_MAX_PASSES 4096
#define PF_BATCH 8
#define SC_ITEMS 4
#define SC_TILE (256 * SC_ITEMS)

struct PfMeta {
  int mn[3], mx[3];                 // ordered-int encodings of the bounding box of the finite points
  unsigned long long n_finite;
  int extent_bad;                   // a finite coordinate with |c| >= 32768 m
  int overflow;                     // PCL's "leaf size is too small for the input dataset"
  int min_b[3], div_b[3];
  int mul1, mul2;
  unsigned int n_voxels;
  unsigned int n_bad_curv;          // points with curvature > curvature_threshold
  unsigned int n_clusters, n_kept;
  unsigned int scan_total;
  unsigned int changed[PF_MAX_PASSES];
};

// ---- generic exclusive scan of unsigned ints (three launches: tile sums, scan of the tile sums, apply) -----------------
__device__ __forceinline__ unsigned int block_excl_scan_256(unsigned int v, unsigned int* s_warp, unsigned int* block_total) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned int incl = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned int o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off) incl += o;
  }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  unsigned int wbase = 0, tot = 0;
#pragma unroll
  for (int ww = 0; ww < 8; ++ww) { const unsigned int c = s_warp[ww]; if (ww < w) wbase += c; tot += c; }
  __syncthreads();                   // s_warp may be reused by the caller
  *block_total = tot;
  return wbase + incl - v;
}

__global__ void __launch_bounds__(256) k_scan_tiles(const unsigned int* __restrict__ in, int n, unsigned int* __restrict__ tile_sums) {
  __shared__ unsigned int s_warp[8];
  const int base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
  unsigned int v = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) if (base + j < n) v += __ldg(&in[base + j]);
  unsigned int tot;
  block_excl_scan_256(v, s_warp, &tot);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// one block: exclusive scan of the tile sums in place, running carry over chunks of 256; total -> *total
__global__ void __launch_bounds__(256) k_scan_top(unsigned int* __restrict__ tile_sums, int n_tiles, unsigned int* total) {
  __shared__ unsigned int s_warp[8];
  unsigned int carry = 0;
  for (int c = 0; c < n_tiles; c += 256) {
    const int i = c + threadIdx.x;
    const unsigned int v = i < n_tiles ? tile_sums[i] : 0u;
    unsigned int tot;
    const unsigned int ex = block_excl_scan_256(v, s_warp, &tot);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) k_scan_apply(const unsigned int* __restrict__ in, int n, const unsigned int* __restrict__ tile_sums,
                                                    unsigned int* __restrict__ out) {
  __shared__ unsigned int s_warp[8];
  const int base = blockIdx.x * SC_TILE + threadIdx.x * SC_ITEMS;
  unsigned int x[SC_ITEMS];
  unsigned int v = 0;
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) { x[j] = base + j < n ? __ldg(&in[base + j]) : 0u; v += x[j]; }
  unsigned int tot;
  unsigned int ex = block_excl_scan_256(v, s_warp, &tot) + __ldg(&tile_sums[blockIdx.x]);
#pragma unroll
  for (int j = 0; j < SC_ITEMS; ++j) { if (base + j < n) out[base + j] = ex; ex += x[j]; }
}

// out[i] = sum of in[0..i), *total = sum of all; tile_scratch: >= ceil(n / SC_TILE) words
int exclusive_scan_u32(Handle* h, const unsigned int* in, unsigned int* out, int n, unsigned int* tile_scratch, unsigned int* total) {
  const int n_tiles = (n + SC_TILE - 1) / SC_TILE;
  k_scan_tiles<<<n_tiles, 256, 0, h->stream>>>(in, n, tile_scratch);
  k_scan_top<<<1, 256, 0, h->stream>>>(tile_scratch, n_tiles, total);
  k_scan_apply<<<n_tiles, 256, 0, h->stream>>>(in, n, tile_scratch, out);
  CUDA_TRY(cudaGetLastError());
  h->launches += 3;
  return AICP_B200_OK;
}

// ---- stage V: pcl::VoxelGrid ---------------------------------------------------------------------------------------------
__global__ void k_pf_meta_init(PfMeta* m) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int d = 0; d < 3; ++d) { m->mn[d] = 0x7FFFFFFF; m->mx[d] = (int)0x80000000; m->min_b[d] = 0; m->div_b[d] = 0; }
    m->n_finite = 0; m->extent_bad = 0; m->overflow = 0; m->mul1 = 0; m->mul2 = 0;
    m->n_voxels = 0; m->n_bad_curv = 0; m->n_clusters = 0; m->n_kept = 0; m->scan_total = 0;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < PF_MAX_PASSES; i += gridDim.x * blockDim.x) m->changed[i] = 0u;
}

__device__ __forceinline__ bool finite3(const float4& p) { return isfinite(p.x) && isfinite(p.y) && isfinite(p.z); }

__global__ void __launch_bounds__(256) k_vg_minmax(const float4* __restrict__ pts, long long n, PfMeta* m) {
  int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
  int hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  unsigned int cnt = 0;
  int bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float4 p = __ldg(&pts[i]);
    if (!finite3(p)) continue;
    const float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (!(fabsf(c[d]) < 32768.f)) bad = 1;
      const int o = float_to_ordered(c[d]);
      lo[d] = min(lo[d], o); hi[d] = max(hi[d], o);
    }
    ++cnt;
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) { lo[d] = __reduce_min_sync(0xFFFFFFFFu, lo[d]); hi[d] = __reduce_max_sync(0xFFFFFFFFu, hi[d]); }
  cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
  bad = __any_sync(0xFFFFFFFFu, bad);
  __shared__ int s_lo[8][3], s_hi[8][3], s_bad[8];
  __shared__ unsigned int s_cnt[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { s_lo[w][d] = lo[d]; s_hi[w][d] = hi[d]; }
    s_cnt[w] = cnt; s_bad[w] = bad;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    int l = s_lo[0][d], hh = s_hi[0][d];
    for (int k = 1; k < 8; ++k) { l = min(l, s_lo[k][d]); hh = max(hh, s_hi[k][d]); }
    atomicMin(&m->mn[d], l);
    atomicMax(&m->mx[d], hh);
  }
  if (threadIdx.x == 3) {
    unsigned long long c = 0; int b = 0;
    for (int k = 0; k < 8; ++k) { c += s_cnt[k]; b |= s_bad[k]; }
    if (c) atomicAdd(&m->n_finite, c);
    if (b) atomicOr(&m->extent_bad, 1);
  }
}

__global__ void k_vg_params(PfMeta* m, float inv) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (m->n_finite == 0ull) return;
  long long dd[3];
  for (int d = 0; d < 3; ++d) {
    const float lo = ordered_to_float(m->mn[d]), hi = ordered_to_float(m->mx[d]);
    dd[d] = (long long)__fmul_rn(__fsub_rn(hi, lo), inv) + 1ll;
    const int min_b = (int)floorf(__fmul_rn(lo, inv)), max_b = (int)floorf(__fmul_rn(hi, inv));
    m->min_b[d] = min_b;
    m->div_b[d] = max_b - min_b + 1;
  }
  if (dd[0] * dd[1] * dd[2] > 2147483647ll) m->overflow = 1;
  m->mul1 = m->div_b[0];
  m->mul2 = m->div_b[0] * m->div_b[1];
}

__global__ void __launch_bounds__(256) k_vg_keys(const float4* __restrict__ pts, int n, const PfMeta* __restrict__ m, float inv,
                                                 unsigned int* __restrict__ keys, unsigned int* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(&pts[i]);
  unsigned int key = 0xFFFFFFFFu;
  if (finite3(p)) {
    const int i0 = (int)__fsub_rn(floorf(__fmul_rn(p.x, inv)), (float)m->min_b[0]);
    const int i1 = (int)__fsub_rn(floorf(__fmul_rn(p.y, inv)), (float)m->min_b[1]);
    const int i2 = (int)__fsub_rn(floorf(__fmul_rn(p.z, inv)), (float)m->min_b[2]);
    key = (unsigned int)(i0 + i1 * m->mul1 + i2 * m->mul2);
  }
  keys[i] = key;
  vals[i] = (unsigned int)i;
}

__global__ void __launch_bounds__(256) k_vg_heads(const unsigned int* __restrict__ keys, int n, unsigned int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned int k = __ldg(&keys[i]);
  flag[i] = (k != 0xFFFFFFFFu && (i == 0 || __ldg(&keys[i - 1]) != k)) ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_vg_centroids(const unsigned int* __restrict__ keys, const unsigned int* __restrict__ vals,
                                                      const unsigned int* __restrict__ flag, const unsigned int* __restrict__ slot,
                                                      const float4* __restrict__ pts, int n, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !__ldg(&flag[i])) return;
  const unsigned int k = __ldg(&keys[i]);
  long long sx = 0, sy = 0, sz = 0;
  int j = i;
  do {
    const float4 p = __ldg(&pts[__ldg(&vals[j])]);
    sx += __double2ll_rn((double)p.x * PF_FIXED);
    sy += __double2ll_rn((double)p.y * PF_FIXED);
    sz += __double2ll_rn((double)p.z * PF_FIXED);
    ++j;
  } while (j < n && __ldg(&keys[j]) == k);
  const double cnt = (double)(j - i);
  const double s = 1.0 / PF_FIXED;
  out[__ldg(&slot[i])] = make_float4((float)(((double)sx / cnt) * s), (float)(((double)sy / cnt) * s), (float)(((double)sz / cnt) * s), 1.0f);
}

// Heads, their ranks and the centroids in ONE pass over the sorted pairs (replaces k_vg_heads + the three scan launches +
// k_vg_centroids): tiles of 2048 sorted pairs (256 threads x 8 rows, row-major = sorted order); per-row warp ballots + a
// 64-entry shared scan rank the heads inside the tile; the tile's base slot comes from a chained scan with decoupled
// look-back over the tiles (the protocol of k_crop_box: tiles take their number from an atomic ticket, so a tile only waits
// for tiles that are already running); the thread of a head sums its run (which may continue into the next tiles).
#define VG_ROWS 8
#define VG_SERIAL_RUN 48     // points of a voxel summed by its head thread before the warp takes over the rest of the run
#define VG_TILE (256 * VG_ROWS)
#define VG_AGG (1ull << 62)
#define VG_PREFIX (2ull << 62)
#define VG_MASK ((1ull << 62) - 1ull)

__global__ void k_vg_fused_reset(unsigned long long* status, int n_tiles, unsigned int* ticket) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_tiles) status[i] = 0ull;
  if (i == 0) *ticket = 0u;
}

__global__ void __launch_bounds__(256) k_vg_fused(const unsigned int* __restrict__ keys, const unsigned int* __restrict__ vals,
                                                  const float4* __restrict__ pts, int n, float4* __restrict__ out,
                                                  unsigned long long* status, unsigned int* ticket, PfMeta* m) {
  __shared__ unsigned int s_tile;
  __shared__ int s_cnt[VG_ROWS * 8 + 1];
  __shared__ unsigned long long s_base;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  const long long base = (long long)tile * VG_TILE;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  unsigned int head = 0;              // bit r: element (row r of this thread) starts a run of equal voxel indices
  int rank[VG_ROWS];
#pragma unroll
  for (int r = 0; r < VG_ROWS; ++r) {
    const long long i = base + r * 256 + threadIdx.x;
    bool hd = false;
    if (i < n) {
      const unsigned int k = __ldg(&keys[i]);
      hd = k != 0xFFFFFFFFu && (i == 0 || __ldg(&keys[i - 1]) != k);
    }
    const unsigned int bal = __ballot_sync(0xFFFFFFFFu, hd);
    rank[r] = __popc(bal & ((1u << lane) - 1u));
    if (hd) head |= 1u << r;
    if (lane == 0) s_cnt[r * 8 + w] = __popc(bal);
  }
  __syncthreads();
  if (w == 0) {
    const int a = s_cnt[2 * lane], b = s_cnt[2 * lane + 1];
    int incl = a + b;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
      if (lane >= off) incl += o;
    }
    const int excl = incl - (a + b);
    s_cnt[2 * lane] = excl; s_cnt[2 * lane + 1] = excl + a;
    const unsigned int tile_total = (unsigned int)__shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long prefix = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(&status[0], VG_PREFIX | (unsigned long long)tile_total);
    } else {
      if (lane == 0) atomicExch(&status[tile], VG_AGG | (unsigned long long)tile_total);
      long long j0 = (long long)tile - 1;
      while (true) {
        const long long j = j0 - lane;
        unsigned long long st = j >= 0 ? *(volatile unsigned long long*)&status[j] : VG_PREFIX;   // nothing before tile 0
        while (__any_sync(0xFFFFFFFFu, (st >> 62) == 0ull))                  // predecessors are running (ticket order)
          if ((st >> 62) == 0ull) st = *(volatile unsigned long long*)&status[j];
        const unsigned int has_prefix = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2ull);
        const int stop = has_prefix ? __ffs(has_prefix) - 1 : 31;
        const unsigned int part = __reduce_add_sync(0xFFFFFFFFu, lane <= stop ? (unsigned int)(st & VG_MASK) : 0u);
        prefix += part;
        if (has_prefix) break;
        j0 -= 32;
      }
      if (lane == 0) atomicExch(&status[tile], VG_PREFIX | (prefix + (unsigned long long)tile_total));
    }
    if (lane == 0) {
      s_base = prefix;
      if (base + VG_TILE >= n) m->n_voxels = (unsigned int)(prefix + (unsigned long long)tile_total);   // the last tile knows the answer
    }
  }
  __syncthreads();
  const unsigned long long obase = s_base;
  const double sc = 1.0 / PF_FIXED;
#pragma unroll
  for (int r = 0; r < VG_ROWS; ++r) {
    const bool is_head = (head & (1u << r)) != 0;
    const int i = (int)(base + r * 256 + threadIdx.x);
    unsigned int k = 0;
    long long sx = 0, sy = 0, sz = 0;
    int j = i;
    bool more = false;
    if (is_head) {
      k = __ldg(&keys[i]);
      do {
        const float4 p = __ldg(&pts[__ldg(&vals[j])]);
        sx += __double2ll_rn((double)p.x * PF_FIXED);
        sy += __double2ll_rn((double)p.y * PF_FIXED);
        sz += __double2ll_rn((double)p.z * PF_FIXED);
        ++j;
      } while (j < n && __ldg(&keys[j]) == k && j - i < VG_SERIAL_RUN);
      more = j < n && __ldg(&keys[j]) == k;
    }
    // A voxel with more than VG_SERIAL_RUN points (a leaf size far above the point spacing, or a dense blob): the rest of its
    // run is summed by the whole warp, 32 points per step, instead of one thread walking it alone -- the sums are exact
    // integers, so the order of the additions does not matter
    unsigned int mm = __ballot_sync(0xFFFFFFFFu, more);
    while (mm) {
      const int src = __ffs(mm) - 1;
      mm &= mm - 1;
      const unsigned int kk = __shfl_sync(0xFFFFFFFFu, k, src);
      int jj = __shfl_sync(0xFFFFFFFFu, j, src) + lane;
      long long ax = 0, ay = 0, az = 0;
      int seen = 0;
      while (true) {
        const bool ok = jj < n && __ldg(&keys[jj]) == kk;
        if (!__any_sync(0xFFFFFFFFu, ok)) break;
        if (ok) {
          const float4 p = __ldg(&pts[__ldg(&vals[jj])]);
          ax += __double2ll_rn((double)p.x * PF_FIXED);
          ay += __double2ll_rn((double)p.y * PF_FIXED);
          az += __double2ll_rn((double)p.z * PF_FIXED);
          ++seen;
        }
        jj += 32;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        ax += __shfl_xor_sync(0xFFFFFFFFu, ax, off); ay += __shfl_xor_sync(0xFFFFFFFFu, ay, off);
        az += __shfl_xor_sync(0xFFFFFFFFu, az, off); seen += __shfl_xor_sync(0xFFFFFFFFu, seen, off);
      }
      if (lane == src) { sx += ax; sy += ay; sz += az; j += seen; }
    }
    if (is_head) {
      const double cnt = (double)(j - i);
      out[obase + (unsigned long long)(s_cnt[r * 8 + w] + rank[r])] =
          make_float4((float)(((double)sx / cnt) * sc), (float)(((double)sy / cnt) * sc), (float)(((double)sz / cnt) * sc), 1.0f);
    }
  }
}

// ---- stage N: pcl::NormalEstimation --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_pf_normals(IndexView ix, int k, const int* __restrict__ knn_pos, float vpx, float vpy, float vpz,
                                                    float curv_thr, float4* __restrict__ normals_morton, PfMeta* m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ix.n) return;
  const int* nb = knn_pos + (size_t)i * k;
  // pcl::computeMeanAndCovarianceMatrix: single-pass float32 accumulators in list order, shifted data
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f, a8 = 0.f;
  const float4 K = __ldg(&ix.pts[__ldg(&nb[0])]);     // shift by the first neighbour (the query point itself), as PCL >= 1.10
  for (int j = 0; j < k; ++j) {
    const float4 p = __ldg(&ix.pts[__ldg(&nb[j])]);
    const float x = __fsub_rn(p.x, K.x), y = __fsub_rn(p.y, K.y), z = __fsub_rn(p.z, K.z);
    a0 = __fadd_rn(a0, __fmul_rn(x, x)); a1 = __fadd_rn(a1, __fmul_rn(x, y)); a2 = __fadd_rn(a2, __fmul_rn(x, z));
    a3 = __fadd_rn(a3, __fmul_rn(y, y)); a4 = __fadd_rn(a4, __fmul_rn(y, z)); a5 = __fadd_rn(a5, __fmul_rn(z, z));
    a6 = __fadd_rn(a6, x); a7 = __fadd_rn(a7, y); a8 = __fadd_rn(a8, z);
  }
  const float kf = (float)k;
  a0 = __fdiv_rn(a0, kf); a1 = __fdiv_rn(a1, kf); a2 = __fdiv_rn(a2, kf); a3 = __fdiv_rn(a3, kf); a4 = __fdiv_rn(a4, kf);
  a5 = __fdiv_rn(a5, kf); a6 = __fdiv_rn(a6, kf); a7 = __fdiv_rn(a7, kf); a8 = __fdiv_rn(a8, kf);
  const float c00 = __fsub_rn(a0, __fmul_rn(a6, a6)), c01 = __fsub_rn(a1, __fmul_rn(a6, a7)), c02 = __fsub_rn(a2, __fmul_rn(a6, a8));
  const float c11 = __fsub_rn(a3, __fmul_rn(a7, a7)), c12 = __fsub_rn(a4, __fmul_rn(a7, a8)), c22 = __fsub_rn(a5, __fmul_rn(a8, a8));
  double a[3][3] = {{(double)c00, (double)c01, (double)c02}, {(double)c01, (double)c11, (double)c12}, {(double)c02, (double)c12, (double)c22}};
  double v[3][3];
  det_jacobi<3>(a, v);
  int smallest = 0; double sv = a[0][0];
  if (a[1][1] < sv) { smallest = 1; sv = a[1][1]; }
  if (a[2][2] < sv) { smallest = 2; sv = a[2][2]; }
  double nx = smallest == 0 ? v[0][0] : (smallest == 1 ? v[0][1] : v[0][2]);
  double ny = smallest == 0 ? v[1][0] : (smallest == 1 ? v[1][1] : v[1][2]);
  double nz = smallest == 0 ? v[2][0] : (smallest == 1 ? v[2][1] : v[2][2]);
  const double nn = sqrt((nx * nx + ny * ny) + nz * nz);
  nx = nx / nn; ny = ny / nn; nz = nz / nn;
  double lead = nx, al = fabs(nx);
  if (fabs(ny) > al) { lead = ny; al = fabs(ny); }
  if (fabs(nz) > al) { lead = nz; al = fabs(nz); }
  if (lead < 0.0) { nx = -nx; ny = -ny; nz = -nz; }
  float fx = (float)nx, fy = (float)ny, fz = (float)nz;
  // pcl::flipNormalTowardsViewpoint
  const float4 q = __ldg(&ix.pts[i]);
  const float vx = __fsub_rn(vpx, q.x), vy = __fsub_rn(vpy, q.y), vz = __fsub_rn(vpz, q.z);
  const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, fx), __fmul_rn(vy, fy)), __fmul_rn(vz, fz));
  if (cos_theta < 0.f) { fx = -fx; fy = -fy; fz = -fz; }
  const float eig_sum = __fadd_rn(__fadd_rn(c00, c11), c22);
  const float curv = eig_sum != 0.f ? fabsf(__fdiv_rn((float)sv, eig_sum)) : 0.f;
  if (curv > curv_thr) atomicAdd(&m->n_bad_curv, 1u);
  normals_morton[i] = make_float4(fx, fy, fz, curv);
}

// ---- stage R: pcl::RegionGrowing -----------------------------------------------------------------------------------------
// sort input in ORIGINAL point order so that the stable sort breaks curvature ties by index: keys[orig] = curvature bits
// (non-negative floats order like their bit patterns), vals[orig] = Morton position
__global__ void __launch_bounds__(256) k_pf_curv_keys(const float4* __restrict__ pts_morton, const float4* __restrict__ normals_morton, int n,
                                                      unsigned int* __restrict__ keys, unsigned int* __restrict__ vals,
                                                      float4* __restrict__ normals_orig) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const int orig = __float_as_int(__ldg(&pts_morton[pos]).w);
  const float4 nm = __ldg(&normals_morton[pos]);
  keys[orig] = __float_as_uint(nm.w);
  vals[orig] = (unsigned int)pos;
  normals_orig[orig] = nm;
}

// after the sort: vals[r] = Morton position of the point of rank r
__global__ void __launch_bounds__(256) k_pf_rank_init(const unsigned int* __restrict__ vals, int n, int* __restrict__ label,
                                                      int* __restrict__ seed_pos) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int pos = (int)__ldg(&vals[r]);
  label[pos] = r;
  seed_pos[r] = pos;
}

__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// mask: neighbours (of the first n_nb) that pass the smoothness test = out-edges of the directed graph;
// mutual: those whose own list contains this point too -- the test is symmetric, so these are two-way edges
__global__ void __launch_bounds__(256) k_pf_edges(const float4* __restrict__ normals_morton, const int* __restrict__ knn_pos, int k, int n_nb,
                                                  int n, float cos_thr, unsigned int* __restrict__ mask, unsigned int* __restrict__ mutual) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const float4 nc = __ldg(&normals_morton[pos]);
  const int* nb = knn_pos + (size_t)pos * k;
  unsigned int mk = 0, mu = 0;
  for (int j = 0; j < n_nb; ++j) {
    const int w = __ldg(&nb[j]);
    const float4 nn = __ldg(&normals_morton[w]);
    const float dot = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(nn.x, nc.x), __fmul_rn(nn.y, nc.y)), __fmul_rn(nn.z, nc.z)));
    if (dot < cos_thr) continue;
    mk |= 1u << j;
    if (w == pos) continue;
    const int* nw = knn_pos + (size_t)w * k;
    bool back = false;
    for (int i = 0; i < n_nb; ++i) back |= __ldg(&nw[i]) == pos;
    if (back) mu |= 1u << j;
  }
  mask[pos] = mk;
  mutual[pos] = mu;
}

// ---- strongly connected shortcut: points joined by two-way edges reach each other, hence have the same ancestors and the
// same final label.  Components of the two-way graph by lock-free union-find (hook the larger root index under the smaller
// with atomicCAS, path halving in find) -- one pass over the edges, no iteration.
// Loads here are plain (L1-cacheable): a stale value is the vertex itself or an older ancestor, which only costs a failed
// CAS or a longer walk; the flatten kernel runs after a kernel boundary and sees the final forest.
__device__ __forceinline__ int cc_find(int* parent, int v) {
  int curr = parent[v];
  if (curr != v) {
    int prev = v, next;
    while (curr > (next = parent[curr])) {
      parent[prev] = next;              // path halving; benign race: always an ancestor
      prev = curr;
      curr = next;
    }
  }
  return curr;
}

// parent[v] = the smallest two-way neighbour below v (most points are hooked before the union pass starts), else v
__global__ void __launch_bounds__(256) k_pf_cc_init(int* __restrict__ parent, int* __restrict__ crank, const unsigned int* __restrict__ mutual,
                                                    const int* __restrict__ knn_pos, int k, int n) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  unsigned int mu = __ldg(&mutual[v]);
  const int* nb = knn_pos + (size_t)v * k;
  int p = v;
  while (mu) {
    const int j = __ffs(mu) - 1;
    mu &= mu - 1u;
    p = min(p, __ldg(&nb[j]));
  }
  parent[v] = p;
  crank[v] = 0x7FFFFFFF;
}

__global__ void __launch_bounds__(256) k_pf_cc_union(int* parent, const unsigned int* __restrict__ mutual, const int* __restrict__ knn_pos,
                                                     int k, int n) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  unsigned int mu = __ldg(&mutual[v]);
  const int* nb = knn_pos + (size_t)v * k;
  int a = cc_find(parent, v);
  while (mu) {
    const int j = __ffs(mu) - 1;
    mu &= mu - 1u;
    const int w = __ldg(&nb[j]);
    if (w > v) continue;                // every two-way edge is seen from both ends: union from the larger one
    int b = cc_find(parent, w);
    bool repeat;
    do {
      repeat = false;
      if (a != b) {
        int ret;
        if (a < b) { if ((ret = atomicCAS(&parent[b], b, a)) != b) { b = ret; repeat = true; } }
        else       { if ((ret = atomicCAS(&parent[a], a, b)) != a) { a = ret; repeat = true; } }
      }
    } while (repeat);
  }
}

// root[v] and the component's lowest rank (crank[root])
__global__ void __launch_bounds__(256) k_pf_cc_flatten(int* parent, const int* __restrict__ rank_of, int* __restrict__ root, int* crank, int n) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int r = cc_find(parent, v);
  root[v] = r;
  atomicMin(&crank[r], __ldg(&rank_of[v]));
}

// out-edges that leave the component; the components' labels start at their lowest rank (crank, in place)
__global__ void __launch_bounds__(256) k_pf_cross(const int* __restrict__ root, const unsigned int* __restrict__ mask, const int* __restrict__ knn_pos,
                                                  int k, int n, unsigned int* __restrict__ cross) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n) return;
  const int rv = __ldg(&root[v]);
  unsigned int mk = __ldg(&mask[v]), cr = 0;
  const int* nb = knn_pos + (size_t)v * k;
  while (mk) {
    const int j = __ffs(mk) - 1;
    mk &= mk - 1u;
    if (__ldg(&root[__ldg(&nb[j])]) != rv) cr |= 1u << j;
  }
  cross[v] = cr;
}

// min-label propagation over the condensed graph: clabel[root] = lowest rank that reaches the component.  A pass pushes
// every component's label along its leaving edges and applies the shortcut clabel[c] <- clabel[component of seed(clabel[c])].
__global__ void __launch_bounds__(256) k_pf_propagate(int* clabel, const int* __restrict__ root, const unsigned int* __restrict__ cross,
                                                      const int* __restrict__ knn_pos, const int* __restrict__ seed_pos, int k, int n,
                                                      unsigned int* changed, int pass) {
  if (pass > 0 && changed[pass - 1] == 0u) return;          // the previous pass changed nothing: fixed point reached
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  bool any = false;
  if (v < n) {
    const int rv = __ldg(&root[v]);
    unsigned int cr = __ldg(&cross[v]);
    if (cr || rv == v) {
      int l = ld_relaxed(&clabel[rv]);
      const int ls = ld_relaxed(&clabel[__ldg(&root[__ldg(&seed_pos[l])])]);   // the seed's ancestors are this component's ancestors
      if (ls < l) { atomicMin(&clabel[rv], ls); l = ls; any = true; }
      const int* nb = knn_pos + (size_t)v * k;
      while (cr) {
        const int j = __ffs(cr) - 1;
        cr &= cr - 1u;
        const int rw = __ldg(&root[__ldg(&nb[j])]);
        if (ld_relaxed(&clabel[rw]) > l) { atomicMin(&clabel[rw], l); any = true; }
      }
    }
  }
  if (__syncthreads_or(any ? 1 : 0) && threadIdx.x == 0) atomicOr(&changed[pass], 1u);
}

__global__ void __launch_bounds__(256) k_pf_labels(const int* __restrict__ clabel, const int* __restrict__ root, int n, int* __restrict__ label) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) label[v] = __ldg(&clabel[__ldg(&root[v])]);
}

// region sizes; lanes of a warp that share a label (Morton neighbours usually do) add once
__global__ void __launch_bounds__(256) k_pf_count(const int* __restrict__ label, int n, unsigned int* __restrict__ count) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  const int l = pos < n ? __ldg(&label[pos]) : -1;
  const unsigned int peers = __match_any_sync(0xFFFFFFFFu, l);
  if (l >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&count[l], (unsigned int)__popc(peers));
}

// per rank r: count[r] is the size of the region seeded by the point of rank r (0 if that point is not a seed)
__global__ void __launch_bounds__(256) k_pf_seed_flags(const unsigned int* __restrict__ count, int n, unsigned int min_size, unsigned int max_size,
                                                       unsigned int* __restrict__ flag, PfMeta* m) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int c = 0;
  if (r < n) {
    c = __ldg(&count[r]);
    if (!(c >= min_size && c <= max_size)) c = 0;
    flag[r] = c ? 1u : 0u;
  }
  const unsigned int s = __reduce_add_sync(0xFFFFFFFFu, c);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(&m->n_kept, s);
}

// sort input in ORIGINAL order again: key = region label (seed rank) of kept regions, 0xFFFFFFFF otherwise; the stable sort
// yields clusters in seed order with ascending point index inside.  Also the per-point cluster ordinal (or -1).
__global__ void __launch_bounds__(256) k_pf_final(const float4* __restrict__ pts_morton, const int* __restrict__ label,
                                                  const unsigned int* __restrict__ flag, const unsigned int* __restrict__ ordinal, int n,
                                                  unsigned int* __restrict__ keys, unsigned int* __restrict__ vals, int* __restrict__ labels_out) {
  const int pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const int orig = __float_as_int(__ldg(&pts_morton[pos]).w);
  const int l = __ldg(&label[pos]);
  const bool keep = __ldg(&flag[l]) != 0u;
  keys[orig] = keep ? (unsigned int)l : 0xFFFFFFFFu;
  vals[orig] = (unsigned int)orig;
  labels_out[orig] = keep ? (int)__ldg(&ordinal[l]) : -1;
}

__global__ void __launch_bounds__(256) k_pf_gather(const unsigned int* __restrict__ vals, const float4* __restrict__ sampled, int n,
                                                   const PfMeta* __restrict__ m, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (unsigned int)i >= m->n_kept) return;
  out[i] = __ldg(&sampled[__ldg(&vals[i])]);
}

// ---- host ----------------------------------------------------------------------------------------------------------------
static int pf_reserve_sort(Handle* h, size_t n) {
  CUDA_TRY(h->pf_keys.reserve(n)); CUDA_TRY(h->pf_keys_alt.reserve(n));
  CUDA_TRY(h->pf_vals.reserve(n)); CUDA_TRY(h->pf_vals_alt.reserve(n));
  CUDA_TRY(h->pf_flag.reserve(n)); CUDA_TRY(h->pf_slot.reserve(n));
  CUDA_TRY(h->pf_tiles.reserve(n / SC_TILE + 2));
  return AICP_B200_OK;
}

static int pf_meta(Handle* h) {
  if (!h->pf_meta) CUDA_TRY(cudaMalloc((void**)&h->pf_meta, sizeof(PfMeta)));
  if (!h->pf_meta_host) CUDA_TRY(cudaMallocHost((void**)&h->pf_meta_host, 256));
  for (int i = 0; i < 2; ++i) if (!h->pf_ev[i]) CUDA_TRY(cudaEventCreate(&h->pf_ev[i]));
  k_pf_meta_init<<<4, 256, 0, h->stream>>>(reinterpret_cast<PfMeta*>(h->pf_meta));
  h->launches += 1;
  return AICP_B200_OK;
}

// pcl::VoxelGrid of `pts` into h->pf_sampled; *n_out = number of voxels (or n when PCL would return the input unchanged)
int run_voxel_grid(Handle* h, const float4* pts, int64_t n64, float leaf, int64_t* n_out) {
  *n_out = 0;
  h->pf_n_sampled = 0;
  if (n64 < 0 || n64 > (1ll << 28)) return fail(h, AICP_B200_ERR_BAD_ARG, "voxel_grid: cloud size %lld out of range [0, 2^28]", (long long)n64);
  if (!(leaf > 0.f)) return fail(h, AICP_B200_ERR_BAD_ARG, "voxel_grid: leaf size must be positive");
  if (n64 == 0) return AICP_B200_OK;
  const int n = (int)n64;
  cudaStream_t s = h->stream;
  int rc;
  if ((rc = pf_meta(h))) return rc;
  if ((rc = pf_reserve_sort(h, (size_t)n))) return rc;
  CUDA_TRY(h->pf_sampled.reserve((size_t)n));
  PfMeta* m = reinterpret_cast<PfMeta*>(h->pf_meta);
  const float inv = 1.0f / leaf;
  const int blocks = (n + 255) / 256;
  k_vg_minmax<<<blocks < 148 * 4 ? blocks : 148 * 4, 256, 0, s>>>(pts, (long long)n, m);
  k_vg_params<<<1, 32, 0, s>>>(m, inv);
  k_vg_keys<<<blocks, 256, 0, s>>>(pts, n, m, inv, h->pf_keys.p, h->pf_vals.p);
  h->launches += 3;
  if ((rc = radix_sort_pairs(h, h->pf_keys.p, h->pf_vals.p, h->pf_keys_alt.p, h->pf_vals_alt.p, n, h->pf_sort_tmp))) return rc;
  {
    const int n_tiles = (n + VG_TILE - 1) / VG_TILE;
    CUDA_TRY(h->pf_status.reserve((size_t)n_tiles + 1));
    unsigned int* ticket = reinterpret_cast<unsigned int*>(h->pf_status.p + n_tiles);
    k_vg_fused_reset<<<(n_tiles + 255) / 256, 256, 0, s>>>(h->pf_status.p, n_tiles, ticket);
    k_vg_fused<<<n_tiles, 256, 0, s>>>(h->pf_keys.p, h->pf_vals.p, pts, n, h->pf_sampled.p, h->pf_status.p, ticket, m);
    h->launches += 2;
  }
  CUDA_TRY(cudaGetLastError());
  PfMeta* mh = reinterpret_cast<PfMeta*>(h->pf_meta_host);
  CUDA_TRY(cudaMemcpyAsync(mh, m, offsetof(PfMeta, changed), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (mh->extent_bad) return fail(h, AICP_B200_ERR_EXTENT, "voxel_grid: a coordinate is 32768 m or more from the origin");
  if (mh->overflow) {                 // PCL: "Leaf size is too small for the input dataset" -> output = input
    if (mh->n_finite != (unsigned long long)n)
      return fail(h, AICP_B200_ERR_NONFINITE_INPUT, "voxel_grid: leaf size too small for the cloud's extent and the cloud has non-finite points");
    CUDA_TRY(cudaMemcpyAsync(h->pf_sampled.p, pts, sizeof(float4) * (size_t)n, cudaMemcpyDeviceToDevice, s));
    *n_out = n;
  } else {
    *n_out = (int64_t)mh->n_voxels;
  }
  h->pf_n_sampled = *n_out;
  return AICP_B200_OK;
}

// the whole pre-filter; result in h->pf_out (n_out points), by-products in pf_sampled / pf_normals_orig / pf_labels
int run_prefilter(Handle* h, const float4* pts, int64_t n, const aicp_b200_prefilter_config* cfg, const float* viewpoint,
                  aicp_b200_prefilter_info* info) {
  cudaStream_t s = h->stream;
  h->pf_n_out = 0; h->pf_n_clusters = 0; h->pf_n_sampled = 0; h->pf_has_segments = false;
  aicp_b200_prefilter_info inf;
  memset(&inf, 0, sizeof(inf));
  const int launches0 = h->launches;
  if (cfg->knn_normals < 3 || cfg->knn_normals > 32 || cfg->n_neighbours < 1 || cfg->n_neighbours > cfg->knn_normals ||
      cfg->min_cluster_size < 1 || cfg->max_cluster_size < cfg->min_cluster_size || !(cfg->leaf_size > 0.f))
    return fail(h, AICP_B200_ERR_BAD_ARG, "prefilter: unsupported configuration (3 <= knn_normals <= 32, 1 <= n_neighbours <= knn_normals, "
                "1 <= min_cluster_size <= max_cluster_size, leaf_size > 0)");
  for (int i = 0; i < 2; ++i) if (!h->pf_ev[i]) CUDA_TRY(cudaEventCreate(&h->pf_ev[i]));
  CUDA_TRY(cudaEventRecord(h->pf_ev[0], s));
  int64_t m64 = 0;
  int rc = run_voxel_grid(h, pts, n, cfg->leaf_size, &m64);
  if (rc) return rc;
  inf.n_sampled = m64;
  const int m = (int)m64;
  if (m <= cfg->knn_normals) {
    // fewer points than neighbours: PCL's regions cannot reach knn_normals + 1 points; with min_cluster_size above that
    // (the reference: 50 > 31) nothing is kept
    if (m >= cfg->min_cluster_size)
      return fail(h, AICP_B200_ERR_KNN_TOO_LARGE, "prefilter: %d sampled points <= knn_normals %d", m, cfg->knn_normals);
    if (info) { inf.gpu_launches = h->launches - launches0; *info = inf; }
    return AICP_B200_OK;
  }
  PfMeta* md = reinterpret_cast<PfMeta*>(h->pf_meta);
  PfMeta* mh = reinterpret_cast<PfMeta*>(h->pf_meta_host);
  if ((rc = build_index(h, h->pf_ix, h->pf_sampled.p, m))) return rc;
  if ((rc = run_knn(h, h->pf_ix, cfg->knn_normals, nullptr))) return rc;
  CUDA_TRY(h->pf_normals.reserve((size_t)m)); CUDA_TRY(h->pf_normals_orig.reserve((size_t)m));
  CUDA_TRY(h->pf_label.reserve((size_t)m)); CUDA_TRY(h->pf_seed_pos.reserve((size_t)m));
  CUDA_TRY(h->pf_mask.reserve((size_t)m)); CUDA_TRY(h->pf_count.reserve((size_t)m)); CUDA_TRY(h->pf_mutual.reserve((size_t)m));
  CUDA_TRY(h->pf_parent.reserve((size_t)m)); CUDA_TRY(h->pf_root.reserve((size_t)m)); CUDA_TRY(h->pf_clabel.reserve((size_t)m));
  CUDA_TRY(h->pf_labels_out.reserve((size_t)m)); CUDA_TRY(h->pf_out.reserve((size_t)m));
  const int blocks = (m + 255) / 256;
  const float vp[3] = {viewpoint ? viewpoint[0] : 0.f, viewpoint ? viewpoint[1] : 0.f, viewpoint ? viewpoint[2] : 0.f};
  const int k = cfg->knn_normals;
  k_pf_normals<<<(m + 127) / 128, 128, 0, s>>>(h->pf_ix.view(), k, h->knn_pos.p, vp[0], vp[1], vp[2], cfg->curvature_threshold,
                                               h->pf_normals.p, md);
  k_pf_curv_keys<<<blocks, 256, 0, s>>>(h->pf_ix.pts.p, h->pf_normals.p, m, h->pf_keys.p, h->pf_vals.p, h->pf_normals_orig.p);
  h->launches += 2;
  if ((rc = radix_sort_pairs(h, h->pf_keys.p, h->pf_vals.p, h->pf_keys_alt.p, h->pf_vals_alt.p, m, h->pf_sort_tmp))) return rc;
  k_pf_rank_init<<<blocks, 256, 0, s>>>(h->pf_vals.p, m, h->pf_label.p, h->pf_seed_pos.p);
  const float cos_thr = cosf(cfg->smoothness_threshold);      // host libm, as RegionGrowing::validatePoint and the oracle
  k_pf_edges<<<blocks, 256, 0, s>>>(h->pf_normals.p, h->knn_pos.p, k, cfg->n_neighbours, m, cos_thr, h->pf_mask.p, h->pf_mutual.p);
  // label currently holds every point's rank; components of the two-way graph, then propagation between components
  int* parent = h->pf_parent.p;
  int* root = h->pf_root.p;
  int* clabel = h->pf_clabel.p;
  k_pf_cc_init<<<blocks, 256, 0, s>>>(parent, clabel, h->pf_mutual.p, h->knn_pos.p, k, m);
  k_pf_cc_union<<<blocks, 256, 0, s>>>(parent, h->pf_mutual.p, h->knn_pos.p, k, m);
  k_pf_cc_flatten<<<blocks, 256, 0, s>>>(parent, h->pf_label.p, root, clabel, m);
  k_pf_cross<<<blocks, 256, 0, s>>>(root, h->pf_mask.p, h->knn_pos.p, k, m, h->pf_mutual.p);     // pf_mutual now: leaving edges
  h->launches += 6;
  CUDA_TRY(cudaGetLastError());
  int passes = 0;
  bool converged = false;
  while (passes < PF_MAX_PASSES && !converged) {
    for (int b = 0; b < PF_BATCH; ++b, ++passes)
      k_pf_propagate<<<blocks, 256, 0, s>>>(clabel, root, h->pf_mutual.p, h->knn_pos.p, h->pf_seed_pos.p, k, m, md->changed, passes);
    h->launches += PF_BATCH;
    CUDA_TRY(cudaGetLastError());
    unsigned int* last = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(mh) + 192);
    CUDA_TRY(cudaMemcpyAsync(last, &md->changed[passes - 1], sizeof(unsigned int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    converged = *last == 0u;
  }
  if (!converged) return fail(h, AICP_B200_ERR_BAD_ARG, "prefilter: region growing did not reach its fixed point in %d passes", passes);
  inf.passes = passes;
  k_pf_labels<<<blocks, 256, 0, s>>>(clabel, root, m, h->pf_label.p);
  h->launches += 1;
  CUDA_TRY(cudaMemsetAsync(h->pf_count.p, 0, sizeof(unsigned int) * (size_t)m, s));
  k_pf_count<<<blocks, 256, 0, s>>>(h->pf_label.p, m, h->pf_count.p);
  k_pf_seed_flags<<<blocks, 256, 0, s>>>(h->pf_count.p, m, (unsigned int)cfg->min_cluster_size, (unsigned int)cfg->max_cluster_size,
                                         h->pf_flag.p, md);
  h->launches += 2;
  if ((rc = exclusive_scan_u32(h, h->pf_flag.p, h->pf_slot.p, m, h->pf_tiles.p, &md->n_clusters))) return rc;
  k_pf_final<<<blocks, 256, 0, s>>>(h->pf_ix.pts.p, h->pf_label.p, h->pf_flag.p, h->pf_slot.p, m, h->pf_keys.p, h->pf_vals.p,
                                    h->pf_labels_out.p);
  h->launches += 1;
  if ((rc = radix_sort_pairs(h, h->pf_keys.p, h->pf_vals.p, h->pf_keys_alt.p, h->pf_vals_alt.p, m, h->pf_sort_tmp))) return rc;
  k_pf_gather<<<blocks, 256, 0, s>>>(h->pf_vals.p, h->pf_sampled.p, m, md, h->pf_out.p);
  h->launches += 1;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(h->pf_ev[1], s));
  CUDA_TRY(cudaMemcpyAsync(mh, md, offsetof(PfMeta, changed), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (mh->n_bad_curv)
    return fail(h, AICP_B200_ERR_EXTENT, "prefilter: %u points have curvature above the threshold %g (numerically broken float32 covariance: "
                "cloud too far from the origin); PCL's non-expanding points are not supported", mh->n_bad_curv, (double)cfg->curvature_threshold);
  h->pf_n_out = (int64_t)mh->n_kept;
  h->pf_n_clusters = (int64_t)mh->n_clusters;
  h->pf_has_segments = true;
  inf.n_clusters = h->pf_n_clusters;
  inf.n_out = h->pf_n_out;
  inf.gpu_launches = h->launches - launches0;
  CUDA_TRY(cudaEventElapsedTime(&inf.ms_total, h->pf_ev[0], h->pf_ev[1]));
  if (info) *info = inf;
  return AICP_B200_OK;
}

}  // namespace aicp
