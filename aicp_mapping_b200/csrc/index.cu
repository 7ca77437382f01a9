// index.cu -- device-side build of the Morton-ordered spatial index (see search.cuh for the layout).
//
//   k_index_stats   one pass over the cloud: bounding box (ordered-int atomics), exact fixed-point centroid sums
//                   (order independent, so the mean is identical on every GPU count), non-finite flag
//   k_morton_keys   30-bit Morton key of the isotropically quantised position + identity permutation
//   radix sort      (key, original index) pairs, 30 significant bits
//   k_gather        Morton-ordered float4 copy with the original index in .w, padded to whole leaves
//   k_refit         leaf boxes + bottom-up union to the root in ONE launch (second-arriver rule on atomic counters)
//
// Algorithmic HBM bytes per reference point: read 16 + key/perm 8 written + 8 read by the gather + 16
// written = 48, plus 4 B/point of node boxes.
#include <cub/device/device_radix_sort.cuh>

#include "handle.cuh"

namespace aicp {

__global__ void k_meta_init(IndexMeta* m) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    for (int d = 0; d < 3; ++d) { m->bmin[d] = 0x7FFFFFFF; m->bmax[d] = (int)0x80000000; m->csum[d] = 0; }
    m->nonfinite = 0;
  }
}

__global__ void __launch_bounds__(256) k_index_stats(const float4* __restrict__ pts, int n, IndexMeta* m) {
  int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF};
  int hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
  long long cs[3] = {0, 0, 0};
  int bad = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float4 p = __ldg(&pts[i]);
    float c[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      if (!isfinite(c[d])) { bad = 1; continue; }
      int o = float_to_ordered(c[d]);
      lo[d] = min(lo[d], o); hi[d] = max(hi[d], o);
      cs[d] += __double2ll_rn((double)c[d] * AICP_CENTROID_SCALE);
    }
  }
  // warp reduce
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = min(lo[d], __shfl_xor_sync(0xFFFFFFFFu, lo[d], off));
      hi[d] = max(hi[d], __shfl_xor_sync(0xFFFFFFFFu, hi[d], off));
      cs[d] += __shfl_xor_sync(0xFFFFFFFFu, cs[d], off);
    }
    bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, off);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      atomicMin(&m->bmin[d], lo[d]);
      atomicMax(&m->bmax[d], hi[d]);
      atomicAdd((unsigned long long*)&m->csum[d], (unsigned long long)cs[d]);
    }
    if (bad) atomicOr(&m->nonfinite, 1);
  }
}

__device__ __forceinline__ unsigned int spread10(unsigned int v) {
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

__global__ void __launch_bounds__(256) k_morton_keys(const float4* __restrict__ pts, int n, const IndexMeta* __restrict__ m,
                                                     unsigned int* __restrict__ keys, unsigned int* __restrict__ vals) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float lx = ordered_to_float(m->bmin[0]), ly = ordered_to_float(m->bmin[1]), lz = ordered_to_float(m->bmin[2]);
  float ex = ordered_to_float(m->bmax[0]) - lx, ey = ordered_to_float(m->bmax[1]) - ly, ez = ordered_to_float(m->bmax[2]) - lz;
  float ext = fmaxf(ex, fmaxf(ey, ez));
  float scale = ext > 0.f ? 1023.0f / ext : 0.f;
  float4 p = __ldg(&pts[i]);
  int qx = min(1023, max(0, (int)((p.x - lx) * scale)));
  int qy = min(1023, max(0, (int)((p.y - ly) * scale)));
  int qz = min(1023, max(0, (int)((p.z - lz) * scale)));
  keys[i] = spread10((unsigned)qx) | (spread10((unsigned)qy) << 1) | (spread10((unsigned)qz) << 2);
  vals[i] = (unsigned)i;
}

__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ pts, const unsigned int* __restrict__ perm, int n,
                                                int n_pad, float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  float4 o;
  if (i < n) {
    unsigned int src = perm[i];
    float4 p = __ldg(&pts[src]);
    o = make_float4(p.x, p.y, p.z, __int_as_float((int)src));
  } else {
    o = make_float4(INFINITY, INFINITY, INFINITY, __int_as_float(0x7FFFFFFF));
  }
  out[i] = o;
}

// One thread per leaf: compute the leaf box, then climb.  At every parent the first child to arrive stops, the
// second (which can see both boxes after the fence) writes the union and continues: one launch builds all levels.
__global__ void __launch_bounds__(256) k_refit(const float4* __restrict__ pts, int n, int first_leaf, float4* node, int* flags) {
  int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= first_leaf) return;
  float3 lo = make_float3(INFINITY, INFINITY, INFINITY), hi = make_float3(-INFINITY, -INFINITY, -INFINITY);
  int base = l * AICP_LEAF;
  if (base < n) {
#pragma unroll
    for (int j = 0; j < AICP_LEAF; ++j) {
      if (base + j < n) {
        float4 p = __ldg(&pts[base + j]);
        lo.x = fminf(lo.x, p.x); lo.y = fminf(lo.y, p.y); lo.z = fminf(lo.z, p.z);
        hi.x = fmaxf(hi.x, p.x); hi.y = fmaxf(hi.y, p.y); hi.z = fmaxf(hi.z, p.z);
      }
    }
  }
  int id = first_leaf + l;
  while (true) {
    __stcg(&node[2 * id], make_float4(lo.x, lo.y, lo.z, 0.f));
    __stcg(&node[2 * id + 1], make_float4(hi.x, hi.y, hi.z, 0.f));
    if (id == 1) break;
    __threadfence();
    int parent = id >> 1;
    if (atomicAdd(&flags[parent], 1) == 0) break;   // sibling not there yet: it will do the parent
    __threadfence();
    int sib = id ^ 1;
    float4 sa = __ldcg(&node[2 * sib]), sb = __ldcg(&node[2 * sib + 1]);
    lo.x = fminf(lo.x, sa.x); lo.y = fminf(lo.y, sa.y); lo.z = fminf(lo.z, sa.z);
    hi.x = fmaxf(hi.x, sb.x); hi.y = fmaxf(hi.y, sb.y); hi.z = fmaxf(hi.z, sb.z);
    id = parent;
  }
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

int build_index(Handle* h, SpatialIndex& ix, const float4* pts_dev, int64_t n64) {
  if (n64 < 1 || n64 > (1ll << 30)) return fail(h, AICP_B200_ERR_BAD_ARG, "cloud size %lld out of range", (long long)n64);
  int n = (int)n64;
  cudaStream_t s = h->stream;
  int leaves = (n + AICP_LEAF - 1) / AICP_LEAF;
  int first_leaf = next_pow2(leaves);
  if (first_leaf < 4) first_leaf = 4;          // the warp k-NN walks 32-point chunks = nodes two levels above the leaves
  int n_pad = first_leaf * AICP_LEAF;
  if (!ix.meta) CUDA_TRY(cudaMalloc((void**)&ix.meta, sizeof(IndexMeta)));
  CUDA_TRY(ix.pts.reserve((size_t)n_pad));
  CUDA_TRY(ix.node.reserve((size_t)4 * first_leaf));
  CUDA_TRY(ix.keys.reserve((size_t)n)); CUDA_TRY(ix.keys_alt.reserve((size_t)n));
  CUDA_TRY(ix.vals.reserve((size_t)n)); CUDA_TRY(ix.vals_alt.reserve((size_t)n));
  CUDA_TRY(ix.flags.reserve((size_t)first_leaf));
  size_t tmp_bytes = 0;
  cub::DoubleBuffer<unsigned int> dk(ix.keys.p, ix.keys_alt.p), dv(ix.vals.p, ix.vals_alt.p);
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n, 0, 30, s));
  CUDA_TRY(ix.sort_tmp.reserve(tmp_bytes));

  k_meta_init<<<1, 32, 0, s>>>(ix.meta);
  int blocks = (n + 255) / 256;
  int stat_blocks = blocks < 148 * 4 ? blocks : 148 * 4;
  k_index_stats<<<stat_blocks, 256, 0, s>>>(pts_dev, n, ix.meta);
  k_morton_keys<<<blocks, 256, 0, s>>>(pts_dev, n, ix.meta, ix.keys.p, ix.vals.p);
  CUDA_TRY(cub::DeviceRadixSort::SortPairs(ix.sort_tmp.p, tmp_bytes, dk, dv, n, 0, 30, s));
  k_gather<<<(n_pad + 255) / 256, 256, 0, s>>>(pts_dev, dv.Current(), n, n_pad, ix.pts.p);
  CUDA_TRY(cudaMemsetAsync(ix.flags.p, 0, sizeof(int) * (size_t)first_leaf, s));
  k_refit<<<(first_leaf + 255) / 256, 256, 0, s>>>(ix.pts.p, n, first_leaf, ix.node.p, ix.flags.p);
  CUDA_TRY(cudaGetLastError());
  h->launches += 5 + 4;    // own kernels + the radix sort's passes (upsweep/scan/downsweep or onesweep)
  ix.n = n; ix.n_pad = n_pad; ix.first_leaf = first_leaf;
  return AICP_B200_OK;
}

}  // namespace aicp
