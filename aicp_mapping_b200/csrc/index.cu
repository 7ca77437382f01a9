// index.cu -- device-side build of the Morton-ordered spatial index (see search.cuh for the layout).
//
//   k_index_stats   one pass over the cloud: bounding box (ordered-int atomics), exact fixed-point centroid sums
//                   (order independent, so the mean is identical on every GPU count), non-finite flag
//   k_morton_keys   30-bit Morton key of the isotropically quantised position + identity permutation
//   radix sort      (key, original index) pairs, 30 significant bits (sort.cu)
//   k_gather        Morton-ordered float4 copy with the original index in .w
//   k_radix_tree    Karras binary radix tree over the sorted keys: one thread per internal node
//   k_chunk_boxes   boxes of every 32 / 1024 / 32768 consecutive sorted points
//   k_refit         one thread per node assembles both child boxes from those chunk boxes (a range query, no bottom-up
//                   dependency chain) and writes the node's 64-byte record
//
// Algorithmic HBM bytes per point: read 16 + key/perm 8 written + 8 read by the gather + 16 written = 48,
// plus 64 B of tree record and 16 B of node range per point.
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
  // warp reduce (integer min / max / add: order independent), then one set of atomics per BLOCK
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo[d] = __reduce_min_sync(0xFFFFFFFFu, lo[d]);
    hi[d] = __reduce_max_sync(0xFFFFFFFFu, hi[d]);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cs[d] += __shfl_xor_sync(0xFFFFFFFFu, cs[d], off);
  }
  bad = __any_sync(0xFFFFFFFFu, bad);
  __shared__ int s_lo[8][3], s_hi[8][3], s_bad[8];
  __shared__ long long s_cs[8][3];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { s_lo[w][d] = lo[d]; s_hi[w][d] = hi[d]; s_cs[w][d] = cs[d]; }
    s_bad[w] = bad;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    int l = s_lo[0][d], h = s_hi[0][d];
    long long c = s_cs[0][d];
    for (int k = 1; k < 8; ++k) { l = min(l, s_lo[k][d]); h = max(h, s_hi[k][d]); c += s_cs[k][d]; }
    atomicMin(&m->bmin[d], l);
    atomicMax(&m->bmax[d], h);
    atomicAdd((unsigned long long*)&m->csum[d], (unsigned long long)c);
  }
  if (threadIdx.x == 3) {
    int b = 0;
    for (int k = 0; k < 8; ++k) b |= s_bad[k];
    if (b) atomicOr(&m->nonfinite, 1);
  }
}

// quantisation parameters from the bounding box (one thread; the searches read them back from meta)
__global__ void k_quant_params(IndexMeta* m) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float lx = ordered_to_float(m->bmin[0]), ly = ordered_to_float(m->bmin[1]), lz = ordered_to_float(m->bmin[2]);
  float ex = ordered_to_float(m->bmax[0]) - lx, ey = ordered_to_float(m->bmax[1]) - ly, ez = ordered_to_float(m->bmax[2]) - lz;
  float ext = fmaxf(ex, fmaxf(ey, ez));
  m->qlo[0] = lx; m->qlo[1] = ly; m->qlo[2] = lz;
  m->qscale = ext > 0.f ? 1023.0f / ext : 0.f;
  // margin by which the per-node cell boxes are shrunk (search.cuh, ball_inside_node): 1 % of a cell + 1e-5 of the
  // largest coordinate magnitude in either frame (the centred frame lies within the bounding box)
  float amax = 0.f;
  for (int d = 0; d < 3; ++d) amax = fmaxf(amax, fmaxf(fabsf(ordered_to_float(m->bmin[d])), fabsf(ordered_to_float(m->bmax[d]))));
  m->cell_margin = (ext > 0.f ? 0.01f * ext / 1023.0f : 0.f) + 1e-5f * (amax + ext);
}

__global__ void __launch_bounds__(256) k_morton_keys(const float4* __restrict__ pts, int n, const IndexMeta* __restrict__ m,
                                                     unsigned int* __restrict__ keys, unsigned int* __restrict__ vals, unsigned int first_index) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lx = m->qlo[0], ly = m->qlo[1], lz = m->qlo[2], scale = m->qscale;
  float4 p = __ldg(&pts[i]);
  keys[i] = morton_key(morton_quant(p.x, lx, scale), morton_quant(p.y, ly, scale), morton_quant(p.z, lz, scale));
  vals[i] = first_index + (unsigned)i;      // original index: appended points continue the numbering (append.cu)
}

// bounding box, exact centroid sums and the non-finite flag of a cloud into *m (reset first)
void launch_index_stats(Handle* h, const float4* pts, int n, IndexMeta* m) {
  k_meta_init<<<1, 32, 0, h->stream>>>(m);
  const int blocks = (n + 255) / 256;
  k_index_stats<<<blocks < 148 * 2 ? blocks : 148 * 2, 256, 0, h->stream>>>(pts, n, m);
}

void launch_morton_keys(Handle* h, const SpatialIndex& ix, const float4* pts, int n, unsigned int* keys, unsigned int* vals, unsigned int first_index) {
  k_morton_keys<<<(n + 255) / 256, 256, 0, h->stream>>>(pts, n, ix.meta, keys, vals, first_index);
}

__global__ void __launch_bounds__(256) k_gather(const float4* __restrict__ pts, const unsigned int* __restrict__ perm, int n,
                                                float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned int src = perm[i];
  float4 p = __ldg(&pts[src]);
  out[i] = make_float4(p.x, p.y, p.z, __int_as_float((int)src));
}

// ---- binary radix tree over the sorted keys (Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees,
// and k-d Trees", HPG 2012).  Keys are made unique by appending the sorted position, so equal Morton codes (many
// points in one 12 cm cell near the sensor) split by position bits.  One thread per internal node, no dependencies.
__device__ __forceinline__ int delta(const unsigned int* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  unsigned int x = __ldg(&keys[i]) ^ __ldg(&keys[j]);
  if (x) return __clz(x);
  return 32 + __clz((unsigned)i ^ (unsigned)j);
}

__global__ void __launch_bounds__(256) k_radix_tree(const unsigned int* __restrict__ keys, int n, int4* __restrict__ meta,
                                                    int* __restrict__ parent_int, int* __restrict__ parent_leaf,
                                                    int* __restrict__ owner8, int* __restrict__ owner32,
                                                    const IndexMeta* __restrict__ im, float4* __restrict__ cellbox) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  int dmin = delta(keys, n, i, i - d);
  int lmax = 2;
  while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1)
    if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  int j = i + l * d;
  int dnode = delta(keys, n, i, j);
  int s = 0, t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  int gamma = i + s * d + min(d, 0);
  int first = min(i, j), last = max(i, j);
  // dnode < 32: the node is a Morton cell (all keys sharing a prefix of the 30 key bits) -- the property the bottom-up
  // searches rely on to stop early; dnode >= 32 means the split is among equal keys (position bits) and proves nothing
  meta[i] = make_int4(first, gamma + 1, last + 1, dnode < 32 ? 1 : 0);
  {
    // float box of the node's Morton cell, shrunk by the margin (empty when the node splits equal keys)
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    if (dnode < 32) {
      const int P = dnode - 2;                                   // known leading bits of the 30-bit key
      const unsigned int key = __ldg(&keys[first]);
      const float scale = im->qscale, margin = im->cell_margin;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        // bit b of axis a sits at key bit 3b + a; it is known when 3b + a >= 30 - P
        int bmin = (30 - P - a + 2) / 3;
        if (bmin < 0) bmin = 0;
        const int na = bmin > 10 ? 0 : 10 - bmin;
        const int shift = 10 - na;
        const int c = (int)morton_compact10(key >> a);
        const int cmin = (c >> shift) << shift, cmax = cmin + (1 << shift) - 1;
        lo[a] = cmin <= 0 ? -INFINITY : im->qlo[a] + (float)cmin / scale + margin;
        hi[a] = cmax >= 1023 ? INFINITY : im->qlo[a] + (float)(cmax + 1) / scale - margin;
      }
    }
    cellbox[2 * (size_t)i] = make_float4(lo[0], lo[1], lo[2], 0.f);
    cellbox[2 * (size_t)i + 1] = make_float4(hi[0], hi[1], hi[2], 0.f);
  }
  if (first == gamma) parent_leaf[gamma] = (i << 1); else parent_int[gamma] = (i << 1);
  if (last == gamma + 1) parent_leaf[gamma + 1] = (i << 1) | 1; else parent_int[gamma + 1] = (i << 1) | 1;
  // owner of a point = the lowest node with MORE than LEAF points above it (code = node << 1 | side): where the
  // bottom-up searches start.  Written by that node for each child that is small enough to be scanned directly.
  const int split = gamma + 1, end = last + 1, count = end - first;
  const int cl = split - first, cr = end - split;
  if (count > AICP_LEAF) {
    if (cl <= AICP_LEAF) for (int p = first; p < split; ++p) owner8[p] = (i << 1);
    if (cr <= AICP_LEAF) for (int p = split; p < end; ++p) owner8[p] = (i << 1) | 1;
  }
  if (count > 32) {
    if (cl <= 32) for (int p = first; p < split; ++p) owner32[p] = (i << 1);
    if (cr <= 32) for (int p = split; p < end; ++p) owner32[p] = (i << 1) | 1;
  }
}

// ---- child boxes without a bottom-up dependency chain -----------------------------------------------------------------
// A node of the radix tree covers a contiguous range of the Morton-ordered points, so its box is a range query.  Three
// levels of chunk boxes (32, 1024 and 32768 consecutive points) are built first; every node then assembles the boxes of
// its two children from at most 31 points + 31 + 31 chunk boxes at either end and the 32768-point chunks in between.
// No atomics, no fences, no level-by-level waiting: the former bottom-up refit spent 55 us of a 131 072-point build
// waiting for ~20 dependent fence / atomic round trips.
struct Box3 {
  float3 lo, hi;
};
__device__ __forceinline__ void box_merge(Box3& b, float4 lo, float4 hi) {
  b.lo.x = fminf(b.lo.x, lo.x); b.lo.y = fminf(b.lo.y, lo.y); b.lo.z = fminf(b.lo.z, lo.z);
  b.hi.x = fmaxf(b.hi.x, hi.x); b.hi.y = fmaxf(b.hi.y, hi.y); b.hi.z = fmaxf(b.hi.z, hi.z);
}

// level 1 (32 points) and level 2 (1024 points) boxes: one block of 1024 threads per level-2 chunk
__global__ void __launch_bounds__(1024) k_chunk_boxes(const float4* __restrict__ pts, int n, float4* __restrict__ l1,
                                                      float4* __restrict__ l2) {
  __shared__ float4 s_lo[32], s_hi[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
  if (i < n) { float4 p = __ldg(&pts[i]); lo = make_float4(p.x, p.y, p.z, 0.f); hi = lo; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo.x = fminf(lo.x, __shfl_xor_sync(0xFFFFFFFFu, lo.x, off)); hi.x = fmaxf(hi.x, __shfl_xor_sync(0xFFFFFFFFu, hi.x, off));
    lo.y = fminf(lo.y, __shfl_xor_sync(0xFFFFFFFFu, lo.y, off)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xFFFFFFFFu, hi.y, off));
    lo.z = fminf(lo.z, __shfl_xor_sync(0xFFFFFFFFu, lo.z, off)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xFFFFFFFFu, hi.z, off));
  }
  if (lane == 0) {
    s_lo[w] = lo; s_hi[w] = hi;
    const int c = blockIdx.x * 32 + w;
    if (c * 32 < n) { l1[2 * (size_t)c] = lo; l1[2 * (size_t)c + 1] = hi; }
  }
  __syncthreads();
  if (w == 0) {
    lo = s_lo[lane]; hi = s_hi[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      lo.x = fminf(lo.x, __shfl_xor_sync(0xFFFFFFFFu, lo.x, off)); hi.x = fmaxf(hi.x, __shfl_xor_sync(0xFFFFFFFFu, hi.x, off));
      lo.y = fminf(lo.y, __shfl_xor_sync(0xFFFFFFFFu, lo.y, off)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xFFFFFFFFu, hi.y, off));
      lo.z = fminf(lo.z, __shfl_xor_sync(0xFFFFFFFFu, lo.z, off)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xFFFFFFFFu, hi.z, off));
    }
    if (lane == 0) { l2[2 * (size_t)blockIdx.x] = lo; l2[2 * (size_t)blockIdx.x + 1] = hi; }
  }
}

// level 3 (32768 points): one warp per chunk folds its <= 32 level-2 boxes
__global__ void __launch_bounds__(256) k_chunk_boxes_top(const float4* __restrict__ l2, int n_l2, float4* __restrict__ l3, int n_l3) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n_l3) return;
  float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.f), hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.f);
  const int j = c * 32 + lane;
  if (j < n_l2) { lo = __ldg(&l2[2 * (size_t)j]); hi = __ldg(&l2[2 * (size_t)j + 1]); }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo.x = fminf(lo.x, __shfl_xor_sync(0xFFFFFFFFu, lo.x, off)); hi.x = fmaxf(hi.x, __shfl_xor_sync(0xFFFFFFFFu, hi.x, off));
    lo.y = fminf(lo.y, __shfl_xor_sync(0xFFFFFFFFu, lo.y, off)); hi.y = fmaxf(hi.y, __shfl_xor_sync(0xFFFFFFFFu, hi.y, off));
    lo.z = fminf(lo.z, __shfl_xor_sync(0xFFFFFFFFu, lo.z, off)); hi.z = fmaxf(hi.z, __shfl_xor_sync(0xFFFFFFFFu, hi.z, off));
  }
  if (lane == 0) { l3[2 * (size_t)c] = lo; l3[2 * (size_t)c + 1] = hi; }
}

// box of the points [a, b), a < b: points up to a multiple of 32, 32-chunks up to a multiple of 1024, 1024-chunks up to a
// multiple of 32768, 32768-chunks, then down again.  Every phase is a counted loop of independent loads, so several are
// in flight at once (a data-dependent walk would serialise ~100 L2 round trips for the top nodes of the tree).
__device__ __forceinline__ Box3 range_box(const float4* __restrict__ pts, const float4* __restrict__ l1, const float4* __restrict__ l2,
                                          const float4* __restrict__ l3, int a, int b) {
  Box3 r{make_float3(INFINITY, INFINITY, INFINITY), make_float3(-INFINITY, -INFINITY, -INFINITY)};
  int i = a, e;
  e = min(b, (i + 31) & ~31);
#pragma unroll 8
  for (; i < e; ++i) { const float4 p = __ldg(&pts[i]); box_merge(r, p, p); }
  e = min(b & ~31, (i + 1023) & ~1023);
#pragma unroll 4
  for (; i < e; i += 32) box_merge(r, __ldg(&l1[2 * (size_t)(i >> 5)]), __ldg(&l1[2 * (size_t)(i >> 5) + 1]));
  e = min(b & ~1023, (i + 32767) & ~32767);
#pragma unroll 4
  for (; i < e; i += 1024) box_merge(r, __ldg(&l2[2 * (size_t)(i >> 10)]), __ldg(&l2[2 * (size_t)(i >> 10) + 1]));
  e = b & ~32767;
#pragma unroll 4
  for (; i < e; i += 32768) box_merge(r, __ldg(&l3[2 * (size_t)(i >> 15)]), __ldg(&l3[2 * (size_t)(i >> 15) + 1]));
  e = b & ~1023;
#pragma unroll 4
  for (; i < e; i += 1024) box_merge(r, __ldg(&l2[2 * (size_t)(i >> 10)]), __ldg(&l2[2 * (size_t)(i >> 10) + 1]));
  e = b & ~31;
#pragma unroll 4
  for (; i < e; i += 32) box_merge(r, __ldg(&l1[2 * (size_t)(i >> 5)]), __ldg(&l1[2 * (size_t)(i >> 5) + 1]));
#pragma unroll 4
  for (; i < b; ++i) { const float4 p = __ldg(&pts[i]); box_merge(r, p, p); }
  return r;
}

// One thread per CHILD of every internal node: the child's box + half of the node's 64-byte record
// (rec[4i+0..1] = left box + first, split;  rec[4i+2..3] = right box + end, up-link).
__global__ void __launch_bounds__(256) k_refit(const float4* __restrict__ pts, int n, const int4* __restrict__ meta,
                                               const int* __restrict__ parent_int, const float4* __restrict__ l1,
                                               const float4* __restrict__ l2, const float4* __restrict__ l3, float4* __restrict__ rec) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = t >> 1, side = t & 1;
  if (i >= n - 1) return;
  const int4 m = __ldg(&meta[i]);                                       // (first, split, end, node is a Morton cell)
  float4* r = rec + 4 * (size_t)i + 2 * side;
  if (side == 0) {
    const Box3 bx = range_box(pts, l1, l2, l3, m.x, m.y);
    r[0] = make_float4(bx.lo.x, bx.lo.y, bx.lo.z, __int_as_float(m.x));
    r[1] = make_float4(bx.hi.x, bx.hi.y, bx.hi.z, __int_as_float(m.y));
  } else {
    const Box3 bx = range_box(pts, l1, l2, l3, m.y, m.z);
    // up-link for the bottom-up searches: (parent << 2 | side-in-parent << 1 | this-node-is-a-Morton-cell), -1 at the root
    const int up = i == 0 ? -1 : ((__ldg(&parent_int[i]) << 1) | m.w);
    r[0] = make_float4(bx.lo.x, bx.lo.y, bx.lo.z, __int_as_float(m.z));
    r[1] = make_float4(bx.hi.x, bx.hi.y, bx.hi.z, __int_as_float(up));
  }
}

int build_index(Handle* h, SpatialIndex& ix, const float4* pts_dev, int64_t n64, bool with_tree, cudaEvent_t after_stats) {
  if (n64 < 1 || n64 > (1ll << 28)) return fail(h, AICP_B200_ERR_BAD_ARG, "cloud size %lld out of range [1, 2^28]", (long long)n64);
  int n = (int)n64;
  cudaStream_t s = h->stream;
  if (!ix.meta) { CUDA_TRY(cudaMalloc((void**)&ix.meta, sizeof(IndexMeta))); g_alloc_generation.fetch_add(1, std::memory_order_relaxed); }
  CUDA_TRY(ix.pts.reserve((size_t)n));
  if (with_tree) {
    CUDA_TRY(ix.rec.reserve((size_t)4 * n));
    CUDA_TRY(ix.node_meta.reserve((size_t)n));
    CUDA_TRY(ix.flags.reserve((size_t)2 * n));          // parent of internal nodes | parent of points
    CUDA_TRY(ix.chunkbox.reserve((size_t)2 * ((size_t)n / 32 + n / 1024 + n / 32768 + 3)));
    CUDA_TRY(ix.owner.reserve((size_t)2 * n));          // owner8 | owner32
    CUDA_TRY(ix.cellbox.reserve((size_t)2 * n));
  }
  CUDA_TRY(ix.keys.reserve((size_t)n)); CUDA_TRY(ix.keys_alt.reserve((size_t)n));
  CUDA_TRY(ix.vals.reserve((size_t)n)); CUDA_TRY(ix.vals_alt.reserve((size_t)n));

  k_meta_init<<<1, 32, 0, s>>>(ix.meta);
  int blocks = (n + 255) / 256;
  int stat_blocks = blocks < 148 * 2 ? blocks : 148 * 2;
  k_index_stats<<<stat_blocks, 256, 0, s>>>(pts_dev, n, ix.meta);
  k_quant_params<<<1, 32, 0, s>>>(ix.meta);
  if (after_stats) CUDA_TRY(cudaEventRecord(after_stats, s));      // bounding box and centroid sums are final: dependants may start
  k_morton_keys<<<blocks, 256, 0, s>>>(pts_dev, n, ix.meta, ix.keys.p, ix.vals.p, 0u);
  int rc = radix_sort_pairs(h, ix.keys.p, ix.vals.p, ix.keys_alt.p, ix.vals_alt.p, n, ix.sort_tmp);    // result in keys / vals
  if (rc) return rc;
  k_gather<<<blocks, 256, 0, s>>>(pts_dev, ix.vals.p, n, ix.pts.p);
  h->launches += 5;
  if (n > 1 && with_tree && (rc = build_tree(h, ix, n))) return rc;
  CUDA_TRY(cudaGetLastError());
  ix.n = n;
  return AICP_B200_OK;
}

// the radix tree, the chunk boxes and the node records over ix.keys (sorted) / ix.pts (Morton-ordered), n > 1; the buffers
// are grown here, so an append that merged more points into the arrays can call it directly
int build_tree(Handle* h, SpatialIndex& ix, int n) {
  cudaStream_t s = h->stream;
  CUDA_TRY(ix.rec.reserve((size_t)4 * n));
  CUDA_TRY(ix.node_meta.reserve((size_t)n));
  CUDA_TRY(ix.flags.reserve((size_t)2 * n));
  CUDA_TRY(ix.chunkbox.reserve((size_t)2 * ((size_t)n / 32 + n / 1024 + n / 32768 + 3)));
  CUDA_TRY(ix.owner.reserve((size_t)2 * n));
  CUDA_TRY(ix.cellbox.reserve((size_t)2 * n));
  const int blocks = (n + 255) / 256;
  int* parent_int = ix.flags.p;
  int* parent_leaf = ix.flags.p + n;
  const int n_l1 = (n + 31) / 32, n_l2 = (n + 1023) / 1024, n_l3 = (n + 32767) / 32768;
  float4* l1 = ix.chunkbox.p;
  float4* l2 = l1 + 2 * (size_t)n_l1;
  float4* l3 = l2 + 2 * (size_t)n_l2;
  k_chunk_boxes<<<n_l2, 1024, 0, s>>>(ix.pts.p, n, l1, l2);
  k_chunk_boxes_top<<<(n_l3 * 32 + 255) / 256, 256, 0, s>>>(l2, n_l2, l3, n_l3);
  k_radix_tree<<<blocks, 256, 0, s>>>(ix.keys.p, n, ix.node_meta.p, parent_int, parent_leaf, ix.owner.p, ix.owner.p + n, ix.meta, ix.cellbox.p);
  k_refit<<<(2 * (n - 1) + 255) / 256, 256, 0, s>>>(ix.pts.p, n, ix.node_meta.p, parent_int, l1, l2, l3, ix.rec.p);
  h->launches += 4;
  CUDA_TRY(cudaGetLastError());
  return AICP_B200_OK;
}

}  // namespace aicp
