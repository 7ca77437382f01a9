// append.cu -- incremental insert into the device-resident reference index.
//
// replaces, for a reference that lives on the GPU across registrations, what App does after every accepted registration:
// merge the aligned cloud into the map and use the bigger map as the next reference (aicp_core/src/registration/app.cpp:
// 476-493, AlignedCloud merging) -- which with aicp_b200_set_reference meant a full rebuild: Morton sort, tree and the
// SurfaceNormal filter over EVERY point (5.7 + 14.5 ms for the 10.5 M-point map).
//
// aicp_b200_reference_append(new cloud) leaves the handle in exactly the state aicp_b200_set_reference(old + new) followed
// by a rebuild would produce -- same Morton order, same tree, same normals, bit for bit -- but does only this:
//   1. stats of the new points; if one lies outside the old bounding box the Morton quantisation changes and everything is
//      rebuilt (the fallback); otherwise the old keys stay valid
//   2. keys of the new points, radix sort of that small set
//   3. MERGE of the two sorted runs: every element finds its slot with one binary search in the OTHER run (old elements
//      precede new ones of the same key: their original indices are smaller, which is the order a stable sort of the whole
//      set gives); points, keys, indices, normals and k-th-neighbour distances move together (one pass over the map)
//   4. radix tree + boxes over the merged arrays (build_tree: the same kernels as a full build)
//   5. the neighbourhoods that changed: a new point p enters the k-NN list of an old point q exactly when it is nearer to q
//      than q's last list entry (rk2[q], kept from when q's normal was computed; p's original index is larger than any
//      old one, so an exact tie changes nothing).  Maxima of rk2 over the chunks of 32 / 1024 / 32768 consecutive
//      points (next to the chunk boxes of the same chunks) let one warp per new point find those q without touching the
//      rest of the map
//   6. exact k-NN + normal for the listed points only (the new ones and the affected old ones), through the same kernels
// The mean of the reference changes with every append (exact fixed-point sums: add the new points' sums); the centred
// copies are refreshed by the next registration.
#include <utility>

#include "handle.cuh"

namespace aicp {

// ---- merge of two sorted runs ------------------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_u32(const unsigned int* __restrict__ a, int n, unsigned int key) {   // first i with a[i] >= key
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&a[mid]) < key) lo = mid + 1; else hi = mid; }
  return lo;
}
__device__ __forceinline__ int upper_bound_u32(const unsigned int* __restrict__ a, int n, unsigned int key) {   // first i with a[i] > key
  int lo = 0, hi = n;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&a[mid]) <= key) lo = mid + 1; else hi = mid; }
  return lo;
}

// thread i < n_old moves old element i, thread n_old + j moves new element j
__global__ void __launch_bounds__(256) k_merge(const unsigned int* __restrict__ keys_old, const unsigned int* __restrict__ vals_old,
                                               const float4* __restrict__ pts_old, const float4* __restrict__ nrm_old,
                                               const float* __restrict__ rk2_old, int n_old, const unsigned int* __restrict__ keys_new,
                                               const unsigned int* __restrict__ vals_new, const float4* __restrict__ new_pts, int m,
                                               unsigned int* __restrict__ keys, unsigned int* __restrict__ vals, float4* __restrict__ pts,
                                               float4* __restrict__ nrm, float* __restrict__ rk2, unsigned int* __restrict__ is_new) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_old) {
    const unsigned int key = __ldg(&keys_old[t]);
    const int dst = t + lower_bound_u32(keys_new, m, key);
    keys[dst] = key; vals[dst] = __ldg(&vals_old[t]); pts[dst] = __ldg(&pts_old[t]); nrm[dst] = __ldg(&nrm_old[t]);
    rk2[dst] = __ldg(&rk2_old[t]); is_new[dst] = 0u;
  } else if (t < n_old + m) {
    const int j = t - n_old;
    const unsigned int key = __ldg(&keys_new[j]);
    const int dst = j + upper_bound_u32(keys_old, n_old, key);
    const unsigned int v = __ldg(&vals_new[j]);                     // original index n_old_total + (position in the appended cloud)
    const float4 p = __ldg(&new_pts[v - (unsigned int)n_old]);
    keys[dst] = key; vals[dst] = v; pts[dst] = make_float4(p.x, p.y, p.z, __int_as_float((int)v));
    nrm[dst] = make_float4(0.f, 0.f, 0.f, 0.f);
    rk2[dst] = -1.f;                                               // no old point's list can be "beaten" through this entry
    is_new[dst] = 1u;
  }
}

// ---- maxima of rk2 over the chunks of 32 / 1024 / 32768 consecutive points (layout of the chunk boxes) ---------------------
__global__ void __launch_bounds__(1024) k_chunk_rmax(const float* __restrict__ rk2, int n, float* __restrict__ l1, float* __restrict__ l2) {
  __shared__ float s_m[32];
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float v = i < n ? __ldg(&rk2[i]) : -1.f;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, off));
  if (lane == 0) { s_m[w] = v; if ((blockIdx.x * 32 + w) * 32 < n) l1[blockIdx.x * 32 + w] = v; }
  __syncthreads();
  if (w == 0) {
    v = s_m[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, off));
    if (lane == 0) l2[blockIdx.x] = v;
  }
}
__global__ void __launch_bounds__(256) k_chunk_rmax_top(const float* __restrict__ l2, int n_l2, float* __restrict__ l3, int n_l3) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= n_l3) return;
  float v = c * 32 + lane < n_l2 ? __ldg(&l2[c * 32 + lane]) : -1.f;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xFFFFFFFFu, v, off));
  if (lane == 0) l3[c] = v;
}

// ---- which old points get a new neighbour: one warp per appended point ---------------------------------------------------
// Descends the three chunk levels; a chunk is entered when the new point is within sqrt(max rk2) of its box
// (box_d2_f <= d2_f of any point inside, so nothing is missed).  flag[q] = 1 for every old q with d2(q, p) <= rk2[q]
// ("<=": a superset of the strict condition; recomputing an unchanged neighbourhood reproduces it).
__global__ void __launch_bounds__(256) k_mark_affected(const float4* __restrict__ pts, const float* __restrict__ rk2, int n,
                                                       const float4* __restrict__ box1, const float4* __restrict__ box2,
                                                       const float4* __restrict__ box3, const float* __restrict__ r1, const float* __restrict__ r2,
                                                       const float* __restrict__ r3, const int* __restrict__ new_pos, int m,
                                                       unsigned int* __restrict__ flag) {
  const int wq = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (wq >= m) return;
  const float4 p = __ldg(&pts[__ldg(&new_pos[wq])]);
  const int n1 = (n + 31) / 32, n2 = (n + 1023) / 1024, n3 = (n + 32767) / 32768;
  auto reach = [&](const float4* box, const float* r, int c, int cn) {
    if (c >= cn) return false;
    const float4 lo = __ldg(&box[2 * (size_t)c]), hi = __ldg(&box[2 * (size_t)c + 1]);
    return box_d2_f(make_float3(lo.x, lo.y, lo.z), make_float3(hi.x, hi.y, hi.z), p.x, p.y, p.z) <= __ldg(&r[c]);
  };
  for (int c3 = 0; c3 < n3; c3 += 32) {
    unsigned int m3 = __ballot_sync(0xFFFFFFFFu, reach(box3, r3, c3 + lane, n3));
    while (m3) {
      const int a3 = c3 + __ffs(m3) - 1; m3 &= m3 - 1;
      unsigned int m2 = __ballot_sync(0xFFFFFFFFu, reach(box2, r2, a3 * 32 + lane, n2));
      while (m2) {
        const int a2 = a3 * 32 + __ffs(m2) - 1; m2 &= m2 - 1;
        unsigned int m1 = __ballot_sync(0xFFFFFFFFu, reach(box1, r1, a2 * 32 + lane, n1));
        while (m1) {
          const int a1 = a2 * 32 + __ffs(m1) - 1; m1 &= m1 - 1;
          const int q = a1 * 32 + lane;
          if (q < n) {
            const float4 o = __ldg(&pts[q]);
            if (d2_f(o.x, o.y, o.z, p.x, p.y, p.z) <= __ldg(&rk2[q])) flag[q] = 1u;
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_flag_new(const unsigned int* __restrict__ is_new, int n, unsigned int* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = __ldg(&is_new[i]);
}
// positions with flag set -> list[scan[i]] (ascending); new_only: the positions of the appended points
__global__ void __launch_bounds__(256) k_compact_positions(const unsigned int* __restrict__ flag, const unsigned int* __restrict__ scan, int n,
                                                           int* __restrict__ list) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && __ldg(&flag[i])) list[__ldg(&scan[i])] = i;
}

__global__ void k_meta_add(IndexMeta* m, const IndexMeta* add) {
  if (threadIdx.x == 0 && blockIdx.x == 0) for (int d = 0; d < 3; ++d) m->csum[d] += add->csum[d];
}

int run_reference_append(Handle* h, const float4* new_pts, int64_t m64, aicp_b200_append_info* info) {
  cudaStream_t s = h->stream;
  SpatialIndex& ix = h->ref_ix;
  const int n_old = ix.n, m = (int)m64, n = n_old + m, knn = h->cfg.knn_normals;
  cudaEvent_t e0 = h->ev[0], e1 = h->ev[3];
  CUDA_TRY(cudaEventRecord(e0, s));
  // ---- 1. are the old keys still valid?
  if (!h->app_meta) CUDA_TRY(cudaMalloc((void**)&h->app_meta, sizeof(IndexMeta)));
  launch_index_stats(h, new_pts, m, h->app_meta);
  IndexMeta hm[2];
  CUDA_TRY(cudaMemcpyAsync(&hm[0], h->app_meta, sizeof(IndexMeta), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(&hm[1], ix.meta, sizeof(IndexMeta), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  h->launches += 2;
  if (hm[0].nonfinite) return fail(h, AICP_B200_ERR_NONFINITE_INPUT, "reference_append: the appended cloud contains non-finite coordinates");
  bool inside = true;
  for (int d = 0; d < 3; ++d) inside = inside && hm[0].bmin[d] >= hm[1].bmin[d] && hm[0].bmax[d] <= hm[1].bmax[d];
  // the original-order copy grows in every case (it is what a full rebuild reads)
  if ((size_t)n > h->ref_in.cap) {
    DevBuf<float4> bigger;
    CUDA_TRY(bigger.reserve((size_t)n + (size_t)n / 4));
    CUDA_TRY(cudaMemcpyAsync(bigger.p, h->ref_in.p, sizeof(float4) * (size_t)n_old, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    h->ref_in.release();
    h->ref_in = bigger;
  }
  CUDA_TRY(cudaMemcpyAsync(h->ref_in.p + n_old, new_pts, sizeof(float4) * (size_t)m, cudaMemcpyDeviceToDevice, s));
  h->n_ref = n;
  if (info) { memset(info, 0, sizeof(*info)); info->n_total = n; }
  if (!inside || n_old <= 32 || knn >= n_old) {
    // the bounding box grows: new quantisation, new keys for every point -- rebuild at the next registration
    h->ref_ready = false;
    CUDA_TRY(cudaStreamSynchronize(s));
    return AICP_B200_OK;
  }
  // ---- 2. keys of the new points, sorted (original indices continue at n_old)
  CUDA_TRY(h->app_keys.reserve((size_t)n)); CUDA_TRY(h->app_vals.reserve((size_t)n));
  CUDA_TRY(h->app_new.reserve((size_t)2 * m));
  unsigned int* kn = h->app_new.p; unsigned int* vn = h->app_new.p + m;               // the new run: keys | original indices
  CUDA_TRY(ix.keys_alt.reserve((size_t)m)); CUDA_TRY(ix.vals_alt.reserve((size_t)m));
  launch_morton_keys(h, ix, new_pts, m, kn, vn, (unsigned int)n_old);
  int rc = radix_sort_pairs(h, kn, vn, ix.keys_alt.p, ix.vals_alt.p, m, ix.sort_tmp);
  if (rc) return rc;
  // ---- 3. merge into fresh arrays, then swap them in
  CUDA_TRY(h->app_pts.reserve((size_t)n)); CUDA_TRY(h->app_normals.reserve((size_t)n)); CUDA_TRY(h->app_rk2.reserve((size_t)n));
  CUDA_TRY(h->app_flag.reserve((size_t)2 * n)); CUDA_TRY(h->app_scan.reserve((size_t)n));
  unsigned int* is_new = h->app_flag.p + n;
  k_merge<<<(n + 255) / 256, 256, 0, s>>>(ix.keys.p, ix.vals.p, ix.pts.p, h->normals.p, h->ref_rk2.p, n_old, kn, vn, new_pts, m,
                                         h->app_keys.p, h->app_vals.p, h->app_pts.p, h->app_normals.p, h->app_rk2.p, is_new);
  CUDA_TRY(cudaGetLastError());
  // the live arrays and the merge targets trade places (nothing is freed here: the kernels already enqueued keep their pointers)
  std::swap(ix.keys, h->app_keys); std::swap(ix.vals, h->app_vals); std::swap(ix.pts, h->app_pts);
  std::swap(h->normals, h->app_normals); std::swap(h->ref_rk2, h->app_rk2);
  ix.n = n;
  k_meta_add<<<1, 32, 0, s>>>(ix.meta, h->app_meta);
  h->launches += 3;
  // ---- 4. tree over the merged arrays
  if ((rc = build_tree(h, ix, n))) return rc;
  // ---- 5. the neighbourhoods the new points enter
  const int n1 = (n + 31) / 32, n2 = (n + 1023) / 1024, n3 = (n + 32767) / 32768;
  CUDA_TRY(h->app_rmax.reserve((size_t)n1 + n2 + n3 + 8));
  float* r1 = h->app_rmax.p; float* r2 = r1 + n1; float* r3 = r2 + n2;
  k_chunk_rmax<<<n2, 1024, 0, s>>>(h->ref_rk2.p, n, r1, r2);
  k_chunk_rmax_top<<<(n3 * 32 + 255) / 256, 256, 0, s>>>(r2, n2, r3, n3);
  CUDA_TRY(h->app_tiles.reserve((size_t)n / 1024 + 8));          // exclusive_scan_u32 works in tiles of 1024
  CUDA_TRY(h->app_list.reserve((size_t)n));
  unsigned int* flag = h->app_flag.p;
  unsigned int* total = h->app_tiles.p + (n / 1024 + 4);
  // positions of the new points (ascending), then the flags: new points + affected old ones
  if ((rc = exclusive_scan_u32(h, is_new, h->app_scan.p, n, h->app_tiles.p, total))) return rc;
  k_compact_positions<<<(n + 255) / 256, 256, 0, s>>>(is_new, h->app_scan.p, n, h->app_list.p);
  k_flag_new<<<(n + 255) / 256, 256, 0, s>>>(is_new, n, flag);
  const float4* b1 = ix.chunkbox.p; const float4* b2 = b1 + 2 * (size_t)n1; const float4* b3 = b2 + 2 * (size_t)n2;
  k_mark_affected<<<(unsigned)(((size_t)m * 32 + 255) / 256), 256, 0, s>>>(ix.pts.p, h->ref_rk2.p, n, b1, b2, b3, r1, r2, r3, h->app_list.p, m, flag);
  if ((rc = exclusive_scan_u32(h, flag, h->app_scan.p, n, h->app_tiles.p, total))) return rc;
  k_compact_positions<<<(n + 255) / 256, 256, 0, s>>>(flag, h->app_scan.p, n, h->app_list.p);
  unsigned int n_list = 0;
  CUDA_TRY(cudaMemcpyAsync(&n_list, total, sizeof(n_list), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaGetLastError());
  h->launches += 6;
  // ---- 6. exact k-NN + normals of the listed points
  if ((rc = run_surface_normals(h, ix, knn, h->normals.p, nullptr, 0, (int)n_list, h->app_list.p, h->ref_rk2.p))) return rc;
  h->ref_recentre = true;
  CUDA_TRY(cudaEventRecord(e1, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (info) {
    info->incremental = 1; info->n_recomputed = (int64_t)n_list;
    cudaEventElapsedTime(&info->ms, e0, e1);
  }
  return AICP_B200_OK;
}

}  // namespace aicp
