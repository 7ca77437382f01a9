// normals.cu -- SurfaceNormalDataPointsFilter{knn, keepNormals, keepDensities} on the GPU.
//
// replaces [UPSTREAM] libpointmatcher DataPointsFilters/SurfaceNormal.cpp as configured by
// aicp_core/config/icp/icp_autotuned.yaml:18-23: per point, the knn nearest neighbours (self
// included, ordered by (d2, index)), C = sum (p - mean)(p - mean)^T, normal = eigenvector of the smallest
// eigenvalue, density = knn / (4/3 pi r_max^3).
//
// Two kernels:
//   k_knn_warp          one WARP per run of 16 consecutive (Morton-ordered) query points, one query at a time; the list
//                       of a query seeds the next one, whose neighbourhood is almost the same, so the pruning bound is
//                       near-final before the tree is touched.  The k best candidates live one per lane; candidates arrive 32 at a
//                       time as a coalesced 512-byte chunk of the Morton-ordered cloud, each lane evaluates one, and
//                       an accepted one replaces the current worst entry, which is re-located with warp reductions
//                       (warp-shuffle top-k); a 15-round bitonic shuffle sort orders the list at the end.  The walk over the radix tree is warp-uniform (every lane reads
//                       the same 64-byte node record, a broadcast access), nearest child first, pruning against the
//                       current k-th distance; subtrees of <= 32 points are scanned as one chunk.  The 32 Morton
//                       neighbours of the query are scanned first, so the bound is tight before the tree is touched.
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
  int worst_lane;
  int k;
  int lane;
};

// The k live entries sit UNSORTED in lanes 0..k-1; only the worst one (largest (d2, id)) is tracked.  An accepted
// candidate replaces the worst entry, then the new worst is found with two warp reductions: ~15 instructions per
// insertion instead of a shuffle-up insertion into a sorted list.  The list is sorted once, at the end.
__device__ __forceinline__ void find_worst(WarpKnn& w) {
  const bool live = w.lane < w.k;
  unsigned key = live ? __float_as_uint(w.ld) : 0u;           // d2 >= 0: the bit pattern orders like the value
  unsigned kmax = __reduce_max_sync(0xFFFFFFFFu, key);
  unsigned tie = __ballot_sync(0xFFFFFFFFu, live && key == kmax);
  w.worst_d = __uint_as_float(kmax);
  if ((tie & (tie - 1)) == 0) {                                // the usual case: a single farthest entry
    w.worst_lane = __ffs(tie) - 1;
    w.worst_id = __shfl_sync(0xFFFFFFFFu, w.lid, w.worst_lane);
  } else {                                                     // equal distances: the largest index is the worst
    int idkey = ((tie >> w.lane) & 1u) ? w.lid : -1;
    w.worst_id = __reduce_max_sync(0xFFFFFFFFu, idkey);
    w.worst_lane = __ffs(__ballot_sync(0xFFFFFFFFu, idkey == w.worst_id)) - 1;
  }
}

__device__ __forceinline__ void warp_insert(WarpKnn& w, float cd, int cid, int cpos) {
  if (w.lane == w.worst_lane) { w.ld = cd; w.lid = cid; w.lpos = cpos; }
  find_worst(w);
}

// up to 32 consecutive Morton-ordered points [first, first+cnt): one candidate per lane, one coalesced load.
// A candidate that beats the current worst entry is inserted unless it is already in the list (the list is seeded with
// points that the walk meets again: the Morton window of the first query of a run, the previous query's neighbours after).
__device__ __forceinline__ void scan_range(const IndexView& ix, WarpKnn& w, int first, int cnt) {
  float d = INFINITY;
  int id = 0x7FFFFFFF;
  const bool ok = w.lane < cnt;
  if (ok) {
    float4 p = __ldg(&ix.pts[first + w.lane]);
    id = __float_as_int(p.w);
    d = d2_f(w.qx, w.qy, w.qz, p.x, p.y, p.z);
  }
  unsigned m = __ballot_sync(0xFFFFFFFFu, ok && cand_less(d, id, w.worst_d, w.worst_id));
  // points of this chunk that are already in the list (the list is seeded with points the walk meets again): one
  // warp-wide OR of their position bits removes them all at once
  unsigned off = (unsigned)(w.lpos - first);
  m &= ~__reduce_or_sync(0xFFFFFFFFu, (w.lane < w.k && off < 32u) ? (1u << off) : 0u);
  while (m) {
    int b = __ffs(m) - 1;
    m &= m - 1;
    float cd = __shfl_sync(0xFFFFFFFFu, d, b);
    int cid = __shfl_sync(0xFFFFFFFFu, id, b);
    if (cand_less(cd, cid, w.worst_d, w.worst_id)) warp_insert(w, cd, cid, first + b);
  }
}

// ascending bitonic sort of one (d2, id, pos) entry per lane.  (d2, id) is packed into one 64-bit key -- d2 >= 0, so
// its bit pattern orders like the value -- which makes a compare-exchange round 3 shuffles + one 64-bit compare.
// ids are unique, so the order is total (padding entries are all (inf, INT_MAX) and never swap among themselves).
__device__ __forceinline__ void warp_sort(WarpKnn& w) {
  unsigned long long key = ((unsigned long long)__float_as_uint(w.ld) << 32) | (unsigned)w.lid;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      unsigned long long pk = __shfl_xor_sync(0xFFFFFFFFu, key, j);
      int ppos = __shfl_xor_sync(0xFFFFFFFFu, w.lpos, j);
      bool keep_min = ((w.lane & k) == 0) == ((w.lane & j) == 0);
      if (keep_min ? (pk < key) : (pk > key)) { key = pk; w.lpos = ppos; }
    }
  }
  w.ld = __uint_as_float((unsigned)(key >> 32));
  w.lid = (int)(unsigned)key;
}

#define KNN_WARPS 8
#define KNN_LEAF 32

// warp-uniform top-down walk of one subtree (code >= 0: internal node, code < 0: chunk [~code, ~code + cnt))
__device__ __forceinline__ void knn_descend(const IndexView& ix, WarpKnn& w, int code, int cnt, int* st_a, int* st_b, float* st_d) {
  int sp = 0;
  while (true) {
    if (code < 0) {
      scan_range(ix, w, ~code, cnt);
    } else {
      const float4* r = ix.rec + 4 * (size_t)code;          // same address in every lane: one broadcast access
      float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
      int first = __float_as_int(r0.w), split = __float_as_int(r1.w), end = __float_as_int(r2.w);
      float dl = box_d2_f(make_float3(r0.x, r0.y, r0.z), make_float3(r1.x, r1.y, r1.z), w.qx, w.qy, w.qz);
      float dr = box_d2_f(make_float3(r2.x, r2.y, r2.z), make_float3(r3.x, r3.y, r3.z), w.qx, w.qy, w.qz);
      int cl = split - first, cr = end - split;
      int code_l = cl <= KNN_LEAF ? ~first : split - 1;
      int code_r = cr <= KNN_LEAF ? ~split : split;
      bool swap = dr < dl;
      int code_n = swap ? code_r : code_l, cnt_n = swap ? cr : cl;
      int code_f = swap ? code_l : code_r, cnt_f = swap ? cl : cr;
      float dn = swap ? dr : dl, df = swap ? dl : dr;
      if (df <= w.worst_d) { st_a[sp] = code_f; st_b[sp] = cnt_f; st_d[sp] = df; ++sp; }
      if (dn <= w.worst_d) { code = code_n; cnt = cnt_n; continue; }
    }
    bool found = false;
    while (sp > 0) {
      --sp;
      if (st_d[sp] <= w.worst_d) { code = st_a[sp]; cnt = st_b[sp]; found = true; break; }
    }
    if (!found) break;
  }
}
#define KNN_RUN 16        // consecutive Morton-ordered queries handled by one warp

__global__ void __launch_bounds__(32 * KNN_WARPS) k_knn_warp(IndexView ix, int k, int* __restrict__ knn_pos,
                                                             int* __restrict__ knn_out_orig, int q0, int q1, const int* __restrict__ qlist) {
  __shared__ int s_a[KNN_WARPS][AICP_STACK];
  __shared__ int s_b[KNN_WARPS][AICP_STACK];
  __shared__ float s_d[KNN_WARPS][AICP_STACK];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int run = blockIdx.x * KNN_WARPS + wid;
  // queries [q0, q1): the whole cloud, this rank's slice of a replicated reference, or (qlist) entries q0 .. q1 - 1 of a list
  // of Morton positions in ascending order -- the points whose neighbourhood an appended cloud changed (append.cu)
  const int i0 = q0 + run * KNN_RUN;
  if (i0 >= q1) return;
  const int i1 = i0 + KNN_RUN < q1 ? i0 + KNN_RUN : q1;
  int* st_a = s_a[wid]; int* st_b = s_b[wid]; float* st_d = s_d[wid];
  WarpKnn w;
  w.k = k; w.lane = lane;
  w.ld = INFINITY; w.lid = 0x7FFFFFFF; w.lpos = -1;
  for (int r = i0; r < i1; ++r) {
    const int i = qlist ? __ldg(&qlist[r]) : r;
    float4 q = __ldg(&ix.pts[i]);
    w.qx = q.x; w.qy = q.y; w.qz = q.z;
    if (r == i0) {
      // first query of the run: seed with the 32 Morton neighbours of the query (itself included) and keep the k nearest
      int pre_lo = i - 16;
      if (pre_lo > ix.n - 32) pre_lo = ix.n - 32;
      if (pre_lo < 0) pre_lo = 0;
      int pre_hi = pre_lo + 32 < ix.n ? pre_lo + 32 : ix.n;
      if (pre_lo + lane < pre_hi) {
        float4 p = __ldg(&ix.pts[pre_lo + lane]);
        w.lid = __float_as_int(p.w);
        w.lpos = pre_lo + lane;
        w.ld = d2_f(w.qx, w.qy, w.qz, p.x, p.y, p.z);
      }
      warp_sort(w);
    } else if (lane < k) {
      // next query of the run: its neighbourhood is almost the previous one, so the previous list re-measured from the
      // new query is a near-final bound before the tree is touched
      float4 p = __ldg(&ix.pts[w.lpos]);
      w.ld = d2_f(w.qx, w.qy, w.qz, p.x, p.y, p.z);
    }
    if (lane >= k) { w.ld = INFINITY; w.lid = 0x7FFFFFFF; w.lpos = -1; }
    find_worst(w);
    if (ix.n > KNN_LEAF) {
      // bottom-up: the query is a point of the cloud, so start at its own chunk and climb; a sibling subtree is entered
      // only if its box can still hold one of the k nearest, and the climb stops once the ball (q, k-th distance) lies
      // inside an ancestor that is a Morton cell (same argument as nn_search_up in search.cuh)
      int own = __ldg(&ix.owner32[i]);
      int node = own >> 1, side = own & 1;
      bool first_level = true;
      while (true) {
        const float4* r = ix.rec + 4 * (size_t)node;            // same address in every lane: one broadcast access
        float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
        int first = __float_as_int(r0.w), split = __float_as_int(r1.w), end = __float_as_int(r2.w), up = __float_as_int(r3.w);
        if (first_level) {
          if (side == 0) scan_range(ix, w, first, split - first); else scan_range(ix, w, split, end - split);
          first_level = false;
        }
        float ds = side == 0 ? box_d2_f(make_float3(r2.x, r2.y, r2.z), make_float3(r3.x, r3.y, r3.z), w.qx, w.qy, w.qz)
                             : box_d2_f(make_float3(r0.x, r0.y, r0.z), make_float3(r1.x, r1.y, r1.z), w.qx, w.qy, w.qz);
        if (ds <= w.worst_d) {
          int sf = side == 0 ? split : first, sc = side == 0 ? end - split : split - first;
          int scode = sc <= KNN_LEAF ? ~sf : (side == 0 ? split : split - 1);
          knn_descend(ix, w, scode, sc, st_a, st_b, st_d);
        }
        if (up < 0) break;
        if (ball_inside_node(ix, node, w.qx, w.qy, w.qz, w.worst_d)) break;
        side = (up >> 1) & 1;
        node = up >> 2;
      }
    } else {
      scan_range(ix, w, 0, ix.n);
    }
    warp_sort(w);                        // neighbours are reported in (d2, id) order
    if (lane < k) {
      knn_pos[(size_t)i * k + lane] = w.lpos;
      if (knn_out_orig) knn_out_orig[(size_t)__float_as_int(q.w) * k + lane] = w.lid;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// k_knn_tile: one WARP per tile of 32 consecutive Morton-ordered points, one QUERY PER LANE.
// The 32 queries of a tile are spatial neighbours, so they need almost the same candidates.  Instead of 32 tree walks
// the warp does ONE: after scanning the tile itself (every lane then holds an upper bound rho_i on its k-th distance) it
// climbs from the tile's leaf to the first ancestor whose Morton cell contains all 32 balls (q_i, rho_i) -- nothing
// outside that ancestor can be a neighbour of any lane -- and walks that subtree once, nearest child first, pruning a
// child when no lane's ball reaches its box.  A surviving chunk (<= 32 points) is staged in shared memory with one
// coalesced load and every lane measures every point of it against its own query: ~12 instructions per candidate for 32
// queries.  Each lane keeps its k best in a private max-heap in shared memory (column layout, conflict free), keyed by
// the 64-bit (d2 bits, original index) so the order is the oracle's; a heap sort at the end lists them ascending.
// Exactness: a candidate can enter lane i's heap only if (d2, id) < current worst; a subtree is pruned only when its box
// distance exceeds every lane's current worst distance; bounds only shrink.  Result = the k smallest (d2, id) per query.
#define TILE_WARPS 4
#define KNN_INF_KEY 0x7F8000007FFFFFFFull

__device__ __forceinline__ unsigned long long knn_key(float d, int id) {
  return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)id;
}

// per-lane max-heap of size k in shared memory; K / P already point at this lane's column (stride 32).
// Places (key, pos) at index i0 and sifts it down.
__device__ __forceinline__ void heap_sift_down(unsigned long long* K, int* P, int k, int i0, unsigned long long key, int pos) {
  int i = i0;
  while (true) {
    int c = 2 * i + 1;
    if (c >= k) break;
    unsigned long long kc = K[c * 32];
    if (c + 1 < k) {
      unsigned long long kr = K[(c + 1) * 32];
      if (kr > kc) { kc = kr; ++c; }
    }
    if (kc <= key) break;
    K[i * 32] = kc; P[i * 32] = P[c * 32];
    i = c;
  }
  K[i * 32] = key; P[i * 32] = pos;
}
__device__ __forceinline__ void heap_replace_top(unsigned long long* K, int* P, int k, unsigned long long key, int pos) {
  heap_sift_down(K, P, k, 0, key, pos);
}

// One accepted candidate.  The first k of a lane are stored as they come and turned into a heap once (Floyd, k/2 short
// sift-downs) instead of k full-depth replacements of infinite sentinels; `fill` counts them (== k afterwards).
__device__ __forceinline__ void heap_accept(unsigned long long* K, int* P, int k, int& fill, unsigned long long& top,
                                            unsigned long long key, int pos) {
  if (fill < k) {
    K[fill * 32] = key; P[fill * 32] = pos;
    if (++fill == k) {
      for (int i = k / 2 - 1; i >= 0; --i) heap_sift_down(K, P, k, i, K[i * 32], P[i * 32]);
      top = K[0];
    }
  } else if (key < top) {
    heap_replace_top(K, P, k, key, pos);
    top = K[0];
  }
}

// every lane measures the points [first, first + cnt) (cnt <= 32) against its own query; positions in [skip_lo, skip_hi)
// were scanned before.  Two steps, so that the expensive heap updates of the 32 lanes run side by side: (1) a
// divergence-free pass marks, per lane, the candidates within the lane's current bound (a 32-bit mask); (2) rounds in
// which every lane with a marked candidate left inserts its next one -- max_i popcount(mask_i) rounds instead of one
// round per candidate that ANY lane accepts.
__device__ __forceinline__ void tile_scan(const IndexView& ix, float4* s_pts, int first, int cnt, int skip_lo, int skip_hi,
                                          float qx, float qy, float qz, bool active, unsigned long long* K, int* P, int k,
                                          int& fill, unsigned long long& top, int lane) {
  if (lane < cnt) s_pts[lane] = __ldg(&ix.pts[first + lane]);
  __syncwarp();
  const float bound = __uint_as_float((unsigned int)(top >> 32));
  unsigned int mask = 0;
  for (int j = 0; j < cnt; ++j) {
    const float4 p = s_pts[j];                                          // broadcast read
    const float d = d2_f(qx, qy, qz, p.x, p.y, p.z);
    if (d <= bound) mask |= 1u << j;
  }
  // drop the positions scanned before (warp-uniform range) and the padding lanes
  const int lo = skip_lo - first, hi = skip_hi - first;
  if (hi > 0 && lo < 32) {
    const unsigned int upto_hi = hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
    const unsigned int upto_lo = lo <= 0 ? 0u : ((1u << lo) - 1u);
    mask &= ~(upto_hi & ~upto_lo);
  }
  if (!active) mask = 0;
  while (__any_sync(0xFFFFFFFFu, mask != 0)) {
    if (mask) {
      const int j = __ffs(mask) - 1;
      mask &= mask - 1;
      const float4 p = s_pts[j];
      const float d = d2_f(qx, qy, qz, p.x, p.y, p.z);
      heap_accept(K, P, k, fill, top, knn_key(d, __float_as_int(p.w)), first + j);
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32 * TILE_WARPS) k_knn_tile(IndexView ix, int k, int* __restrict__ knn_pos,
                                                              int* __restrict__ knn_out_orig, int q0, int q1) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int tile = blockIdx.x * TILE_WARPS + wid;
  const int i0 = q0 + tile * 32;                                        // q0 is a multiple of 32, q1 a multiple of 32 or ix.n
  if (i0 >= q1) return;                                                 // whole warp; no block-level barrier below
  const size_t per_warp = (size_t)k * 32 * 12 + 512 + AICP_STACK * 4;
  unsigned char* base = smem + per_warp * wid;
  unsigned long long* K = reinterpret_cast<unsigned long long*>(base) + lane;
  int* P = reinterpret_cast<int*>(base + (size_t)k * 32 * 8) + lane;
  float4* s_pts = reinterpret_cast<float4*>(base + (size_t)k * 32 * 12);
  int* stack = reinterpret_cast<int*>(base + (size_t)k * 32 * 12 + 512);
  const int cnt0 = ix.n - i0 < 32 ? ix.n - i0 : 32;
  const bool active = lane < cnt0;
  const float4 q = __ldg(&ix.pts[i0 + (active ? lane : 0)]);
  unsigned long long top = KNN_INF_KEY;      // worst kept key; infinite until the lane holds k candidates
  int fill = 0;
  tile_scan(ix, s_pts, i0, cnt0, 0, 0, q.x, q.y, q.z, active, K, P, k, fill, top, lane);
  // the two Morton-adjacent tiles as well: a tile that straddles a jump of the Morton curve holds too few points near
  // each of its queries, and its bounds would otherwise cover the gap
  int s_lo = i0, s_hi = i0 + cnt0;                                     // positions scanned so far
  if (i0 >= 32) { tile_scan(ix, s_pts, i0 - 32, 32, 0, 0, q.x, q.y, q.z, active, K, P, k, fill, top, lane); s_lo = i0 - 32; }
  if (i0 + 32 < ix.n) {
    const int c1 = ix.n - (i0 + 32) < 32 ? ix.n - (i0 + 32) : 32;
    tile_scan(ix, s_pts, i0 + 32, c1, 0, 0, q.x, q.y, q.z, active, K, P, k, fill, top, lane);
    s_hi = i0 + 32 + c1;
  }
  if (ix.n > s_hi - s_lo) {
    // box around the 32 balls (q_i, rho_i): radii rounded up, ends rounded outward
    float blx = INFINITY, bly = INFINITY, blz = INFINITY, bhx = -INFINITY, bhy = -INFINITY, bhz = -INFINITY;
    if (active) {
      const float r = __fsqrt_ru(__uint_as_float((unsigned int)(top >> 32)));
      blx = __fsub_rd(q.x, r); bly = __fsub_rd(q.y, r); blz = __fsub_rd(q.z, r);
      bhx = __fadd_ru(q.x, r); bhy = __fadd_ru(q.y, r); bhz = __fadd_ru(q.z, r);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      blx = fminf(blx, __shfl_xor_sync(0xFFFFFFFFu, blx, off)); bhx = fmaxf(bhx, __shfl_xor_sync(0xFFFFFFFFu, bhx, off));
      bly = fminf(bly, __shfl_xor_sync(0xFFFFFFFFu, bly, off)); bhy = fmaxf(bhy, __shfl_xor_sync(0xFFFFFFFFu, bhy, off));
      blz = fminf(blz, __shfl_xor_sync(0xFFFFFFFFu, blz, off)); bhz = fmaxf(bhz, __shfl_xor_sync(0xFFFFFFFFu, bhz, off));
    }
    // climb to the first ancestor whose (shrunk) Morton cell holds the whole box: no point outside it is in any ball
    int node = __ldg(&ix.owner32[i0]) >> 1;
    while (true) {
      const int up = __float_as_int(__ldg(&ix.rec[4 * (size_t)node + 3]).w);
      if (up < 0) break;
      const float4 clo = __ldg(&ix.cellbox[2 * (size_t)node]), chi = __ldg(&ix.cellbox[2 * (size_t)node + 1]);
      if (blx >= clo.x && bly >= clo.y && blz >= clo.z && bhx <= chi.x && bhy <= chi.y && bhz <= chi.z) break;
      node = up >> 2;
    }
    // one walk of that subtree for the whole tile
    int sp = 0, code = node;
    while (true) {
      const float4* r = ix.rec + 4 * (size_t)code;                      // same address in every lane: broadcast
      const float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
      const int first = __float_as_int(r0.w), split = __float_as_int(r1.w), end = __float_as_int(r2.w);
      const float dl = box_d2_f(make_float3(r0.x, r0.y, r0.z), make_float3(r1.x, r1.y, r1.z), q.x, q.y, q.z);
      const float dr = box_d2_f(make_float3(r2.x, r2.y, r2.z), make_float3(r3.x, r3.y, r3.z), q.x, q.y, q.z);
      // nearest child first (by the closest lane); children inside the already scanned positions are skipped
      const unsigned int ml = __reduce_min_sync(0xFFFFFFFFu, active ? __float_as_uint(dl) : 0xFFFFFFFFu);
      const unsigned int mr = __reduce_min_sync(0xFFFFFFFFu, active ? __float_as_uint(dr) : 0xFFFFFFFFu);
      const bool lfirst = ml <= mr;
      const bool own_l = first >= s_lo && split <= s_hi, own_r = split >= s_lo && end <= s_hi;
      int next = -1;
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const bool left = (pass == 0) == lfirst;
        if (left ? own_l : own_r) continue;
        const float dc = left ? dl : dr;
        if (__ballot_sync(0xFFFFFFFFu, active && dc <= __uint_as_float((unsigned int)(top >> 32))) == 0u) continue;
        const int cf = left ? first : split, cc = left ? split - first : end - split;
        if (cc <= 32) tile_scan(ix, s_pts, cf, cc, s_lo, s_hi, q.x, q.y, q.z, active, K, P, k, fill, top, lane);
        else {
          const int child = left ? split - 1 : split;
          if (next < 0) next = child; else stack[sp++] = child;        // the nearer internal child goes first
        }
      }
      if (next >= 0) code = next;
      else if (sp > 0) code = stack[--sp];
      else break;
    }
  }
  // heap sort: ascending (d2, id) in slots 0..k-1
  for (int m = k - 1; m > 0; --m) {
    const unsigned long long km = K[m * 32];
    const int pm = P[m * 32];
    K[m * 32] = K[0]; P[m * 32] = P[0];
    heap_replace_top(K, P, m, km, pm);
  }
  if (active) {
    int* out = knn_pos + (size_t)(i0 + lane) * k;
    for (int s = 0; s < k; ++s) out[s] = P[s * 32];
    if (knn_out_orig) {
      int* oo = knn_out_orig + (size_t)__float_as_int(q.w) * k;
      for (int s = 0; s < k; ++s) oo[s] = (int)(unsigned int)K[s * 32];
    }
  }
}

__global__ void __launch_bounds__(128) k_normals_from_knn(IndexView ix, int k, const int* __restrict__ knn_pos,
                                                          float4* __restrict__ normals_morton, int q0, int q1,
                                                          const int* __restrict__ qlist, float* __restrict__ rk2) {
  int i = q0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q1) return;
  if (qlist) i = __ldg(&qlist[i]);
  if (rk2) {
    // squared distance to the LAST neighbour of the list (lists are ascending in (d2, id)): a later point enters this list
    // exactly when it is nearer than that, which is how an append finds the neighbourhoods it changed (append.cu)
    const float4 a = __ldg(&ix.pts[i]), b = __ldg(&ix.pts[__ldg(&knn_pos[(size_t)i * k + k - 1])]);
    rk2[i] = d2_f(a.x, a.y, a.z, b.x, b.y, b.z);
  }
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

// exact k-NN lists (positions in Morton order, ascending (d2, original index)) into h->knn_pos; shared by the SurfaceNormal
// filter of the ICP chain and by the pre-filter (prefilter.cu: pcl::NormalEstimation + pcl::RegionGrowing neighbourhoods)
int run_knn(Handle* h, const SpatialIndex& ix, int knn, int* knn_out_orig, int q0, int q1, const int* qlist) {
  if (knn < 1 || knn > 32) return fail(h, AICP_B200_ERR_BAD_ARG, "k-NN search: knn %d outside [1,32]", knn);
  if (knn >= ix.n) return fail(h, AICP_B200_ERR_KNN_TOO_LARGE, "k-NN search: knn %d >= %d points", knn, ix.n);
  if (q1 < 0) q1 = ix.n;
  CUDA_TRY(h->knn_pos.reserve((size_t)ix.n * knn));
  if (q0 >= q1) return AICP_B200_OK;
  const int nq = q1 - q0;
  // Two schedules of the same exact search (identical output): the tile kernel executes ~45 % fewer instructions and wins
  // whenever the GPU is full (batched registrations, large clouds); the warp-per-query kernel has twice as many, shorter
  // warps and wins the latency of ONE lidar-sized cloud on an otherwise idle GPU.  A query list goes to the warp kernel.
  const bool tile = !qlist && (h->knn_schedule == 2 || (h->knn_schedule == 0 && (h->batch_worker || nq >= (1 << 20))));
  if (!tile) {
    k_knn_warp<<<(unsigned)((nq + KNN_WARPS * KNN_RUN - 1) / (KNN_WARPS * KNN_RUN)), 32 * KNN_WARPS, 0, h->stream>>>(ix.view(), knn, h->knn_pos.p, knn_out_orig, q0, q1, qlist);
  } else {
    const size_t smem = ((size_t)knn * 32 * 12 + 512 + AICP_STACK * 4) * TILE_WARPS;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(k_knn_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (nq + 31) / 32;
    k_knn_tile<<<(unsigned)((tiles + TILE_WARPS - 1) / TILE_WARPS), 32 * TILE_WARPS, smem, h->stream>>>(ix.view(), knn, h->knn_pos.p, knn_out_orig, q0, q1);
  }
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return AICP_B200_OK;
}

// [q0, q1): the Morton positions whose normals are computed (q1 < 0: all) -- a sharded registration gives every rank one
// slice of the replicated reference and all-gathers the results (icp.cu); q0 must be a multiple of 32.  qlist: [q0, q1) are
// entries of a list of positions instead.  rk2 (nullable): per-point squared distance to the last neighbour of the list.
int run_surface_normals(Handle* h, const SpatialIndex& ix, int knn, float4* normals_morton, int* knn_out_orig, int q0, int q1,
                        const int* qlist, float* rk2) {
  if (q1 < 0) q1 = ix.n;
  int rc = run_knn(h, ix, knn, knn_out_orig, q0, q1, qlist);
  if (rc) return rc;
  if (q0 >= q1) return AICP_B200_OK;
  k_normals_from_knn<<<(q1 - q0 + 127) / 128, 128, 0, h->stream>>>(ix.view(), knn, h->knn_pos.p, normals_morton, q0, q1, qlist, rk2);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return AICP_B200_OK;
}

}  // namespace aicp
