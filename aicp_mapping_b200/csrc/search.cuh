// search.cuh -- exact nearest-neighbour search over the Morton-ordered spatial index.
//
// Index layout (built by index.cu), all in HBM, L2-resident for lidar-sized clouds:
//   pts : float4[n]        points in Morton order; .w carries the ORIGINAL index (int bits)
//   rec : float4[4*(n-1)]  binary radix tree over the sorted Morton keys (Karras 2012 topology: every node is an octree
//                          cell prefix, so sibling boxes are disjoint cells -- on lidar clouds this visits ~4x fewer nodes
//                          and ~8x fewer points than a fixed-fan-out tree over the same order).  One 64-byte record per
//                          internal node holds BOTH child boxes plus the node's point range:
//                            rec[4i+0] = (lo_left.xyz , first)      rec[4i+1] = (hi_left.xyz , split)
//                            rec[4i+2] = (lo_right.xyz, end)        rec[4i+3] = (hi_right.xyz, unused)
//                          left child = node `split-1` over [first, split), right child = node `split` over [split, end).
//                          A child with <= LEAF points is scanned directly instead of being entered, so the per-thread
//                          1-NN walk (LEAF 8) and the warp k-NN walk (LEAF 32) share one tree.
//                          rec[4i+3].w is the up-link (parent << 2 | side in parent << 1 | node-is-a-Morton-cell), -1 at
//                          the root: searches that start from a known nearby point climb instead of descending.
//
// Exactness: box_d2_f(child) <= d2_f(p) for every point p stored below it, in float arithmetic, because float
// subtraction, multiplication and addition are monotone and both sides use the same operation order.  A subtree is
// skipped only when box_d2 > best_d2 (strictly), so every point that could win on (d2, original index) is visited:
// the result equals libnabo's exact search (epsilon = 0) with ties resolved to the lowest reference index.
#pragma once

#include "common.cuh"

namespace aicp {

// per-index reduction results and quantisation parameters (device memory, written by index.cu)
struct IndexMeta {
  int bmin[3], bmax[3];          // ordered-int encoded bounding box
  long long csum[3];             // sum of round(coord * 2^16)
  int nonfinite;
  int pad;
  float qlo[3];                  // Morton quantisation: cell = (int)((x - qlo) * qscale), clamped to [0, 1023]
  float qscale;
  float cell_margin;             // how far the stored cell boxes are shrunk (see ball_inside_node)
};

struct IndexView {
  const float4* __restrict__ pts;
  const float4* __restrict__ rec;
  const int* __restrict__ owner8;     // per point: (node << 1 | side) of the lowest node with more than 8 points above it
  const int* __restrict__ owner32;    // same for 32 points (warp k-NN chunks)
  const float4* __restrict__ cellbox; // per internal node: [2i] = lo.xyz, [2i+1] = hi.xyz of the node's Morton cell, shrunk
  int n;
};

__device__ __forceinline__ unsigned int morton_spread10(unsigned int v) {
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}
__device__ __forceinline__ unsigned int morton_compact10(unsigned int v) {
  v &= 0x09249249u;
  v = (v | (v >> 2)) & 0x030C30C3u;
  v = (v | (v >> 4)) & 0x0300F00Fu;
  v = (v | (v >> 8)) & 0x030000FFu;
  v = (v | (v >> 16)) & 0x000003FFu;
  return v;
}

// the quantisation of index.cu's k_morton_keys: monotone non-decreasing in x
__device__ __forceinline__ int morton_quant(float x, float lo, float scale) {
  return min(1023, max(0, (int)((x - lo) * scale)));
}
__device__ __forceinline__ unsigned int morton_key(int qx, int qy, int qz) {
  return morton_spread10((unsigned)qx) | (morton_spread10((unsigned)qy) << 1) | (morton_spread10((unsigned)qz) << 2);
}

// Stop test of the bottom-up searches: "can the ball (q, sqrt(d2)) reach a point outside node c?"
// A node whose keys share a prefix of the 30 key bits and that is a Morton cell holds ALL indexed points whose key
// carries that prefix, i.e. all points whose quantised coordinates fall into an axis-aligned range of cells.  index.cu
// stores that range as a float box per node, shrunk by a margin (1 % of a quantisation cell + 1e-5 of the coordinate
// magnitude) that exceeds every rounding between a coordinate and its cell (the quantisation, the centring shift of the
// reference frame): a point whose key lies outside the prefix is strictly outside the stored box.  If the ball, with the
// radius rounded up and the ends rounded outward, lies inside the box, no point outside the node can beat or tie the
// current bound.  Sides that coincide with the clamped ends of the quantisation range are infinite; a node that splits
// equal keys is not a cell and stores an empty box.
__device__ __forceinline__ bool ball_inside_node(const IndexView& ix, int node, float qx, float qy, float qz, float d2) {
  const float4 lo = __ldg(&ix.cellbox[2 * (size_t)node]), hi = __ldg(&ix.cellbox[2 * (size_t)node + 1]);
  const float r = __fsqrt_ru(d2);
  return __fsub_rd(qx, r) >= lo.x && __fadd_ru(qx, r) <= hi.x && __fsub_rd(qy, r) >= lo.y && __fadd_ru(qy, r) <= hi.y &&
         __fsub_rd(qz, r) >= lo.z && __fadd_ru(qz, r) <= hi.z;
}

#define AICP_STACK 64

// experiment builds (-DAICP_DEBUG_WARP_TIMES=<iteration>): per-query work counters threaded through the 1-NN search
#ifdef AICP_DEBUG_WARP_TIMES
struct DbgCnt { int levels, nodes, points, descents; };
#define AICP_DBG_PARAM , DbgCnt* dbg
#define AICP_DBG_ARG , dbg
#define AICP_DBG(x) do { if (dbg) { x; } } while (0)
#else
#define AICP_DBG_PARAM
#define AICP_DBG_ARG
#define AICP_DBG(x) do { } while (0)
#endif

__device__ __forceinline__ bool cand_less(float d2a, int ia, float d2b, int ib) {
  return d2a < d2b || (d2a == d2b && ia < ib);
}

struct NnBest {
  float d;
  int id;
  int pos;
};

// Leaf scan.  The loads are issued AICP_SCAN_BATCH at a time before the first distance is needed: the search is bound by
// the latency of its slowest warp (ncu: 26 of 39 stall cycles per issue are barrier waits for it, issue slots 18 % busy), and
// a scan that waits for every point in turn is a chain of up to 16 L1/L2 round trips.  Points past the end of the range are
// replaced by the last one (re-evaluating a point cannot change the result).
#ifndef AICP_SCAN_BATCH
#define AICP_SCAN_BATCH 4
#endif
__device__ __forceinline__ void nn_scan(const IndexView& ix, float qx, float qy, float qz, int first, int cnt, NnBest& b AICP_DBG_PARAM) {
  const int last = first + cnt - 1;
  AICP_DBG(dbg->points += cnt);
  for (int j = first; j <= last; j += AICP_SCAN_BATCH) {
    float4 p[AICP_SCAN_BATCH];
#pragma unroll
    for (int u = 0; u < AICP_SCAN_BATCH; ++u) p[u] = __ldg(&ix.pts[min(j + u, last)]);
#pragma unroll
    for (int u = 0; u < AICP_SCAN_BATCH; ++u) {
      const float d = d2_f(qx, qy, qz, p[u].x, p[u].y, p[u].z);
      const int id = __float_as_int(p[u].w);
      if (d < b.d || (d == b.d && id < b.id)) { b.d = d; b.id = id; b.pos = min(j + u, last); }
    }
  }
}

// top-down walk of one subtree (code >= 0: internal node, code < 0: scan range [~code, ~code + cnt))
__device__ inline void nn_descend(const IndexView& ix, float qx, float qy, float qz, int code, int cnt, NnBest& b AICP_DBG_PARAM) {
  int st_a[AICP_STACK];      // internal node index, or ~first for a scan range
  int st_b[AICP_STACK];      // point count of a scan range
  float st_d[AICP_STACK];
  int sp = 0;
  while (true) {
    if (code < 0) {
      nn_scan(ix, qx, qy, qz, ~code, cnt, b AICP_DBG_ARG);
    } else {
      AICP_DBG(dbg->nodes += 1);
      const float4* r = ix.rec + 4 * (size_t)code;
      float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
      int first = __float_as_int(r0.w), split = __float_as_int(r1.w), end = __float_as_int(r2.w);
      float dl = box_d2_f(make_float3(r0.x, r0.y, r0.z), make_float3(r1.x, r1.y, r1.z), qx, qy, qz);
      float dr = box_d2_f(make_float3(r2.x, r2.y, r2.z), make_float3(r3.x, r3.y, r3.z), qx, qy, qz);
      int cl = split - first, cr = end - split;
      int code_l = cl <= AICP_LEAF ? ~first : split - 1;
      int code_r = cr <= AICP_LEAF ? ~split : split;
      // nearer child next, farther child on the stack
      bool swap = dr < dl;
      int code_n = swap ? code_r : code_l, cnt_n = swap ? cr : cl;
      int code_f = swap ? code_l : code_r, cnt_f = swap ? cl : cr;
      float dn = swap ? dr : dl, df = swap ? dl : dr;
      if (df <= b.d) { st_a[sp] = code_f; st_b[sp] = cnt_f; st_d[sp] = df; ++sp; }
      if (dn <= b.d) { code = code_n; cnt = cnt_n; continue; }
    }
    // pop the next subtree that can still improve on (best, best_id)
    bool found = false;
    while (sp > 0) {
      --sp;
      if (st_d[sp] <= b.d) { code = st_a[sp]; cnt = st_b[sp]; found = true; break; }
    }
    if (!found) break;
  }
}

// 1-NN from the root: returns position in the Morton-ordered array (so the caller can gather normals) and d2.
__device__ inline void nn_search(const IndexView& ix, float qx, float qy, float qz, int* out_pos, float* out_d2 AICP_DBG_PARAM) {
  NnBest b{INFINITY, 0x7FFFFFFF, -1};
  if (ix.n <= AICP_LEAF) nn_scan(ix, qx, qy, qz, 0, ix.n, b AICP_DBG_ARG);
  else nn_descend(ix, qx, qy, qz, 0, ix.n, b AICP_DBG_ARG);
  *out_pos = b.pos;
  *out_d2 = b.d;
}

// 1-NN seeded with a reference point known to be near (the previous ICP iteration's match).  The walk starts at the
// seed's leaf and CLIMBS: at every ancestor the sibling subtree is entered only if its box can still beat the bound, and
// the climb stops as soon as the ball (q, best) lies inside an ancestor that is a Morton cell -- every point outside that
// ancestor has a key outside the cell, hence (monotone quantisation) lies outside the ball.  Once ICP has pulled the
// clouds together the ball is a few centimetres wide and the climb ends 3-5 levels above the leaf instead of walking a
// root-to-leaf path; the result is identical to nn_search (the same candidates can win).
__device__ inline void nn_search_up(const IndexView& ix, float qx, float qy, float qz, int seed_pos, int* out_pos, float* out_d2 AICP_DBG_PARAM) {
  NnBest b{INFINITY, 0x7FFFFFFF, -1};
  if (ix.n <= AICP_LEAF) {
    nn_scan(ix, qx, qy, qz, 0, ix.n, b AICP_DBG_ARG);
  } else {
    int own = __ldg(&ix.owner8[seed_pos]);
    int node = own >> 1, side = own & 1;
    bool first_level = true;
    while (true) {
      AICP_DBG(dbg->levels += 1);
      const float4* r = ix.rec + 4 * (size_t)node;
      float4 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2), r3 = __ldg(r + 3);
      int first = __float_as_int(r0.w), split = __float_as_int(r1.w), end = __float_as_int(r2.w), up = __float_as_int(r3.w);
      if (first_level) {
        // the seed's own small range first: it holds the seed, so the bound is at most the seed's distance afterwards
        if (side == 0) nn_scan(ix, qx, qy, qz, first, split - first, b AICP_DBG_ARG); else nn_scan(ix, qx, qy, qz, split, end - split, b AICP_DBG_ARG);
        first_level = false;
      }
      // sibling subtree of the side we came from
      float ds = side == 0 ? box_d2_f(make_float3(r2.x, r2.y, r2.z), make_float3(r3.x, r3.y, r3.z), qx, qy, qz)
                           : box_d2_f(make_float3(r0.x, r0.y, r0.z), make_float3(r1.x, r1.y, r1.z), qx, qy, qz);
      if (ds <= b.d) {
        int sf = side == 0 ? split : first, sc = side == 0 ? end - split : split - first;
        int scode = sc <= AICP_LEAF ? ~sf : (side == 0 ? split : split - 1);
        AICP_DBG(dbg->descents += 1);
        nn_descend(ix, qx, qy, qz, scode, sc, b AICP_DBG_ARG);
      }
      if (up < 0) break;                                   // root done
      // stop once the ball (q, best) cannot reach outside this node (see ball_inside_node)
      if (ball_inside_node(ix, node, qx, qy, qz, b.d)) break;
      side = (up >> 1) & 1;
      node = up >> 2;
    }
  }
  *out_pos = b.pos;
  *out_d2 = b.d;
}

}  // namespace aicp
