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
//
// Exactness: box_d2_f(child) <= d2_f(p) for every point p stored below it, in float arithmetic, because float
// subtraction, multiplication and addition are monotone and both sides use the same operation order.  A subtree is
// skipped only when box_d2 > best_d2 (strictly), so every point that could win on (d2, original index) is visited:
// the result equals libnabo's exact search (epsilon = 0) with ties resolved to the lowest reference index.
#pragma once

#include "common.cuh"

namespace aicp {

struct IndexView {
  const float4* __restrict__ pts;
  const float4* __restrict__ rec;
  int n;
};

#define AICP_STACK 64

__device__ __forceinline__ bool cand_less(float d2a, int ia, float d2b, int ib) {
  return d2a < d2b || (d2a == d2b && ia < ib);
}

// 1-NN: returns position in the Morton-ordered array (so the caller can gather normals) and the squared distance.
// warm_pos >= 0 seeds the search with a known reference point (the previous iteration's match): the bound starts at its
// distance, so most of the tree is pruned at the root.  The seed is an ordinary candidate, so the result is unchanged.
__device__ inline void nn_search(const IndexView& ix, float qx, float qy, float qz, int* out_pos, float* out_d2, int warm_pos = -1) {
  int st_a[AICP_STACK];      // internal node index, or ~first for a scan range
  int st_b[AICP_STACK];      // point count of a scan range
  float st_d[AICP_STACK];
  int sp = 0;
  float best = INFINITY;
  int best_id = 0x7FFFFFFF;
  int best_pos = -1;
  if (warm_pos >= 0) {
    float4 p = __ldg(&ix.pts[warm_pos]);
    best = d2_f(qx, qy, qz, p.x, p.y, p.z);
    best_id = __float_as_int(p.w);
    best_pos = warm_pos;
  }
  int code = (ix.n <= AICP_LEAF) ? ~0 : 0;      // start at the root, or scan everything when the cloud is one leaf
  int cnt = ix.n;
  while (true) {
    if (code < 0) {
      int first = ~code;
      for (int j = 0; j < cnt; ++j) {
        float4 p = __ldg(&ix.pts[first + j]);
        float d = d2_f(qx, qy, qz, p.x, p.y, p.z);
        int id = __float_as_int(p.w);
        if (d < best || (d == best && id < best_id)) { best = d; best_id = id; best_pos = first + j; }
      }
    } else {
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
      if (df <= best) { st_a[sp] = code_f; st_b[sp] = cnt_f; st_d[sp] = df; ++sp; }
      if (dn <= best) { code = code_n; cnt = cnt_n; continue; }
    }
    // pop the next subtree that can still improve on (best, best_id)
    bool found = false;
    while (sp > 0) {
      --sp;
      if (st_d[sp] <= best) { code = st_a[sp]; cnt = st_b[sp]; found = true; break; }
    }
    if (!found) break;
  }
  *out_pos = best_pos;
  *out_d2 = best;
}

}  // namespace aicp
