// search.cuh -- exact nearest-neighbour search over the Morton-ordered spatial index.
//
// Index layout (built by index.cu), all in HBM, L2-resident for lidar-sized clouds:
//   pts  : float4[n_pad]   reference points in Morton order; .w carries the ORIGINAL index (int bits);
//                          padded to a multiple of AICP_LEAF with (+inf,+inf,+inf, INT_MAX)
//   node : float4[2 * 2*Lp] implicit complete binary tree over Lp = pow2 >= ceil(n/AICP_LEAF) leaves,
//                          node i (1-based heap order) = { lo.xyz , hi.xyz } in two float4; children 2i, 2i+1 are
//                          adjacent, so one visit loads 64 contiguous bytes; a leaf loads one 128-byte line of points.
//
// Exactness: box_d2_f(node) <= d2_f(p) for every point p stored below the node, in float arithmetic, because float
// subtraction, multiplication and addition are monotone and both sides use the same operation order.  A subtree is
// skipped only when box_d2 > best_d2 (strictly), so every point that could win on (d2, original index) is visited:
// the result equals libnabo's exact search (epsilon = 0) with ties resolved to the lowest reference index.
#pragma once

#include "common.cuh"

namespace aicp {

struct IndexView {
  const float4* __restrict__ pts;
  const float4* __restrict__ node;
  int n;          // real points
  int first_leaf; // Lp: heap index of leaf 0
};

#define AICP_STACK 48

__device__ __forceinline__ float node_d2(const IndexView& ix, int node, float qx, float qy, float qz) {
  float4 a = __ldg(&ix.node[2 * node]);
  float4 b = __ldg(&ix.node[2 * node + 1]);
  return box_d2_f(make_float3(a.x, a.y, a.z), make_float3(b.x, b.y, b.z), qx, qy, qz);
}

// 1-NN: returns position in the Morton-ordered array (so the caller can gather normals) and the squared distance.
// warm_pos >= 0 seeds the search with a known reference point (the previous iteration's match): the bound starts at its
// distance, so most of the tree is pruned at the root.  The seed is an ordinary candidate, so the result is unchanged.
__device__ inline void nn_search(const IndexView& ix, float qx, float qy, float qz, int* out_pos, float* out_d2, int warm_pos = -1) {
  int stack_n[AICP_STACK];
  float stack_d[AICP_STACK];
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
  int node = 1;
  while (true) {
    if (node >= ix.first_leaf) {
      int base = (node - ix.first_leaf) * AICP_LEAF;
#pragma unroll
      for (int j = 0; j < AICP_LEAF; ++j) {
        float4 p = __ldg(&ix.pts[base + j]);
        float d = d2_f(qx, qy, qz, p.x, p.y, p.z);
        int id = __float_as_int(p.w);
        if (d < best || (d == best && id < best_id)) { best = d; best_id = id; best_pos = base + j; }
      }
      node = 0;
    } else {
      int l = 2 * node, r = l + 1;
      float4 la = __ldg(&ix.node[2 * l]), lb = __ldg(&ix.node[2 * l + 1]);
      float4 ra = __ldg(&ix.node[2 * r]), rb = __ldg(&ix.node[2 * r + 1]);
      float dl = box_d2_f(make_float3(la.x, la.y, la.z), make_float3(lb.x, lb.y, lb.z), qx, qy, qz);
      float dr = box_d2_f(make_float3(ra.x, ra.y, ra.z), make_float3(rb.x, rb.y, rb.z), qx, qy, qz);
      int nn = l, fn = r;
      float nd = dl, fd = dr;
      if (dr < dl) { nn = r; fn = l; nd = dr; fd = dl; }
      // empty (padding) subtrees have d2 = +inf and are never entered
      bool take_near = nd <= best && nd < INFINITY;
      bool take_far = fd <= best && fd < INFINITY;
      if (take_far) { stack_n[sp] = fn; stack_d[sp] = fd; ++sp; }
      node = take_near ? nn : 0;
    }
    if (node == 0) {
      // pop the next subtree that can still improve on (best, best_id)
      while (sp > 0) {
        --sp;
        if (stack_d[sp] <= best) { node = stack_n[sp]; break; }
      }
      if (node == 0) break;
    }
  }
  *out_pos = best_pos;
  *out_d2 = best;
}

// k-NN into a caller-provided ascending list (d2, original id, position), strided so that consecutive threads touch
// consecutive shared-memory banks.  Entries are ordered by (d2, id); the list is complete (k entries) on return when
// n >= k.
struct KnnList {
  float* d2;   // [k * stride]
  int* id;     // [k * stride]
  int* pos;    // [k * stride]
  int stride;
  int k;
};

__device__ __forceinline__ bool cand_less(float d2a, int ia, float d2b, int ib) {
  return d2a < d2b || (d2a == d2b && ia < ib);
}

__device__ inline void knn_insert(const KnnList& L, int& cnt, float d, int id, int pos) {
  int n = cnt;
  const int s = L.stride;
  if (n == L.k) {
    if (!cand_less(d, id, L.d2[(L.k - 1) * s], L.id[(L.k - 1) * s])) return;
    n = L.k - 1;
  }
  int at = n;
  while (at > 0 && cand_less(d, id, L.d2[(at - 1) * s], L.id[(at - 1) * s])) {
    L.d2[at * s] = L.d2[(at - 1) * s];
    L.id[at * s] = L.id[(at - 1) * s];
    L.pos[at * s] = L.pos[(at - 1) * s];
    --at;
  }
  L.d2[at * s] = d; L.id[at * s] = id; L.pos[at * s] = pos;
  cnt = n + 1;
}

__device__ inline void knn_search(const IndexView& ix, float qx, float qy, float qz, const KnnList& L) {
  int stack_n[AICP_STACK];
  float stack_d[AICP_STACK];
  int sp = 0;
  int cnt = 0;
  int node = 1;
  const int last = (L.k - 1) * L.stride;
  while (true) {
    if (node >= ix.first_leaf) {
      int base = (node - ix.first_leaf) * AICP_LEAF;
#pragma unroll
      for (int j = 0; j < AICP_LEAF; ++j) {
        float4 p = __ldg(&ix.pts[base + j]);
        int id = __float_as_int(p.w);
        if (id != 0x7FFFFFFF) {
          float d = d2_f(qx, qy, qz, p.x, p.y, p.z);
          knn_insert(L, cnt, d, id, base + j);
        }
      }
      node = 0;
    } else {
      int l = 2 * node, r = l + 1;
      float dl = node_d2(ix, l, qx, qy, qz);
      float dr = node_d2(ix, r, qx, qy, qz);
      int nn = l, fn = r;
      float nd = dl, fd = dr;
      if (dr < dl) { nn = r; fn = l; nd = dr; fd = dl; }
      float worst = (cnt == L.k) ? L.d2[last] : INFINITY;
      bool take_near = nd <= worst && nd < INFINITY;
      bool take_far = fd <= worst && fd < INFINITY;
      if (take_far) { stack_n[sp] = fn; stack_d[sp] = fd; ++sp; }
      node = take_near ? nn : 0;
    }
    if (node == 0) {
      float worst = (cnt == L.k) ? L.d2[last] : INFINITY;
      while (sp > 0) {
        --sp;
        if (stack_d[sp] <= worst) { node = stack_n[sp]; break; }
      }
      if (node == 0) break;
    }
  }
}

}  // namespace aicp
