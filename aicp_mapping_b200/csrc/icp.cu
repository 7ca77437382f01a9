// icp.cu -- the device-resident point-to-plane ICP loop.
//
// replaces T = icp_(reading, reference, init) at aicp_core/src/registration/pointmatcher_registration.cpp:111, i.e.
// [UPSTREAM] PM::ICP::operator() with the chain of aicp_core/config/icp/icp_autotuned.yaml:27-51:
//   KDTreeMatcher{knn 1, epsilon 0}  ->  TrimmedDistOutlierFilter{ratio}  ->  PointToPlaneErrorMinimizer
//   ->  CounterTransformationChecker + DifferentialTransformationChecker.
//
// One iteration = three launches on one stream, no host synchronisation anywhere in the loop:
//   k_match       reading' -> T_iter * reading' (float, fixed order), exact NN in the centred reference index,
//                 writes (position, d2), builds the first radix-select histogram; its last block picks the digit
//   k_select23    compacts the d2 keys that carry that digit (a few thousand) into a candidate list; its last block
//                 finishes the radix select (digits 2 and 3) over the list in shared memory -> exact k-th order statistic
//   k_accumulate  weights (d2 <= limit), F = [p x n; n], exact fixed-point sums of F F^T and F (delta.n) reduced
//                 with a transposed warp shuffle and 128-bit integer atomics; its last block solves the 6x6 system in
//                 float64, updates T_iter and evaluates both transformation checkers (-> st->done)
// Every kernel returns immediately once st->done is set.  The host enqueues the first smoothLength iterations blindly (the
// differential checker cannot stop earlier) and then stays at most two iterations ahead of the device, reading the
// iteration counter the solve publishes in mapped pinned memory -- a registration that converges after 8 of 20
// iterations does not pay for 36 empty launches.
#include <atomic>
#include <thread>

#include "detmath.cuh"
#include "handle.cuh"

#ifndef AICP_SIDE_STREAM
#define AICP_SIDE_STREAM 1      // 0: reading-side setup behind the reference's (A/B builds)
#endif

namespace aicp {

__device__ __forceinline__ int ld_int(const int* p) { return *(const volatile int*)p; }

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// loop progress for the host (mapped pinned memory): [0] = iterations completed, [1] = loop finished
__device__ __forceinline__ void publish_progress(volatile int* progress, int iter, int done) {
  if (!progress) return;
  progress[0] = iter;
  if (done) progress[1] = 1;
  __threadfence_system();
}

// early-exit path of the loop kernels: the loop is over (converged, counter, or a status was raised)
__device__ __forceinline__ void publish_done(volatile int* progress) {
  if (progress && blockIdx.x == 0 && threadIdx.x == 0) { progress[1] = 1; __threadfence_system(); }
}

__device__ __forceinline__ void raise_status(DeviceState* st, int code) {
  atomicCAS(&st->status, 0, code);
  *(volatile int*)&st->done = 1;
}

// ---- setup --------------------------------------------------------------------------------------------------------
__global__ void k_loop_init(DeviceState* st, const IndexMeta* __restrict__ meta, long long n_ref, int have_init) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->status = 0; st->done = 0; st->done_at = -1; st->stop_reason = 0; st->iter = 0; st->hist_n = 1;
  st->prefix = 0; st->k_rem = 0; st->n_valid = 0; st->limit = 0.f; st->n_used_last = 0;
  for (int i = 0; i < 4; ++i) st->ticket[i] = 0;
  st->cand_n = 0;
  for (int i = 0; i < 3; ++i) st->tail_ns[i] = 0;
  st->bar_arrive = 0; st->bar_release = 0; st->t_mark = 0; st->n_read_total = 0;
  for (int i = 0; i < 4; ++i) st->phase_ns[i] = 0;
  for (int i = 0; i < AICP_NSUM; ++i) { st->sum_lo[i] = 0; st->sum_hi[i] = 0; }
  if (meta->nonfinite) { st->status = AICP_B200_ERR_NONFINITE_INPUT; st->done = 1; }
  // A.1 step 2: centre on the mean of the (filtered) reference; exact fixed-point sum -> order independent
  for (int d = 0; d < 3; ++d) st->mu[d] = (float)((double)meta->csum[d] / (AICP_CENTROID_SCALE * (double)n_ref));
  st->mu[3] = 0.f;
  for (int d = 0; d < 3; ++d) {
    float lo = __fsub_rn(ordered_to_float(meta->bmin[d]), st->mu[d]);
    float hi = __fsub_rn(ordered_to_float(meta->bmax[d]), st->mu[d]);
    if (!(fabsf(lo) <= 1024.f) || !(fabsf(hi) <= 1024.f)) { if (!st->status) st->status = AICP_B200_ERR_EXTENT; st->done = 1; }
  }
  if (!have_init) for (int i = 0; i < 16; ++i) st->T_init[i] = (i % 5 == 0) ? 1.f : 0.f;
  // A.1 step 5: T_refMean_dataIn = T_refIn_refMean^-1 * T_init
  for (int i = 0; i < 16; ++i) st->M0[i] = st->T_init[i];
  for (int d = 0; d < 3; ++d) st->M0[12 + d] = __fsub_rn(st->T_init[12 + d], st->mu[d]);
  st->M0[3] = st->M0[7] = st->M0[11] = 0.f; st->M0[15] = 1.f;
  for (int i = 0; i < 16; ++i) st->T_iter[i] = (i % 5 == 0) ? 1.f : 0.f;
  det_quat_from_T(st->T_iter, st->quat_hist[0]);
  st->tr_hist[0][0] = st->tr_hist[0][1] = st->tr_hist[0][2] = 0.0;
}

// reference' = reference - mean: points and tree records (float subtraction is monotone, so the boxes stay valid)
__global__ void __launch_bounds__(256) k_centre(const float4* __restrict__ pts, int n, const float4* __restrict__ rec,
                                                int n_rec4, const float4* __restrict__ cell, const DeviceState* __restrict__ st,
                                                float4* __restrict__ pts_c, float4* __restrict__ rec_c, float4* __restrict__ cell_c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float mx = st->mu[0], my = st->mu[1], mz = st->mu[2];
  if (i < n) {
    float4 p = __ldg(&pts[i]);
    pts_c[i] = make_float4(__fsub_rn(p.x, mx), __fsub_rn(p.y, my), __fsub_rn(p.z, mz), p.w);
  }
  if (i < n_rec4) {
    float4 b = __ldg(&rec[i]);
    rec_c[i] = make_float4(__fsub_rn(b.x, mx), __fsub_rn(b.y, my), __fsub_rn(b.z, mz), b.w);
  }
  if (i < n_rec4 / 2) {                                    // cell boxes: 2 float4 per node; infinite sides stay infinite
    float4 b = __ldg(&cell[i]);
    cell_c[i] = make_float4(__fsub_rn(b.x, mx), __fsub_rn(b.y, my), __fsub_rn(b.z, mz), b.w);
  }
}

// reading' = T_refMean_dataIn * reading; optional initialised reading (init_T * reading) for getInitializedReading
__global__ void __launch_bounds__(256) k_read_prepare(const float4* __restrict__ read, int n, DeviceState* st,
                                                      float4* __restrict__ read0, float4* __restrict__ read_init) {
  __shared__ float sM[16], sI[16];
  if (threadIdx.x < 16) { sM[threadIdx.x] = st->M0[threadIdx.x]; sI[threadIdx.x] = st->T_init[threadIdx.x]; }
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(&read[i]);
  if (!isfinite(p.x) || !isfinite(p.y) || !isfinite(p.z)) { raise_status(st, AICP_B200_ERR_NONFINITE_INPUT); return; }
  float3 q = xform_f(sM, p.x, p.y, p.z);
  if (!(fabsf(q.x) <= 1024.f) || !(fabsf(q.y) <= 1024.f) || !(fabsf(q.z) <= 1024.f)) raise_status(st, AICP_B200_ERR_EXTENT);
  read0[i] = make_float4(q.x, q.y, q.z, p.w);
  if (read_init) {
    float3 r = xform_f(sI, p.x, p.y, p.z);
    read_init[i] = make_float4(r.x, r.y, r.z, p.w);
  }
}

// ---- trimmed quantile: 3-digit radix select on the bit pattern of d2 (positive floats order like their bits) --------
// digit 1 = bits 30..20, digit 2 = bits 19..9, digit 3 = bits 8..0
__device__ __forceinline__ bool d2_valid(float d) { return d > 0.f && d < INFINITY; }   // Matches::getDistsQuantile filter

// Rank search in a 2048-bin histogram held 8 consecutive bins per thread (256 threads): the bin that holds rank
// `target`, the rank inside that bin, and the histogram total.  pass 1 derives the target from the total (A.4):
// index = size_t(float(n_valid) * ratio), clamped to n_valid - 1; ratio == 1 -> the maximum.
__device__ __forceinline__ void block_pick(const unsigned int (&h)[8], int pass, float ratio, unsigned long long k_rem_in,
                                           unsigned int* out_bin, unsigned long long* out_rem, unsigned long long* out_total) {
  __shared__ unsigned long long s_warp[8];
  __shared__ unsigned long long s_target, s_rem, s_total;
  __shared__ unsigned int s_bin;
  const int t = threadIdx.x;
  unsigned long long local = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) local += h[j];
  unsigned long long incl = local;
  const int lane = t & 31, w = t >> 5;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
    if (lane >= off) incl += o;
  }
  if (lane == 31) s_warp[w] = incl;
  if (t == 0) { s_bin = 0; s_rem = 0; }
  __syncthreads();
  unsigned long long base = 0, total = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { if (i < w) base += s_warp[i]; total += s_warp[i]; }
  const unsigned long long excl = base + incl - local;
  if (t == 0) {
    unsigned long long target;
    if (pass == 1) {
      if (total == 0) target = 0;
      else if (ratio == 1.0f) target = total - 1;
      else {
        float fi = __fmul_rn(__ull2float_rn(total), ratio);      // size_t * float in float32, truncated (A.4)
        target = (unsigned long long)fi;
        if (target > total - 1) target = total - 1;
      }
    } else {
      target = k_rem_in;
    }
    s_target = target; s_total = total;
  }
  __syncthreads();
  const unsigned long long target = s_target;
  unsigned long long cum = excl;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (h[j] != 0 && cum <= target && target < cum + h[j]) { s_bin = (unsigned)(t * 8 + j); s_rem = target - cum; }
    cum += h[j];
  }
  __syncthreads();
  *out_bin = s_bin; *out_rem = s_rem; *out_total = s_total;
  __syncthreads();                                       // the scratch is reused by the next call
}

// executed by all 256 threads of the last block of a histogram pass over the GLOBAL histogram (which it zeroes again)
__device__ void select_pick(DeviceState* st, unsigned int* hist, int pass, float ratio) {
  const int t = threadIdx.x;
  unsigned int h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { h[j] = __ldcg(&hist[t * 8 + j]); hist[t * 8 + j] = 0; }
  unsigned int bin; unsigned long long rem, total;
  block_pick(h, pass, ratio, st->k_rem, &bin, &rem, &total);
  if (t == 0) {
    if (pass == 1) {
      st->n_valid = total;
      if (total == 0) raise_status(st, AICP_B200_ERR_NO_VALID_MATCH);
    }
    st->k_rem = rem;
    if (pass == 1) st->prefix = bin;
    else if (pass == 2) st->prefix = (st->prefix << 11) | bin;
    else st->limit = __uint_as_float((st->prefix << 9) | bin);
  }
}

__device__ __forceinline__ bool block_is_last(unsigned int* ticket) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int t = atomicAdd(ticket, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *ticket = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

__device__ __forceinline__ void hist_flush(const unsigned int* sh, unsigned int* hist) {
  for (int b = threadIdx.x; b < AICP_HIST_BINS; b += blockDim.x)
    if (sh[b]) atomicAdd(&hist[b], sh[b]);
}

#ifdef AICP_DEBUG_WARP_TIMES
__device__ uint4 g_query_stats[131072];      // (climb levels | sibling descents << 16, nodes visited, points scanned, ns)
extern "C" int aicp_b200_debug_query_stats(unsigned int* out, int n_words) {
  return (int)cudaMemcpyFromSymbol(out, g_query_stats, sizeof(unsigned int) * (size_t)n_words);
}
#endif
// One query per thread: T_iter * reading' (float, fixed order) -> exact 1-NN, cold (root descent) in the first iteration,
// bottom-up from the previous match afterwards.  Writes (position, d2), the optional trace and the first radix-select
// digit into the block's shared histogram.  match_pos / d2out are read back later by the same launch of the persistent
// loop kernel, so they carry no __restrict__ / const (no non-coherent loads of data this kernel writes).
__device__ __forceinline__ void match_thread(const IndexView& ix, const float* sT, const float4* __restrict__ read0, int n, int i,
                                             int iter, int* match_pos, float* d2out, int* trace_idx, unsigned int* sh) {
  if (i >= n) return;
  float4 r = __ldg(&read0[i]);
  float3 p = xform_f(sT, r.x, r.y, r.z);
  int pos; float d;
#ifdef AICP_DEBUG_WARP_TIMES
  DbgCnt cnt{0, 0, 0, 0};
  DbgCnt* dbg = &cnt;
  const unsigned long long q0 = global_ns();
#endif
  if (iter > 0) nn_search_up(ix, p.x, p.y, p.z, match_pos[i], &pos, &d AICP_DBG_ARG);
  else nn_search(ix, p.x, p.y, p.z, &pos, &d AICP_DBG_ARG);
#ifdef AICP_DEBUG_WARP_TIMES
  if (iter == AICP_DEBUG_WARP_TIMES && i < 131072) {
    g_query_stats[i] = make_uint4((unsigned)cnt.levels | ((unsigned)cnt.descents << 16), (unsigned)cnt.nodes, (unsigned)cnt.points, (unsigned)(global_ns() - q0));
  }
#endif
  match_pos[i] = pos;
  d2out[i] = d;
  if (trace_idx) trace_idx[(size_t)iter * n + __float_as_int(r.w)] = __float_as_int(__ldg(&ix.pts[pos]).w);
  if (d2_valid(d)) atomicAdd(&sh[__float_as_uint(d) >> 20], 1u);
}

__global__ void __launch_bounds__(256) k_match(IndexView ix, const float4* __restrict__ read0, int n, DeviceState* st,
                                               int* match_pos, float* d2out,
                                               unsigned int* hist, int* trace_idx, float ratio, int tail,
                                               volatile int* progress) {
  if (ld_int(&st->done)) { publish_done(progress); return; }
  __shared__ unsigned int sh[AICP_HIST_BINS];
  __shared__ float sT[16];
  for (int b = threadIdx.x; b < AICP_HIST_BINS; b += blockDim.x) sh[b] = 0;
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T_iter[threadIdx.x];
  __syncthreads();
  match_thread(ix, sT, read0, n, blockIdx.x * blockDim.x + threadIdx.x, st->iter, match_pos, d2out, trace_idx, sh);
  __syncthreads();
  hist_flush(sh, hist);
  if (tail && block_is_last(&st->ticket[0])) {
    const unsigned long long t0 = global_ns();
    select_pick(st, hist, 1, ratio);
    if (threadIdx.x == 0) st->tail_ns[0] += global_ns() - t0;
  }
}

// k_match_tile: the same correspondence search with one WARP per tile of 32 consecutive (Morton-ordered) reading points,
// one query per lane, and ONE shared walk of the reference tree per tile.  The per-thread walk of k_match keeps only a
// third of the lanes busy (every lane follows its own path); here all lanes execute the same path: the tile climbs from
// a seed to the first ancestor whose Morton cell contains every lane's ball (p_i, best_i), then walks that subtree once,
// nearest child first, entering a child when ANY lane's ball reaches its box.  A leaf range is staged in shared memory
// with one coalesced load and every lane measures every point of it.  More distance evaluations than the per-thread
// walk, but no divergence and no per-lane stacks.  Same candidates can win, same (d2, id) order: identical result.
#define MATCH_CHUNK 32            // children with at most this many points are scanned, larger ones are entered (A/B: 8 -7 %, 16 -3 %)

// warp-wide: `i` is this lane's query (may be >= n in the last tile); s_pts / stack are the warp's staging area and stack
__device__ __forceinline__ void match_tile(const IndexView& ix, const float* sT, const float4* __restrict__ read0, int n, int i, int iter,
                                           int* match_pos, float* d2out, int* trace_idx, unsigned int* sh, float4* s_pts, int* stack) {
  const int lane = threadIdx.x & 31;
  const bool active = i < n;
  if (__ballot_sync(0xFFFFFFFFu, active) != 0u) {
    const float4 r = __ldg(&read0[active ? i : n - 1]);
    const float3 p = xform_f(sT, r.x, r.y, r.z);
    float bd = INFINITY; int bid = 0x7FFFFFFF, bpos = -1;
    int seed = -1;
    if (iter > 0 && active) {
      seed = match_pos[i];
      const float4 mp = __ldg(&ix.pts[seed]);
      bd = d2_f(p.x, p.y, p.z, mp.x, mp.y, mp.z); bid = __float_as_int(mp.w); bpos = seed;
    }
    // scan [first, first + cnt), cnt <= 32: every lane against its own query
    auto scan = [&](int first, int cnt) {
      if (lane < cnt) s_pts[lane] = __ldg(&ix.pts[first + lane]);
      __syncwarp();
#pragma unroll 4
      for (int j = 0; j < cnt; ++j) {
        const float4 c = s_pts[j];
        const float d = d2_f(p.x, p.y, p.z, c.x, c.y, c.z);
        const int id = __float_as_int(c.w);
        if (d < bd || (d == bd && id < bid)) { bd = d; bid = id; bpos = first + j; }
      }
      __syncwarp();
    };
    if (ix.n <= 32) {
      scan(0, ix.n);
    } else {
      int node = 0;
      if (iter > 0) {
        // box around the balls (p_i, best_i), radii rounded up and ends rounded outward; then climb from lane 0's seed to
        // the first ancestor whose (shrunk) Morton cell holds the box: nothing outside it can beat or tie any lane
        float blx = INFINITY, bly = INFINITY, blz = INFINITY, bhx = -INFINITY, bhy = -INFINITY, bhz = -INFINITY;
        if (active) {
          const float rr = __fsqrt_ru(bd);
          blx = __fsub_rd(p.x, rr); bly = __fsub_rd(p.y, rr); blz = __fsub_rd(p.z, rr);
          bhx = __fadd_ru(p.x, rr); bhy = __fadd_ru(p.y, rr); bhz = __fadd_ru(p.z, rr);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          blx = fminf(blx, __shfl_xor_sync(0xFFFFFFFFu, blx, off)); bhx = fmaxf(bhx, __shfl_xor_sync(0xFFFFFFFFu, bhx, off));
          bly = fminf(bly, __shfl_xor_sync(0xFFFFFFFFu, bly, off)); bhy = fmaxf(bhy, __shfl_xor_sync(0xFFFFFFFFu, bhy, off));
          blz = fminf(blz, __shfl_xor_sync(0xFFFFFFFFu, blz, off)); bhz = fmaxf(bhz, __shfl_xor_sync(0xFFFFFFFFu, bhz, off));
        }
        node = __ldg(&ix.owner32[__shfl_sync(0xFFFFFFFFu, seed, 0)]) >> 1;        // lane 0 is active in every live warp
        while (true) {
          const int up = __float_as_int(__ldg(&ix.rec[4 * (size_t)node + 3]).w);
          if (up < 0) break;
          const float4 clo = __ldg(&ix.cellbox[2 * (size_t)node]), chi = __ldg(&ix.cellbox[2 * (size_t)node + 1]);
          if (blx >= clo.x && bly >= clo.y && blz >= clo.z && bhx <= chi.x && bhy <= chi.y && bhz <= chi.z) break;
          node = up >> 2;
        }
      }
      int sp = 0, code = node;
      while (true) {
        const float4* rc = ix.rec + 4 * (size_t)code;                   // same address in every lane: broadcast
        const float4 r0 = __ldg(rc), r1 = __ldg(rc + 1), r2 = __ldg(rc + 2), r3 = __ldg(rc + 3);
        const int first = __float_as_int(r0.w), split = __float_as_int(r1.w), end = __float_as_int(r2.w);
        const float dl = box_d2_f(make_float3(r0.x, r0.y, r0.z), make_float3(r1.x, r1.y, r1.z), p.x, p.y, p.z);
        const float dr = box_d2_f(make_float3(r2.x, r2.y, r2.z), make_float3(r3.x, r3.y, r3.z), p.x, p.y, p.z);
        const unsigned int ml = __reduce_min_sync(0xFFFFFFFFu, active ? __float_as_uint(dl) : 0xFFFFFFFFu);
        const unsigned int mr = __reduce_min_sync(0xFFFFFFFFu, active ? __float_as_uint(dr) : 0xFFFFFFFFu);
        const bool lfirst = ml <= mr;
        int next = -1;
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
          const bool left = (pass == 0) == lfirst;
          const float dc = left ? dl : dr;
          if (__ballot_sync(0xFFFFFFFFu, active && dc <= bd) == 0u) continue;     // no lane's ball reaches this child
          const int cf = left ? first : split, cc = left ? split - first : end - split;
          if (cc <= MATCH_CHUNK) scan(cf, cc);
          else {
            const int child = left ? split - 1 : split;
            if (next < 0) next = child; else stack[sp++] = child;
          }
        }
        if (next >= 0) code = next;
        else if (sp > 0) code = stack[--sp];
        else break;
      }
    }
    if (active) {
      match_pos[i] = bpos;
      d2out[i] = bd;
      if (trace_idx) trace_idx[(size_t)iter * n + __float_as_int(r.w)] = bid;
      if (d2_valid(bd)) atomicAdd(&sh[__float_as_uint(bd) >> 20], 1u);
    }
  }
}

__global__ void __launch_bounds__(256) k_match_tile(IndexView ix, const float4* __restrict__ read0, int n, DeviceState* st,
                                                    int* match_pos, float* d2out,
                                                    unsigned int* hist, int* trace_idx, float ratio, int tail,
                                                    volatile int* progress) {
  if (ld_int(&st->done)) { publish_done(progress); return; }
  __shared__ unsigned int sh[AICP_HIST_BINS];
  __shared__ float sT[16];
  __shared__ float4 s_stage[8][32];
  __shared__ int s_stack[8][AICP_STACK];
  for (int b = threadIdx.x; b < AICP_HIST_BINS; b += blockDim.x) sh[b] = 0;
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T_iter[threadIdx.x];
  __syncthreads();
  const int w = threadIdx.x >> 5;
  match_tile(ix, sT, read0, n, blockIdx.x * blockDim.x + threadIdx.x, st->iter, match_pos, d2out, trace_idx, sh, s_stage[w], s_stack[w]);
  __syncthreads();
  hist_flush(sh, hist);
  if (tail && block_is_last(&st->ticket[0])) {
    const unsigned long long t0 = global_ns();
    select_pick(st, hist, 1, ratio);
    if (threadIdx.x == 0) st->tail_ns[0] += global_ns() - t0;
  }
}

// Digits 2 and 3 of the trimmed quantile in ONE launch.  Grid: every d2 key whose first digit is the one k_match picked
// (typically a few thousand of the n keys) is appended to `cand` (warp-aggregated atomics).  Last block: radix select of
// the remaining 20 bits over the candidate list with shared-memory histograms -> st->limit, the exact k-th smallest.
__global__ void __launch_bounds__(256) k_select23(const float* __restrict__ d2, int n, DeviceState* st, unsigned int* cand,
                                                  volatile int* progress) {
  if (ld_int(&st->done)) { publish_done(progress); return; }
  __shared__ unsigned int sh[AICP_HIST_BINS];
  const unsigned int prefix = st->prefix;
  const int lane = threadIdx.x & 31;
  for (int base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {     // warp-uniform trip count
    const int i = base + threadIdx.x;
    unsigned int key = 0;
    bool hit = false;
    if (i < n) {
      float d = __ldg(&d2[i]);
      key = __float_as_uint(d);
      hit = d2_valid(d) && (key >> 20) == prefix;
    }
    const unsigned int m = __ballot_sync(0xFFFFFFFFu, hit);
    if (m) {
      const int leader = __ffs(m) - 1;
      unsigned int off = 0;
      if (lane == leader) off = atomicAdd(&st->cand_n, (unsigned)__popc(m));
      off = __shfl_sync(0xFFFFFFFFu, off, leader);
      if (hit) cand[off + __popc(m & ((1u << lane) - 1u))] = key;
    }
  }
  if (!block_is_last(&st->ticket[1])) return;
  const unsigned long long t0 = global_ns();
  const int t = threadIdx.x;
  const unsigned int c = *(volatile unsigned int*)&st->cand_n;
  unsigned int h[8], bin; unsigned long long rem, total;
  // digit 2
  for (int b = t; b < AICP_HIST_BINS; b += 256) sh[b] = 0;
  __syncthreads();
  for (unsigned int j = t; j < c; j += 256) atomicAdd(&sh[(__ldcg(&cand[j]) >> 9) & 2047u], 1u);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = sh[t * 8 + j];
  block_pick(h, 2, 0.f, st->k_rem, &bin, &rem, &total);
  const unsigned int prefix22 = (prefix << 11) | bin;
  // digit 3
  for (int b = t; b < AICP_HIST_BINS; b += 256) sh[b] = 0;
  __syncthreads();
  for (unsigned int j = t; j < c; j += 256) {
    unsigned int key = __ldcg(&cand[j]);
    if ((key >> 9) == prefix22) atomicAdd(&sh[key & 511u], 1u);
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = sh[t * 8 + j];
  block_pick(h, 3, 0.f, rem, &bin, &rem, &total);
  if (t == 0) {
    st->prefix = prefix22;
    st->k_rem = rem;
    st->limit = __uint_as_float((prefix22 << 9) | bin);
    st->cand_n = 0;
    st->tail_ns[1] += global_ns() - t0;
  }
}

// pass 1: plain histogram of digit 1 (stage entry point); pass 2 / 3: next digits among keys matching the prefix
__global__ void __launch_bounds__(256) k_select(const float* __restrict__ d2, int n, DeviceState* st, unsigned int* hist,
                                                int pass, float ratio, int tail) {
  if (ld_int(&st->done)) return;
  __shared__ unsigned int sh[AICP_HIST_BINS];
  for (int b = threadIdx.x; b < AICP_HIST_BINS; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  const unsigned int prefix = st->prefix;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float d = __ldg(&d2[i]);
    if (!d2_valid(d)) continue;
    unsigned int key = __float_as_uint(d);
    if (pass == 1) atomicAdd(&sh[key >> 20], 1u);
    else if (pass == 2) { if ((key >> 20) == prefix) atomicAdd(&sh[(key >> 9) & 2047u], 1u); }
    else { if ((key >> 9) == prefix) atomicAdd(&sh[key & 511u], 1u); }
  }
  __syncthreads();
  hist_flush(sh, hist);
  if (tail && block_is_last(&st->ticket[1])) select_pick(st, hist, pass, ratio);
}

// sharded registration: the histogram is all-reduced over the ranks first, then every rank picks the same digit
__global__ void __launch_bounds__(256) k_pick(DeviceState* st, unsigned int* hist, int pass, float ratio) {
  if (ld_int(&st->done)) {
    for (int b = threadIdx.x; b < AICP_HIST_BINS; b += blockDim.x) hist[b] = 0;
    return;
  }
  select_pick(st, hist, pass, ratio);
}

// ---- normal equations ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long fixed_term(float a, float b) {
  return __double2ll_rn(((double)a * (double)b) * AICP_FIXED_SCALE);
}

__device__ __forceinline__ void atomic_add_128(unsigned long long* lo, long long* hi, long long v) {
  unsigned long long vlo = (unsigned long long)v;
  unsigned long long old = atomicAdd(lo, vlo);
  unsigned long long carry = (old + vlo < old) ? 1ull : 0ull;
  unsigned long long vhi = (v < 0 ? ~0ull : 0ull) + carry;
  if (vhi) atomicAdd((unsigned long long*)hi, vhi);
}

// A.5 solve + pose update, A.7 checkers.  Called by ALL 32 lanes of one warp: the sums are converted and the 6x6 system
// solved cooperatively (det_solve6_warp), the pose update and the checkers -- short scalar chains -- run in lane 0.
// lane s < 28 passes sum s as (lo, hi)
__device__ __noinline__ void solve_and_check(DeviceState* st, const LoopParams lp, int n_read, unsigned long long lo, long long hi) {
  const int lane = threadIdx.x & 31;
  double v = 0.0;
  long long n_used = 0;
  if (lane < AICP_NSUM) {
    v = fixed128_to_double(hi, lo);
    n_used = (long long)lo;
  }
  n_used = __shfl_sync(0xFFFFFFFFu, n_used, 27);
  double x[6];
  det_solve6_warp(v, x);
  if (lane != 0) return;
  float dT[16], T[16];
  det_pose_increment(x, dT);
  for (int i = 0; i < 16; ++i) T[i] = st->T_iter[i];
  mat4_mul_f(dT, T, T);                                    // T_iter = dT * T_iter
  for (int i = 0; i < 16; ++i) st->T_iter[i] = T[i];
  int it = st->iter;
  aicp_b200_iter_trace* tr = &st->trace[it];
  for (int i = 0; i < 16; ++i) tr->T_iter[i] = T[i];
  tr->limit_d2 = st->limit; tr->n_valid = (long long)st->n_valid; tr->n_used = n_used;
  tr->rot_err = __longlong_as_double(0x7FF8000000000000ll); tr->trans_err = tr->rot_err;
  st->n_used_last = n_used;
  ++it;
  st->iter = it;
  bool has_nan = false;
  for (int i = 0; i < 16; ++i) if (T[i] != T[i]) has_nan = true;
  if (has_nan) { raise_status(st, AICP_B200_ERR_NAN); return; }
  bool iterate = true;
  if (it >= lp.max_iterations) { iterate = false; st->stop_reason = AICP_B200_STOP_COUNTER; }
  int hn = st->hist_n;
  det_quat_from_T(T, st->quat_hist[hn]);
  st->tr_hist[hn][0] = (double)T[12]; st->tr_hist[hn][1] = (double)T[13]; st->tr_hist[hn][2] = (double)T[14];
  {
    // the newest consecutive pair; older pairs were computed by earlier iterations (same values, same order of summation)
    st->ang_step[hn] = det_quat_angular_distance(st->quat_hist[hn], st->quat_hist[hn - 1]);
    double dx = st->tr_hist[hn][0] - st->tr_hist[hn - 1][0], dy = st->tr_hist[hn][1] - st->tr_hist[hn - 1][1],
           dz = st->tr_hist[hn][2] - st->tr_hist[hn - 1][2];
    st->trn_step[hn] = sqrt((dx * dx + dy * dy) + dz * dz);
  }
  ++hn;
  st->hist_n = hn;
  if (hn > lp.smooth_length) {
    double re = 0.0, te = 0.0;
    for (int i = hn - 1; i >= hn - lp.smooth_length; --i) {
      re = re + st->ang_step[i];
      te = te + st->trn_step[i];
    }
    re = re / (double)lp.smooth_length;
    te = te / (double)lp.smooth_length;
    tr->rot_err = re; tr->trans_err = te;
    if (re < (double)lp.min_diff_rot && te < (double)lp.min_diff_trans) {
      if (iterate) st->stop_reason = AICP_B200_STOP_DIFFERENTIAL;
      iterate = false;
    }
  }
  if (!iterate) *(volatile int*)&st->done = 1;
}

// llrint(a * b * 2^30) of two floats: the product and the scaling are exact in double; below 2^51 the rounding to an
// integer is done by ONE double addition (x + 1.5 * 2^52 lands in the binade where the spacing is 1, rounding to nearest
// even as cvt.rni does), which replaces the slow F2I.S64.F64 of __double2ll_rn -- same bits
__device__ __forceinline__ long long fixed_term_d(double a, double b, bool fast) {
  const double x = (a * b) * AICP_FIXED_SCALE;
  if (fast) return __double_as_longlong(x + 6755399441055744.0) - 0x4338000000000000ll;
  return __double2ll_rn(x);
}

// warp sum of a 64-bit term through three 32-bit redux.sync (signed high word, two 16-bit halves of the low word); the
// total is added to lane `slot`'s accumulator -- one register pair per lane holds the warp's 28 sums
__device__ __forceinline__ void warp_add_term(long long term, int slot, int lane, long long& acc) {
  const int hi = (int)(term >> 32);
  const unsigned int lo = (unsigned int)term;
  const int s_hi = __reduce_add_sync(0xFFFFFFFFu, hi);
  const unsigned int s_l1 = __reduce_add_sync(0xFFFFFFFFu, lo >> 16);
  const unsigned int s_l0 = __reduce_add_sync(0xFFFFFFFFu, lo & 0xFFFFu);
  const long long tot = ((long long)s_hi << 32) + ((long long)s_l1 << 16) + (long long)s_l0;
  if (lane == slot) acc += tot;
}

// Warp-wide: the 28 exact fixed-point terms of this lane's correspondence -- F F^T (21), F res (6), 1 -- with F = [p x n; n],
// res = (p - q) . n (A.5) are summed over the warp and added to the lane-indexed accumulators (lane s holds slot s).
// `in` false (outlier, padding): the lane contributes zeros.  Shared by k_accumulate and the persistent loop kernel.
__device__ __forceinline__ void acc_point_terms(bool in, const float4& r, const float4& q, const float4& nr, const float* sT, int lane,
                                                long long& acc) {
  if (!__any_sync(0xFFFFFFFFu, in)) return;
  double F[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  bool small = true;
  if (in) {
    const float3 p = xform_f(sT, r.x, r.y, r.z);
    const float c0 = __fsub_rn(__fmul_rn(p.y, nr.z), __fmul_rn(p.z, nr.y));      // c = p x n
    const float c1 = __fsub_rn(__fmul_rn(p.z, nr.x), __fmul_rn(p.x, nr.z));
    const float c2 = __fsub_rn(__fmul_rn(p.x, nr.y), __fmul_rn(p.y, nr.x));
    const float ddx = __fsub_rn(p.x, q.x), ddy = __fsub_rn(p.y, q.y), ddz = __fsub_rn(p.z, q.z);
    float res = __fmul_rn(ddx, nr.x);
    res = __fadd_rn(res, __fmul_rn(ddy, nr.y));
    res = __fadd_rn(res, __fmul_rn(ddz, nr.z));
    // every |factor| < 1448 = sqrt(2^21): all products stay below 2^51 after scaling by 2^30
    small = fabsf(c0) < 1448.f && fabsf(c1) < 1448.f && fabsf(c2) < 1448.f && fabsf(nr.x) < 1448.f && fabsf(nr.y) < 1448.f &&
            fabsf(nr.z) < 1448.f && fabsf(res) < 1448.f;
    F[0] = c0; F[1] = c1; F[2] = c2; F[3] = nr.x; F[4] = nr.y; F[5] = nr.z; F[6] = res;
  }
  const bool fast = __all_sync(0xFFFFFFFFu, small);
  int s = 0;
#pragma unroll
  for (int x = 0; x < 6; ++x)
#pragma unroll
    for (int y = x; y < 6; ++y) { warp_add_term(fixed_term_d(F[x], F[y], fast), s, lane, acc); ++s; }
#pragma unroll
  for (int x = 0; x < 6; ++x) { warp_add_term(fixed_term_d(F[x], F[6], fast), s, lane, acc); ++s; }
  const unsigned int cnt = __popc(__ballot_sync(0xFFFFFFFFu, in));
  if (lane == 27) acc += cnt;
}

#define ACC_SLOTS 64          // partial-sum slots: block b adds into slot b % ACC_SLOTS, the last block folds the slots
#define ACC_PTS 2             // reading points per thread

__global__ void __launch_bounds__(256, 2) k_accumulate(const float4* __restrict__ refc, const float4* __restrict__ normals,
                                                    const float4* __restrict__ read0, const int* __restrict__ match_pos,
                                                    const float* __restrict__ d2, int n, DeviceState* st, LoopParams lp, int tail,
                                                    volatile int* progress, unsigned long long* slots) {
  if (ld_int(&st->done)) { publish_done(progress); return; }
  __shared__ float sT[16];
  __shared__ long long s_part[8][32];
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T_iter[threadIdx.x];
  __syncthreads();
  const float limit = st->limit;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // all independent loads first (distance, reading point, match), then the two gathers of the inliers: two memory
  // round trips per thread instead of four
  float d[ACC_PTS]; float4 r[ACC_PTS]; int pos[ACC_PTS]; bool in[ACC_PTS];
#pragma unroll
  for (int u = 0; u < ACC_PTS; ++u) {
    const int i = blockIdx.x * (256 * ACC_PTS) + u * 256 + threadIdx.x;
    d[u] = INFINITY; pos[u] = 0; r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) { d[u] = __ldg(&d2[i]); pos[u] = __ldg(&match_pos[i]); r[u] = __ldg(&read0[i]); }
  }
  float4 q[ACC_PTS], nr[ACC_PTS];
#pragma unroll
  for (int u = 0; u < ACC_PTS; ++u) {
    in[u] = d[u] <= limit;                                     // TrimmedDist weight (A.4); false for NaN and padding
    q[u] = nr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in[u]) { q[u] = __ldg(&refc[pos[u]]); nr[u] = __ldg(&normals[pos[u]]); }
  }
  // exact warp sums through redux.sync, one 64-bit accumulator per lane (lane s = slot s): 48 registers instead of the 128
  // of the transposed-shuffle reduction over 32 accumulators per thread that this replaces
  long long acc = 0;
#pragma unroll
  for (int u = 0; u < ACC_PTS; ++u) acc_point_terms(in[u], r[u], q[u], nr[u], sT, lane, acc);
  s_part[w][lane] = acc;
  __syncthreads();
  if (w == 0) {
    // |term| < 2^52.6 and a block holds 512 points, so the block total fits in 64 bits; the slots are 128-bit
    long long tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += s_part[k][lane];
    unsigned long long* slot = slots + ((size_t)(blockIdx.x % ACC_SLOTS) * 32 + lane) * 2;
    if (lane < AICP_NSUM && tot != 0) atomic_add_128(slot, (long long*)(slot + 1), tot);
  }
  if (!block_is_last(&st->ticket[2])) return;
  // last block: fold the slots (and zero them for the next launch) into st->sum_*, then solve
  const unsigned long long t0 = global_ns();
  __shared__ unsigned long long s_lo[8][32];
  __shared__ long long s_hi[8][32];
  {
    const int nslots = gridDim.x < ACC_SLOTS ? (int)gridDim.x : ACC_SLOTS;
    unsigned long long lo = 0; long long hi = 0;
    if (lane < AICP_NSUM) {
      for (int g = w; g < nslots; g += 8) {
        unsigned long long* slot = slots + ((size_t)g * 32 + lane) * 2;
        unsigned long long l = __ldcg(slot); long long hh = (long long)__ldcg(slot + 1);
        slot[0] = 0; slot[1] = 0;
        unsigned long long nl = lo + l;
        hi = hi + hh + (nl < lo ? 1 : 0);
        lo = nl;
      }
    }
    s_lo[w][lane] = lo; s_hi[w][lane] = hi;
  }
  __syncthreads();
  if (w == 0 && lane < AICP_NSUM) {
    unsigned long long lo = 0; long long hi = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned long long nl = lo + s_lo[k][lane];
      hi = hi + s_hi[k][lane] + (nl < lo ? 1 : 0);
      lo = nl;
    }
    st->sum_lo[lane] = lo; st->sum_hi[lane] = hi;
  }
  __threadfence();
  __syncthreads();
  if (tail && threadIdx.x < 32) {
    const bool has = lane < AICP_NSUM;
    const unsigned long long slo = has ? __ldcg(&st->sum_lo[lane]) : 0ull;
    const long long shi = has ? __ldcg(&st->sum_hi[lane]) : 0ll;
    if (has) { st->sum_lo[lane] = 0; st->sum_hi[lane] = 0; }
    solve_and_check(st, lp, n, slo, shi);
    if (threadIdx.x == 0) {
      st->tail_ns[2] += global_ns() - t0;
      publish_progress(progress, st->iter, *(volatile int*)&st->done);
    }
  }
}

// ---- the persistent loop kernel -----------------------------------------------------------------------------------------
// ONE cooperative launch runs the whole ICP loop: search -> trimmed quantile -> normal equations -> solve -> checkers,
// iteration after iteration, until a checker (or a status) ends it -- no launch gaps, no host polling, and loop control
// that never leaves the device.  The three phases of an iteration are the parallel parts of k_match* / k_select23 /
// k_accumulate; what those kernels do in their LAST block (digit pick, digits 2+3, fold + solve) is the serial section of a
// grid barrier here: the last block to arrive runs it and only then releases the others.  Three barriers per iteration
// (measured on B200: 1.2 - 2.3 us each for 74 - 592 blocks; cooperative kernels of different streams do run concurrently,
// profiles/round2_a_coop_probe.txt).  Reading points are handled in tiles of 256 (tile t -> block t mod gridDim), the
// same thread touches the same points in every phase.
//
// Sharded registration (reading split over GPUs, comm.cu): the serial sections also exchange with the peers -- the
// digit-1 histogram, the candidate keys of the picked bin, the 28 partial sums -- by storing straight into every peer's
// inbox over NVLink (peer-mapped memory) and spinning on sequence-stamped flags: three exchanges per iteration inside the
// kernel instead of four ncclAllReduce launches between nine kernels.
// Spin loads are RELAXED: an acquire load is a load + fence, and a gpu-scope fence invalidates the SM's whole L1
// (MEMBAR + CCTL.IVALL in SASS) -- a block polling with ld.acquire wipes the L1 under the blocks that are still searching
// (measured: the search phase of the loop kernel 70 us instead of 59).  One fence follows the spin instead.
__device__ __forceinline__ unsigned int ld_relaxed_gpu_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int atom_add_release_gpu_u32(unsigned int* p, unsigned int v) {
  unsigned int old;
  asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// grid barrier with a serial section: every block arrives; the last one runs fn() (all its threads) and then releases
template <typename F>
__device__ __forceinline__ void grid_serial(DeviceState* st, unsigned int& epoch, F&& fn) {
  __shared__ bool s_last;
  ++epoch;
  __syncthreads();
  if (threadIdx.x == 0) {
    // release only (the block's writes, ordered before this by the barrier above, become visible before the count): no L1
    // invalidation while other blocks of this SM are still at work
    const unsigned int t = atom_add_release_gpu_u32(&st->bar_arrive, 1u);
    s_last = (t == epoch * gridDim.x - 1u);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();                                 // acquire side of every other block's arrival
    fn();
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu_u32(&st->bar_release, epoch);
  } else if (threadIdx.x == 0) {
    while (ld_relaxed_gpu_u32(&st->bar_release) < epoch) {}
    __threadfence();
  }
  __syncthreads();
}

// ---- peer exchange (sharded registration): flag-in-data words, see common.cuh
__device__ __forceinline__ unsigned int peer_stamp(const PeerView& pv, int iter, int round) {
  return (unsigned int)(pv.epoch * 2048ull) + (unsigned int)(iter * 4 + round + 1);
}
__device__ __forceinline__ void ll_store(unsigned long long* p, unsigned int payload, unsigned int stamp) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(((unsigned long long)stamp << 32) | payload) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// spin until word *p carries `stamp`; false when the time limit passes first
__device__ __forceinline__ bool ll_wait(const unsigned long long* p, unsigned int stamp, unsigned int* payload) {
  unsigned long long v = ll_load(p);
  if ((unsigned int)(v >> 32) != stamp) {
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while ((unsigned int)((v = ll_load(p)) >> 32) != stamp)
      if ((++spins & 1023u) == 0 && global_ns() - t0 > AICP_PEER_TIMEOUT_NS) return false;
  }
  *payload = (unsigned int)v;
  return true;
}
// block-wide verdict of an exchange: a peer reported a status, or some word never arrived -> every rank ends the loop
// (AICP_B200_ERR_COMM; the rank that raised the original status keeps its own code)
__device__ __forceinline__ void peer_verdict(DeviceState* st, int bad, unsigned long long t0) {
  bad = __syncthreads_or(bad);
  if (threadIdx.x == 0) {
    st->phase_ns[3] += global_ns() - t0;
    if (bad) raise_status(st, AICP_B200_ERR_COMM);
  }
}

// serial section 1: the digit-1 histogram is complete in `hist` (and, sharded, is summed over the ranks) -> pick the digit
__device__ void loop_pick(DeviceState* st, unsigned int* hist, float ratio, const PeerView& pv, int iter) {
  const int t = threadIdx.x;
  unsigned int h[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { h[j] = __ldcg(&hist[t * 8 + j]); hist[t * 8 + j] = 0; }
  if (pv.n_ranks > 1) {
    const unsigned int stamp = peer_stamp(pv, iter, 0);
    const unsigned int my_status = *(volatile int*)&st->status != 0 ? 1u : 0u;
    for (int r = 0; r < pv.n_ranks; ++r) {
      if (r == pv.rank) continue;
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(pv.inbox[r] + AICP_INBOX_HIST_OFF + (size_t)pv.rank * AICP_INBOX_HIST_STRIDE);
      // word j * 256 + t carries bin 8 t + j: the 32 lanes of a warp store 256 contiguous bytes (the layout bin -> word 8 t + j
      // made every lane's 8-byte store its own NVLink packet, 2048 packets per peer: 20 us per exchange at 8 ranks)
#pragma unroll
      for (int j = 0; j < 8; ++j) ll_store(dst + j * 256 + t, h[j], stamp);
      if (t == 0) ll_store(dst + AICP_HIST_BINS, my_status, stamp);
    }
    const unsigned long long t0 = global_ns();
    int bad = 0;
    for (int r = 0; r < pv.n_ranks; ++r) {
      if (r == pv.rank) continue;
      const unsigned long long* in = reinterpret_cast<const unsigned long long*>(pv.inbox[pv.rank] + AICP_INBOX_HIST_OFF + (size_t)r * AICP_INBOX_HIST_STRIDE);
      unsigned long long v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = ll_load(in + j * 256 + t);            // all eight in flight; late words are re-polled below
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        unsigned int x = (unsigned int)v[j];
        if ((unsigned int)(v[j] >> 32) != stamp && !ll_wait(in + j * 256 + t, stamp, &x)) bad = 1;
        h[j] += x;
      }
      if (t == 0) { unsigned int x = 0; if (!ll_wait(in + AICP_HIST_BINS, stamp, &x) || x) bad = 1; }
    }
    peer_verdict(st, bad, t0);
  }
  unsigned int bin; unsigned long long rem, total;
  block_pick(h, 1, ratio, 0, &bin, &rem, &total);
  if (t == 0) {
    st->n_valid = total;
    if (total == 0) raise_status(st, AICP_B200_ERR_NO_VALID_MATCH);
    st->k_rem = rem;
    st->prefix = bin;
  }
}

// serial section 2: digits 2 and 3 over the candidate keys of the picked bin (sharded: over every rank's list) -> st->limit
__device__ void loop_select23(DeviceState* st, unsigned int* cand, unsigned int* sh, const PeerView& pv, int iter) {
  const int t = threadIdx.x;
  const unsigned int c_local = *(volatile unsigned int*)&st->cand_n;
  const unsigned int prefix = *(volatile unsigned int*)&st->prefix;
  const unsigned int stamp = peer_stamp(pv, iter, 1);
  __shared__ unsigned int s_cnt[AICP_MAX_RANKS];
  if (pv.n_ranks > 1) {
    // more candidates than a peer's inbox holds (a shard of over cand_cap points whose distances share one 11-bit digit):
    // raised as a status, which the header carries to every rank, so all of them end the loop together
    const bool fits = c_local <= pv.cand_cap;
    if (!fits && t == 0) raise_status(st, AICP_B200_ERR_COMM);
    const unsigned int my_status = (!fits || *(volatile int*)&st->status != 0) ? 1u : 0u;
    for (int r = 0; r < pv.n_ranks; ++r) {
      if (r == pv.rank) continue;
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(pv.inbox[r] + AICP_INBOX_CAND_OFF + (size_t)pv.rank * pv.cand_stride);
      if (t == 0) ll_store(dst, (fits ? c_local : 0u) | (my_status << 31), stamp);
      for (unsigned int j = t; j < c_local && fits; j += 256) ll_store(dst + 1 + j, __ldcg(&cand[j]), stamp);
    }
    const unsigned long long t0 = global_ns();
    int bad = 0;
    if (t < pv.n_ranks && t != pv.rank) {
      unsigned int hdr = 0;
      if (!ll_wait(reinterpret_cast<const unsigned long long*>(pv.inbox[pv.rank] + AICP_INBOX_CAND_OFF + (size_t)t * pv.cand_stride), stamp, &hdr) ||
          (hdr >> 31)) bad = 1;
      s_cnt[t] = bad ? 0u : (hdr & 0x7FFFFFFFu);
    }
    peer_verdict(st, bad, t0);
  }
  // visit every key of every list: the local one, then each peer's copy in this rank's inbox (words polled until stamped)
  auto for_each_key = [&](auto&& fn) {
    for (unsigned int j = t; j < c_local; j += 256) fn(__ldcg(&cand[j]));
    for (int r = 0; r < pv.n_ranks; ++r) {
      if (r == pv.rank || pv.n_ranks == 1) continue;
      const unsigned long long* in = reinterpret_cast<const unsigned long long*>(pv.inbox[pv.rank] + AICP_INBOX_CAND_OFF + (size_t)r * pv.cand_stride) + 1;
      const unsigned int c = s_cnt[r];
      // four words per thread in flight (a polled word is an uncached L2 round trip: one at a time costs 0.7 us each)
      for (unsigned int j = t; j < c; j += 1024) {
        unsigned long long v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = j + 256 * u < c ? ll_load(in + j + 256 * u) : 0ull;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (j + 256 * u >= c) continue;
          unsigned int key = (unsigned int)v[u];
          if ((unsigned int)(v[u] >> 32) != stamp && !ll_wait(in + j + 256 * u, stamp, &key)) { raise_status(st, AICP_B200_ERR_COMM); continue; }
          fn(key);
        }
      }
    }
  };
  unsigned int h[8], bin; unsigned long long rem, total;
  for (int b = t; b < AICP_HIST_BINS; b += 256) sh[b] = 0;
  __syncthreads();
  for_each_key([&](unsigned int key) { atomicAdd(&sh[(key >> 9) & 2047u], 1u); });
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = sh[t * 8 + j];
  block_pick(h, 2, 0.f, *(volatile unsigned long long*)&st->k_rem, &bin, &rem, &total);
  const unsigned int prefix22 = (prefix << 11) | bin;
  for (int b = t; b < AICP_HIST_BINS; b += 256) sh[b] = 0;
  __syncthreads();
  for_each_key([&](unsigned int key) { if ((key >> 9) == prefix22) atomicAdd(&sh[key & 511u], 1u); });
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) h[j] = sh[t * 8 + j];
  block_pick(h, 3, 0.f, rem, &bin, &rem, &total);
  if (t == 0) {
    st->prefix = prefix22;
    st->k_rem = rem;
    st->limit = __uint_as_float((prefix22 << 9) | bin);
    st->cand_n = 0;
  }
}

// serial section 3: fold the partial-sum slots (zeroing them), add the peers' sums, solve, update T_iter, run the checkers
__device__ void loop_fold_solve(DeviceState* st, unsigned long long* slots, const LoopParams& lp, int n, const PeerView& pv, int iter,
                                unsigned long long (*s_lo)[32], long long (*s_hi)[32]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  {
    const int nslots = gridDim.x < ACC_SLOTS ? (int)gridDim.x : ACC_SLOTS;
    unsigned long long lo = 0; long long hi = 0;
    if (lane < AICP_NSUM) {
      for (int g = w; g < nslots; g += 8) {
        unsigned long long* slot = slots + ((size_t)g * 32 + lane) * 2;
        unsigned long long l = __ldcg(slot); long long hh = (long long)__ldcg(slot + 1);
        slot[0] = 0; slot[1] = 0;
        unsigned long long nl = lo + l;
        hi = hi + hh + (nl < lo ? 1 : 0);
        lo = nl;
      }
    }
    s_lo[w][lane] = lo; s_hi[w][lane] = hi;
  }
  __syncthreads();
  unsigned long long lo = 0; long long hi = 0;
  if (w == 0 && lane < AICP_NSUM) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      unsigned long long nl = lo + s_lo[k][lane];
      hi = hi + s_hi[k][lane] + (nl < lo ? 1 : 0);
      lo = nl;
    }
  }
  if (pv.n_ranks > 1) {
    const unsigned int stamp = peer_stamp(pv, iter, 2);
    if (w == 0) {
      // this rank's 28 sums as four 32-bit limbs each, its share of the reading (lane 28) and its status (lane 31)
      unsigned int limb[4] = {(unsigned int)lo, (unsigned int)(lo >> 32), (unsigned int)hi, (unsigned int)((unsigned long long)hi >> 32)};
      if (lane == AICP_NSUM) { limb[0] = (unsigned int)n; limb[1] = 0; limb[2] = 0; limb[3] = 0; }
      if (lane == 31) { limb[0] = 0; limb[1] = 0; limb[2] = 0; limb[3] = *(volatile int*)&st->status != 0 ? 1u : 0u; }
      if (lane <= AICP_NSUM || lane == 31) {
        for (int r = 0; r < pv.n_ranks; ++r) {
          if (r == pv.rank) continue;
          unsigned long long* dst = reinterpret_cast<unsigned long long*>(pv.inbox[r] + AICP_INBOX_SUMS_OFF + (size_t)pv.rank * AICP_INBOX_SUMS_STRIDE) + lane;
#pragma unroll
          for (int k = 0; k < 4; ++k) ll_store(dst + 32 * k, limb[k], stamp);          // limb k of lane's sum at word 32 k + lane: coalesced
        }
      }
    }
    const unsigned long long t0 = global_ns();
    int bad = 0;
    unsigned long long n_all = (unsigned long long)n;
    if (w == 0 && (lane <= AICP_NSUM || lane == 31)) {
      for (int r = 0; r < pv.n_ranks; ++r) {
        if (r == pv.rank) continue;
        const unsigned long long* in = reinterpret_cast<const unsigned long long*>(pv.inbox[pv.rank] + AICP_INBOX_SUMS_OFF + (size_t)r * AICP_INBOX_SUMS_STRIDE) + lane;
        unsigned int x[4] = {0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < 4; ++k) if (!ll_wait(in + 32 * k, stamp, &x[k])) bad = 1;
        if (lane < AICP_NSUM) {
          const unsigned long long plo = (unsigned long long)x[0] | ((unsigned long long)x[1] << 32);
          const long long phi = (long long)((unsigned long long)x[2] | ((unsigned long long)x[3] << 32));
          const unsigned long long nl = lo + plo;
          hi = hi + phi + (nl < lo ? 1 : 0);
          lo = nl;
        } else if (lane == AICP_NSUM) n_all += x[0];
        else if (x[3]) bad = 1;
      }
      if (lane == AICP_NSUM) st->n_read_total = n_all;
    }
    peer_verdict(st, bad, t0);
  }
  // the sums go from the fold to the solve in registers (lane s of warp 0 holds sum s)
  if (w == 0 && !ld_int(&st->done)) solve_and_check(st, lp, n, lo, hi);
}

struct LoopArgs {
  IndexView ix;                  // centred reference index
  const float4* normals;
  const float4* read0;           // T_refMean_dataIn * reading, Morton order
  int n;
  DeviceState* st;
  int* match_pos;
  float* d2;
  unsigned int* hist;
  unsigned int* cand;
  int* trace_idx;
  unsigned long long* slots;
  LoopParams lp;
  PeerView pv;
  int spread;                    // search phase: one query per `spread` lanes (1, 2, 4, 8), see launch_loop
};

// phase 2 of an iteration: the d2 keys whose first digit is the picked one are appended to the candidate list.  A block that
// owns several tiles (the 74-block launches of batch workers: 7 tiles each) reads four of them before it touches the first:
// the phase is a chain of L2 round trips otherwise (57 us alone for 131 072 points, profiles/round2_g_ncu_table_batch_schedule.txt)
__device__ __forceinline__ void loop_candidates(const LoopArgs& a, int Q, int n_tiles) {
  DeviceState* st = a.st;
  const int tid = threadIdx.x, lane = tid & 31, n = a.n;
  const unsigned int prefix = __ldcg(&st->prefix);
  for (int t0 = blockIdx.x; t0 < n_tiles; t0 += 4 * gridDim.x) {
    float d[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = t0 + u * gridDim.x;
      const int i = (t < n_tiles && tid < Q) ? t * Q + tid : n;
      d[u] = i < n ? __ldcg(&a.d2[i]) : __int_as_float(0x7FC00000);          // NaN: never a candidate
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (t0 + u * (int)gridDim.x >= n_tiles) break;                          // block-uniform
      const unsigned int key = __float_as_uint(d[u]);
      const bool hit = d2_valid(d[u]) && (key >> 20) == prefix;
      const unsigned int m = __ballot_sync(0xFFFFFFFFu, hit);
      if (m) {
        const int leader = __ffs(m) - 1;
        unsigned int off = 0;
        if (lane == leader) off = atomicAdd(&st->cand_n, (unsigned)__popc(m));
        off = __shfl_sync(0xFFFFFFFFu, off, leader);
        if (hit) a.cand[off + __popc(m & ((1u << lane) - 1u))] = key;
      }
    }
  }
}

// phase 3 of an iteration: exact fixed-point normal equations of the inliers (d2 <= limit) into the 128-bit slots; two tiles
// of a block are in flight at a time (their five loads per point each)
__device__ __forceinline__ void loop_accumulate(const LoopArgs& a, int Q, int n_tiles, const float* sT, long long (*s_hi)[32]) {
  DeviceState* st = a.st;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, n = a.n;
  const float limit = __ldcg(&st->limit);
  long long acc = 0;
  int held = 0;                                  // tiles folded into acc since the last flush
  auto flush = [&] {
    // |term| < 2^52.6 and at most 4 tiles (1024 points) per flush: the block total fits in 64 bits; the slots are 128-bit
    s_hi[w][lane] = acc;
    __syncthreads();
    if (w == 0) {
      long long tot = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += s_hi[k][lane];
      unsigned long long* slot = a.slots + ((size_t)(blockIdx.x % ACC_SLOTS) * 32 + lane) * 2;
      if (lane < AICP_NSUM && tot != 0) atomic_add_128(slot, (long long*)(slot + 1), tot);
    }
    __syncthreads();
    acc = 0; held = 0;
  };
  for (int t0 = blockIdx.x; t0 < n_tiles; t0 += 2 * gridDim.x) {
    bool in[2]; int pos[2]; int idx[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int t = t0 + u * gridDim.x;
      idx[u] = (t < n_tiles && tid < Q) ? t * Q + tid : n;
      in[u] = false; pos[u] = 0;
      if (idx[u] < n) { in[u] = __ldcg(&a.d2[idx[u]]) <= limit; pos[u] = __ldcg(&a.match_pos[idx[u]]); }   // TrimmedDist weight (A.4); false for NaN
    }
    float4 r[2], q[2], nr[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      r[u] = q[u] = nr[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (in[u]) { r[u] = __ldg(&a.read0[idx[u]]); q[u] = __ldg(&a.ix.pts[pos[u]]); nr[u] = __ldg(&a.normals[pos[u]]); }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (t0 + u * (int)gridDim.x >= n_tiles) break;                          // block-uniform
      acc_point_terms(in[u], r[u], q[u], nr[u], sT, lane, acc);
      if (++held == 4) flush();
    }
  }
  if (held) flush();
}

#ifdef AICP_DEBUG_WARP_TIMES
// experiment builds only (tools/warp_times_probe.py): per warp of the search phase of iteration AICP_DEBUG_WARP_TIMES, its duration
// in ns, its start offset from the phase start, how many of its lanes were outliers of the previous iteration, and its SM
__device__ unsigned int g_warp_times[4 * 8192];
extern "C" int aicp_b200_debug_warp_times(unsigned int* out, int n_words) {
  return (int)cudaMemcpyFromSymbol(out, g_warp_times, sizeof(unsigned int) * (size_t)n_words);
}
#endif

template <bool TILE>
__global__ void __launch_bounds__(256, 4) k_icp_loop(const __grid_constant__ LoopArgs a) {
  __shared__ unsigned int sh[AICP_HIST_BINS];
  __shared__ float sT[16];
  __shared__ float4 s_stage[TILE ? 8 : 1][32];
  __shared__ int s_stack[TILE ? 8 : 1][AICP_STACK];
  __shared__ unsigned long long s_lo[8][32];
  __shared__ long long s_hi[8][32];
  DeviceState* st = a.st;
  const PeerView& pv = a.pv;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  // a tile = the Q reading points one block handles at a time.  The search is bound by the latency of its slowest warp, and
  // that latency by how many divergent tree walks a warp serialises -- not by throughput (65 536 queries take as long as
  // 131 072).  When the reading is small enough to leave resident warps idle anyway, the queries are spread over more
  // warps: one query per S lanes (S <= 8), so a warp serialises 32 / S walks.  The dense phases use the first Q threads.
  const int S = TILE ? 1 : a.spread, Q = 256 / S;
  const int n = a.n, n_tiles = (n + Q - 1) / Q;
  const bool sharded = pv.n_ranks > 1;
  // a rank whose setup raised a status still joins the first exchange, so that every rank ends the loop together
  const bool dead = ld_int(&st->done) != 0;
  if (dead && !sharded) return;
  unsigned int epoch = 0;
  if (blockIdx.x == 0 && tid == 0) st->t_mark = global_ns();
  for (int it = 0; it <= AICP_B200_MAX_ITERS; ++it) {
    // ---- phase 1: correspondences + digit-1 histogram
    for (int b = tid; b < AICP_HIST_BINS; b += 256) sh[b] = 0;
    if (tid < 16) sT[tid] = __ldcg(&st->T_iter[tid]);
    __syncthreads();
    if (!dead) {
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (TILE && it > 0) {
          match_tile(a.ix, sT, a.read0, n, t * 256 + tid, it, a.match_pos, a.d2, a.trace_idx, sh, s_stage[TILE ? w : 0], s_stack[TILE ? w : 0]);
        } else {
#ifdef AICP_DEBUG_WARP_TIMES
          const unsigned long long w0 = global_ns();
          bool hard = false;
          if (it > 0 && t * Q + tid / S < n) hard = a.d2[t * Q + tid / S] > __ldcg(&st->limit);
          const unsigned int n_hard = __popc(__ballot_sync(0xFFFFFFFFu, hard));
#endif
          if ((tid & (S - 1)) == 0) match_thread(a.ix, sT, a.read0, n, t * Q + tid / S, it, a.match_pos, a.d2, a.trace_idx, sh);
#ifdef AICP_DEBUG_WARP_TIMES
          __syncwarp();
          if (it == AICP_DEBUG_WARP_TIMES && lane == 0 && t * 8 + w < 8192) {
            unsigned int smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned int* o = g_warp_times + 4 * (t * 8 + w);
            o[0] = (unsigned int)(global_ns() - w0); o[1] = (unsigned int)(w0 - st->t_mark); o[2] = n_hard; o[3] = smid;
          }
#endif
        }
      }
    }
    __syncthreads();
    hist_flush(sh, a.hist);
    grid_serial(st, epoch, [&] {
      const unsigned long long t0 = global_ns();
      loop_pick(st, a.hist, a.lp.ratio, pv, it);
      if (tid == 0) { st->tail_ns[0] += global_ns() - t0; st->phase_ns[0] += t0 - st->t_mark; st->t_mark = t0; }
    });
    if (ld_int(&st->done)) break;
    // ---- phase 2: candidate keys of the picked bin -> digits 2 and 3
    loop_candidates(a, Q, n_tiles);
    grid_serial(st, epoch, [&] {
      const unsigned long long t0 = global_ns();
      loop_select23(st, a.cand, sh, pv, it);
      if (tid == 0) { const unsigned long long t1 = global_ns(); st->tail_ns[1] += t1 - t0; st->phase_ns[1] += t1 - st->t_mark; st->t_mark = t1; }
    });
    if (ld_int(&st->done)) break;
    // ---- phase 3: normal equations of the inliers (d2 <= limit), exact fixed point
    loop_accumulate(a, Q, n_tiles, sT, s_hi);
    grid_serial(st, epoch, [&] {
      const unsigned long long t0 = global_ns();
      loop_fold_solve(st, a.slots, a.lp, n, pv, it, s_lo, s_hi);
      if (tid == 0) { const unsigned long long t1 = global_ns(); st->tail_ns[2] += t1 - t0; st->phase_ns[2] += t1 - st->t_mark; st->t_mark = t1; }
    });
    if (ld_int(&st->done)) break;
  }
}

// The multi-launch loop (batch workers) after k_match*: digits 2 + 3 of the trimmed quantile AND the normal equations +
// solve in ONE small cooperative launch -- candidates -> grid barrier whose serial section finishes the radix select ->
// exact sums -> the last block folds and solves.  Replaces k_select23 + k_accumulate (two launches, each costing ~45 us of
// queueing while eight registrations share the GPU).  `seq` (iteration + 1) is what the barrier's release word counts up to.
__global__ void __launch_bounds__(256, 4) k_quantile_accumulate(const __grid_constant__ LoopArgs a, unsigned int seq, int tail,
                                                                volatile int* progress) {
  DeviceState* st = a.st;
  if (ld_int(&st->done)) { publish_done(progress); return; }
  __shared__ unsigned int sh[AICP_HIST_BINS];
  __shared__ float sT[16];
  __shared__ unsigned long long s_lo[8][32];
  __shared__ long long s_hi[8][32];
  __shared__ bool s_last;
  const int tid = threadIdx.x;
  const int n_tiles = (a.n + 255) / 256;
  if (tid < 16) sT[tid] = st->T_iter[tid];
  loop_candidates(a, 256, n_tiles);
  // grid barrier with a serial section (see grid_serial); the arrival counter is re-armed by the block that runs the section
  __syncthreads();
  if (tid == 0) s_last = atom_add_release_gpu_u32(&st->bar_arrive, 1u) == gridDim.x - 1u;
  __syncthreads();
  if (s_last) {
    __threadfence();
    const unsigned long long t0 = global_ns();
    loop_select23(st, a.cand, sh, a.pv, 0);
    if (tid == 0) { st->tail_ns[1] += global_ns() - t0; st->bar_arrive = 0u; }
    __threadfence();
    __syncthreads();
    if (tid == 0) st_release_gpu_u32(&st->bar_release, seq);
  } else if (tid == 0) {
    while (ld_relaxed_gpu_u32(&st->bar_release) < seq) {}
    __threadfence();
  }
  __syncthreads();
  loop_accumulate(a, 256, n_tiles, sT, s_hi);
  if (!block_is_last(&st->ticket[2])) return;
  const unsigned long long t0 = global_ns();
  if (tail) {
    loop_fold_solve(st, a.slots, a.lp, a.n, a.pv, 0, s_lo, s_hi);
    if (tid == 0) {
      st->tail_ns[2] += global_ns() - t0;
      publish_progress(progress, st->iter, *(volatile int*)&st->done);
    }
  }
}

// ---- sharded registration: exchange of the normal-equation partials -----------------------------------------------
// 128-bit two's complement sums are all-reduced as four 32-bit limbs held in uint64 (no carry can be lost for fewer
// than 2^32 ranks); limb 4*AICP_NSUM carries "some rank raised a status".  Integer addition is associative, so every
// rank count gives bit-identical normal equations.
__global__ void k_sums_to_limbs(DeviceState* st, unsigned long long* limbs) {
  int i = threadIdx.x;
  if (i < AICP_NSUM) {
    unsigned long long lo = st->sum_lo[i], hi = (unsigned long long)st->sum_hi[i];
    limbs[4 * i + 0] = lo & 0xFFFFFFFFull; limbs[4 * i + 1] = lo >> 32;
    limbs[4 * i + 2] = hi & 0xFFFFFFFFull; limbs[4 * i + 3] = hi >> 32;
    st->sum_lo[i] = 0; st->sum_hi[i] = 0;
  }
  if (i == AICP_NSUM) limbs[4 * AICP_NSUM] = st->status != 0 ? 1ull : 0ull;
}

// Sharded loop control: every rank holds the same state after the exchange, so "the loop ended at iteration D" is the same
// fact on every rank.  The kernel publishes, in this order, (D + 1) and then the number of iterations whose exchange has
// completed; the hosts enqueue iteration `it` unless the loop had ended by iteration it - 2 -- a rule that depends only on
// D, not on when a host happens to look, so all ranks issue the same number of NCCL calls.
__global__ void k_limbs_solve(DeviceState* st, const unsigned long long* limbs, LoopParams lp, int n, int it, volatile int* progress) {
  if (threadIdx.x >= 32 || blockIdx.x != 0) return;
  const int lane = threadIdx.x;
  if (lane == 0 && limbs[4 * AICP_NSUM] != 0 && st->status == 0) raise_status(st, AICP_B200_ERR_COMM);
  __syncwarp();
  if (!ld_int(&st->done)) {
    if (lane < AICP_NSUM) {
      const int i = lane;
      unsigned long long l0 = limbs[4 * i], l1 = limbs[4 * i + 1], l2 = limbs[4 * i + 2], l3 = limbs[4 * i + 3];
      unsigned long long c = l0 >> 32;  unsigned long long w0 = l0 & 0xFFFFFFFFull;
      l1 += c; c = l1 >> 32;            unsigned long long w1 = l1 & 0xFFFFFFFFull;
      l2 += c; c = l2 >> 32;            unsigned long long w2 = l2 & 0xFFFFFFFFull;
      l3 += c;                          unsigned long long w3 = l3 & 0xFFFFFFFFull;      // modulo 2^128
      st->sum_lo[i] = w0 | (w1 << 32);
      st->sum_hi[i] = (long long)(w2 | (w3 << 32));
    }
    __threadfence();
    __syncwarp();
    const bool has = lane < AICP_NSUM;
    const unsigned long long slo = has ? __ldcg(&st->sum_lo[lane]) : 0ull;
    const long long shi = has ? __ldcg(&st->sum_hi[lane]) : 0ll;
    if (has) { st->sum_lo[lane] = 0; st->sum_hi[lane] = 0; }
    solve_and_check(st, lp, n, slo, shi);
  }
  __syncwarp();
  if (lane != 0) return;
  if (ld_int(&st->done) && st->done_at < 0) st->done_at = it;
  if (progress) {
    if (st->done_at >= 0) { progress[1] = st->done_at + 1; __threadfence_system(); }
    progress[0] = it + 1;
    __threadfence_system();
  }
}

// ---- epilogue ----------------------------------------------------------------------------------------------------
// A.1 step 7: T = T_refIn_refMean * T_iter * T_refMean_dataIn, evaluated left to right
__global__ void k_finalize(DeviceState* st) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float Tmu[16], Y[16];
  for (int i = 0; i < 16; ++i) Tmu[i] = (i % 5 == 0) ? 1.f : 0.f;
  Tmu[12] = st->mu[0]; Tmu[13] = st->mu[1]; Tmu[14] = st->mu[2];
  mat4_mul_f(Tmu, st->T_iter, Y);
  mat4_mul_f(Y, st->M0, st->T_final);
}

// pointmatcher_registration.cpp:128-131: out_read_cloud_ = T * reading (the unfiltered copy)
__global__ void __launch_bounds__(256) k_transform_out(const float4* __restrict__ read, int n, const DeviceState* __restrict__ st,
                                                       float4* __restrict__ out) {
  __shared__ float sT[16];
  if (threadIdx.x < 16) sT[threadIdx.x] = st->T_final[threadIdx.x];
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(&read[i]);
  float3 q = xform_f(sT, p.x, p.y, p.z);
  out[i] = make_float4(q.x, q.y, q.z, p.w);
}

// un-permute Morton-ordered normals (api: get_reference_normals)
__global__ void __launch_bounds__(256) k_scatter_normals(const float4* __restrict__ pts, const float4* __restrict__ normals, int n,
                                                         float4* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[__float_as_int(__ldg(&pts[i]).w)] = normals[i];
}

// stage entry point: plain exact NN of queries (no transform, no histogram)
__global__ void __launch_bounds__(256) k_match_plain(IndexView ix, const float4* __restrict__ qry, int n, int* __restrict__ out_idx,
                                                     float* __restrict__ out_d2) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 q = __ldg(&qry[i]);
  int pos; float d;
#ifdef AICP_DEBUG_WARP_TIMES
  DbgCnt* dbg = nullptr;
#endif
  nn_search(ix, q.x, q.y, q.z, &pos, &d AICP_DBG_ARG);
  out_idx[i] = __float_as_int(__ldg(&ix.pts[pos]).w);
  out_d2[i] = d;
}

__global__ void k_stage_reset(DeviceState* st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    st->status = 0; st->done = 0; st->prefix = 0; st->k_rem = 0; st->n_valid = 0; st->limit = 0.f;
    for (int i = 0; i < 4; ++i) st->ticket[i] = 0;
  }
}

// ---- host orchestration ------------------------------------------------------------------------------------------------
static int ensure_state(Handle* h) {
  if (!h->st) {
    CUDA_TRY(cudaMalloc((void**)&h->st, sizeof(DeviceState)));
    CUDA_TRY(cudaMemsetAsync(h->st, 0, sizeof(DeviceState), h->stream));
  }
  if (!h->st_host) CUDA_TRY(cudaMallocHost((void**)&h->st_host, sizeof(DeviceState)));
  if (!h->progress_host) {
    CUDA_TRY(cudaHostAlloc((void**)&h->progress_host, 2 * sizeof(int), cudaHostAllocMapped));
    CUDA_TRY(cudaHostGetDevicePointer((void**)&h->progress_dev, (void*)h->progress_host, 0));
    h->progress_host[0] = 0; h->progress_host[1] = 0;
  }
  if (h->hist.cap < AICP_HIST_BINS) {
    CUDA_TRY(h->hist.reserve(AICP_HIST_BINS));
    CUDA_TRY(cudaMemsetAsync(h->hist.p, 0, sizeof(unsigned int) * h->hist.cap, h->stream));   // cudaMalloc does not zero
  }
  if (h->acc_slots.cap < (size_t)ACC_SLOTS * 64) {
    CUDA_TRY(h->acc_slots.reserve((size_t)ACC_SLOTS * 64));
    CUDA_TRY(cudaMemsetAsync(h->acc_slots.p, 0, sizeof(unsigned long long) * h->acc_slots.cap, h->stream));
  }
  static_assert(AICP_HIST_BINS == 2048, "select_pick assumes 256 threads x 8 bins");
  return AICP_B200_OK;
}

static int status_to_error(Handle* h, int status) {
  switch (status) {
    case AICP_B200_OK: return AICP_B200_OK;
    case AICP_B200_ERR_NONFINITE_INPUT: return fail(h, status, "input cloud contains non-finite coordinates");
    case AICP_B200_ERR_EXTENT: return fail(h, status, "cloud extends more than 1024 m from the reference centroid");
    case AICP_B200_ERR_NO_VALID_MATCH: return fail(h, status, "TrimmedDistOutlierFilter: no finite positive distance (ConvergenceError)");
    case AICP_B200_ERR_NAN: return fail(h, status, "NaN in the iteration transform (ConvergenceError)");
    default: return fail(h, status, "device raised status %d", status);
  }
}

// one cooperative launch of the whole loop; the grid is what can be co-resident (a batch worker takes its share of the GPU)
static int launch_loop(Handle* h, LoopArgs& la, bool tile) {
  void* fn = tile ? (void*)k_icp_loop<true> : (void*)k_icp_loop<false>;
  if (!h->loop_occ[tile ? 1 : 0]) {
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 256, 0));
    CUDA_TRY(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device));
    if (occ < 1) return fail(h, AICP_B200_ERR_CUDA, "k_icp_loop cannot be resident on this device");
    h->loop_occ[tile ? 1 : 0] = occ;
  }
  int cap = h->loop_occ[tile ? 1 : 0] * h->n_sm;
  if (h->batch_worker && h->batch_streams > 1) cap /= h->batch_streams;
  if (cap < 1) cap = 1;
  int spread = 1;                                           // widest spread whose tiles are still all resident at once
  while (!tile && spread < 8 && (la.n + 256 / (2 * spread) - 1) / (256 / (2 * spread)) <= cap) spread *= 2;
  if (h->loop_spread > 0) spread = tile ? 1 : h->loop_spread;
  la.spread = spread;
  const int q = 256 / spread;
  int grid = (la.n + q - 1) / q;
  if (grid > cap) grid = cap;
  void* params[] = {&la};
  CUDA_TRY(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(256), params, 0, h->stream));
  return AICP_B200_OK;
}

int run_registration(Handle* h, const float* init_T_host, bool rebuild_reference, aicp_b200_stats* stats, float* out_T) {
  int rc = ensure_state(h);
  if (rc) return rc;
  cudaStream_t s = h->stream;
  const aicp_b200_icp_config& cfg = h->cfg;
  if (cfg.max_iterations < 1 || cfg.max_iterations > AICP_B200_MAX_ITERS || cfg.smooth_length < 1 ||
      cfg.smooth_length > AICP_B200_MAX_ITERS)
    return fail(h, AICP_B200_ERR_BAD_ARG, "maxIterationCount / smoothLength outside [1,%d]", AICP_B200_MAX_ITERS);
  if (!(cfg.ratio > 0.f) || cfg.ratio > 1.f) return fail(h, AICP_B200_ERR_BAD_ARG, "TrimmedDistOutlierFilter ratio %g outside (0,1]", cfg.ratio);
  const int n_read = (int)h->n_read, n_ref = (int)h->n_ref;
  h->launches = 0;
  const int prof = h->profiling;
  if (prof) {
    size_t need = 3 + 4 * (size_t)cfg.max_iterations;
    while (h->prof_ev.size() < need) { cudaEvent_t e; CUDA_TRY(cudaEventCreate(&e)); h->prof_ev.push_back(e); }
  }
  auto mark = [&](size_t i) { if (prof >= 2) cudaEventRecord(h->prof_ev[i], s); };
  auto mark_match = [&](size_t i) { if (prof >= 1) cudaEventRecord(h->prof_ev[i], s); };
  CUDA_TRY(cudaEventRecord(h->ev[0], s));
  mark(0);

  // One registration at a time: the reading-side setup (T_refMean_dataIn * reading, Morton sort of the reading) only needs
  // the reference's centroid, which the first kernel of its index build delivers -- it runs on a side stream beside the
  // reference's tree and SurfaceNormal filter instead of behind them (~50 us of a 1.27 ms registration).  Batch workers
  // share the GPU with seven other registrations and gain nothing from it.
  const bool fork = rebuild_reference && !h->batch_worker && !cfg.reading_normals && AICP_SIDE_STREAM;
  if (fork && !h->side) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_init, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  cudaStream_t rs = fork ? h->side : s;                   // the stream of the reading-side setup
  // buffers of the reading side and of the loop (grow-only: no-ops in the steady state)
  CUDA_TRY(h->read0.reserve((size_t)n_read));
  CUDA_TRY(h->read_out.reserve((size_t)n_read));
  CUDA_TRY(h->match_pos.reserve((size_t)n_read));
  CUDA_TRY(h->d2.reserve((size_t)n_read));
  float4* read_init = nullptr;
  if (init_T_host) { CUDA_TRY(h->read_init.reserve((size_t)n_read)); read_init = h->read_init.p; }
  h->has_init_reading = init_T_host != nullptr;
  int* trace_idx = nullptr;
  if (h->trace_matches) {
    CUDA_TRY(h->trace_idx.reserve((size_t)cfg.max_iterations * n_read));
    trace_idx = h->trace_idx.p;
    h->trace_iters = cfg.max_iterations; h->trace_n = n_read;
  }
  const int blocks = (n_read + 255) / 256;
  if (init_T_host) memcpy(h->st_host->T_init, init_T_host, 16 * sizeof(float));

  // everything between the inputs and the loop: index + normals of the reference, loop state, centred copies, the reading
  // transformed into the centred frame and Morton-ordered.  Stream work and grow-only reservations only, no host round trip.
  auto setup = [&]() -> int {
    int rc = AICP_B200_OK;
    if (rebuild_reference) {
      h->ref_ready = false;
      rc = build_index(h, h->ref_ix, h->ref_in.p, n_ref, true, fork ? h->ev_fork : nullptr);
      if (rc) return rc;
      mark(1);
      CUDA_TRY(h->refc_pts.reserve((size_t)h->ref_ix.n));
      CUDA_TRY(h->refc_rec.reserve((size_t)4 * h->ref_ix.n));
      CUDA_TRY(h->refc_cell.reserve((size_t)2 * h->ref_ix.n));
      if (h->comm && h->ref_ix.n >= 4096) {
        // sharded registration: every rank holds the same reference (hence the same Morton order), computes the normals of
        // its slice and the slices are all-gathered -- same bits as computing them all, 1 / n_ranks of the k-NN work
        int q0, q1, per;
        comm_slice(h, h->ref_ix.n, &q0, &q1, &per);
        CUDA_TRY(h->normals.reserve((size_t)per * comm_ranks(h)));
        rc = run_surface_normals(h, h->ref_ix, cfg.knn_normals, h->normals.p, nullptr, q0, q1);
        if (!rc) rc = comm_allgather_bytes(h, h->normals.p, (size_t)per * sizeof(float4));
      } else {
        CUDA_TRY(h->normals.reserve((size_t)h->ref_ix.n));
        CUDA_TRY(h->ref_rk2.reserve((size_t)h->ref_ix.n));      // k-th neighbour distances: what aicp_b200_reference_append needs later
        rc = run_surface_normals(h, h->ref_ix, cfg.knn_normals, h->normals.p, nullptr, 0, -1, nullptr, h->ref_rk2.p);
      }
      if (rc) return rc;
      h->ref_knn = cfg.knn_normals;
      mark(2);
    } else {
      mark(1); mark(2);
      if (h->ref_recentre) {                  // the reference grew by an append: new mean, new centred copies
        CUDA_TRY(h->refc_pts.reserve((size_t)h->ref_ix.n));
        CUDA_TRY(h->refc_rec.reserve((size_t)4 * h->ref_ix.n));
        CUDA_TRY(h->refc_cell.reserve((size_t)2 * h->ref_ix.n));
      }
    }
    if (cfg.reading_normals) {
      // the reference runs the same filter on the reading (icp_autotuned.yaml:9-14); PointToPlane never reads the result
      rc = build_index(h, h->tmp_ix, h->read_in.p, n_read);
      if (rc) return rc;
      CUDA_TRY(h->tmp_a.reserve((size_t)h->tmp_ix.n));
      rc = run_surface_normals(h, h->tmp_ix, cfg.knn_normals, h->tmp_a.p, nullptr);
      if (rc) return rc;
    }
    if (fork) CUDA_TRY(cudaStreamWaitEvent(rs, h->ev_fork, 0));
    if (init_T_host) CUDA_TRY(cudaMemcpyAsync(h->st->T_init, h->st_host->T_init, 16 * sizeof(float), cudaMemcpyHostToDevice, rs));
    CUDA_TRY(cudaMemsetAsync(h->hist.p, 0, sizeof(unsigned int) * AICP_HIST_BINS, s));
    k_loop_init<<<1, 32, 0, rs>>>(h->st, h->ref_ix.meta, (long long)n_ref, init_T_host ? 1 : 0);
    if (fork) {
      CUDA_TRY(cudaEventRecord(h->ev_init, rs));
      CUDA_TRY(cudaStreamWaitEvent(s, h->ev_init, 0));      // k_centre reads the mean
    }
    if (rebuild_reference || h->ref_recentre) {
      h->ref_recentre = false;
      int n4 = 4 * (h->ref_ix.n - 1);
      int m = h->ref_ix.n > n4 ? h->ref_ix.n : n4;
      k_centre<<<(m + 255) / 256, 256, 0, s>>>(h->ref_ix.pts.p, h->ref_ix.n, h->ref_ix.rec.p, n4, h->ref_ix.cellbox.p, h->st,
                                              h->refc_pts.p, h->refc_rec.p, h->refc_cell.p);
      h->launches += 1;
    }
    k_read_prepare<<<blocks, 256, 0, rs>>>(h->read_in.p, n_read, h->st, h->read0.p, read_init);
    h->launches += 2;
    // Morton-order the reading so that the 32 queries of a warp walk the same part of the reference tree
    h->stream = rs;                                           // build_index works on the handle's stream
    rc = build_index(h, h->read_ix, h->read0.p, n_read, false);
    h->stream = s;
    if (rc) return rc;
    if (fork) {
      CUDA_TRY(cudaEventRecord(h->ev_join, rs));
      CUDA_TRY(cudaStreamWaitEvent(s, h->ev_join, 0));
    }
    return AICP_B200_OK;
  };

  // The setup is replayed as ONE CUDA graph: its ~30 small launches are what the GPU's command fetch spends its time on while
  // host clouds are being uploaded next door (batches: +2.5 % device-resident, +4 % from host buffers; one registration at a
  // time: -20 us).  The graph holds raw pointers and sizes, so it is captured the second time a handle
  // meets the same sizes with no buffer (re)allocated anywhere in between, and replayed for as long as that stays true; all
  // host-side state the setup leaves behind is then already in place from the run before.  The side stream of a single
  // registration is part of the capture (fork / join through its events).  AICP_B200_SETUP_GRAPH=0 off, 1 batch workers only.
  static const int setup_graphs = [] { const char* e = getenv("AICP_B200_SETUP_GRAPH"); return e ? atoi(e) : 2; }();
  const bool graph_ok = (h->batch_worker ? setup_graphs >= 1 : setup_graphs >= 2) && !h->setup_graph_off && rebuild_reference && prof < 2 && !h->comm &&
                        !cfg.reading_normals && !h->trace_matches && !h->ref_recentre && cfg.knn_normals <= 24;
  SetupKey key;
  key.n_ref = n_ref; key.n_read = n_read; key.knn = cfg.knn_normals; key.has_init = init_T_host ? 1 : 0;
  key.knn_schedule = h->knn_schedule | (h->batch_worker ? 16 : 0);      // both decide which kernels the setup launches
  key.generation = g_alloc_generation.load(std::memory_order_relaxed);
  if (graph_ok && h->setup_exec && key == h->setup_key) {
    h->ref_ready = false;
    CUDA_TRY(cudaGraphLaunch(h->setup_exec, s));
    h->launches += h->setup_launches;
  } else if (graph_ok && key == h->setup_seen) {
    if (h->setup_exec) { cudaGraphExecDestroy(h->setup_exec); h->setup_exec = nullptr; }
    const int launches_before = h->launches;
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    rc = setup();
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    const bool moved = g_alloc_generation.load(std::memory_order_relaxed) != key.generation;    // somebody allocated meanwhile
    if (!rc && ce == cudaSuccess && graph && !moved && cudaGraphInstantiate(&h->setup_exec, graph, 0) == cudaSuccess) {
      h->setup_key = key;
      h->setup_launches = h->launches - launches_before;
      cudaGraphDestroy(graph);
      CUDA_TRY(cudaGraphLaunch(h->setup_exec, s));
    } else {
      // nothing has run yet: drop the capture and do the setup with plain launches
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      h->setup_exec = nullptr;
      if (rc || ce != cudaSuccess) h->setup_graph_off = true;
      h->launches = launches_before;
      if ((rc = setup())) return rc;
      h->setup_seen = key;
      h->setup_seen.generation = g_alloc_generation.load(std::memory_order_relaxed);
    }
  } else {
    if ((rc = setup())) return rc;
    h->setup_seen = key;
    h->setup_seen.generation = g_alloc_generation.load(std::memory_order_relaxed);     // after the setup's own reservations
  }
  const float4* read_s = h->read_ix.pts.p;
  CUDA_TRY(cudaEventRecord(h->ev[1], s));

  // the loop schedule is decided first (the description is at `persistent` below): it decides the carrier of a sharded exchange
  const bool want_persistent = h->loop_schedule == 2 || (h->loop_schedule == 0 && !h->batch_worker);
  if (h->comm && (rc = comm_begin_registration(h, n_read, want_persistent))) return rc;
  IndexView cix{h->refc_pts.p, h->refc_rec.p, h->ref_ix.owner.p, h->ref_ix.owner.p + h->ref_ix.n, h->refc_cell.p, h->ref_ix.n};
  LoopParams lp{cfg.ratio, cfg.max_iterations, cfg.min_diff_rot, cfg.min_diff_trans, cfg.smooth_length};
  const int sel_blocks = blocks < 148 * 2 ? blocks : 148 * 2;
  const int acc_blocks = (n_read + 256 * ACC_PTS - 1) / (256 * ACC_PTS);
  CUDA_TRY(h->cand.reserve((size_t)n_read));
  // Loop control lives on the device.  The host enqueues the iterations the differential checker cannot stop before
  // (smoothLength) blindly, then stays at most LOOKAHEAD iterations ahead of the device by watching the iteration
  // counter that the solve publishes in mapped pinned memory; iterations enqueued after convergence return at once.
  const int LOOKAHEAD = 2;
  // two schedules of the same exact search: the tile kernel issues ~5 % fewer instructions (throughput of batched
  // registrations), the per-thread kernel finishes one launch on an idle GPU ~25 % sooner (latency)
  const bool tile_match = h->match_schedule == 2 || (h->match_schedule == 0 && h->batch_worker);
  volatile int* prog = h->progress_host;
  int* prog_dev = h->comm ? nullptr : h->progress_dev;
  prog[0] = 0; prog[1] = 0;           // the stream is idle here: every earlier call ended with a synchronisation
  int enqueued = 0;
  PeerView pv;
  const bool peer_exchange = comm_peer_view(h, &pv);
  // Schedule of the loop.  One registration at a time (the production path, app.cpp:528-550) and the sharded registration
  // run the persistent kernel: 1.31 vs 1.38 ms for the C3 pair on B200, and the search phase alone 47 vs 59 us per iteration
  // (the L1 stays warm across iterations).  Batch workers keep three launches per iteration: eight resident loop kernels
  // hold every register file of the GPU while they wait at their barriers, which starves the other streams' setup kernels and
  // loses the block scheduler's dynamic load balance (measured: 1905 vs 2340 registrations/s).
  const bool persistent = want_persistent && (!h->comm || peer_exchange);
  if (persistent) {
    LoopArgs la{cix, h->normals.p, read_s, n_read, h->st, h->match_pos.p, h->d2.p, h->hist.p, h->cand.p, trace_idx, h->acc_slots.p, lp, pv, 1};
    if ((rc = launch_loop(h, la, tile_match))) return rc;
    h->launches += 1;
  }
  for (int it = 0; it < cfg.max_iterations && !persistent; ++it) {
    if (it >= cfg.smooth_length && it >= LOOKAHEAD) {
      unsigned spins = 0;
      // busy-wait on purpose: sleeping on a blocking-sync event instead was measured 6 % slower (wake-up latency), also with
      // four ranks on one host; after a short spin the thread yields so that oversubscribed hosts degrade gracefully
      while ((h->comm || prog[1] == 0) && prog[0] < it - LOOKAHEAD + 1) {
        if ((++spins & 0x3FFu) == 0 && cudaStreamQuery(s) != cudaErrorNotReady) break;   // stream drained or failed: stop waiting
        if (spins > 256) std::this_thread::yield();
#if defined(__x86_64__)
        else __builtin_ia32_pause();
#endif
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      const int ended = prog[1];
      if (!h->comm) { if (ended != 0) break; }
      else if (ended != 0 && ended - 1 <= it - LOOKAHEAD) break;       // same decision on every rank (see k_limbs_solve)
    }
    mark_match(3 + 4 * (size_t)it);
    if (!h->comm) {
      if (!tile_match || it == 0)      // the cold first search has no per-lane bounds to share: per-thread descent (26 % fewer instructions)
        k_match<<<blocks, 256, 0, s>>>(cix, read_s, n_read, h->st, h->match_pos.p, h->d2.p, h->hist.p, trace_idx, cfg.ratio, 1, prog_dev);
      else
        k_match_tile<<<blocks, 256, 0, s>>>(cix, read_s, n_read, h->st, h->match_pos.p, h->d2.p, h->hist.p, trace_idx, cfg.ratio, 1, prog_dev);
      mark_match(4 + 4 * (size_t)it);
      if (h->fused_tail) {
        // quantile digits 2 + 3, normal equations and solve in one cooperative launch (k_quantile_accumulate)
        LoopArgs la{cix, h->normals.p, read_s, n_read, h->st, h->match_pos.p, h->d2.p, h->hist.p, h->cand.p, trace_idx, h->acc_slots.p, lp, pv, 1};
        unsigned int seq = (unsigned int)it + 1u;
        int one = 1;
        void* params[] = {&la, &seq, &one, &prog_dev};
        int qgrid = blocks < 74 ? blocks : 74;
        mark(5 + 4 * (size_t)it);
        CUDA_TRY(cudaLaunchCooperativeKernel((void*)k_quantile_accumulate, dim3((unsigned)qgrid), dim3(256), params, 0, s));
        mark(6 + 4 * (size_t)it);
        h->launches += 2;
      } else {
        k_select23<<<sel_blocks, 256, 0, s>>>(h->d2.p, n_read, h->st, h->cand.p, prog_dev);
        mark(5 + 4 * (size_t)it);
        k_accumulate<<<acc_blocks, 256, 0, s>>>(h->refc_pts.p, h->normals.p, read_s, h->match_pos.p, h->d2.p, n_read, h->st, lp, 1, prog_dev, h->acc_slots.p);
        mark(6 + 4 * (size_t)it);
        h->launches += 3;
      }
    } else {
      // reading sharded over the ranks: the trimmed quantile is GLOBAL, so each radix-select digit is
      // picked from the all-reduced histogram; then the 27 normal-equation partials (+ count) are all-reduced
      if (!tile_match || it == 0)
        k_match<<<blocks, 256, 0, s>>>(cix, read_s, n_read, h->st, h->match_pos.p, h->d2.p, h->hist.p, trace_idx, cfg.ratio, 0, nullptr);
      else
        k_match_tile<<<blocks, 256, 0, s>>>(cix, read_s, n_read, h->st, h->match_pos.p, h->d2.p, h->hist.p, trace_idx, cfg.ratio, 0, nullptr);
      if ((rc = comm_allreduce_u32(h, h->hist.p, AICP_HIST_BINS))) return rc;
      k_pick<<<1, 256, 0, s>>>(h->st, h->hist.p, 1, cfg.ratio);
      mark_match(4 + 4 * (size_t)it);
      for (int pass = 2; pass <= 3; ++pass) {
        k_select<<<sel_blocks, 256, 0, s>>>(h->d2.p, n_read, h->st, h->hist.p, pass, cfg.ratio, 0);
        if ((rc = comm_allreduce_u32(h, h->hist.p, AICP_HIST_BINS))) return rc;
        k_pick<<<1, 256, 0, s>>>(h->st, h->hist.p, pass, cfg.ratio);
      }
      mark(5 + 4 * (size_t)it);
      k_accumulate<<<acc_blocks, 256, 0, s>>>(h->refc_pts.p, h->normals.p, read_s, h->match_pos.p, h->d2.p, n_read, h->st, lp, 0, nullptr, h->acc_slots.p);
      unsigned long long* limbs = comm_limbs(h);
      k_sums_to_limbs<<<1, 32, 0, s>>>(h->st, limbs);
      if ((rc = comm_allreduce_u64(h, limbs, 4 * AICP_NSUM + 1))) return rc;
      k_limbs_solve<<<1, 32, 0, s>>>(h->st, limbs, lp, n_read, it, h->progress_dev);
      mark(6 + 4 * (size_t)it);
      h->launches += 9;
    }
    ++enqueued;
  }
  CUDA_TRY(cudaEventRecord(h->ev[2], s));
  k_finalize<<<1, 32, 0, s>>>(h->st);
  k_transform_out<<<blocks, 256, 0, s>>>(h->read_in.p, n_read, h->st, h->read_out.p);
  h->launches += 2;
  CUDA_TRY(cudaMemcpyAsync(h->st_host, h->st, sizeof(DeviceState), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaEventRecord(h->ev[3], s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaGetLastError());

  const DeviceState* hs = h->st_host;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->iterations = hs->iter;
    stats->stop_reason = hs->stop_reason;
    const long long n_all = !h->comm ? (long long)n_read : (peer_exchange ? (long long)hs->n_read_total : comm_total_reading(h));
    stats->weighted_point_used_ratio = (float)hs->n_used_last / (float)n_all;
    for (int d = 0; d < 3; ++d) stats->mean_ref[d] = hs->mu[d];
    stats->n_ref = n_ref; stats->n_read = n_read;
    cudaEventElapsedTime(&stats->ms_total, h->ev[0], h->ev[3]);
    cudaEventElapsedTime(&stats->ms_setup, h->ev[0], h->ev[1]);
    cudaEventElapsedTime(&stats->ms_iterations, h->ev[1], h->ev[2]);
    stats->gpu_launches = h->launches;
    if (prof) {
      stats->profiled = prof;
      float ms = 0.f;
      if (prof >= 2) {
        cudaEventElapsedTime(&stats->ms_index, h->prof_ev[0], h->prof_ev[1]);
        cudaEventElapsedTime(&stats->ms_normals, h->prof_ev[1], h->prof_ev[2]);
      }
      for (int it = 0; it < hs->iter && it < enqueued; ++it) {
        cudaEventElapsedTime(&ms, h->prof_ev[3 + 4 * it], h->prof_ev[4 + 4 * it]); stats->ms_match += ms;
        if (prof < 2) continue;
        cudaEventElapsedTime(&ms, h->prof_ev[4 + 4 * it], h->prof_ev[5 + 4 * it]); stats->ms_select += ms;
        cudaEventElapsedTime(&ms, h->prof_ev[5 + 4 * it], h->prof_ev[6 + 4 * it]); stats->ms_accumulate += ms;
      }
    }
    if (persistent) {
      // phase clocks of the loop kernel (globaltimer, taken at its grid barriers) instead of CUDA events around launches
      stats->profiled = 2;
      stats->ms_match = (float)((double)hs->phase_ns[0] * 1e-6);
      stats->ms_select = (float)((double)hs->phase_ns[1] * 1e-6);
      stats->ms_accumulate = (float)((double)hs->phase_ns[2] * 1e-6);
      stats->ms_exchange = (float)((double)hs->phase_ns[3] * 1e-6);
    }
    stats->ms_tail_pick = (float)((double)hs->tail_ns[0] * 1e-6);
    stats->ms_tail_select = (float)((double)hs->tail_ns[1] * 1e-6);
    stats->ms_tail_solve = (float)((double)hs->tail_ns[2] * 1e-6);
    int nt = hs->iter < AICP_B200_MAX_ITERS ? hs->iter : AICP_B200_MAX_ITERS;
    memcpy(stats->trace, hs->trace, sizeof(aicp_b200_iter_trace) * (size_t)nt);
  }
  if (hs->status) return status_to_error(h, hs->status);
  h->ref_ready = true;
  if (out_T) memcpy(out_T, hs->T_final, 16 * sizeof(float));
  return AICP_B200_OK;
}

int run_match_stage(Handle* h, const SpatialIndex& ix, const float4* qry, int64_t n_qry, int* out_idx, float* out_d2) {
  if (n_qry == 0) return AICP_B200_OK;
  k_match_plain<<<(unsigned)((n_qry + 255) / 256), 256, 0, h->stream>>>(ix.view(), qry, (int)n_qry, out_idx, out_d2);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return AICP_B200_OK;
}

int run_trim_stage(Handle* h, const float* d2_dev, int64_t n64, float ratio, float* out_limit, int64_t* out_n_valid) {
  int rc = ensure_state(h);
  if (rc) return rc;
  if (!(ratio > 0.f) || ratio > 1.f) return fail(h, AICP_B200_ERR_BAD_ARG, "ratio %g outside (0,1]", ratio);
  cudaStream_t s = h->stream;
  int n = (int)n64;
  int blocks = (n + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 2) blocks = 148 * 2;
  k_stage_reset<<<1, 32, 0, s>>>(h->st);
  CUDA_TRY(cudaMemsetAsync(h->hist.p, 0, sizeof(unsigned int) * AICP_HIST_BINS, s));
  for (int pass = 1; pass <= 3; ++pass) k_select<<<blocks, 256, 0, s>>>(d2_dev, n, h->st, h->hist.p, pass, ratio, 1);
  CUDA_TRY(cudaMemcpyAsync(h->st_host, h->st, sizeof(DeviceState), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaGetLastError());
  if (out_n_valid) *out_n_valid = (int64_t)h->st_host->n_valid;
  if (h->st_host->status) return status_to_error(h, h->st_host->status);
  *out_limit = h->st_host->limit;
  return AICP_B200_OK;
}

int scatter_normals(Handle* h, const float4* pts_morton, const float4* normals_morton, int n, float4* out_dev) {
  k_scatter_normals<<<(n + 255) / 256, 256, 0, h->stream>>>(pts_morton, normals_morton, n, out_dev);
  CUDA_TRY(cudaGetLastError());
  return AICP_B200_OK;
}

}  // namespace aicp
