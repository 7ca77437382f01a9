// common.cuh -- shared device/host declarations of libaicp_b200 (sm_100a only).
//
// Numerical contract (DESIGN.md "Arithmetic"): float32 geometry with a fixed operation order and no FMA contraction
// (the library is compiled with -fmad=false and the distance code uses __fmul_rn/__fadd_rn explicitly), exact
// fixed-point reductions, float64 solve built from + - * / sqrt only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aicp_b200.h"

#ifndef AICP_LEAF
#define AICP_LEAF 16                // points per tree leaf scanned by the per-thread 1-NN walk (8: -2 %, 4: -5 % throughput)
#endif
#define AICP_FIXED_SCALE 1073741824.0   // 2^30, normal-equation fixed point (matches the oracle's contract)
#define AICP_CENTROID_SCALE 65536.0     // 2^16
#define AICP_MAX_KNN 64
#define AICP_NSUM 28                // 21 (A upper triangle) + 6 (g) + 1 (count)
#define AICP_HIST_BINS 2048

#define CUDA_TRY(expr)                                                     \
  do {                                                                     \
    cudaError_t e__ = (expr);                                              \
    if (e__ != cudaSuccess) return aicp::fail_cuda(h, e__, #expr, __LINE__); \
  } while (0)

namespace aicp {

// ---- device-resident state of one registration -------------------------------------------------------------------
struct DeviceState {
  // geometry
  float T_iter[16];                 // column-major, centred frame
  float M0[16];                     // T_refMean_dataIn
  float T_final[16];
  float T_init[16];
  float mu[4];
  int   bbox_min[3], bbox_max[3];   // order-preserving int encodings of the reference bounding box
  int   rbox_min[3], rbox_max[3];   // reading' bounding box
  long long centroid_sum[3];
  // loop control
  int   iter;
  int   done;
  int   done_at;                    // sharded loop: index of the iteration whose solve (or status exchange) ended the loop, -1 before
  int   stop_reason;
  int   status;                     // AICP_B200_* error raised on device
  int   hist_n;
  // trimmed quantile (radix select)
  unsigned int prefix;
  unsigned long long k_rem;
  unsigned long long n_valid;
  float limit;
  unsigned int ticket[4];
  unsigned int cand_n;              // candidate keys appended by k_select23
  unsigned long long tail_ns[3];    // time spent in the single-block tails: digit-1 pick, digits 2+3, solve + checkers
  // persistent loop kernel (k_icp_loop): grid barrier words and per-phase clocks (globaltimer ns, taken by the block that
  // runs a phase's serial section): [0] search, [1] quantile, [2] normal equations + solve, [3] peer exchanges
  unsigned int bar_arrive, bar_release;
  unsigned long long phase_ns[4];
  unsigned long long t_mark;
  unsigned long long n_read_total;  // sharded registration: reading points over all ranks (travels with the sums)
  // normal equations, 128-bit two's complement fixed point
  unsigned long long sum_lo[AICP_NSUM];
  long long sum_hi[AICP_NSUM];
  long long n_used_last;
  // checkers
  double quat_hist[AICP_B200_MAX_ITERS + 1][4];
  double tr_hist[AICP_B200_MAX_ITERS + 1][3];
  double ang_step[AICP_B200_MAX_ITERS + 1];   // |angularDistance(q_i, q_{i-1})|, cached so each iteration computes one
  double trn_step[AICP_B200_MAX_ITERS + 1];   // ||t_i - t_{i-1}||
  aicp_b200_iter_trace trace[AICP_B200_MAX_ITERS];
};

// ---- sharded registration over peer-mapped memory (comm.cu, icp.cu) ---------------------------------------------------
// Every rank owns one INBOX in its own HBM; every peer maps it (CUDA IPC between processes, peer access inside one) and
// stores its contribution for this rank there.  The protocol is "flag in data" (what NCCL calls LL): every 64-bit word a
// peer writes carries 32 bits of payload and, in its upper half, the 32-bit sequence stamp of the exchange
// (2048 * epoch + 4 * iteration + round + 1); a 64-bit store is single-copy atomic, so a reader that sees the stamp sees
// the payload -- no fence on the sending side, no separate flag, one NVLink write latency per exchange (the first version
// used plain stores + __threadfence_system + a flag + a fence after the flag: 16 us per exchange at 8 ranks).
// Layout of an inbox, per SOURCE rank s (all 64-bit words):
//   sums   @ SUMS_OFF + 1024 s         word 32 k + l: limb k of sum l (l < 28), l = 28: points of the source's shard, l = 31: its status
//   hist   @ HIST_OFF + (2048 + 8) 8 s word 256 j + t: bin 8 t + j (a warp stores 256 contiguous bytes), then one status word
//   cand   @ CAND_OFF + stride s       header (count | status << 31), then the source's candidate keys (capacity cand_cap)
#define AICP_MAX_RANKS 16
#define AICP_INBOX_SUMS_OFF 0
#define AICP_INBOX_SUMS_STRIDE 1024
#define AICP_INBOX_HIST_OFF (AICP_INBOX_SUMS_OFF + AICP_INBOX_SUMS_STRIDE * AICP_MAX_RANKS)
#define AICP_INBOX_HIST_STRIDE ((AICP_HIST_BINS + 8) * 8)
#define AICP_INBOX_CAND_OFF (AICP_INBOX_HIST_OFF + AICP_INBOX_HIST_STRIDE * AICP_MAX_RANKS)
#define AICP_PEER_TIMEOUT_NS 10000000000ull      // a peer that does not show up within 10 s ends the loop with AICP_B200_ERR_COMM

struct PeerView {
  int rank, n_ranks;                       // n_ranks <= 1: no exchange
  unsigned long long epoch;                // registration sequence number, the same on every rank
  unsigned char* inbox[AICP_MAX_RANKS];    // inbox[r] = rank r's inbox as mapped into this process (inbox[rank]: local memory)
  size_t cand_stride;                      // bytes between two sources' candidate areas
  unsigned int cand_cap;                   // keys a source's candidate area holds
};

struct LoopParams {
  float ratio;
  int   max_iterations;
  float min_diff_rot, min_diff_trans;
  int   smooth_length;
};

// ---- float helpers (never contracted) -------------------------------------------------------------------------------
__device__ __forceinline__ float d2_f(float qx, float qy, float qz, float px, float py, float pz) {
  float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
  float d = __fmul_rn(dx, dx);
  d = __fadd_rn(d, __fmul_rn(dy, dy));
  d = __fadd_rn(d, __fmul_rn(dz, dz));
  return d;
}

// squared distance from q to the box [lo,hi]; never larger than d2_f(q, p) for any p inside (monotone rounding)
__device__ __forceinline__ float box_d2_f(float3 lo, float3 hi, float qx, float qy, float qz) {
  float dx = fmaxf(0.f, fmaxf(__fsub_rn(lo.x, qx), __fsub_rn(qx, hi.x)));
  float dy = fmaxf(0.f, fmaxf(__fsub_rn(lo.y, qy), __fsub_rn(qy, hi.y)));
  float dz = fmaxf(0.f, fmaxf(__fsub_rn(lo.z, qz), __fsub_rn(qz, hi.z)));
  float d = __fmul_rn(dx, dx);
  d = __fadd_rn(d, __fmul_rn(dy, dy));
  d = __fadd_rn(d, __fmul_rn(dz, dz));
  return d;
}

// out = T * (x,y,z,1), column-major T, k ascending
__device__ __forceinline__ float3 xform_f(const float* __restrict__ T, float x, float y, float z) {
  float3 o;
  float a;
  a = __fmul_rn(T[0], x); a = __fadd_rn(a, __fmul_rn(T[4], y)); a = __fadd_rn(a, __fmul_rn(T[8], z)); o.x = __fadd_rn(a, T[12]);
  a = __fmul_rn(T[1], x); a = __fadd_rn(a, __fmul_rn(T[5], y)); a = __fadd_rn(a, __fmul_rn(T[9], z)); o.y = __fadd_rn(a, T[13]);
  a = __fmul_rn(T[2], x); a = __fadd_rn(a, __fmul_rn(T[6], y)); a = __fadd_rn(a, __fmul_rn(T[10], z)); o.z = __fadd_rn(a, T[14]);
  return o;
}

// C = A * B for rigid 4x4 (bottom row fixed), float, k ascending.  C may alias A or B.
__device__ inline void mat4_mul_f(const float* A, const float* B, float* C) {
  float R[16];
  for (int c = 0; c < 4; ++c) {
    for (int r = 0; r < 3; ++r) {
      float acc = __fmul_rn(A[0 * 4 + r], B[c * 4 + 0]);
      acc = __fadd_rn(acc, __fmul_rn(A[1 * 4 + r], B[c * 4 + 1]));
      acc = __fadd_rn(acc, __fmul_rn(A[2 * 4 + r], B[c * 4 + 2]));
      if (c == 3) acc = __fadd_rn(acc, A[3 * 4 + r]);
      R[c * 4 + r] = acc;
    }
    R[c * 4 + 3] = (c == 3) ? 1.f : 0.f;
  }
  for (int i = 0; i < 16; ++i) C[i] = R[i];
}

// order-preserving float <-> int (for atomicMin / atomicMax on floats)
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __host__ __forceinline__ float ordered_to_float(int i) {
  int b = i >= 0 ? i : i ^ 0x7FFFFFFF;
#ifdef __CUDA_ARCH__
  return __int_as_float(b);
#else
  float f; memcpy(&f, &b, 4); return f;
#endif
}

// ---- host-side plumbing ------------------------------------------------------------------------------------------
struct Handle;
int fail(Handle* h, int code, const char* fmt, ...);
int fail_cuda(Handle* h, cudaError_t e, const char* expr, int line);

}  // namespace aicp
