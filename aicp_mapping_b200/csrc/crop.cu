// crop.cu -- getPointsInOrientedBox on the GPU: the oriented crop of the (prior or built) map around the robot.
//
// replaces getPointsInOrientedBox (aicp_core/src/utils/filteringUtils.cpp:621-637), i.e. pcl::CropBox with
// min = (m,m,m), max = (M,M,M), rotation = origin.block<3,3>(0,0).eulerAngles(0,1,2), translation = origin.col(3), which
// App runs on the whole map before every registration against it (app.cpp:41-69; +-15 m, aicp.launch:56).
// [UPSTREAM, recalled] CropBox keeps a point when  min <= R^-1 (p - t) <= max  with R = pcl::getTransformation(rpy).
//
// One pass over the map, HBM-bound: a 10 485 760-point map is 168 MB read + 16 B per kept point written.
//   k_crop_box  tiles of 2048 points (256 threads x 8 rows, coalesced float4 loads kept in registers); per-row warp
//               ballots rank the kept points, a 64-entry shared scan orders (row, warp) pairs so that the output keeps the
//               INPUT ORDER (CropBox semantics); the tile's base offset comes from a single-pass chained scan with
//               decoupled look-back (tiles take their number from an atomic ticket, so a tile only ever waits for tiles
//               that are already running).
// Arithmetic: local_k = (m_k0*dx + m_k1*dy) + m_k2*dz in float32 without FMA, inverse rotation = transpose; identical to
// oracle/aicp_oracle_filters.c, so the kept set is bit-for-bit the oracle's.
#include <math.h>

#include "handle.cuh"

namespace aicp {

#define CROP_ROWS 8
#define CROP_TILE (256 * CROP_ROWS)
// tile status word: bits 63..62 = 0 not ready, 1 aggregate of this tile only, 2 inclusive prefix; bits 61..0 = count
#define CROP_AGG (1ull << 62)
#define CROP_PREFIX (2ull << 62)
#define CROP_MASK ((1ull << 62) - 1ull)

struct CropBox9 {
  float m[9];        // inverse rotation (row-major): local = m * (p - t)
  float t[3];
  float bmin, bmax;
  float reach;       // sqrt(3) * max(|bmin|, |bmax|) * (1 + 1e-4): no point with a larger |p_k - t_k| can be inside
};

__global__ void k_crop_reset(unsigned long long* status, int n_tiles, unsigned int* ticket, unsigned long long* total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_tiles) status[i] = 0ull;
  if (i == 0) { *ticket = 0u; *total = 0ull; }
}

__global__ void __launch_bounds__(256) k_crop_box(const float4* __restrict__ pts, long long n, CropBox9 box, float4* __restrict__ out,
                                                  unsigned long long* status, unsigned int* ticket, unsigned long long* total) {
  __shared__ unsigned int s_tile;
  __shared__ int s_cnt[CROP_ROWS * 8 + 1];
  __shared__ unsigned long long s_base;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned int tile = s_tile;
  const long long base = (long long)tile * CROP_TILE;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // all loads of the tile first (8 independent 16-byte loads in flight per thread), then the tests
  float4 p[CROP_ROWS];
#pragma unroll
  for (int r = 0; r < CROP_ROWS; ++r) {
    const long long i = base + r * 256 + threadIdx.x;
    p[r] = i < n ? __ldg(&pts[i]) : make_float4(NAN, NAN, NAN, 0.f);          // padding behaves like a non-finite point
  }
  unsigned int keep = 0;              // bit r: row r of this thread is inside the box
  int rank[CROP_ROWS];
#pragma unroll
  for (int r = 0; r < CROP_ROWS; ++r) {
    bool in = false;
    const float dx = __fsub_rn(p[r].x, box.t[0]), dy = __fsub_rn(p[r].y, box.t[1]), dz = __fsub_rn(p[r].z, box.t[2]);
    // cheap conservative reject before the rotation: inside needs |local_k| <= B on every axis, hence |d|_2 <= sqrt(3) B;
    // a single |d_k| beyond that bound (with 1e-4 of slack, 100x the float error of the rotated values) decides "outside".
    // NaN and infinite coordinates fail this test too (the comparison is false for NaN), so no separate isfinite pass.
    if (fabsf(dx) <= box.reach && fabsf(dy) <= box.reach && fabsf(dz) <= box.reach) {
      const float lx = __fadd_rn(__fadd_rn(__fmul_rn(box.m[0], dx), __fmul_rn(box.m[1], dy)), __fmul_rn(box.m[2], dz));
      const float ly = __fadd_rn(__fadd_rn(__fmul_rn(box.m[3], dx), __fmul_rn(box.m[4], dy)), __fmul_rn(box.m[5], dz));
      const float lz = __fadd_rn(__fadd_rn(__fmul_rn(box.m[6], dx), __fmul_rn(box.m[7], dy)), __fmul_rn(box.m[8], dz));
      in = !(lx < box.bmin || ly < box.bmin || lz < box.bmin || lx > box.bmax || ly > box.bmax || lz > box.bmax);
    }
    const unsigned int bal = __ballot_sync(0xFFFFFFFFu, in);
    rank[r] = __popc(bal & ((1u << lane) - 1u));
    if (in) keep |= 1u << r;
    if (lane == 0) s_cnt[r * 8 + w] = __popc(bal);
  }
  __syncthreads();
  // exclusive scan of the 64 (row, warp) counts, row-major = input order; warp 0, two entries per lane
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
    // chained scan with decoupled look-back, one warp wide: publish this tile's aggregate, then read the status words of
    // the 32 preceding tiles at once, add the aggregates down to the nearest tile that already knows its inclusive prefix
    unsigned long long prefix = 0;
    if (tile == 0) {
      if (lane == 0) atomicExch(&status[0], CROP_PREFIX | (unsigned long long)tile_total);
    } else {
      if (lane == 0) atomicExch(&status[tile], CROP_AGG | (unsigned long long)tile_total);
      long long j0 = (long long)tile - 1;
      while (true) {
        const long long j = j0 - lane;
        unsigned long long st = j >= 0 ? *(volatile unsigned long long*)&status[j] : CROP_PREFIX;   // nothing before tile 0
        while (__any_sync(0xFFFFFFFFu, (st >> 62) == 0ull))                  // predecessors are running (ticket order)
          if ((st >> 62) == 0ull) st = *(volatile unsigned long long*)&status[j];
        const unsigned int has_prefix = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2ull);
        const int stop = has_prefix ? __ffs(has_prefix) - 1 : 31;
        const unsigned int part = __reduce_add_sync(0xFFFFFFFFu, lane <= stop ? (unsigned int)(st & CROP_MASK) : 0u);
        prefix += part;
        if (has_prefix) break;
        j0 -= 32;
      }
      if (lane == 0) atomicExch(&status[tile], CROP_PREFIX | (prefix + (unsigned long long)tile_total));
    }
    if (lane == 0) {
      s_cnt[CROP_ROWS * 8] = (int)tile_total;
      s_base = prefix;
      if (base + CROP_TILE >= n) *total = prefix + (unsigned long long)tile_total;     // the last tile knows the answer
    }
  }
  __syncthreads();
  const unsigned long long obase = s_base;
#pragma unroll
  for (int r = 0; r < CROP_ROWS; ++r)
    if (keep & (1u << r)) out[obase + (unsigned long long)(s_cnt[r * 8 + w] + rank[r])] = p[r];
}

// ---- the same pass as a persistent, TMA-fed pipeline (sm_100a) ---------------------------------------------------------
// k_crop_box issues eight 16-byte loads per thread and then waits for them: between two tiles of a CTA nothing is in flight,
// and the kernel reaches 55 % of the copy peak (ncu, round 1).  Here two CTAs per SM stay resident, every CTA owns a
// CONTIGUOUS run of tiles and keeps CROP_STAGES tiles of 32 KB in flight with the bulk-copy engine: one elected thread arms
// an mbarrier with the byte count (mbarrier.arrive.expect_tx) and issues cp.async.bulk.shared::cluster.global (SASS:
// UBLKCP), the 256 threads sleep on the barrier (try_wait), read their eight rows from shared memory (conflict-free
// 16-byte rows) and hand the stage back.  A CTA appends what it keeps to its own stretch of a staging buffer -- no
// inter-CTA dependency in the streaming pass (a first version kept the chained scan and a tile ticket inside the
// persistent loop: the look-back latency of every tile sat on its CTA's critical path, 97 us instead of 48) -- and
// k_crop_gather then moves the stretches (2 % of the map) to their final places in input order.  No reset kernel, and
// the total goes to mapped host memory instead of a copy-back.
#ifndef CROP_STAGES
#define CROP_STAGES 3          // tiles in flight per CTA
#endif
#ifndef CROP_CTAS
#define CROP_CTAS 2            // resident CTAs per SM (CROP_STAGES x 32 KB of shared memory each)
#endif

__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned int bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int parity) {
  unsigned int ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

// CTA c: tiles [c * tiles_per_cta, (c + 1) * tiles_per_cta); kept points -> stash[first point of the run + rank], count -> counts[c]
__global__ void __launch_bounds__(256, CROP_CTAS) k_crop_box_bulk(const float4* __restrict__ pts, long long n, CropBox9 box, float4* __restrict__ stash,
                                                          unsigned int* __restrict__ counts, int tiles_per_cta) {
  extern __shared__ __align__(128) unsigned char dyn_smem[];
  float4* stage_buf = reinterpret_cast<float4*>(dyn_smem);
  __shared__ __align__(8) unsigned long long mbar[CROP_STAGES];
  __shared__ int s_cnt[2][CROP_ROWS * 8];
  __shared__ unsigned int s_base[2];
  const long long n_tiles = (n + CROP_TILE - 1) / CROP_TILE;
  const long long t0 = (long long)blockIdx.x * tiles_per_cta;
  long long t1 = t0 + tiles_per_cta;
  if (t1 > n_tiles) t1 = n_tiles;
  const int my_tiles = t1 > t0 ? (int)(t1 - t0) : 0;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float4* my_out = stash + t0 * CROP_TILE;
  // elected thread: start the copy of this CTA's k-th tile into stage k % CROP_STAGES
  auto issue = [&](int k) {
    if (k >= my_tiles) return;
    const int s = k % CROP_STAGES;
    const long long base = (t0 + k) * CROP_TILE;
    const long long left = n - base;
    const unsigned int bytes = (unsigned int)((left < CROP_TILE ? left : (long long)CROP_TILE) * 16);
    mbar_expect_tx(&mbar[s], bytes);
    bulk_load(stage_buf + (size_t)s * CROP_TILE, pts + base, bytes, &mbar[s]);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < CROP_STAGES; ++s) mbar_init(&mbar[s], 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int k = 0; k < CROP_STAGES; ++k) issue(k);
  }
  __syncthreads();
  unsigned int running = 0;                                // warp 0: points this CTA has kept so far
  for (int k = 0; k < my_tiles; ++k) {
    const int s = k % CROP_STAGES, b2 = k & 1;
    mbar_wait(&mbar[s], (unsigned int)(k / CROP_STAGES) & 1u);
    const long long base = (t0 + k) * CROP_TILE;
    const float4* buf = stage_buf + (size_t)s * CROP_TILE;
    float4 p[CROP_ROWS];
#pragma unroll
    for (int r = 0; r < CROP_ROWS; ++r)
      p[r] = base + r * 256 + threadIdx.x < n ? buf[r * 256 + threadIdx.x] : make_float4(NAN, NAN, NAN, 0.f);
    unsigned int keep = 0;
    int rank[CROP_ROWS];
#pragma unroll
    for (int r = 0; r < CROP_ROWS; ++r) {
      bool in = false;
      const float dx = __fsub_rn(p[r].x, box.t[0]), dy = __fsub_rn(p[r].y, box.t[1]), dz = __fsub_rn(p[r].z, box.t[2]);
      if (fabsf(dx) <= box.reach && fabsf(dy) <= box.reach && fabsf(dz) <= box.reach) {      // see k_crop_box
        const float lx = __fadd_rn(__fadd_rn(__fmul_rn(box.m[0], dx), __fmul_rn(box.m[1], dy)), __fmul_rn(box.m[2], dz));
        const float ly = __fadd_rn(__fadd_rn(__fmul_rn(box.m[3], dx), __fmul_rn(box.m[4], dy)), __fmul_rn(box.m[5], dz));
        const float lz = __fadd_rn(__fadd_rn(__fmul_rn(box.m[6], dx), __fmul_rn(box.m[7], dy)), __fmul_rn(box.m[8], dz));
        in = !(lx < box.bmin || ly < box.bmin || lz < box.bmin || lx > box.bmax || ly > box.bmax || lz > box.bmax);
      }
      const unsigned int bal = __ballot_sync(0xFFFFFFFFu, in);
      rank[r] = __popc(bal & ((1u << lane) - 1u));
      if (in) keep |= 1u << r;
      if (lane == 0) s_cnt[b2][r * 8 + w] = __popc(bal);
    }
    __syncthreads();                                       // the rows are in registers: stage s is free again
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(k + CROP_STAGES);
    }
    if (w == 0) {
      // exclusive scan of the 64 (row, warp) counts, row-major = input order
      const int a = s_cnt[b2][2 * lane], b = s_cnt[b2][2 * lane + 1];
      int incl = a + b;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += o;
      }
      const int excl = incl - (a + b);
      s_cnt[b2][2 * lane] = excl; s_cnt[b2][2 * lane + 1] = excl + a;
      if (lane == 0) s_base[b2] = running;
      running += (unsigned int)__shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    __syncthreads();                                       // s_cnt / s_base are double-buffered: the next tile writes the other set
    const unsigned int obase = s_base[b2];
#pragma unroll
    for (int r = 0; r < CROP_ROWS; ++r)
      if (keep & (1u << r)) my_out[obase + (unsigned int)(s_cnt[b2][r * 8 + w] + rank[r])] = p[r];
  }
  if (threadIdx.x == 0) counts[blockIdx.x] = running;
}

// CTA c moves its stretch of the stash to out[sum of the counts before c ...]: the output keeps the input order
__global__ void __launch_bounds__(256) k_crop_gather(const float4* __restrict__ stash, const unsigned int* __restrict__ counts, int n_ctas,
                                                     int tiles_per_cta, float4* __restrict__ out, unsigned long long* total,
                                                     volatile unsigned long long* total_host) {
  __shared__ unsigned long long s_warp[8];
  __shared__ unsigned long long s_prefix, s_total;
  unsigned long long before = 0, all = 0;
  for (int j = threadIdx.x; j < n_ctas; j += 256) {
    const unsigned int c = __ldg(&counts[j]);
    all += c;
    if (j < (int)blockIdx.x) before += c;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    before += __shfl_xor_sync(0xFFFFFFFFu, before, off);
    all += __shfl_xor_sync(0xFFFFFFFFu, all, off);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) s_warp[w] = before;
  __syncthreads();
  if (threadIdx.x == 0) { unsigned long long t = 0; for (int k = 0; k < 8; ++k) t += s_warp[k]; s_prefix = t; }
  __syncthreads();
  if (lane == 0) s_warp[w] = all;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int k = 0; k < 8; ++k) t += s_warp[k];
    s_total = t;
    if (blockIdx.x == 0) { *total = t; *total_host = t; __threadfence_system(); }
  }
  __syncthreads();
  const unsigned int mine = __ldg(&counts[blockIdx.x]);
  const float4* src = stash + (long long)blockIdx.x * tiles_per_cta * CROP_TILE;
  float4* dst = out + s_prefix;
  for (unsigned int j = threadIdx.x; j < mine; j += 256) dst[j] = __ldg(&src[j]);
}

int run_crop_box(Handle* h, const float4* pts, int64_t n, float bmin, float bmax, const float* rpy, const float* translation,
                 float4* out_dev, int64_t* n_out) {
  cudaStream_t s = h->stream;
  *n_out = 0;
  if (n == 0) return AICP_B200_OK;
  // pcl::getTransformation(0,0,0,roll,pitch,yaw) in float (host libm, the same calls as the oracle), then the transpose
  const float A = cosf(rpy[2]), B = sinf(rpy[2]), C = cosf(rpy[1]), D = sinf(rpy[1]), E = cosf(rpy[0]), F = sinf(rpy[0]);
  const float DE = D * E, DF = D * F;
  const float R[9] = {A * C, A * DF - B * E, B * F + A * DE, B * C, A * E + B * DF, B * DE - A * F, -D, C * F, C * E};
  CropBox9 box;
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) box.m[3 * r + c] = R[3 * c + r];
  for (int d = 0; d < 3; ++d) box.t[d] = translation[d];
  box.bmin = bmin; box.bmax = bmax;
  box.reach = 1.7320508f * fmaxf(fabsf(bmin), fabsf(bmax)) * 1.0001f;
  if (!(box.reach < INFINITY)) box.reach = 3.0e38f;      // finite, so that non-finite coordinates are still rejected
  const long long n_tiles = (n + CROP_TILE - 1) / CROP_TILE;
  if (!h->crop_legacy && (reinterpret_cast<uintptr_t>(pts) & 15u) == 0) {
    // persistent TMA-fed kernel + gather; the per-handle control block and the mapped total are set up once
    if (!h->crop_total_host) {
      CUDA_TRY(cudaHostAlloc((void**)&h->crop_total_host, sizeof(unsigned long long), cudaHostAllocMapped));
      CUDA_TRY(cudaHostGetDevicePointer((void**)&h->crop_total_dev, (void*)h->crop_total_host, 0));
      CUDA_TRY(cudaFuncSetAttribute(k_crop_box_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, CROP_STAGES * CROP_TILE * 16));
      CUDA_TRY(cudaDeviceGetAttribute(&h->n_sm, cudaDevAttrMultiProcessorCount, h->device));
    }
    long long grid = (long long)CROP_CTAS * h->n_sm;
    if (grid > n_tiles) grid = n_tiles;
    const int tiles_per_cta = (int)((n_tiles + grid - 1) / grid);
    grid = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    CUDA_TRY(h->crop_stash.reserve((size_t)n_tiles * CROP_TILE));
    CUDA_TRY(h->crop_status.reserve((size_t)grid + 2));
    unsigned int* counts = reinterpret_cast<unsigned int*>(h->crop_status.p + 1);
    unsigned long long* total = h->crop_status.p;
    k_crop_box_bulk<<<(unsigned)grid, 256, CROP_STAGES * CROP_TILE * 16, s>>>(pts, (long long)n, box, h->crop_stash.p, counts, tiles_per_cta);
    k_crop_gather<<<(unsigned)grid, 256, 0, s>>>(h->crop_stash.p, counts, (int)grid, tiles_per_cta, out_dev, total, h->crop_total_dev);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaStreamSynchronize(s));
    h->launches += 2;
    *n_out = (int64_t)*h->crop_total_host;
    return AICP_B200_OK;
  }
  CUDA_TRY(h->crop_status.reserve((size_t)n_tiles + 2));
  unsigned long long* status = h->crop_status.p;
  unsigned long long* total = status + n_tiles;
  unsigned int* ticket = reinterpret_cast<unsigned int*>(status + n_tiles + 1);
  k_crop_reset<<<(unsigned)((n_tiles + 255) / 256), 256, 0, s>>>(status, (int)n_tiles, ticket, total);
  k_crop_box<<<(unsigned)n_tiles, 256, 0, s>>>(pts, (long long)n, box, out_dev, status, ticket, total);
  CUDA_TRY(cudaGetLastError());
  unsigned long long host_total = 0;
  CUDA_TRY(cudaMemcpyAsync(&host_total, total, sizeof(host_total), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  h->launches += 2;
  *n_out = (int64_t)host_total;
  return AICP_B200_OK;
}

}  // namespace aicp
