// sort.cu -- least-significant-digit radix sort of (30-bit Morton key, 32-bit index) pairs, hand-written for this library
// (it replaces the CUB DeviceRadixSort call the index build started with; same launch count, no library kernel left on the
// path).
//
//   k_radix_hist   one pass over the keys: the four 256-bin digit histograms (shared-memory atomics, flushed per block)
//   k_radix_pass   x4, one launch per 8-bit digit, a single sweep over the data ("onesweep"): a tile of 2048 pairs per
//                  block; inside a warp equal digits find each other with __match_any_sync (rank among equal digits =
//                  popcount of the lower lanes, stable), per-warp digit counters in shared memory order the warps, and
//                  the tile's base offset per digit comes from a chained scan over the tiles with decoupled look-back --
//                  256 chains, one per digit and thread.  Tiles take their number from an atomic ticket, so a tile only
//                  waits for tiles that are already running.
// Stable (LSD needs it): ranks preserve the input order within a digit at every level (lane < lane, item < item,
// warp < warp, tile < tile).  Algorithmic bytes: 8 B read for the histograms + 4 x (8 read + 8 written) per pair.
#include "handle.cuh"

namespace aicp {

#define RS_ITEMS 8
#define RS_WARPS 8
#define RS_TILE (RS_WARPS * 32 * RS_ITEMS)
#define RS_AGG (1u << 30)
#define RS_PREFIX (2u << 30)
#define RS_MASK ((1u << 30) - 1u)

__global__ void __launch_bounds__(256) k_radix_hist(const unsigned int* __restrict__ keys, int n, unsigned int* ghist) {
  __shared__ unsigned int sh[4 * 256];
  for (int b = threadIdx.x; b < 4 * 256; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned int k = __ldg(&keys[i]);
    atomicAdd(&sh[k & 255u], 1u);
    atomicAdd(&sh[256 + ((k >> 8) & 255u)], 1u);
    atomicAdd(&sh[512 + ((k >> 16) & 255u)], 1u);
    atomicAdd(&sh[768 + (k >> 24)], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 4 * 256; b += blockDim.x)
    if (sh[b]) atomicAdd(&ghist[b], sh[b]);
}

__global__ void __launch_bounds__(256) k_radix_pass(const unsigned int* __restrict__ keys_in, const unsigned int* __restrict__ vals_in,
                                                    unsigned int* __restrict__ keys_out, unsigned int* __restrict__ vals_out, int n,
                                                    int shift, const unsigned int* __restrict__ ghist, unsigned int* status,
                                                    unsigned int* ticket) {
  __shared__ unsigned int whist[RS_WARPS][256];
  __shared__ unsigned int tile_off[256];
  __shared__ unsigned int s_warp_sum[8];
  __shared__ unsigned int s_tile;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
  for (int b = threadIdx.x; b < RS_WARPS * 256; b += 256) (&whist[0][0])[b] = 0;
  __syncthreads();
  const unsigned int tile = s_tile;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long base = (long long)tile * RS_TILE + w * (32 * RS_ITEMS);
  unsigned int key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const long long i = base + j * 32 + lane;
    key[j] = i < n ? __ldg(&keys_in[i]) : 0xFFFFFFFFu;
    val[j] = i < n ? __ldg(&vals_in[i]) : 0u;
  }
  // rank of every pair among the pairs of this WARP with the same digit, in input order (item, then lane)
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    const bool active = base + j * 32 + lane < n;
    const unsigned int d = active ? ((key[j] >> shift) & 255u) : 256u;           // padding forms its own group
    const unsigned int m = __match_any_sync(0xFFFFFFFFu, d);
    const int leader = __ffs(m) - 1;
    unsigned int old = 0;
    if (active && lane == leader) { old = whist[w][d]; whist[w][d] = old + (unsigned int)__popc(m); }
    old = __shfl_sync(0xFFFFFFFFu, old, leader);
    rank[j] = old + (unsigned int)__popc(m & ((1u << lane) - 1u));
    __syncwarp();
  }
  __syncthreads();
  // thread d: order the warps for digit d, then the tiles (chained scan with decoupled look-back), then the digits
  {
    const int d = threadIdx.x;
    unsigned int running = 0;
#pragma unroll
    for (int ww = 0; ww < RS_WARPS; ++ww) { const unsigned int c = whist[ww][d]; whist[ww][d] = running; running += c; }
    const unsigned int tile_count = running;
    unsigned int* st = status + d;                                   // status[tile * 256 + d]
    unsigned int prefix = 0;
    if (tile == 0) {
      atomicExch(&st[0], RS_PREFIX | tile_count);
    } else {
      atomicExch(&st[(size_t)tile * 256], RS_AGG | tile_count);
      long long t = (long long)tile - 1;
      while (true) {
        // four predecessors per step: the loads are independent, so one L2 round trip serves up to four hops of the walk.
        // Tile 0 only ever publishes a PREFIX, which ends the walk, so a word below tile 0 (read as 0 = "not published") is
        // never consumed: s1 is looked at only after s0 was an aggregate, i.e. t >= 1, and so on.
        const unsigned int s0 = *(volatile unsigned int*)&st[(size_t)t * 256];
        const unsigned int s1 = t >= 1 ? *(volatile unsigned int*)&st[(size_t)(t - 1) * 256] : 0u;
        const unsigned int s2 = t >= 2 ? *(volatile unsigned int*)&st[(size_t)(t - 2) * 256] : 0u;
        const unsigned int s3 = t >= 3 ? *(volatile unsigned int*)&st[(size_t)(t - 3) * 256] : 0u;
        if ((s0 >> 30) == 0u) continue;                              // predecessor running (ticket order): not published yet
        prefix += s0 & RS_MASK;
        if ((s0 >> 30) == 2u) break;
        if ((s1 >> 30) == 0u) { t -= 1; continue; }
        prefix += s1 & RS_MASK;
        if ((s1 >> 30) == 2u) break;
        if ((s2 >> 30) == 0u) { t -= 2; continue; }
        prefix += s2 & RS_MASK;
        if ((s2 >> 30) == 2u) break;
        if ((s3 >> 30) == 0u) { t -= 3; continue; }
        prefix += s3 & RS_MASK;
        if ((s3 >> 30) == 2u) break;
        t -= 4;
      }
      atomicExch(&st[(size_t)tile * 256], RS_PREFIX | (prefix + tile_count));
    }
    // exclusive scan of the global digit histogram over the 256 digits (every block, redundantly: 256 values)
    const unsigned int g = __ldg(&ghist[d]);
    unsigned int incl = g;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned int o = __shfl_up_sync(0xFFFFFFFFu, incl, off);
      if (lane >= off) incl += o;
    }
    if (lane == 31) s_warp_sum[w] = incl;
    __syncthreads();
    unsigned int wbase = 0;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) if (ww < w) wbase += s_warp_sum[ww];
    tile_off[d] = wbase + incl - g + prefix;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RS_ITEMS; ++j) {
    if (base + j * 32 + lane < n) {
      const unsigned int d = (key[j] >> shift) & 255u;
      const unsigned int pos = tile_off[d] + whist[w][d] + rank[j];
      keys_out[pos] = key[j];
      vals_out[pos] = val[j];
    }
  }
}

// Sorts (keys, vals) by the low 30 bits of the key.  Ping-pongs between (keys, vals) and (keys_alt, vals_alt); after the
// four passes the result is back in (keys, vals).  scratch: >= 4 * 256 + 4 + 4 * n_tiles * 256 words.
int radix_sort_pairs(Handle* h, unsigned int* keys, unsigned int* vals, unsigned int* keys_alt, unsigned int* vals_alt, int n,
                     DevBuf<unsigned int>& scratch) {
  cudaStream_t s = h->stream;
  const int n_tiles = (n + RS_TILE - 1) / RS_TILE;
  const size_t words = 4 * 256 + 4 + (size_t)4 * n_tiles * 256;
  CUDA_TRY(scratch.reserve(words));
  unsigned int* ghist = scratch.p;
  unsigned int* tickets = scratch.p + 4 * 256;
  unsigned int* status = scratch.p + 4 * 256 + 4;
  CUDA_TRY(cudaMemsetAsync(scratch.p, 0, sizeof(unsigned int) * words, s));
  int hb = (n + 255) / 256;
  if (hb > 148 * 4) hb = 148 * 4;
  k_radix_hist<<<hb, 256, 0, s>>>(keys, n, ghist);
  unsigned int *ki = keys, *vi = vals, *ko = keys_alt, *vo = vals_alt;
  for (int pass = 0; pass < 4; ++pass) {
    k_radix_pass<<<n_tiles, 256, 0, s>>>(ki, vi, ko, vo, n, 8 * pass, ghist + 256 * pass, status + (size_t)pass * n_tiles * 256, tickets + pass);
    unsigned int* t = ki; ki = ko; ko = t;
    t = vi; vi = vo; vo = t;
  }
  CUDA_TRY(cudaGetLastError());
  h->launches += 5;
  return AICP_B200_OK;
}

}  // namespace aicp
