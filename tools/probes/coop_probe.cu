// coop_probe.cu -- do cooperative (grid-synchronising) kernels on DIFFERENT streams of one GPU run concurrently?
// Each kernel: G blocks x 256 threads, R rounds of {spin ~20 us, grid barrier}.  Prints the time of 1 kernel alone and of
// S kernels launched on S streams; concurrent execution gives about the same time, serialised execution S times it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coop_probe coop_probe.cu && ./coop_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256, 4) k_spin(unsigned int* bar, int rounds, long long spin_ns, int* out) {
  unsigned int epoch = 0;
  for (int r = 0; r < rounds; ++r) {
    if (threadIdx.x == 0) {
      unsigned long long t0, t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while ((long long)(t - t0) < spin_ns);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      ++epoch;
      __threadfence();
      atomicAdd(bar, 1u);
      while (ld_acquire(bar) < epoch * gridDim.x) {}
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = rounds;
}

int main() {
  int dev = 0, coop = 0, sms = 0;
  cudaSetDevice(dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spin, 256, 0);
  printf("cooperative launch %d, %d SMs, %d blocks/SM\n", coop, sms, occ);
  const int S = 8;
  cudaStream_t st[S];
  unsigned int* bars; int* outs;
  cudaMalloc(&bars, S * 256); cudaMalloc(&outs, S * sizeof(int));
  for (int i = 0; i < S; ++i) cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  int rounds = 50; long long spin = 20000;
  for (int grid : {74, 148, 296, 592}) {
    for (int ns : {1, 2, 4, 8}) {
      if ((long long)grid * ns > (long long)occ * sms) { printf("grid %4d x %d streams: exceeds co-residency, skipped\n", grid, ns); continue; }
      cudaMemset(bars, 0, S * 256);
      cudaDeviceSynchronize();
      cudaEventRecord(e0, 0);
      for (int i = 0; i < ns; ++i) cudaStreamWaitEvent(st[i], e0, 0);
      for (int i = 0; i < ns; ++i) {
        unsigned int* b = bars + i * 64; int* o = outs + i;
        void* args[] = {&b, &rounds, &spin, &o};
        cudaError_t e = cudaLaunchCooperativeKernel((void*)k_spin, dim3(grid), dim3(256), args, 0, st[i]);
        if (e != cudaSuccess) printf("launch failed: %s\n", cudaGetErrorString(e));
      }
      cudaEvent_t done[S];
      for (int i = 0; i < ns; ++i) { cudaEventCreate(&done[i]); cudaEventRecord(done[i], st[i]); cudaStreamWaitEvent(0, done[i], 0); }
      cudaEventRecord(e1, 0);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
      printf("grid %4d x %d streams: %.3f ms (%s) -> %.2f us per round\n", grid, ns, ms, cudaGetErrorString(e), ms * 1e3 / rounds);
      for (int i = 0; i < ns; ++i) cudaEventDestroy(done[i]);
    }
  }
  // barrier cost alone: no spinning
  for (int grid : {74, 148, 296, 592}) {
    cudaMemset(bars, 0, S * 256);
    int r2 = 1000; long long s2 = 0;
    unsigned int* b = bars; int* o = outs;
    void* args[] = {&b, &r2, &s2, &o};
    cudaDeviceSynchronize();
    cudaEventRecord(e0, st[0]);
    cudaLaunchCooperativeKernel((void*)k_spin, dim3(grid), dim3(256), args, 0, st[0]);
    cudaEventRecord(e1, st[0]);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("barrier only, grid %4d: %.3f us per barrier\n", grid, ms * 1e3 / r2);
  }
  return 0;
}
