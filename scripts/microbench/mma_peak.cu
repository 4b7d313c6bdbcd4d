// Microbenchmark: issue-rate / throughput of tcgen05.mma for kind::tf32 and kind::f16 (bf16) on one SM and
// on the whole chip.  Operands are whatever is in shared memory (values irrelevant).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_peak mma_peak.cu && ./mma_peak
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dmd_era5_b200/csrc/tc_common.cuh"
using namespace era5svd::tc;

__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode 0: tf32 SS, 1: tf32 TS, 2: bf16 SS
__global__ void __launch_bounds__(128, 1) peak_kernel(int mode, int N, int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = slot;
  if (warp == 0 && lane == 0) {
    uint32_t idesc;
    if (mode == 2) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    else idesc = make_idesc_tf32(128, N, 0, 0);
    const uint64_t a = make_smem_desc(base, 16, 1024), b = make_smem_desc(base + 32768, 16, 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tm + (uint32_t)((i & 1) * 256);      // two accumulators, alternate (no RAW chain)
      if (mode == 0) umma_tf32_ss(d, a, b, idesc, 1);
      else if (mode == 1) umma_tf32_ts(d, tm + 256 + 128, b, idesc, 1);
      else umma_f16_ss(d, a, b, idesc, 1);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) cycles[0] = t1 - t0;
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  long long* d_c; cudaMalloc(&d_c, 8);
  cudaFuncSetAttribute(peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[3] = {"tf32 SS", "tf32 TS", "bf16 SS"};
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  for (int grid : {1, 148})
    for (int mode = 0; mode < 3; ++mode)
      for (int N : {256, 128, 112, 64}) {
        const int iters = 20000;
        peak_kernel<<<grid, 128, 100 * 1024>>>(mode, N, 100, d_c);   // warm
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        peak_kernel<<<grid, 128, 100 * 1024>>>(mode, N, iters, d_c);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
        const int K = mode == 2 ? 16 : 8;
        double flops = 2.0 * 128 * N * K * (double)iters * grid;
        printf("grid %3d %-8s N=%3d: %7.1f clk/MMA, %8.3f ms, %8.1f TFLOP/s  (err=%s)\n", grid, names[mode], N,
               (double)c / iters, ms, flops / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
