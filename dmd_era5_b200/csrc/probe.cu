// Measured tensor peaks for the roofline denominators ("measure TF32 and FP64 peaks on the box
// first"; MEASURED_PEAKS.json only has bf16 and HBM).  bench.py calls these once per run, outside the timed region.
//   era5svd_probe_tf32_tflops : every SM issues back-to-back tcgen05.mma kind::tf32 (M = 128, K = 8, width N, A from
//                               TMEM or shared memory) from one elected thread in the uniform datapath - the issue
//                               pattern of the tall kernels; dense TFLOP/s = 2 * 128 * N * 8 * instructions / time.
//   era5svd_probe_dmma_tflops : every warp issues independent mma.sync.m8n8k4.f64 chains (the FP64 tensor path of
//                               gemm_simt.cu); TFLOP/s = 2 * 8 * 8 * 4 * instructions / time.
// Operand values are whatever the memories hold; only the rate is measured.  Synchronous (cudaEvent timing inside).
#include "common.cuh"
#include "tc_common.cuh"

namespace era5svd {
namespace tc {

__global__ void __launch_bounds__(128, 1) probe_tf32_kernel(int form, int N, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = slot;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_idesc_tf32(128, N, 0, 0);
    const uint64_t a = make_smem_desc(base, 16, 1024), b = make_smem_desc(base + 32768, 16, 1024);
    const uint32_t at = tm + 448;
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {                      // operands walk a 4-step ring, two alternating accumulators
        const uint32_t d = tm + (uint32_t)((u & 1) * (N <= 224 ? 224 : 0));
        if (form == 0) umma_tf32_ss(d, a + 2 * u, b + 2 * u, idesc, 1);
        else umma_tf32_ts(d, at + u * 8, b + 2 * u, idesc, 1);
      }
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

__global__ void __launch_bounds__(512, 1) probe_dmma_kernel(int iters, double* sink) {
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j][0] = 0.0; c[j][1] = 0.0; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
  if (s == 12345.678) sink[0] = s;                        // keeps the chains alive
}

}  // namespace tc
}  // namespace era5svd

extern "C" int era5svd_probe_tf32_tflops(int form, int N, double seconds, double* tflops) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(tflops && (form == 0 || form == 1), "probe_tf32: bad arguments");
  ERA5SVD_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0, "probe_tf32: N must be a multiple of 16 in [16, 256]");
  const int smem = 100 * 1024;
  ERA5SVD_CUDA(cudaFuncSetAttribute(tc::probe_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int grid = sm_count();
  cudaEvent_t e0, e1;
  ERA5SVD_CUDA(cudaEventCreate(&e0));
  ERA5SVD_CUDA(cudaEventCreate(&e1));
  int iters = 20000;
  float ms = 0.f;
  for (int rep = 0; rep < 2; ++rep) {      // first round calibrates the instruction count to the requested duration
    tc::probe_tf32_kernel<<<grid, 128, smem>>>(form, N, 400);
    ERA5SVD_CUDA(cudaEventRecord(e0));
    tc::probe_tf32_kernel<<<grid, 128, smem>>>(form, N, iters);
    ERA5SVD_CUDA(cudaEventRecord(e1));
    ERA5SVD_CUDA(cudaEventSynchronize(e1));
    ERA5SVD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep == 0 && seconds > 0.0) {
      double want = seconds * 1e3 / (ms > 0.f ? ms : 1.f) * iters;
      if (want > 4e8) want = 4e8;
      if (want < 4000) want = 4000;
      iters = (int)(want / 4) * 4;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  int rc = check_launch("probe_tf32_kernel");
  if (rc) return rc;
  *tflops = 2.0 * 128 * N * 8 * (double)iters * grid / (ms * 1e-3) / 1e12;
  return ERA5SVD_OK;
}

extern "C" int era5svd_probe_dmma_tflops(double seconds, double* tflops) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(tflops, "probe_dmma: null pointer");
  double* sink = nullptr;
  ERA5SVD_CUDA(cudaMalloc(&sink, 8));
  const int grid = sm_count() * 2;
  cudaEvent_t e0, e1;
  ERA5SVD_CUDA(cudaEventCreate(&e0));
  ERA5SVD_CUDA(cudaEventCreate(&e1));
  int iters = 4000;
  float ms = 0.f;
  for (int rep = 0; rep < 2; ++rep) {
    tc::probe_dmma_kernel<<<grid, 512>>>(100, sink);
    ERA5SVD_CUDA(cudaEventRecord(e0));
    tc::probe_dmma_kernel<<<grid, 512>>>(iters, sink);
    ERA5SVD_CUDA(cudaEventRecord(e1));
    ERA5SVD_CUDA(cudaEventSynchronize(e1));
    ERA5SVD_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (rep == 0 && seconds > 0.0) {
      double want = seconds * 1e3 / (ms > 0.f ? ms : 1.f) * iters;
      if (want > 2e8) want = 2e8;
      if (want < 1000) want = 1000;
      iters = (int)want;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  int rc = check_launch("probe_dmma_kernel");
  if (rc) return rc;
  *tflops = 2.0 * 8 * 8 * 4 * 8.0 * (double)iters * (512 / 32) * grid / (ms * 1e-3) / 1e12;
  return ERA5SVD_OK;
}
