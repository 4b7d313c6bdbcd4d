// Probes of tcgen05.mma kind::tf32 behaviour that the tall kernels rely on (one CTA per SM, one issuing thread):
//   1. conversion rule fp32 -> tf32 of the operands (truncate or round?)
//   2. accumulator column base: is D at column 224 (not a multiple of N or of a power of two) legal?
//   3. cost per k-step of the issue patterns used by the sketch / project kernels, TS form (A in TMEM):
//        v2: 3 x N          (lo*hi, hi*lo, hi*hi into one accumulator)
//        v3: N2 = 2N + N    (A_hi x [B_hi | B_lo], A_lo x B_hi)
//      and of single instructions for a range of N.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../dmd_era5_b200/csrc/tc_common.cuh"
using namespace era5svd::tc;

// every row r of the A tile holds a_val(r) in all K positions, every row n of the B tile holds b_val(n):
// D[r][n] = 8 * a(r) * b(n) whatever the swizzle.
__global__ void __launch_bounds__(128, 1) value_kernel(int test, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  float* sm = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  // A tile at base (128 rows x 128 B), B tile at base + 32768 (256 rows x 128 B)
  const float odd = 1.0f + 1.0f / 2048 + 1.0f / 4096;      // between two tf32 values, above the midpoint
  for (int i = threadIdx.x; i < 128 * 32; i += 128) sm[i] = test == 0 ? odd : test == 1 ? (float)(i / 32 + 1) : (float)((i / 32) % 8 + 1);
  for (int i = threadIdx.x; i < 256 * 32; i += 128) sm[8192 + i] = test == 0 ? 1.0f : test == 1 ? (float)(i / 32 + 1) : (float)((i / 32) % 16 + 1);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = slot;
  // A copy in TMEM columns [480, 488)
  {
    uint32_t v[16];
    for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(test == 0 ? odd : test == 1 ? (float)(threadIdx.x + 1) : (float)(threadIdx.x % 8 + 1));
    tmem_st16(tm + 480 + ((uint32_t)(warp * 32) << 16), v);
    tmem_wait_st();
  }
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint64_t a = make_smem_desc(base, 16, 1024), b = make_smem_desc(base + 32768, 16, 1024);
  if (warp == 0 && elect_one()) {
    if (test == 0) {
      umma_tf32_ss(tm, a, b, make_idesc_tf32(128, 16, 0, 0), 0);            // SS at column 0
      umma_tf32_ts(tm + 16, tm + 480, b, make_idesc_tf32(128, 16, 0, 0), 0);  // TS at column 16
    } else if (test == 1) {
      umma_tf32_ts(tm, tm + 480, b, make_idesc_tf32(128, 224, 0, 0), 0);        // base 0
      umma_tf32_ts(tm + 224, tm + 480, b, make_idesc_tf32(128, 224, 0, 0), 0);  // base 224
    } else {
      // test 2: 64 k-steps of the merged pattern issued back to back into ONE accumulator:
      // D[:, 0:224) += A B (N = 224) ; D[:, 0:112) += A B (N = 112)
      for (int i = 0; i < 64; ++i) {
        umma_tf32_ts(tm, tm + 480, b, make_idesc_tf32(128, 224, 0, 0), i != 0);
        umma_tf32_ts(tm, tm + 480, b, make_idesc_tf32(128, 112, 0, 0), 1);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tcgen05_fence_after();
  const int ncol = test == 0 ? 32 : 448;
  for (int c0 = 0; c0 < ncol; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tm + c0 + ((uint32_t)(warp * 32) << 16), v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) out[(size_t)threadIdx.x * 448 + c0 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// pattern 0: single instruction of width N, alternating accumulators; 1: v2 (3 x N, one accumulator);
// 2: v3 (2N + N, one accumulator); form: 0 = SS, 1 = TS
__global__ void __launch_bounds__(128, 1) rate_kernel(int pattern, int form, int N, int iters, long long* cycles) {
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
    const uint32_t id1 = make_idesc_tf32(128, N, 0, 0), id2 = make_idesc_tf32(128, 2 * N <= 256 ? 2 * N : 256, 0, 0);
    const uint64_t a = make_smem_desc(base, 16, 1024), a2 = make_smem_desc(base + 16384, 16, 1024);
    const uint64_t b = make_smem_desc(base + 32768, 16, 1024);
    const uint32_t at = tm + 448;
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      // four k-steps per iteration; operands move through a 4-step ring (different addresses per k-step)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint64_t bb = b + 2 * u, aa = a + 2 * u, aa2 = a2 + 2 * u;
        const uint32_t t_hi = at + u * 8, t_lo = at + 32 + u * 8;
        if (pattern == 0) {
          const uint32_t d = tm + (uint32_t)((u & 1) * 224);
          if (form == 0) umma_tf32_ss(d, aa, bb, id1, 1); else umma_tf32_ts(d, t_hi, bb, id1, 1);
        } else if (pattern == 1) {
          if (form == 0) { umma_tf32_ss(tm, aa2, bb, id1, 1); umma_tf32_ss(tm, aa, bb, id1, 1); umma_tf32_ss(tm, aa, bb, id1, 1); }
          else { umma_tf32_ts(tm, t_lo, bb, id1, 1); umma_tf32_ts(tm, t_hi, bb, id1, 1); umma_tf32_ts(tm, t_hi, bb, id1, 1); }
        } else {
          if (form == 0) { umma_tf32_ss(tm, aa, bb, id2, 1); umma_tf32_ss(tm, aa2, bb, id1, 1); }
          else { umma_tf32_ts(tm, t_hi, bb, id2, 1); umma_tf32_ts(tm, t_lo, bb, id1, 1); }
        }
      }
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
  float* d_out; cudaMalloc(&d_out, 128 * 448 * 4);
  float* h = (float*)malloc(128 * 448 * 4);
  cudaFuncSetAttribute(value_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);

  value_kernel<<<1, 128, 100 * 1024>>>(0, d_out);
  cudaMemcpy(h, d_out, 128 * 448 * 4, cudaMemcpyDeviceToHost);
  printf("conversion probe: a = 1 + 2^-11 + 2^-12, b = 1, K = 8: truncation -> 8.000000, round-to-nearest -> 8.007812\n");
  printf("  SS: D = %.6f   TS: D = %.6f   (err=%s)\n", h[0], h[16], cudaGetErrorString(cudaGetLastError()));

  value_kernel<<<1, 128, 100 * 1024>>>(1, d_out);
  cudaMemcpy(h, d_out, 128 * 448 * 4, cudaMemcpyDeviceToHost);
  long bad0 = 0, bad1 = 0;
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < 224; ++n) {
      const float want = 8.0f * (r + 1) * (n + 1);
      bad0 += h[r * 448 + n] != want;
      bad1 += h[r * 448 + 224 + n] != want;
    }
  printf("accumulator base probe (N = 224): mismatches at column base 0: %ld, at column base 224: %ld  (err=%s)\n", bad0, bad1,
         cudaGetErrorString(cudaGetLastError()));

  for (int rep = 0; rep < 3; ++rep) {
    value_kernel<<<148, 128, 100 * 1024>>>(2, d_out);
    cudaMemcpy(h, d_out, 128 * 448 * 4, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int r = 0; r < 128; ++r)
      for (int n = 0; n < 224; ++n) bad += h[r * 448 + n] != 64.0f * 8.0f * (r % 8 + 1) * (n % 16 + 1) * (n < 112 ? 2 : 1);
    printf("merged-pattern chain (64 k-steps of N=224 then N=112 on one accumulator, back to back): mismatches %ld  (err=%s)\n",
           bad, cudaGetErrorString(cudaGetLastError()));
  }
  const char* forms[2] = {"SS", "TS"};
  const int iters = 20000;
  for (int form = 0; form < 2; ++form)
    for (int N : {16, 32, 64, 96, 112, 128, 160, 192, 224, 256}) {
      rate_kernel<<<148, 128, 100 * 1024>>>(0, form, N, 100, d_c);
      rate_kernel<<<148, 128, 100 * 1024>>>(0, form, N, iters, d_c);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
      printf("single %s N=%3d: %7.1f clk/MMA  (err=%s)\n", forms[form], N, (double)c / iters, cudaGetErrorString(cudaGetLastError()));
    }
  for (int form = 0; form < 2; ++form)
    for (int N : {112, 128})
      for (int pattern = 1; pattern <= 2; ++pattern) {
        rate_kernel<<<148, 128, 100 * 1024>>>(pattern, form, N, 100, d_c);
        rate_kernel<<<148, 128, 100 * 1024>>>(pattern, form, N, iters, d_c);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
        printf("k-step %s N=%3d pattern %s: %7.1f clk/k-step  (err=%s)\n", forms[form], N,
               pattern == 1 ? "v2 (3 x N)   " : "v3 (2N + N)  ", (double)c / iters, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
