// Layout probe for a kind::f16 (bfloat16) tcgen05.mma in TS form: A (M = 128, K = 16) from TENSOR MEMORY as packed bf16
// pairs (8 32-bit columns), B (N x 16) from shared memory, K-major, no swizzle (8 x 16-byte core matrices).
// Question answered: which half of a 32-bit TMEM column holds the even k, and which of (LBO, SBO) is the K-direction
// stride of the core matrices.  Integer-valued operands, so the expected D is exact.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bf16_probe bf16_probe.cu && ./bf16_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../dmd_era5_b200/csrc/tc_common.cuh"
using namespace era5svd::tc;

__device__ __forceinline__ void umma_bf16_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

__host__ __device__ inline int aval(int r, int k) { return (r * 3 + k * 5) % 7 + 1; }
__host__ __device__ inline int bval(int n, int k) { return (n * 2 + k * 3) % 5 + 1; }

// low_even: 1 -> k even in the low 16 bits of a column; lbo / sbo in bytes; N = 16 .. 128
__global__ void __launch_bounds__(128, 1) probe(int low_even, int lbo, int sbo, int N, int swap_kn, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&slot), 256);
  // B[n][k] (bf16) at (n / 8) * s_n + (k / 8) * s_k + (n % 8) * 16 + (k % 8) * 2 with (s_k, s_n) = (lbo, sbo) or swapped
  const int s_k = swap_kn ? sbo : lbo, s_n = swap_kn ? lbo : sbo;
  for (int i = threadIdx.x; i < N * 16; i += 128) {
    const int n = i / 16, k = i % 16;
    *reinterpret_cast<__nv_bfloat16*>(sm + (n / 8) * s_n + (k / 8) * s_k + (n % 8) * 16 + (k % 8) * 2) = __float2bfloat16((float)bval(n, k));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  const uint32_t tm = slot;
  {
    const int r = threadIdx.x;
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) {
      const uint32_t e = __bfloat16_as_ushort(__float2bfloat16((float)aval(r, 2 * j)));
      const uint32_t o = __bfloat16_as_ushort(__float2bfloat16((float)aval(r, 2 * j + 1)));
      v[j] = low_even ? (e | (o << 16)) : (o | (e << 16));
    }
    tmem_st8(tm + 128 + ((uint32_t)(warp * 32) << 16), v);
    tmem_wait_st();
  }
  tcgen05_fence_before(); __syncthreads(); tcgen05_fence_after();
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t b = make_smem_desc(base, (uint32_t)lbo, (uint32_t)sbo, 0 /* no swizzle */);
    umma_bf16_ts(tm, tm + 128, b, idesc, 0);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tcgen05_fence_after();
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tm + c0 + ((uint32_t)(warp * 32) << 16), v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) out[(size_t)threadIdx.x * 128 + c0 + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 256);
}

int main() {
  float* d_out; cudaMalloc(&d_out, 128 * 128 * 4);
  float* h = (float*)malloc(128 * 128 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int N : {16, 112})
    for (int low_even = 0; low_even < 2; ++low_even)
      for (int swap = 0; swap < 2; ++swap) {
        // core matrices of 128 B; K direction: 2 of them, N direction: N / 8
        const int lbo = 128, sbo = 256;                  // descriptor fields; `swap` says which one the DATA uses for K
        cudaMemset(d_out, 0, 128 * 128 * 4);
        probe<<<1, 128, 64 * 1024>>>(low_even, lbo, sbo, N, swap, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d_out, 128 * 128 * 4, cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int r = 0; r < 128; ++r)
          for (int n = 0; n < N; ++n) {
            int want = 0;
            for (int k = 0; k < 16; ++k) want += aval(r, k) * bval(n, k);
            bad += h[r * 128 + n] != (float)want;
          }
        printf("N=%3d  TMEM pair low half = %s k,  data K-stride = %s (desc LBO=128 SBO=256): mismatches %ld of %d  D[0][0..3] = %g %g %g %g  (%s)\n",
               N, low_even ? "even" : "odd ", swap ? "SBO" : "LBO", bad, 128 * N, h[0], h[1], h[2], h[3], cudaGetErrorString(e));
      }
  return 0;
}
