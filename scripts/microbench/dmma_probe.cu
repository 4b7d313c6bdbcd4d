// FP64 tensor-core (mma.sync.m8n8k4.f64) issue-rate probe: TFLOP/s against warps per SM and independent accumulator
// chains per warp, operands from registers or re-loaded from shared memory per instruction (like the DMMA passes of
// gemm_simt.cu: 2 A + 14 B fragment loads per 28 DMMAs with 8 warps, 2 A + 7 B per 14 DMMAs with 16 warps).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench/dmma_probe scripts/microbench/dmma_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

template <int CH, bool LDS>
__global__ void probe(int iters, double* sink) {
  __shared__ double sb[4][120];
  if (threadIdx.x < 120) for (int i = 0; i < 4; ++i) sb[i][threadIdx.x] = 1.0 + 1e-9 * threadIdx.x;
  __syncthreads();
  const int lane = threadIdx.x % 32, g = lane / 4, tig = lane % 4;
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  double c[CH][2];
#pragma unroll
  for (int j = 0; j < CH; ++j) { c[j][0] = 0.0; c[j][1] = 0.0; }
  for (int i = 0; i < iters; ++i) {
    if (LDS) {
      const double a0 = sb[tig][g + (i & 7)], a1 = sb[tig][8 + g + (i & 7)];
#pragma unroll
      for (int j = 0; j < CH / 2; ++j) {
        const double bb = sb[tig][j * 8 + g];
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(c[2 * j][0]), "+d"(c[2 * j][1]) : "d"(a0), "d"(bb));
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(c[2 * j + 1][0]), "+d"(c[2 * j + 1][1]) : "d"(a1), "d"(bb));
      }
    } else {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0.0;
#pragma unroll
  for (int j = 0; j < CH; ++j) s += c[j][0] + c[j][1];
  if (s == 12345.678) sink[0] = s;
}

template <int CH, bool LDS>
static void run(int threads, int sms, double* sink) {
  const int iters = 40000 / CH * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<CH, LDS><<<sms, threads>>>(100, sink);
  cudaEventRecord(e0);
  probe<CH, LDS><<<sms, threads>>>(iters, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double tf = 2.0 * 256 * (double)CH * iters * (threads / 32) * sms / (ms * 1e-3) / 1e12;
  printf("warps/SM %2d  chains %2d  operands %-9s  %6.2f ms  %6.2f TFLOP/s  (%s)\n", threads / 32, CH, LDS ? "smem" : "registers", ms, tf,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* sink;
  cudaMalloc(&sink, 8);
  for (int threads : {128, 256, 512, 1024}) {
    run<8, false>(threads, sms, sink);
    run<14, false>(threads, sms, sink);
    run<28, false>(threads, sms, sink);
    run<14, true>(threads, sms, sink);
    run<28, true>(threads, sms, sink);
  }
  return 0;
}
