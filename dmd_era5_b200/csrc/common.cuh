// Shared helpers for libera5svd (sm_100a).  Not part of the public ABI.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/era5svd.h"

namespace era5svd {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Check the launch that just happened (no sync).  Returns an era5svd_status.
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return ERA5SVD_ERR_CUDA;
  }
  count_launch();
  return ERA5SVD_OK;
}

#define ERA5SVD_REQUIRE(cond, ...)   \
  do {                               \
    if (!(cond)) {                   \
      era5svd::set_error(__VA_ARGS__); \
      return ERA5SVD_ERR_ARG;        \
    }                                \
  } while (0)

#define ERA5SVD_CUDA(call)                                                      \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      era5svd::set_error("%s failed: %s", #call, cudaGetErrorString(e_));       \
      return ERA5SVD_ERR_CUDA;                                                  \
    }                                                                           \
  } while (0)

inline size_t dtype_size(int dtype) { return dtype == ERA5SVD_F64 ? 8 : 4; }
inline bool valid_dtype(int dtype) { return dtype == ERA5SVD_F32 || dtype == ERA5SVD_F64; }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Number of SMs of the current device (cached per device).
int sm_count();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel, size): the attribute is sticky, so repeating
// the driver call on every launch is pure host overhead.  Returns a cudaError_t.
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes);

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace era5svd
