// tcgen05 / TMA path for the tall GEMM passes (3xTF32 split).  Placeholder until the kernels land:
// the entry points exist so that the ABI is stable, and fail loudly.
#include "common.cuh"

namespace era5svd {

int sketch_tf32x3(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, float*,
                  int64_t, cudaStream_t) {
  set_error("sketch: TF32X3 path not built into this library");
  return ERA5SVD_ERR_UNSUPPORTED;
}

int project_tf32x3(const float*, int64_t, int64_t, int64_t, const float*, int64_t, int64_t, double*,
                   int64_t, int, void*, size_t, cudaStream_t) {
  set_error("project: TF32X3 path not built into this library");
  return ERA5SVD_ERR_UNSUPPORTED;
}

size_t project_tf32x3_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }

}  // namespace era5svd
