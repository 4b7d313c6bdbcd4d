// svd_flip, u-based (sklearn/utils/extmath.py:964-972): the sign of each singular pair is chosen
// so that the entry of largest magnitude in the column of U is positive; numpy's argmax returns
// the FIRST maximum, so ties go to the lowest (global) row.  U is row-sharded across GPUs, so the
// decision is a MAXLOC reduction: per-block candidates -> per-rank candidates -> (all-gather) ->
// combine.  HBM-bound: one read of U for the scan, one read+write for the scaling.
#include <cstdint>

#include "common.cuh"

namespace era5svd {

struct Cand {
  double a;     // |value|
  int64_t row;  // global row
  double sgn;   // +1 / -1 / 0
};

__device__ __forceinline__ bool better(double a, int64_t row, double b, int64_t brow) {
  return a > b || (a == b && row < brow);
}

constexpr int AM_COLS = 32;   // columns per block-row of threads
constexpr int AM_ROWS = 8;    // thread rows
constexpr int AM_MAXBLOCKS = 592;   // 4 x 148

// grid.x = row chunks, grid.y = column groups of 32
template <typename T>
__global__ void __launch_bounds__(AM_COLS * AM_ROWS)
col_absmax_kernel(const T* __restrict__ U, int64_t m, int64_t k, int64_t ldu, int64_t row_offset,
                  int64_t rows_per_block, double* __restrict__ ws_a, int64_t* __restrict__ ws_row,
                  double* __restrict__ ws_sgn) {
  const int tx = threadIdx.x % AM_COLS, ty = threadIdx.x / AM_COLS;
  const int64_t c = (int64_t)blockIdx.y * AM_COLS + tx;
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(m, r_begin + rows_per_block);
  double best = -1.0, bsgn = 0.0;
  int64_t brow = INT64_MAX;
  if (c < k) {
    for (int64_t r = r_begin + ty; r < r_end; r += AM_ROWS) {
      double v = (double)U[r * ldu + c];
      double a = fabs(v);
      if (a > best) {  // rows increase within a thread: strict '>' keeps the first maximum
        best = a;
        brow = r;
        bsgn = (v > 0.0) ? 1.0 : ((v < 0.0) ? -1.0 : 0.0);
      }
    }
  }
  __shared__ double s_a[AM_ROWS][AM_COLS];
  __shared__ long long s_r[AM_ROWS][AM_COLS];
  __shared__ double s_s[AM_ROWS][AM_COLS];
  s_a[ty][tx] = best;
  s_r[ty][tx] = brow;
  s_s[ty][tx] = bsgn;
  __syncthreads();
  if (ty == 0 && c < k) {
    for (int i = 1; i < AM_ROWS; ++i)
      if (better(s_a[i][tx], s_r[i][tx], best, brow)) {
        best = s_a[i][tx];
        brow = s_r[i][tx];
        bsgn = s_s[i][tx];
      }
    int64_t o = (int64_t)blockIdx.x * k + c;
    ws_a[o] = best;
    ws_row[o] = (brow == INT64_MAX) ? INT64_MAX : brow + row_offset;
    ws_sgn[o] = bsgn;
  }
}

// Reduce R candidate sets [R x k] with the first-maximum rule: one warp per column, lanes stride
// over the candidate sets, shuffle tree at the end.
__global__ void __launch_bounds__(128)
maxloc_reduce_kernel(const double* __restrict__ a, const int64_t* __restrict__ row,
                     const double* __restrict__ sgn, int64_t R, int64_t k,
                     double* __restrict__ a_out, int64_t* __restrict__ row_out,
                     double* __restrict__ sgn_out) {
  const int lane = threadIdx.x % 32;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (c >= k) return;
  double best = -1.0, bsgn = 0.0;
  long long brow = INT64_MAX;
  for (int64_t i = lane; i < R; i += 32) {
    double ai = a[i * k + c];
    long long ri = row[i * k + c];
    if (better(ai, ri, best, brow)) {
      best = ai;
      brow = ri;
      bsgn = sgn[i * k + c];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double oa = __shfl_xor_sync(0xffffffffu, best, o);
    long long orow = __shfl_xor_sync(0xffffffffu, brow, o);
    double os = __shfl_xor_sync(0xffffffffu, bsgn, o);
    if (better(oa, orow, best, brow)) {
      best = oa;
      brow = orow;
      bsgn = os;
    }
  }
  if (lane == 0) {
    if (a_out) a_out[c] = best;
    if (row_out) row_out[c] = brow;
    sgn_out[c] = bsgn;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
scale_cols_kernel(T* __restrict__ U, int64_t m, int64_t k, int64_t ldu,
                  const double* __restrict__ scale) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t total = m * k;
  for (; idx < total; idx += stride) {
    int64_t r = idx / k, c = idx - r * k;
    T* p = U + r * ldu + c;
    *p = (T)((double)(*p) * scale[c]);
  }
}

// float32 U with 16-byte aligned rows (the tensor-core path's pitch): one float4 per thread, lanes along the row,
// no index divisions (the scalar kernel above spends a 64-bit division per element and reaches ~3.2 TB/s).
constexpr int SC_VX = 32;    // float4 columns per block row (k <= 128)
constexpr int SC_RY = 8;     // rows per block iteration
__global__ void __launch_bounds__(SC_VX * SC_RY)
scale_cols_f32v4_kernel(float* __restrict__ U, int64_t m, int kv, int64_t ldu, const double* __restrict__ scale) {
  const int vx = threadIdx.x % SC_VX, ry = threadIdx.x / SC_VX;
  if (vx >= kv) return;
  const float4 sc = make_float4((float)scale[4 * vx], (float)scale[4 * vx + 1], (float)scale[4 * vx + 2], (float)scale[4 * vx + 3]);
  const int64_t step = (int64_t)gridDim.x * SC_RY;
  int64_t r = (int64_t)blockIdx.x * SC_RY + ry;
  for (; r + 3 * step < m; r += 4 * step) {          // four independent 16-byte loads in flight per thread
    float4* p0 = reinterpret_cast<float4*>(U + r * ldu) + vx;
    float4* p1 = reinterpret_cast<float4*>(U + (r + step) * ldu) + vx;
    float4* p2 = reinterpret_cast<float4*>(U + (r + 2 * step) * ldu) + vx;
    float4* p3 = reinterpret_cast<float4*>(U + (r + 3 * step) * ldu) + vx;
    float4 a = *p0, b = *p1, c = *p2, d = *p3;
    *p0 = make_float4(a.x * sc.x, a.y * sc.y, a.z * sc.z, a.w * sc.w);
    *p1 = make_float4(b.x * sc.x, b.y * sc.y, b.z * sc.z, b.w * sc.w);
    *p2 = make_float4(c.x * sc.x, c.y * sc.y, c.z * sc.z, c.w * sc.w);
    *p3 = make_float4(d.x * sc.x, d.y * sc.y, d.z * sc.z, d.w * sc.w);
  }
  for (; r < m; r += step) {
    float4* p0 = reinterpret_cast<float4*>(U + r * ldu) + vx;
    float4 a = *p0;
    *p0 = make_float4(a.x * sc.x, a.y * sc.y, a.z * sc.z, a.w * sc.w);
  }
}

static int64_t absmax_blocks(int64_t m) {
  int64_t b = ceil_div(m, 256);
  if (b > AM_MAXBLOCKS) b = AM_MAXBLOCKS;
  if (b < 1) b = 1;
  return b;
}

}  // namespace era5svd

extern "C" {

size_t era5svd_col_absmax_workspace_bytes(int64_t m, int64_t k) {
  using namespace era5svd;
  if (m <= 0 || k <= 0) return 0;
  return (size_t)(absmax_blocks(m) * k) * (2 * sizeof(double) + sizeof(int64_t));
}

int era5svd_col_absmax(const void* U, int dtype, int64_t m, int64_t k, int64_t ldu,
                       int64_t row_offset, double* absmax, int64_t* row, double* sign,
                       void* workspace, size_t workspace_bytes, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(U && absmax && row && sign, "col_absmax: null pointer");
  ERA5SVD_REQUIRE(valid_dtype(dtype), "col_absmax: bad dtype");
  ERA5SVD_REQUIRE(m > 0 && k > 0 && ldu >= k, "col_absmax: bad shape");
  size_t need = era5svd_col_absmax_workspace_bytes(m, k);
  if (!workspace || workspace_bytes < need) {
    set_error("col_absmax: workspace too small (%zu < %zu)", workspace_bytes, need);
    return ERA5SVD_ERR_WORKSPACE;
  }
  const int64_t nb = absmax_blocks(m);
  const int64_t rpb = ceil_div(m, nb);
  double* ws_a = (double*)workspace;
  int64_t* ws_row = (int64_t*)(ws_a + nb * k);
  double* ws_sgn = (double*)(ws_row + nb * k);
  cudaStream_t st = as_stream(stream);
  dim3 grid((unsigned)nb, (unsigned)ceil_div(k, AM_COLS));
  if (dtype == ERA5SVD_F32)
    col_absmax_kernel<float><<<grid, AM_COLS * AM_ROWS, 0, st>>>((const float*)U, m, k, ldu, row_offset, rpb, ws_a, ws_row, ws_sgn);
  else
    col_absmax_kernel<double><<<grid, AM_COLS * AM_ROWS, 0, st>>>((const double*)U, m, k, ldu, row_offset, rpb, ws_a, ws_row, ws_sgn);
  int rc = check_launch("col_absmax_kernel");
  if (rc) return rc;
  maxloc_reduce_kernel<<<(unsigned)ceil_div(k, 4), 128, 0, st>>>(ws_a, ws_row, ws_sgn, nb, k, absmax, row, sign);
  return check_launch("maxloc_reduce_kernel");
}

int era5svd_maxloc_combine(const double* absmax, const int64_t* row, const double* sign, int64_t R,
                           int64_t k, double* sign_out, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(absmax && row && sign && sign_out, "maxloc_combine: null pointer");
  ERA5SVD_REQUIRE(R > 0 && k > 0, "maxloc_combine: bad shape");
  maxloc_reduce_kernel<<<(unsigned)ceil_div(k, 4), 128, 0, as_stream(stream)>>>(absmax, row, sign, R, k, nullptr, nullptr, sign_out);
  return check_launch("maxloc_reduce_kernel");
}

int era5svd_scale_cols(void* U, int dtype, int64_t m, int64_t k, int64_t ldu, const double* scale,
                       void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(U && scale, "scale_cols: null pointer");
  ERA5SVD_REQUIRE(valid_dtype(dtype), "scale_cols: bad dtype");
  ERA5SVD_REQUIRE(m > 0 && k > 0 && ldu >= k, "scale_cols: bad shape");
  cudaStream_t st = as_stream(stream);
  int64_t blocks = ceil_div(m * k, 256 * 8);
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (dtype == ERA5SVD_F32 && k % 4 == 0 && k <= 4 * SC_VX && ldu % 4 == 0 && (reinterpret_cast<uintptr_t>(U) & 15u) == 0) {
    int64_t vb = ceil_div(m, (int64_t)SC_RY * 4);
    if (vb > cap) vb = cap;
    if (vb < 1) vb = 1;
    scale_cols_f32v4_kernel<<<(unsigned)vb, SC_VX * SC_RY, 0, st>>>((float*)U, m, (int)(k / 4), ldu, scale);
    return check_launch("scale_cols_f32v4_kernel");
  }
  if (dtype == ERA5SVD_F32)
    scale_cols_kernel<float><<<(unsigned)blocks, 256, 0, st>>>((float*)U, m, k, ldu, scale);
  else
    scale_cols_kernel<double><<<(unsigned)blocks, 256, 0, st>>>((double*)U, m, k, ldu, scale);
  return check_launch("scale_cols_kernel");
}

}  // extern "C"
