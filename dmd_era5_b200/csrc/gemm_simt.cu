// (b) Tall GEMM passes, native-precision path: float64 on the FP64 tensor cores (mma.sync m8n8k4 DMMA),
// float32 by FFMA on the CUDA cores.
//
//   sketch :  Y[m x l]  = X[m x n] * Om[n x l]              (`A @ Q`,   sklearn/utils/extmath.py:378,383)
//   project:  Z[n x l] += X[m x n]^T * Y[m x l]  (float64)  (`A.T @ Q`, extmath.py:379; `Q.T @ M`, :606)
//
// This is the accuracy-first path (FP64 parity mode, and the checker for the tcgen05 path in
// gemm_tc.cu); a 128 x (16*TN) x 16 tiled GEMM, 256 threads, register prefetch of the next k-tile; the float32
// variant is register-tiled SIMT (8 x TN outputs per thread), the float64 variant feeds the same staged tiles to DMMA.  Both passes share one inner product core because
// both stage their operands k-major in shared memory:
//   sketch : As[k][row]  <- X[row][k]   (transposed on the way in)   Bs[k][col] <- Om[k][col]
//   project: As[k][time] <- X[k=row][time] (straight copy)           Bs[k][col] <- Y[k=row][col]
// project splits the row (k) range over blockIdx.z; partial tiles go to the workspace in the
// accumulate type and reduce_partials_kernel sums them in float64 in a fixed order.
#include <stdlib.h>

#include "comm.cuh"
#include "common.cuh"

namespace era5svd {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NTHREADS = 256;
constexpr int PAD = 4;

template <typename T, int TN>
struct Tile {
  static constexpr int BN = 16 * TN;
  T As[BK][BM + PAD];
  T Bs[BK][BN + PAD];     // the pad keeps the DMMA fragment loads (4 k-rows x 8 columns per warp) conflict free
};

// FP64 tensor-core product of one staged tile (float64 operands only): mma.sync.m8n8k4.f64.  Warp w owns rows
// [16 w, 16 w + 16) of the 128-row tile and every 8-column tile; per k4-step it loads 2 A and 2 TN B fragments
// (one double each) and issues 4 TN DMMAs.  Fragment layout (PTX ISA, m8n8k4 .f64): g = lane / 4, tig = lane % 4;
// a = A[g][tig], b = B[tig][g], c = C[g][2 tig + {0, 1}].
template <int TN>
__device__ __forceinline__ void dmma_tile(const Tile<double, TN>& s, double (&acc)[2][2 * TN][2], int warp, int lane) {
  const int g = lane / 4, tig = lane % 4;
#pragma unroll
  for (int k0 = 0; k0 < BK; k0 += 4) {
    const double a0 = s.As[k0 + tig][warp * 16 + g], a1 = s.As[k0 + tig][warp * 16 + 8 + g];
#pragma unroll
    for (int nt = 0; nt < 2 * TN; ++nt) {
      const double b = s.Bs[k0 + tig][nt * 8 + g];
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(acc[0][nt][0]), "+d"(acc[0][nt][1]) : "d"(a0), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(acc[1][nt][0]), "+d"(acc[1][nt][1]) : "d"(a1), "d"(b));
    }
  }
}

template <typename T, int TN>
__device__ __forceinline__ void mma_tile(const Tile<T, TN>& s, T (&acc)[8][TN], int ty, int tx) {
#pragma unroll
  for (int kk = 0; kk < BK; ++kk) {
    T a[8], b[TN];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = s.As[kk][ty * 8 + i];
#pragma unroll
    for (int j = 0; j < TN; ++j) b[j] = s.Bs[kk][tx + 16 * j];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
}

// ---------------------------------------------------------------------------------------------
// sketch
// ---------------------------------------------------------------------------------------------
template <typename T, int TN>
__global__ void __launch_bounds__(NTHREADS)
sketch_kernel(const T* __restrict__ X, int64_t m, int64_t n, int64_t ldx, const T* __restrict__ Om,
              int64_t l, int64_t ldo, T* __restrict__ Y, int64_t ldy) {
  __shared__ Tile<T, TN> s;
  constexpr int BN = 16 * TN;
  const int t = threadIdx.x;
  const int ty = t / 16, tx = t % 16;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int64_t col0 = (int64_t)blockIdx.y * BN;

  // A loader: thread -> (row = t % 128, 8 consecutive k starting at (t / 128) * 8)
  const int a_row = t % BM, a_k0 = (t / BM) * 8;
  const int64_t g_row = row0 + a_row;
  const T* a_ptr = X + (g_row < m ? g_row : 0) * ldx;
  // B loader: thread -> (k = t / 16, cols tx + 16 * jj)
  const int b_k = t / 16;

  T a_reg[8], b_reg[TN];
  T acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

  auto load = [&](int64_t k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int64_t k = k0 + a_k0 + i;
      a_reg[i] = (g_row < m && k < n) ? a_ptr[k] : T(0);
    }
    int64_t kb = k0 + b_k;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t c = col0 + tx + 16 * j;
      b_reg[j] = (kb < n && c < l) ? Om[kb * ldo + c] : T(0);
    }
  };

  load(0);
  for (int64_t k0 = 0; k0 < n; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s.As[a_k0 + i][a_row] = a_reg[i];
#pragma unroll
    for (int j = 0; j < TN; ++j) s.Bs[b_k][tx + 16 * j] = b_reg[j];
    __syncthreads();
    if (k0 + BK < n) load(k0 + BK);
    mma_tile<T, TN>(s, acc, ty, tx);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t r = row0 + ty * 8 + i;
    if (r < m) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int64_t c = col0 + tx + 16 * j;
        if (c < l) Y[r * ldy + c] = acc[i][j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// project
// ---------------------------------------------------------------------------------------------
template <typename T, int TN>
__global__ void __launch_bounds__(NTHREADS)
project_kernel(const T* __restrict__ X, int64_t m, int64_t n, int64_t ldx, const T* __restrict__ Y,
               int64_t l, int64_t ldy, T* __restrict__ part, int64_t rows_per_split) {
  __shared__ Tile<T, TN> s;
  constexpr int BN = 16 * TN;
  const int t = threadIdx.x;
  const int ty = t / 16, tx = t % 16;
  const int64_t t0 = (int64_t)blockIdx.x * BM;     // time tile
  const int64_t col0 = (int64_t)blockIdx.y * BN;   // sketch-column tile
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(m, r_begin + rows_per_split);

  // A loader: thread -> (k = t / 16, 8 consecutive time values starting at (t % 16) * 8)
  const int a_k = t / 16, a_i0 = (t % 16) * 8;
  const int b_k = t / 16;

  T a_reg[8], b_reg[TN];
  T acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

  auto load = [&](int64_t r0) {
    int64_t r = r0 + a_k;
    const T* xr = X + (r < r_end ? r : r_begin) * ldx;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int64_t tt = t0 + a_i0 + i;
      a_reg[i] = (r < r_end && tt < n) ? xr[tt] : T(0);
    }
    const T* yr = Y + (r < r_end ? r : r_begin) * ldy;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t c = col0 + tx + 16 * j;
      b_reg[j] = (r < r_end && c < l) ? yr[c] : T(0);
    }
  };

  if (r_begin < r_end) {
    load(r_begin);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s.As[a_k][a_i0 + i] = a_reg[i];
#pragma unroll
      for (int j = 0; j < TN; ++j) s.Bs[b_k][tx + 16 * j] = b_reg[j];
      __syncthreads();
      if (r0 + BK < r_end) load(r0 + BK);
      mma_tile<T, TN>(s, acc, ty, tx);
      __syncthreads();
    }
  }
  // partial tile -> workspace [split][n][l]
  T* out = part + (int64_t)blockIdx.z * n * l;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int64_t tt = t0 + ty * 8 + i;
    if (tt < n) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        int64_t c = col0 + tx + 16 * j;
        if (c < l) out[tt * l + c] = acc[i][j];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FP64 on the tensor cores (DMMA): same staging as sketch_kernel / project_kernel<double>, product by dmma_tile.
// ---------------------------------------------------------------------------------------------
// COAL (the staging of the X tile, third session of round 2): false = every thread loads 8 consecutive k of ITS row
// (64 bytes per lane, 32 sectors per warp-level LDG.64); true = lane-contiguous - 16 lanes read the 128 contiguous bytes
// of one row's k-tile, two rows per warp instruction (8 sectors) - and the transposing shared-memory store takes the
// bank conflicts instead (4-way on a 64-bit store).
template <int TN, bool COAL>
__global__ void __launch_bounds__(NTHREADS)
sketch_dmma_kernel(const double* __restrict__ X, int64_t m, int64_t n, int64_t ldx, const double* __restrict__ Om,
                   int64_t l, int64_t ldo, double* __restrict__ Y, int64_t ldy) {
  __shared__ Tile<double, TN> s;
  constexpr int BN = 16 * TN;
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int tx = t % 16;
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int64_t col0 = (int64_t)blockIdx.y * BN;
  const int a_row = t % BM, a_k0 = (t / BM) * 8;      // strided staging: row a_row, k in [a_k0, a_k0 + 8)
  const int c_k = t % 16, c_r0 = t / 16;              // coalesced staging: k = c_k, rows c_r0 + 16 i
  const int64_t g_row = row0 + a_row;
  const double* a_ptr = X + (g_row < m ? g_row : 0) * ldx;
  const int b_k = t / 16;
  double a_reg[8], b_reg[TN];
  double acc[2][2 * TN][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2 * TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  auto load = [&](int64_t k0) {
    if constexpr (COAL) {
      const int64_t k = k0 + c_k;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t r = row0 + c_r0 + 16 * i;
        a_reg[i] = (r < m && k < n) ? X[r * ldx + k] : 0.0;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int64_t k = k0 + a_k0 + i;
        a_reg[i] = (g_row < m && k < n) ? a_ptr[k] : 0.0;
      }
    }
    int64_t kb = k0 + b_k;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t c = col0 + tx + 16 * j;
      b_reg[j] = (kb < n && c < l) ? Om[kb * ldo + c] : 0.0;
    }
  };
  load(0);
  for (int64_t k0 = 0; k0 < n; k0 += BK) {
    if constexpr (COAL) {
#pragma unroll
      for (int i = 0; i < 8; ++i) s.As[c_k][c_r0 + 16 * i] = a_reg[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) s.As[a_k0 + i][a_row] = a_reg[i];
    }
#pragma unroll
    for (int j = 0; j < TN; ++j) s.Bs[b_k][tx + 16 * j] = b_reg[j];
    __syncthreads();
    if (k0 + BK < n) load(k0 + BK);
    dmma_tile<TN>(s, acc, warp, lane);
    __syncthreads();
  }
  const int g = lane / 4, tig = lane % 4;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int64_t r = row0 + warp * 16 + mt * 8 + g;
    if (r < m) {
#pragma unroll
      for (int nt = 0; nt < 2 * TN; ++nt) {
        const int64_t c = col0 + nt * 8 + 2 * tig;
        if (c < l) Y[r * ldy + c] = acc[mt][nt][0];
        if (c + 1 < l) Y[r * ldy + c + 1] = acc[mt][nt][1];
      }
    }
  }
}

template <int TN, bool COAL>
__global__ void __launch_bounds__(NTHREADS)
project_dmma_kernel(const double* __restrict__ X, int64_t m, int64_t n, int64_t ldx, const double* __restrict__ Y,
                    int64_t l, int64_t ldy, double* __restrict__ part, int64_t rows_per_split) {
  __shared__ Tile<double, TN> s;
  constexpr int BN = 16 * TN;
  const int t = threadIdx.x, warp = t / 32, lane = t % 32;
  const int tx = t % 16;
  const int64_t t0 = (int64_t)blockIdx.x * BM;     // time tile
  const int64_t col0 = (int64_t)blockIdx.y * BN;   // sketch-column tile
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t r_end = min(m, r_begin + rows_per_split);
  const int a_k = t / 16, a_i0 = (t % 16) * 8;        // strided staging: row a_k, time columns [a_i0, a_i0 + 8)
  const int c_c = t % BM, c_k0 = t / BM;              // coalesced staging: time column c_c, rows c_k0 + 2 i
  double a_reg[8], b_reg[TN];
  double acc[2][2 * TN][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 2 * TN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  auto load = [&](int64_t r0) {
    int64_t r = r0 + a_k;
    if constexpr (COAL) {
      const int64_t tt = t0 + c_c;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int64_t rr = r0 + c_k0 + 2 * i;
        a_reg[i] = (rr < r_end && tt < n) ? X[rr * ldx + tt] : 0.0;
      }
    } else {
      const double* xr = X + (r < r_end ? r : r_begin) * ldx;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int64_t tt = t0 + a_i0 + i;
        a_reg[i] = (r < r_end && tt < n) ? xr[tt] : 0.0;
      }
    }
    const double* yr = Y + (r < r_end ? r : r_begin) * ldy;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int64_t c = col0 + tx + 16 * j;
      b_reg[j] = (r < r_end && c < l) ? yr[c] : 0.0;
    }
  };
  if (r_begin < r_end) {
    load(r_begin);
    for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
      if constexpr (COAL) {
#pragma unroll
        for (int i = 0; i < 8; ++i) s.As[c_k0 + 2 * i][c_c] = a_reg[i];
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) s.As[a_k][a_i0 + i] = a_reg[i];
      }
#pragma unroll
      for (int j = 0; j < TN; ++j) s.Bs[a_k][tx + 16 * j] = b_reg[j];
      __syncthreads();
      if (r0 + BK < r_end) load(r0 + BK);
      dmma_tile<TN>(s, acc, warp, lane);
      __syncthreads();
    }
  }
  double* out = part + (int64_t)blockIdx.z * n * l;
  const int g = lane / 4, tig = lane % 4;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int64_t tt = t0 + warp * 16 + mt * 8 + g;
    if (tt < n) {
#pragma unroll
      for (int nt = 0; nt < 2 * TN; ++nt) {
        const int64_t c = col0 + nt * 8 + 2 * tig;
        if (c < l) out[tt * l + c] = acc[mt][nt][0];
        if (c + 1 < l) out[tt * l + c + 1] = acc[mt][nt][1];
      }
    }
  }
}

// 64 outputs x 4 split groups per CTA: group g sums its contiguous quarter of the splits (eight independent loads
// in flight), the four group sums are added through shared memory in a fixed order (deterministic).  One thread
// per output walking all ~300 splits alone was latency bound (~37 dependent rounds of HBM latency).
constexpr int RP_OUT = 64;
constexpr int RP_GROUPS = 4;

template <typename T>
__global__ void __launch_bounds__(RP_OUT * RP_GROUPS)
reduce_partials_kernel(const T* __restrict__ part, int64_t splits, int64_t n, int64_t l, int64_t lp,
                       double* __restrict__ Z, int64_t ldz, int accumulate) {
  // part[split][n][lp]: lp >= l is the column pitch of a partial tile (columns l..lp-1 are padding)
  __shared__ double sm[RP_GROUPS][RP_OUT];
  const int o = threadIdx.x % RP_OUT, g = threadIdx.x / RP_OUT;
  const int64_t idx = (int64_t)blockIdx.x * RP_OUT + o;
  const int64_t total = n * lp;
  const int64_t per = (splits + RP_GROUPS - 1) / RP_GROUPS;
  const int64_t k_end = (g + 1) * per < splits ? (g + 1) * per : splits;
  double s = 0.0;
  if (idx < total) {
    int64_t k = g * per;
    for (; k + 8 <= k_end; k += 8) {
      T v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = part[(k + u) * total + idx];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    for (; k < k_end; ++k) s += (double)part[k * total + idx];
  }
  sm[g][o] = s;
  __syncthreads();
  if (g == 0 && idx < total) {
    double t = sm[0][o];
#pragma unroll
    for (int j = 1; j < RP_GROUPS; ++j) t += sm[j][o];
    const int64_t r = idx / lp, c = idx % lp;
    if (c < l) {
      double* z = Z + r * ldz + c;
      *z = accumulate ? (*z + t) : t;
    }
  }
}

// The same reduction fused with the all-reduce over the row shards (comm.cuh): phase 1 sums this rank's partial tiles into
// its peer-mapped slot (compact n x l), one grid-wide flag round (comm_grid_exchange), phase 2 adds the ranks' slots in
// rank order into Z.  The grid is bounded (<= 4 CTAs per SM, all co-resident).
template <typename T>
__global__ void __launch_bounds__(RP_OUT * RP_GROUPS, 4)
reduce_partials_allreduce_kernel(const T* __restrict__ part, int64_t splits, int64_t n, int64_t l, int64_t lp,
                                 double* __restrict__ Z, int64_t ldz, int accumulate, CommDev c) {
  __shared__ double sm[RP_GROUPS][RP_OUT];
  const int o = threadIdx.x % RP_OUT, g = threadIdx.x / RP_OUT;
  const int64_t total = n * lp;
  const int64_t nchunks = (total + RP_OUT - 1) / RP_OUT;
  const int64_t per = (splits + RP_GROUPS - 1) / RP_GROUPS;
  const int64_t k_end = (g + 1) * per < splits ? (g + 1) * per : splits;
  double* mine = c.slot[c.rank];
  for (int64_t chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const int64_t idx = chunk * RP_OUT + o;
    double s = 0.0;
    if (idx < total) {
      int64_t k = g * per;
      for (; k + 8 <= k_end; k += 8) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = part[(k + u) * total + idx];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
      }
      for (; k < k_end; ++k) s += (double)part[k * total + idx];
    }
    sm[g][o] = s;
    __syncthreads();
    if (g == 0 && idx < total) {
      double t = sm[0][o];
#pragma unroll
      for (int j = 1; j < RP_GROUPS; ++j) t += sm[j][o];
      const int64_t r = idx / lp, cc = idx % lp;
      if (cc < l) mine[r * l + cc] = accumulate ? (Z[r * ldz + cc] + t) : t;
    }
    __syncthreads();
  }
  comm_grid_exchange(c);
  const int64_t my_chunks = (int64_t)blockIdx.x < nchunks ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  for (int64_t e = threadIdx.x; e < my_chunks * RP_OUT; e += blockDim.x) {
    const int64_t idx = ((int64_t)blockIdx.x + (e / RP_OUT) * gridDim.x) * RP_OUT + e % RP_OUT;
    if (idx < total) {
      const int64_t r = idx / lp, cc = idx % lp;
      if (cc < l) Z[r * ldz + cc] = comm_sum_ranks(c, r * l + cc);
    }
  }
}

// Sum of the partial tiles -> Z; fused with the all-reduce over the ranks when a communicator was bound for this call
// (era5svd_comm_fuse_next_project).
template <typename T>
static int launch_reduce_partials(const T* part, int64_t splits, int64_t n, int64_t l, int64_t lp, double* Z, int64_t ldz,
                                  int accumulate, cudaStream_t st) {
  int64_t bn = 0, bl = 0;
  Comm* cm = take_bound_comm(&bn, &bl);
  const int64_t nchunks = ceil_div(n * lp, (int64_t)RP_OUT);
  if (cm) {
    if (bn != n || bl != l) {
      set_error("fused all-reduce was requested for a %lld x %lld projection, got %lld x %lld", (long long)bn,
                (long long)bl, (long long)n, (long long)l);
      return ERA5SVD_ERR_ARG;
    }
    CommDev d;
    if (!comm_next(cm, n * l, &d)) return ERA5SVD_ERR_ARG;
    const int64_t grid = nchunks < comm_max_grid() ? nchunks : comm_max_grid();
    reduce_partials_allreduce_kernel<T><<<(unsigned)grid, RP_OUT * RP_GROUPS, 0, st>>>(part, splits, n, l, lp, Z, ldz,
                                                                                      accumulate, d);
    comm_note_fused(cm);
    return check_launch("reduce_partials_allreduce_kernel");
  }
  reduce_partials_kernel<T><<<(unsigned)nchunks, RP_OUT * RP_GROUPS, 0, st>>>(part, splits, n, l, lp, Z, ldz, accumulate);
  return check_launch("reduce_partials_kernel");
}

struct ProjectPlan {
  int64_t splits;
  int64_t rows_per_split;
  size_t bytes;
};

// Rows per split: at most 4096 (bounds the length of any float32 running sum), but never so
// many splits that the partial-tile workspace exceeds 256 MiB.
static ProjectPlan project_plan(int dtype, int64_t m, int64_t n, int64_t l) {
  const int64_t cap = (int64_t)256 << 20;
  const int64_t tile_bytes = n * l * (int64_t)dtype_size(dtype);
  int64_t max_splits = cap / (tile_bytes > 0 ? tile_bytes : 1);
  if (max_splits < 1) max_splits = 1;
  int64_t splits = ceil_div(m, 4096);
  if (splits > max_splits) splits = max_splits;
  if (splits > 65535) splits = 65535;
  if (splits < 1) splits = 1;
  int64_t rps = ceil_div(ceil_div(m, splits), BK) * BK;
  splits = ceil_div(m, rps);
  ProjectPlan p;
  p.splits = splits;
  p.rows_per_split = rps;
  p.bytes = (size_t)(splits * tile_bytes);
  return p;
}

// Staging of the X tile in the DMMA kernels: lane-contiguous loads unless ERA5SVD_DMMA_STAGING=strided asks for the first
// version (kept for A/B timing).
static bool dmma_coalesced() {
  static const bool on = [] { const char* e = getenv("ERA5SVD_DMMA_STAGING"); return !(e && e[0] == 's'); }();
  return on;
}

// Column tile width: 112 (TN = 7) or 128 (TN = 8), whichever pads l = k + 10 less.
static bool use_tn7(int64_t l) { return ceil_div(l, 112) * 112 <= ceil_div(l, 128) * 128; }

template <typename T>
int sketch_native(const void* X, int64_t m, int64_t n, int64_t ldx, const void* Om, int64_t l,
                  int64_t ldo, void* Y, int64_t ldy, cudaStream_t st) {
  if constexpr (sizeof(T) == 8) {      // float64: FP64 tensor cores (mma.sync m8n8k4)
    const bool tn7 = use_tn7(l);
    dim3 grid((unsigned)ceil_div(m, BM), (unsigned)ceil_div(l, tn7 ? 112 : 128));
    auto kern = tn7 ? (dmma_coalesced() ? sketch_dmma_kernel<7, true> : sketch_dmma_kernel<7, false>)
                    : (dmma_coalesced() ? sketch_dmma_kernel<8, true> : sketch_dmma_kernel<8, false>);
    kern<<<grid, NTHREADS, 0, st>>>((const double*)X, m, n, ldx, (const double*)Om, l, ldo, (double*)Y, ldy);
    return check_launch("sketch_dmma_kernel");
  }
  if (use_tn7(l)) {
    dim3 grid((unsigned)ceil_div(m, BM), (unsigned)ceil_div(l, 112));
    sketch_kernel<T, 7><<<grid, NTHREADS, 0, st>>>((const T*)X, m, n, ldx, (const T*)Om, l, ldo, (T*)Y, ldy);
  } else {
    dim3 grid((unsigned)ceil_div(m, BM), (unsigned)ceil_div(l, 128));
    sketch_kernel<T, 8><<<grid, NTHREADS, 0, st>>>((const T*)X, m, n, ldx, (const T*)Om, l, ldo, (T*)Y, ldy);
  }
  return check_launch("sketch_kernel");
}

template <typename T>
int project_native(const void* X, int64_t m, int64_t n, int64_t ldx, const void* Y, int64_t l,
                   int64_t ldy, double* Z, int64_t ldz, int accumulate, void* ws,
                   const ProjectPlan& plan, cudaStream_t st) {
  if constexpr (sizeof(T) == 8) {      // float64: FP64 tensor cores (mma.sync m8n8k4)
    const bool tn7 = use_tn7(l);
    dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(l, tn7 ? 112 : 128), (unsigned)plan.splits);
    auto kern = tn7 ? (dmma_coalesced() ? project_dmma_kernel<7, true> : project_dmma_kernel<7, false>)
                    : (dmma_coalesced() ? project_dmma_kernel<8, true> : project_dmma_kernel<8, false>);
    kern<<<grid, NTHREADS, 0, st>>>((const double*)X, m, n, ldx, (const double*)Y, l, ldy, (double*)ws, plan.rows_per_split);
    int rc = check_launch("project_dmma_kernel");
    if (rc) return rc;
    return launch_reduce_partials<double>((const double*)ws, plan.splits, n, l, l, Z, ldz, accumulate, st);
  }
  if (use_tn7(l)) {
    dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(l, 112), (unsigned)plan.splits);
    project_kernel<T, 7><<<grid, NTHREADS, 0, st>>>((const T*)X, m, n, ldx, (const T*)Y, l, ldy, (T*)ws, plan.rows_per_split);
  } else {
    dim3 grid((unsigned)ceil_div(n, BM), (unsigned)ceil_div(l, 128), (unsigned)plan.splits);
    project_kernel<T, 8><<<grid, NTHREADS, 0, st>>>((const T*)X, m, n, ldx, (const T*)Y, l, ldy, (T*)ws, plan.rows_per_split);
  }
  int rc = check_launch("project_kernel");
  if (rc) return rc;
  return launch_reduce_partials<T>((const T*)ws, plan.splits, n, l, l, Z, ldz, accumulate, st);
}

// shared with the tcgen05 path (gemm_tc.cu): float32 partial tiles -> float64 sum
int launch_reduce_partials_f32(const float* part, int64_t splits, int64_t n, int64_t l, int64_t lp, double* Z,
                               int64_t ldz, int accumulate, cudaStream_t st) {
  return launch_reduce_partials<float>(part, splits, n, l, lp, Z, ldz, accumulate, st);
}

}  // namespace era5svd

extern "C" {

int era5svd_sketch(const void* X, int dtype, int64_t m, int64_t n, int64_t ldx, const void* Om,
                   int64_t l, int64_t ldo, void* Y, int64_t ldy, int precision, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(X && Om && Y, "sketch: null pointer");
  ERA5SVD_REQUIRE(valid_dtype(dtype), "sketch: bad dtype %d", dtype);
  ERA5SVD_REQUIRE(m > 0 && n > 0 && l > 0 && ldx >= n && ldo >= l && ldy >= l,
                  "sketch: bad shape m=%lld n=%lld l=%lld ldx=%lld ldo=%lld ldy=%lld", (long long)m,
                  (long long)n, (long long)l, (long long)ldx, (long long)ldo, (long long)ldy);
  ERA5SVD_REQUIRE(ceil_div(l, 112) <= 65535, "sketch: l too large");
  cudaStream_t st = as_stream(stream);
  if (precision == ERA5SVD_PREC_TF32X3) {
    set_error("sketch: the tensor-core path takes pre-split operands: use era5svd_sketch_tf32x3");
    return ERA5SVD_ERR_UNSUPPORTED;
  }
  ERA5SVD_REQUIRE(precision == ERA5SVD_PREC_NATIVE, "sketch: bad precision %d", precision);
  if (dtype == ERA5SVD_F32) return sketch_native<float>(X, m, n, ldx, Om, l, ldo, Y, ldy, st);
  return sketch_native<double>(X, m, n, ldx, Om, l, ldo, Y, ldy, st);
}

size_t era5svd_project_workspace_bytes(int dtype, int64_t m, int64_t n, int64_t l, int precision) {
  using namespace era5svd;
  if (!valid_dtype(dtype) || m <= 0 || n <= 0 || l <= 0) return 0;
  if (precision == ERA5SVD_PREC_TF32X3) return era5svd_project_tf32x3_workspace_bytes(m, n, l);
  return project_plan(dtype, m, n, l).bytes;
}

int era5svd_project(const void* X, int dtype, int64_t m, int64_t n, int64_t ldx, const void* Y,
                    int64_t l, int64_t ldy, double* Z, int64_t ldz, int accumulate, int precision,
                    void* workspace, size_t workspace_bytes, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(X && Y && Z, "project: null pointer");
  ERA5SVD_REQUIRE(valid_dtype(dtype), "project: bad dtype %d", dtype);
  ERA5SVD_REQUIRE(m > 0 && n > 0 && l > 0 && ldx >= n && ldy >= l && ldz >= l,
                  "project: bad shape m=%lld n=%lld l=%lld ldx=%lld ldy=%lld ldz=%lld", (long long)m,
                  (long long)n, (long long)l, (long long)ldx, (long long)ldy, (long long)ldz);
  cudaStream_t st = as_stream(stream);
  if (precision == ERA5SVD_PREC_TF32X3) {
    set_error("project: the tensor-core path takes pre-split operands: use era5svd_project_tf32x3");
    return ERA5SVD_ERR_UNSUPPORTED;
  }
  ERA5SVD_REQUIRE(precision == ERA5SVD_PREC_NATIVE, "project: bad precision %d", precision);
  ERA5SVD_REQUIRE(ceil_div(l, 112) <= 65535, "project: l too large");
  ProjectPlan plan = project_plan(dtype, m, n, l);
  if (!workspace || workspace_bytes < plan.bytes) {
    set_error("project: workspace too small (%zu < %zu)", workspace_bytes, plan.bytes);
    return ERA5SVD_ERR_WORKSPACE;
  }
  if (dtype == ERA5SVD_F32)
    return project_native<float>(X, m, n, ldx, Y, l, ldy, Z, ldz, accumulate, workspace, plan, st);
  return project_native<double>(X, m, n, ldx, Y, l, ldy, Z, ldz, accumulate, workspace, plan, st);
}

}  // extern "C"
