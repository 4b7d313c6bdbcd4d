// (d) Symmetric eigensolver for TIME-SIZED matrices (n up to ~16 k): the eigensolve behind the Gram-route standard
// SVD (np.linalg.svd(X, full_matrices=False) + [:k] truncation, src/dmd_era5/era5_svd/era5_svd.py:249-254; LAPACK
// gesdd on the host).  Only the k LARGEST eigenpairs are ever used (the reference truncates to n_components), so:
//
//   1. era5svd_tridiag_reduce_f64     A = Q T Q^T  Householder tridiagonalisation, all SMs, memory bound: ONE fused
//                                     pass per column (rank-2 update of the trailing block + symv of the next
//                                     column on the updated values): 16 B per trailing element and column
//   2. era5svd_tridiag_eig_topk_f64   k largest eigenvalues of T by bisection on Sturm counts (one thread each),
//                                     vectors by inverse iteration (tridiagonal LU with partial pivoting)
//   3. era5svd_tridiag_apply_f64      Y = T Z   (for the Rayleigh-Ritz clean-up of close eigenvalues, done by the
//                                     host driver with the small float64 kernels)
//   4. era5svd_tridiag_backtransform_f64   V = Q Z : reflectors applied in reverse, one CTA per eigenvector with the
//                                     vector resident in shared memory
//
// The one-CTA Jacobi solver (small_f64.cu) stays the solver for sketch-sized (l x l) matrices.
#include "common.cuh"

namespace era5svd {

namespace {

constexpr int TD_THREADS = 256;
constexpr int TD_WARPS = TD_THREADS / 32;

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int w = threadIdx.x / 32, l = threadIdx.x % 32;
  __syncthreads();                 // red may still be in use by a previous reduction
  if (l == 0) red[w] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < TD_WARPS; ++i) s += red[i];   // same order in every thread and every CTA: deterministic
  return s;
}

// Householder vector for column j (LAPACK dlarfg on x = A[j, j+1:]):  H = I - tau v v^T, v[0] = 1,
// H x = beta e_1.  Every CTA recomputes (beta, tau, scale) redundantly from the row (deterministic).
struct House {
  double beta, tau, scale;   // v[c] = c == 0 ? 1 : x[c] * scale
};

__device__ House householder(const double* __restrict__ x, int r, double* red) {
  double s = 0.0;
  for (int c = 1 + threadIdx.x; c < r; c += TD_THREADS) s = fma(x[c], x[c], s);
  const double sigma = block_sum(s, red);
  const double x0 = x[0];
  House h;
  if (sigma == 0.0) {          // already in tridiagonal form in this column
    h.beta = x0; h.tau = 0.0; h.scale = 0.0;
  } else {
    const double nrm = sqrt(fma(x0, x0, sigma));
    h.beta = x0 >= 0.0 ? -nrm : nrm;
    h.tau = (h.beta - x0) / h.beta;
    h.scale = 1.0 / (x0 - h.beta);
  }
  return h;
}

// Fused pass for column j: rank-2 update of the trailing block with (v_j, w_j) AND the symv of column j + 1 on the
// updated values, 16 B of HBM traffic per trailing element instead of 24 B in two launches.
//   A22 = A[j+1:, j+1:] (r x r).  w = p - (tau/2)(p^T v) v.  Row 0 of the updated block gives the next Householder
//   vector v' (every CTA recomputes beta', tau', scale' from a_0c - v_0 w_c - w_0 v_c: three r-long vector reads),
//   rows i >= 1 are updated in place and dotted with v':  p'_{i-1} = tau' sum_{c>=1} a'_ic v'_{c-1}.
// A warp owns TD_ROWS rows; v, w, v' are staged chunk by chunk in shared memory, so an element costs one global load,
// one global store and three shared-memory reads.  has_update == 0 (first call, j = -1): symv only.
constexpr int TD_ROWS = 2;          // rows per warp
constexpr int TD_CHUNK = 1024;      // staged columns per step

__global__ void __launch_bounds__(TD_THREADS)
tridiag_fused_kernel(double* __restrict__ A, int64_t lda, int n, int j, int has_update, const double* __restrict__ vbuf,
                     const double* __restrict__ pbuf, double* __restrict__ vnext, double* __restrict__ pnext,
                     double* __restrict__ d, double* __restrict__ e, double* __restrict__ tau) {
  __shared__ double red[TD_WARPS];
  __shared__ double sv[TD_CHUNK], sw[TD_CHUNK], sn[TD_CHUNK];
  const int r = n - j - 1;                      // order of the trailing block A22 (rows / columns j+1 .. n-1)
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  double* A22 = A + (int64_t)(j + 1) * lda + (j + 1);
  const double t = has_update ? tau[j] : 0.0;
  double coef = 0.0;
  if (t != 0.0) {
    double s = 0.0;
    for (int c = threadIdx.x; c < r; c += TD_THREADS) s = fma(pbuf[c], vbuf[c], s);
    coef = 0.5 * t * block_sum(s, red);
  }
  const double v0 = t != 0.0 ? vbuf[0] : 0.0, w0 = t != 0.0 ? pbuf[0] - coef * v0 : 0.0;
  auto new_row0 = [&](int c) -> double {        // updated A22[0][c]
    double a = A22[c];
    if (t != 0.0) {
      const double vc = vbuf[c], wc = pbuf[c] - coef * vc;
      a -= fma(v0, wc, w0 * vc);
    }
    return a;
  };
  // Householder vector of the next column from x' = A22'[0, 1:] (length r - 1); nothing to do when r - 1 < 2
  const int rn = r - 1;
  const bool do_next = rn >= 2;
  double beta = 0.0, tnext = 0.0, scale = 0.0;
  if (do_next) {
    double s = 0.0;
    for (int c = 2 + threadIdx.x; c < r; c += TD_THREADS) { const double x = new_row0(c); s = fma(x, x, s); }
    const double sigma = block_sum(s, red);
    const double x0 = new_row0(1);
    if (sigma == 0.0) {
      beta = x0; tnext = 0.0; scale = 0.0;
    } else {
      const double nrm = sqrt(fma(x0, x0, sigma));
      beta = x0 >= 0.0 ? -nrm : nrm;
      tnext = (beta - x0) / beta;
      scale = 1.0 / (x0 - beta);
    }
    if (blockIdx.x == 0) {
      for (int c = 1 + threadIdx.x; c < r; c += TD_THREADS) vnext[c - 1] = c == 1 ? 1.0 : new_row0(c) * scale;
      if (threadIdx.x == 0) { tau[j + 1] = tnext; e[j + 1] = beta; }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) d[j + 1] = new_row0(0);   // row 0 of the block is never written back
  if (blockIdx.x == 0 && has_update) {          // reflector j goes to its final place: row j, columns j+1..
    double* x = A + (int64_t)j * lda + j + 1;
    for (int c = threadIdx.x; c < r; c += TD_THREADS) x[c] = vbuf[c];
  }
  if (t == 0.0 && !do_next) return;
  // rows 1 .. r-1 of A22: warp owns TD_ROWS consecutive rows
  const int row_first = 1 + (blockIdx.x * TD_WARPS + warp) * TD_ROWS;
  double vi[TD_ROWS], wi[TD_ROWS], acc[TD_ROWS];
#pragma unroll
  for (int u = 0; u < TD_ROWS; ++u) {
    const int i = row_first + u;
    vi[u] = (t != 0.0 && i < r) ? vbuf[i] : 0.0;
    wi[u] = (t != 0.0 && i < r) ? pbuf[i] - coef * vi[u] : 0.0;
    acc[u] = 0.0;
  }
  for (int c0 = 0; c0 < r; c0 += TD_CHUNK) {
    const int cn = min(TD_CHUNK, r - c0);
    __syncthreads();
    for (int c = threadIdx.x; c < cn; c += TD_THREADS) {
      const int gc = c0 + c;
      const double vc = t != 0.0 ? vbuf[gc] : 0.0;
      sv[c] = vc;
      sw[c] = t != 0.0 ? pbuf[gc] - coef * vc : 0.0;
      sn[c] = (do_next && gc >= 1) ? (gc == 1 ? 1.0 : new_row0(gc) * scale) : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < TD_ROWS; ++u) {
      const int i = row_first + u;
      if (i >= r) continue;
      double* row = A22 + (int64_t)i * lda + c0;
      int c = lane;
      for (; c + 96 < cn; c += 128) {          // four independent read-modify-writes in flight
        double a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) a[q] = row[c + 32 * q];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int cc = c + 32 * q;
          a[q] -= fma(vi[u], sw[cc], wi[u] * sv[cc]);
          acc[u] = fma(a[q], sn[cc], acc[u]);
        }
        if (t != 0.0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) row[c + 32 * q] = a[q];
        }
      }
      for (; c < cn; c += 32) {
        const double a = row[c] - fma(vi[u], sw[c], wi[u] * sv[c]);
        acc[u] = fma(a, sn[c], acc[u]);
        if (t != 0.0) row[c] = a;
      }
    }
  }
  if (do_next) {
#pragma unroll
    for (int u = 0; u < TD_ROWS; ++u) {
      const int i = row_first + u;
      const double s = warp_sum(acc[u]);
      if (lane == 0 && i < r) pnext[i - 1] = tnext * s;
    }
  }
}

// last 2 x 2 block: d[n-2] was written by the final fused call (n >= 3); the off-diagonal and d[n-1] are read from the
// updated last ROW (the fused kernel does not write row 0 of a block back)
__global__ void tridiag_tail_kernel(const double* __restrict__ A, int64_t lda, int n, double* __restrict__ d,
                                    double* __restrict__ e, double* __restrict__ tau) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (n >= 2) {
    if (n == 2) d[0] = A[0];
    e[n - 2] = A[(int64_t)(n - 1) * lda + n - 2];
    tau[n - 2] = 0.0;
  }
  d[n - 1] = A[(int64_t)(n - 1) * lda + n - 1];
  e[n - 1] = 0.0;
  tau[n - 1] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// k largest eigenvalues by bisection (Sturm counts, LAPACK dstebz recurrence with pivmin guard) and the
// corresponding vectors by inverse iteration (LAPACK dstein / dgttrf-style tridiagonal LU with partial
// pivoting).  One thread per eigenpair; per-thread work arrays are interleaved ([i * k + t]) so that the
// sequential sweeps over i are coalesced across threads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int sturm_count(const double* __restrict__ d, const double* __restrict__ e2, int n,
                                           double x, double pivmin) {
  int cnt = 0;
  double q = d[0] - x;
  if (fabs(q) < pivmin) q = -pivmin;
  cnt += q < 0.0;
  for (int i = 1; i < n; ++i) {
    q = d[i] - x - e2[i - 1] / q;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
  }
  return cnt;    // number of eigenvalues < x
}

__global__ void tridiag_e2_bounds_kernel(const double* __restrict__ d, const double* __restrict__ e, int n,
                                         double* __restrict__ e2, double* __restrict__ bounds) {
  // single CTA: e2 = e^2, Gershgorin interval, pivmin
  __shared__ double smin[TD_WARPS], smax[TD_WARPS], semax[TD_WARPS];
  double lo = 1e300, hi = -1e300, em = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double el = i > 0 ? fabs(e[i - 1]) : 0.0, er = i < n - 1 ? fabs(e[i]) : 0.0;
    lo = fmin(lo, d[i] - el - er);
    hi = fmax(hi, d[i] + el + er);
    if (i < n - 1) { const double v = e[i] * e[i]; e2[i] = v; em = fmax(em, v); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    em = fmax(em, __shfl_xor_sync(0xffffffffu, em, o));
  }
  const int w = threadIdx.x / 32;
  if (threadIdx.x % 32 == 0) { smin[w] = lo; smax[w] = hi; semax[w] = em; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x / 32); ++i) { lo = fmin(lo, smin[i]); hi = fmax(hi, smax[i]); em = fmax(em, semax[i]); }
    const double tn = fmax(fabs(lo), fabs(hi));
    bounds[0] = lo - 2.2e-16 * tn * n - 1e-300;
    bounds[1] = hi + 2.2e-16 * tn * n + 1e-300;
    bounds[2] = fmax(2.2250738585072014e-308 * fmax(em, 1.0), 1e-290);   // pivmin
    bounds[3] = tn;
  }
}

__global__ void __launch_bounds__(32)
tridiag_bisect_kernel(const double* __restrict__ d, const double* __restrict__ e2, int n, int k,
                      const double* __restrict__ bounds, double* __restrict__ W) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k) return;
  const int target = n - 1 - t;           // ascending index of the t-th largest eigenvalue
  double lo = bounds[0], hi = bounds[1];
  const double pivmin = bounds[2];
  for (int it = 0; it < 128; ++it) {
    const double mid = 0.5 * (lo + hi);
    if (mid <= lo || mid >= hi) break;    // interval no longer splittable in float64
    if (sturm_count(d, e2, n, mid, pivmin) > target) hi = mid; else lo = mid;
  }
  W[t] = 0.5 * (lo + hi);
}

// work: 5 arrays of n x k doubles (dd, du, du2, dl, rhs), pivot flags n x k ints
__global__ void __launch_bounds__(32)
tridiag_invit_kernel(const double* __restrict__ d, const double* __restrict__ e, int n, int k,
                     const double* __restrict__ W, const double* __restrict__ bounds, double* __restrict__ Z,
                     int64_t ldz, double* __restrict__ work, int* __restrict__ piv) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= k) return;
  const double tn = bounds[3];
  const double eps = 2.220446049250313e-16;
  // close eigenvalues get distinct shifts (dstein perturbs them too); the driver re-orthogonalises and applies a
  // Rayleigh-Ritz step afterwards, so only independence of the vectors matters here
  double lam = W[t];
  if (t > 0 && fabs(W[t - 1] - lam) < 10.0 * eps * tn) lam -= 10.0 * eps * tn * (1 + t % 7);
  const size_t nk = (size_t)n * k;
  double* dd = work;  double* du = work + nk;  double* du2 = work + 2 * nk;  double* dl = work + 3 * nk;
  double* x = work + 4 * nk;
#define AT(p, i) p[(size_t)(i) * k + t]
  // LU of (T - lam I) with partial pivoting (dgttrf)
  for (int i = 0; i < n; ++i) {
    AT(dd, i) = d[i] - lam;
    AT(du, i) = i < n - 1 ? e[i] : 0.0;
    AT(dl, i) = i < n - 1 ? e[i] : 0.0;
    AT(du2, i) = 0.0;
  }
  const double tiny = eps * tn + 1e-300;
  for (int i = 0; i < n - 1; ++i) {
    const double dii = AT(dd, i), dli = AT(dl, i);
    if (fabs(dii) >= fabs(dli)) {
      piv[(size_t)i * k + t] = 0;
      const double pivot = fabs(dii) < tiny ? (dii < 0 ? -tiny : tiny) : dii;
      AT(dd, i) = pivot;
      const double f = dli / pivot;
      AT(dl, i) = f;
      AT(dd, i + 1) -= f * AT(du, i);
    } else {
      piv[(size_t)i * k + t] = 1;          // swap rows i and i + 1
      const double f = dii / dli;
      AT(dd, i) = dli;
      AT(dl, i) = f;
      const double tmp = AT(du, i);
      AT(du, i) = AT(dd, i + 1);
      AT(dd, i + 1) = tmp - f * AT(dd, i + 1);
      if (i < n - 2) {
        AT(du2, i) = AT(du, i + 1);
        AT(du, i + 1) = -f * AT(du, i + 1);
      }
    }
  }
  {
    const double dnn = AT(dd, n - 1);
    if (fabs(dnn) < tiny) AT(dd, n - 1) = dnn < 0 ? -tiny : tiny;
  }
  // start vector: deterministic pseudo-random entries in (-1, 1) (different for every eigenpair)
  uint64_t st = 0x9E3779B97F4A7C15ull * (uint64_t)(t + 1);
  for (int i = 0; i < n; ++i) {
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    AT(x, i) = ((double)(st >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0;
  }
  for (int iter = 0; iter < 4; ++iter) {
    // forward: L y = P b
    for (int i = 0; i < n - 1; ++i) {
      if (piv[(size_t)i * k + t]) {
        const double tmp = AT(x, i);
        AT(x, i) = AT(x, i + 1);
        AT(x, i + 1) = tmp - AT(dl, i) * AT(x, i + 1);
      } else {
        AT(x, i + 1) -= AT(dl, i) * AT(x, i);
      }
    }
    // backward: U z = y
    double nrm2 = 0.0;
    for (int i = n - 1; i >= 0; --i) {
      double v = AT(x, i);
      if (i < n - 1) v -= AT(du, i) * AT(x, i + 1);
      if (i < n - 2) v -= AT(du2, i) * AT(x, i + 2);
      v /= AT(dd, i);
      AT(x, i) = v;
      nrm2 = fma(v, v, nrm2);
      if (nrm2 > 1e200) {                 // rescale the part computed so far to avoid overflow
        for (int q = i; q < n; ++q) AT(x, q) *= 1e-100;
        nrm2 *= 1e-200;
      }
    }
    const double inv = rsqrt(nrm2);
    for (int i = 0; i < n; ++i) AT(x, i) *= inv;
  }
  for (int i = 0; i < n; ++i) Z[(size_t)i * ldz + t] = AT(x, i);
#undef AT
}

// Y[n x k] = T Z
__global__ void __launch_bounds__(256)
tridiag_apply_kernel(const double* __restrict__ d, const double* __restrict__ e, int n, int k,
                     const double* __restrict__ Z, int64_t ldz, double* __restrict__ Y, int64_t ldy) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n * k) return;
  const int i = (int)(idx / k), c = (int)(idx % k);
  double v = d[i] * Z[(int64_t)i * ldz + c];
  if (i > 0) v = fma(e[i - 1], Z[(int64_t)(i - 1) * ldz + c], v);
  if (i < n - 1) v = fma(e[i], Z[(int64_t)(i + 1) * ldz + c], v);
  Y[(int64_t)i * ldy + c] = v;
}

// V[:, c] = H_0 H_1 ... H_{n-3} Z[:, c]; one CTA per column, the vector lives in shared memory.
__global__ void __launch_bounds__(TD_THREADS)
tridiag_backtransform_kernel(const double* __restrict__ A, int64_t lda, const double* __restrict__ tau, int n,
                             const double* __restrict__ Z, int64_t ldz, double* __restrict__ V, int64_t ldv) {
  extern __shared__ double z[];
  __shared__ double red[TD_WARPS];
  const int c = blockIdx.x;
  for (int i = threadIdx.x; i < n; i += TD_THREADS) z[i] = Z[(int64_t)i * ldz + c];
  __syncthreads();
  for (int j = n - 3; j >= 0; --j) {
    const double t = tau[j];
    if (t == 0.0) continue;
    const int r = n - j - 1;
    const double* v = A + (int64_t)j * lda + j + 1;     // v[0] = 1 stored explicitly
    double s = 0.0;
    for (int i = threadIdx.x; i < r; i += TD_THREADS) s = fma(v[i], z[j + 1 + i], s);
    const double f = t * block_sum(s, red);
    for (int i = threadIdx.x; i < r; i += TD_THREADS) z[j + 1 + i] = fma(-f, v[i], z[j + 1 + i]);
    __syncthreads();            // the next reflector's dot product reads z
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += TD_THREADS) V[(int64_t)i * ldv + c] = z[i];
}

}  // namespace
}  // namespace era5svd

extern "C" {

int era5svd_tridiag_reduce_f64(double* A, int64_t n, int64_t lda, double* d, double* e, double* tau,
                               void* workspace, size_t workspace_bytes, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(A && d && e && tau, "tridiag_reduce: null pointer");
  ERA5SVD_REQUIRE(n > 0 && n <= 32768 && lda >= n, "tridiag_reduce: bad shape n=%lld", (long long)n);
  const size_t need = (size_t)(4 * n) * sizeof(double);
  if (!workspace || workspace_bytes < need) {
    set_error("tridiag_reduce: workspace too small (%zu < %zu)", workspace_bytes, need);
    return ERA5SVD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  double* vb[2] = {(double*)workspace, (double*)workspace + n};
  double* pb[2] = {(double*)workspace + 2 * n, (double*)workspace + 3 * n};
  // call j = -1 computes the reflector and the symv of column 0; call j >= 0 applies reflector j and prepares j + 1
  for (int j = -1; j + 2 < (int)n; ++j) {
    const int r = (int)n - j - 1;
    const int cur = (j + 2) & 1, nxt = cur ^ 1;
    const int grid = (int)ceil_div(r > 1 ? r - 1 : 1, TD_WARPS * TD_ROWS);
    tridiag_fused_kernel<<<grid, TD_THREADS, 0, st>>>(A, lda, (int)n, j, j >= 0 ? 1 : 0, vb[cur], pb[cur], vb[nxt], pb[nxt],
                                                       d, e, tau);
    count_launch(1);
  }
  tridiag_tail_kernel<<<1, 32, 0, st>>>(A, lda, (int)n, d, e, tau);
  return check_launch("tridiag_reduce");
}

size_t era5svd_tridiag_reduce_workspace_bytes(int64_t n) { return n > 0 ? (size_t)(4 * n) * sizeof(double) : 0; }

size_t era5svd_tridiag_eig_topk_workspace_bytes(int64_t n, int64_t k) {
  if (n <= 0 || k <= 0) return 0;
  return (size_t)(5 * n * k) * sizeof(double) + (size_t)(n * k) * sizeof(int) + (size_t)(n + 8) * sizeof(double);
}

int era5svd_tridiag_eig_topk_f64(const double* d, const double* e, int64_t n, int64_t k, double* W, double* Z,
                                 int64_t ldz, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(d && e && W && Z, "tridiag_eig_topk: null pointer");
  ERA5SVD_REQUIRE(n > 0 && k > 0 && k <= n && ldz >= k, "tridiag_eig_topk: bad shape n=%lld k=%lld", (long long)n, (long long)k);
  const size_t need = era5svd_tridiag_eig_topk_workspace_bytes(n, k);
  if (!workspace || workspace_bytes < need) {
    set_error("tridiag_eig_topk: workspace too small (%zu < %zu)", workspace_bytes, need);
    return ERA5SVD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  double* work = (double*)workspace;
  double* e2 = work + 5 * n * k;
  double* bounds = e2 + n;
  int* piv = (int*)(bounds + 8);
  tridiag_e2_bounds_kernel<<<1, 256, 0, st>>>(d, e, (int)n, e2, bounds);
  int rc;
  if ((rc = check_launch("tridiag_e2_bounds_kernel"))) return rc;
  const unsigned blocks = (unsigned)ceil_div(k, 32);
  tridiag_bisect_kernel<<<blocks, 32, 0, st>>>(d, e2, (int)n, (int)k, bounds, W);
  if ((rc = check_launch("tridiag_bisect_kernel"))) return rc;
  tridiag_invit_kernel<<<blocks, 32, 0, st>>>(d, e, (int)n, (int)k, W, bounds, Z, ldz, work, piv);
  return check_launch("tridiag_invit_kernel");
}

int era5svd_tridiag_apply_f64(const double* d, const double* e, int64_t n, int64_t k, const double* Z, int64_t ldz,
                              double* Y, int64_t ldy, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(d && e && Z && Y, "tridiag_apply: null pointer");
  ERA5SVD_REQUIRE(n > 0 && k > 0 && ldz >= k && ldy >= k, "tridiag_apply: bad shape");
  tridiag_apply_kernel<<<(unsigned)ceil_div(n * k, 256), 256, 0, as_stream(stream)>>>(d, e, (int)n, (int)k, Z, ldz, Y, ldy);
  return check_launch("tridiag_apply_kernel");
}

int era5svd_tridiag_backtransform_f64(const double* A, int64_t n, int64_t lda, const double* tau, int64_t k,
                                      const double* Z, int64_t ldz, double* V, int64_t ldv, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(A && tau && Z && V, "tridiag_backtransform: null pointer");
  ERA5SVD_REQUIRE(n > 0 && k > 0 && lda >= n && ldz >= k && ldv >= k, "tridiag_backtransform: bad shape");
  const size_t smem = (size_t)n * sizeof(double);
  ERA5SVD_REQUIRE(smem <= 200 * 1024, "tridiag_backtransform: n = %lld does not fit in shared memory", (long long)n);
  ERA5SVD_CUDA(cudaFuncSetAttribute(tridiag_backtransform_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tridiag_backtransform_kernel<<<(unsigned)k, TD_THREADS, smem, as_stream(stream)>>>(A, lda, tau, (int)n, Z, ldz, V, ldv);
  return check_launch("tridiag_backtransform_kernel");
}

}  // extern "C"
