// (d) Symmetric eigensolver for TIME-SIZED matrices (n up to ~16 k): the eigensolve behind the Gram-route standard
// SVD (np.linalg.svd(X, full_matrices=False) + [:k] truncation, src/dmd_era5/era5_svd/era5_svd.py:249-254; LAPACK
// gesdd on the host).  Only the k LARGEST eigenpairs are ever used (the reference truncates to n_components), so:
//
//   1. era5svd_tridiag_reduce_f64     A = Q T Q^T  Householder tridiagonalisation, all SMs, memory bound:
//                                     per column one symv pass (read the trailing block) and one rank-2 update
//                                     pass (read + write it); 24 B per trailing element and column, 8 n^3 B total
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

// pass A of column j:  p = tau * A22 v   (A22 = trailing r x r block, r = n - j - 1), one warp per row.
// CTA 0 also publishes v, tau, d[j], e[j].
__global__ void __launch_bounds__(TD_THREADS)
tridiag_symv_kernel(const double* __restrict__ A, int64_t lda, int n, int j, double* __restrict__ vbuf,
                    double* __restrict__ pbuf, double* __restrict__ d, double* __restrict__ e,
                    double* __restrict__ tau) {
  __shared__ double red[TD_WARPS];
  const int r = n - j - 1;
  const double* x = A + (int64_t)j * lda + j + 1;
  const House h = householder(x, r, red);
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < r; c += TD_THREADS) vbuf[c] = c == 0 ? 1.0 : x[c] * h.scale;
    if (threadIdx.x == 0) { tau[j] = h.tau; e[j] = h.beta; d[j] = A[(int64_t)j * lda + j]; }
  }
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = blockIdx.x * TD_WARPS + warp; i < r; i += gridDim.x * TD_WARPS) {
    const double* row = A + (int64_t)(j + 1 + i) * lda + j + 1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int c = lane;
    for (; c + 96 < r; c += 128) {          // four independent 256-byte loads in flight per warp
      const double a0 = row[c], a1 = row[c + 32], a2 = row[c + 64], a3 = row[c + 96];
      const double v0 = c == 0 ? 1.0 : x[c] * h.scale, v1 = x[c + 32] * h.scale, v2 = x[c + 64] * h.scale,
                   v3 = x[c + 96] * h.scale;
      s0 = fma(a0, v0, s0); s1 = fma(a1, v1, s1); s2 = fma(a2, v2, s2); s3 = fma(a3, v3, s3);
    }
    for (; c < r; c += 32) s0 = fma(row[c], c == 0 ? 1.0 : x[c] * h.scale, s0);
    const double s = warp_sum((s0 + s1) + (s2 + s3));
    if (lane == 0) pbuf[i] = h.tau * s;
  }
}

// pass B of column j:  w = p - (tau/2)(p^T v) v ;  A22 -= v w^T + w v^T ; CTA 0 stores v into row j of A.
__global__ void __launch_bounds__(TD_THREADS)
tridiag_rank2_kernel(double* __restrict__ A, int64_t lda, int n, int j, const double* __restrict__ vbuf,
                     const double* __restrict__ pbuf, const double* __restrict__ tau) {
  __shared__ double red[TD_WARPS];
  const int r = n - j - 1;
  const double t = tau[j];
  if (t == 0.0) return;                       // H = I: nothing to update (row j already holds v = e_1 pattern)
  double s = 0.0;
  for (int c = threadIdx.x; c < r; c += TD_THREADS) s = fma(pbuf[c], vbuf[c], s);
  const double coef = 0.5 * t * block_sum(s, red);
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int i = blockIdx.x * TD_WARPS + warp; i < r; i += gridDim.x * TD_WARPS) {
    double* row = A + (int64_t)(j + 1 + i) * lda + j + 1;
    const double vi = vbuf[i], wi = pbuf[i] - coef * vi;
    int c = lane;
    for (; c + 96 < r; c += 128) {          // four independent read-modify-writes in flight per warp
      double a[4], vc[4], pc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = row[c + 32 * u]; vc[u] = vbuf[c + 32 * u]; pc[u] = pbuf[c + 32 * u]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) row[c + 32 * u] = a[u] - fma(vi, pc[u] - coef * vc[u], wi * vc[u]);
    }
    for (; c < r; c += 32) {
      const double vc = vbuf[c], wc = pbuf[c] - coef * vc;
      row[c] -= fma(vi, wc, wi * vc);
    }
  }
  if (blockIdx.x == 0) {
    double* x = A + (int64_t)j * lda + j + 1;
    for (int c = threadIdx.x; c < r; c += TD_THREADS) x[c] = vbuf[c];
  }
}

__global__ void tridiag_tail_kernel(const double* __restrict__ A, int64_t lda, int n, double* __restrict__ d,
                                    double* __restrict__ e, double* __restrict__ tau) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (n >= 2) {
    d[n - 2] = A[(int64_t)(n - 2) * lda + n - 2];
    e[n - 2] = A[(int64_t)(n - 2) * lda + n - 1];
    tau[n - 2] = 0.0;
  }
  d[n - 1] = A[(int64_t)(n - 1) * lda + n - 1];
  if (n >= 1) { e[n - 1] = 0.0; tau[n - 1] = 0.0; }
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
  const size_t need = (size_t)(2 * n) * sizeof(double);
  if (!workspace || workspace_bytes < need) {
    set_error("tridiag_reduce: workspace too small (%zu < %zu)", workspace_bytes, need);
    return ERA5SVD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  double* vbuf = (double*)workspace;
  double* pbuf = vbuf + n;
  const int sms = sm_count();
  for (int j = 0; j + 2 < (int)n; ++j) {
    const int r = (int)n - j - 1;
    int grid = (int)ceil_div(r, TD_WARPS);
    if (grid > 8 * sms) grid = 8 * sms;      // 64 resident warps per SM: one row per warp up to n ~ 9.5 k
    tridiag_symv_kernel<<<grid, TD_THREADS, 0, st>>>(A, lda, (int)n, j, vbuf, pbuf, d, e, tau);
    tridiag_rank2_kernel<<<grid, TD_THREADS, 0, st>>>(A, lda, (int)n, j, vbuf, pbuf, tau);
    count_launch(2);
  }
  tridiag_tail_kernel<<<1, 32, 0, st>>>(A, lda, (int)n, d, e, tau);
  return check_launch("tridiag_reduce");
}

size_t era5svd_tridiag_reduce_workspace_bytes(int64_t n) { return n > 0 ? (size_t)(2 * n) * sizeof(double) : 0; }

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
