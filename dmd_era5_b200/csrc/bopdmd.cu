// (d') Optimized DMD / BOP-DMD on the SVD-projected coefficients, batched over bagging trials.
//
// NOT in the reference (ClimeTrend/DMD-ERA5 only cites it: README.md:85, :139); named by BASELINE.json's north_star
// ("BOP-DMD bagging trials on the projected coefficients run as batched per-CTA kernels", configs[4]).  Algorithm:
// variable projection with Levenberg-Marquardt (Askham & Kutz 2018) restated in oracle/bopdmd_np.py; bagging as in
// Sashidhar & Kutz 2022.  Model per trial k (snapshot subset idx[k][0..p)):
//
//     H[idx] ~= Phi B,   Phi[i][j] = exp(alpha_j t_i)            H: n_time x N real (N = n_components), alpha in C^r
//
// Everything an LM iteration needs follows from five reductions over the p snapshots of a trial,
//     G = Phi^H Phi,  F = Phi^H T Phi,  E2 = Phi^H T^2 Phi   (r x r),     C = Phi^H H,  Ct = Phi^H T H   (r x N),
// which are REAL float64 GEMMs on Psi = [Re Phi | Im Phi] (p x 2r):  S_w = Psi^T T^w Psi,  R_w = Psi^T T^w H.
//
//   bop_phi_kernel   : Psi for every trial                                   (memory bound, written once per iteration)
//   bop_gram_kernel  : batched 64 x 64-tile FP64 GEMM, all T^w weights from one pass over the tiles    (FP64-FMA bound)
//   bop_step_kernel  : ONE CTA PER TRIAL: complex Cholesky of G in shared memory, B = G^-1 C, P = G^-1 F, rho,
//                      accept / reject against the trial's accepted state, J^H J = (E2 - F^H P) o conj(B B^H),
//                      rhs = rowsum(conj(B) o (Ct - F^H B)), LM solve, next candidate alpha
// Trials that have converged (done[k] != 0) are skipped by every kernel.
#include "common.cuh"

namespace era5svd {
namespace {

struct cd {
  double x, y;
};
__device__ __forceinline__ cd cmul(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cmulc(cd a, cd b) { return {a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }   // a * conj(b)
__device__ __forceinline__ cd cconjmul(cd a, cd b) { return {a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x}; } // conj(a) * b
__device__ __forceinline__ cd csub(cd a, cd b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cd cadd(cd a, cd b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cd cscale(cd a, double s) { return {a.x * s, a.y * s}; }

// Psi[k][i][j] = Re exp(alpha_kj t_i), Psi[k][i][r + j] = Im exp(alpha_kj t_i),  t_i = t[idx[k][i]]
__global__ void __launch_bounds__(256)
bop_phi_kernel(const cd* __restrict__ alpha, const double* __restrict__ t, const int* __restrict__ idx, int K, int p,
               int r, double* __restrict__ Psi, const int* __restrict__ done) {
  const int k = blockIdx.y;
  if (done && done[k]) return;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)p * r) return;
  const int i = (int)(e / r), j = (int)(e % r);
  const double ti = t[idx[(int64_t)k * p + i]];
  const cd a = alpha[(int64_t)k * r + j];
  const double mag = exp(a.x * ti);
  double sn, cs;
  sincos(a.y * ti, &sn, &cs);
  double* row = Psi + ((int64_t)k * p + i) * (2 * r);
  row[j] = mag * cs;
  row[r + j] = mag * sn;
}

// C[k][w][a][b] = sum_i Psi[k][i][a] * t_i^w * Bm[i][b],  w < W;   Bm = Psi[k] (nb = 2r) or the gathered H rows (nb = N)
template <int W, bool GATHER>
__global__ void __launch_bounds__(256)
bop_gram_kernel(const double* __restrict__ Psi, const double* __restrict__ t, const int* __restrict__ idx,
                const double* __restrict__ H, int64_t ldh, int K, int p, int na, int nb, double* __restrict__ C,
                const int* __restrict__ done) {
  constexpr int TS = 64, KS = 16;
  const int k = blockIdx.z;
  if (done && done[k]) return;
  __shared__ double As[KS][TS + 1];
  __shared__ double Bs[KS][TS + 1];
  __shared__ double Ts[KS];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  const int a0 = blockIdx.y * TS, b0 = blockIdx.x * TS;
  const double* Pk = Psi + (int64_t)k * p * na;
  const int* ik = idx + (int64_t)k * p;
  double acc[W][4][4];
#pragma unroll
  for (int w = 0; w < W; ++w)
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[w][u][v] = 0.0;
  for (int i0 = 0; i0 < p; i0 += KS) {
    // 16 x 64 elements of each operand, 4 per thread; rows beyond p read as zero
    for (int e = threadIdx.x; e < KS * TS; e += 256) {
      const int ii = e / TS, cc = e % TS;
      const int i = i0 + ii;
      double av = 0.0, bv = 0.0;
      if (i < p) {
        if (a0 + cc < na) av = Pk[(int64_t)i * na + a0 + cc];
        if (b0 + cc < nb) bv = GATHER ? H[(int64_t)ik[i] * ldh + b0 + cc] : Pk[(int64_t)i * na + b0 + cc];
      }
      As[ii][cc] = av;
      Bs[ii][cc] = bv;
    }
    if (threadIdx.x < KS) Ts[threadIdx.x] = (i0 + threadIdx.x < p) ? t[ik[i0 + threadIdx.x]] : 0.0;
    __syncthreads();
#pragma unroll
    for (int ii = 0; ii < KS; ++ii) {
      double a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { a[u] = As[ii][ty + 16 * u]; b[u] = Bs[ii][tx + 16 * u]; }
      double wgt = 1.0;
#pragma unroll
      for (int w = 0; w < W; ++w) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double aw = a[u] * wgt;
#pragma unroll
          for (int v = 0; v < 4; ++v) acc[w][u][v] = fma(aw, b[v], acc[w][u][v]);
        }
        wgt *= Ts[ii];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int w = 0; w < W; ++w) {
    double* Ck = C + ((int64_t)k * W + w) * na * nb;
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int a = a0 + ty + 16 * u, b = b0 + tx + 16 * v;
        if (a < na && b < nb) Ck[(int64_t)a * nb + b] = acc[w][u][v];
      }
  }
}

// hn2[k] = sum over the trial's snapshots of ||h(t_i)||^2
__global__ void __launch_bounds__(256)
bop_hnorm_kernel(const double* __restrict__ H, int64_t ldh, const int* __restrict__ idx, int p, int N,
                 double* __restrict__ hn2) {
  __shared__ double red[8];
  const int k = blockIdx.x;
  double s = 0.0;
  for (int64_t e = threadIdx.x; e < (int64_t)p * N; e += 256) {
    const double v = H[(int64_t)idx[(int64_t)k * p + e / N] * ldh + e % N];
    s = fma(v, v, s);
  }
  s = warp_sum(s);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += red[i];
    hn2[k] = tot;
  }
}

constexpr int ST_THREADS = 1024;
constexpr int ST_SPLIT = 4;            // threads per right-hand side in the triangular solves

__device__ double block_sum_st(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = v;
  __syncthreads();
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < ST_THREADS / 32; ++i) s += red[i];
  return s;
}

// In-place complex Cholesky A = L L^H of a Hermitian positive definite r x r matrix, L[i * ld + j], lower triangle.
// *ok (shared) becomes 0 on a non-positive pivot.  Right-looking, two barriers per column.
__device__ void chol_lower(cd* L, int r, int ld, int* ok) {
  for (int j = 0; j < r; ++j) {
    __syncthreads();
    const double djj = L[j * ld + j].x;
    if (!(djj > 0.0) || !*ok) {
      if (threadIdx.x == 0) *ok = 0;
      __syncthreads();
      return;
    }
    const double inv = rsqrt(djj);
    for (int i = j + 1 + threadIdx.x; i < r; i += ST_THREADS) L[i * ld + j] = cscale(L[i * ld + j], inv);
    __syncthreads();
    if (threadIdx.x == 0) L[j * ld + j] = {sqrt(djj), 0.0};   // after the barrier: every thread has read d_jj above
    const int m = r - j - 1;
    for (int e = threadIdx.x; e < m * m; e += ST_THREADS) {
      const int i = j + 1 + e / m, c = j + 1 + e % m;
      if (c <= i) L[i * ld + c] = csub(L[i * ld + c], cmulc(L[i * ld + j], L[c * ld + j]));
    }
  }
  __syncthreads();
}

// X <- (L L^H)^-1 X for the ncol columns of the ROW-MAJOR matrix X[row * ldx + col] (global memory).  ST_SPLIT adjacent
// lanes share one column: each takes every ST_SPLIT-th term of the substitution sum (independent, coalesced loads of
// X rows; L broadcast from shared memory) and the partial sums meet through shuffles.
__device__ void chol_solve_rows(const cd* L, int r, int ld, cd* X, int ldx, int ncol) {
  const int part = threadIdx.x % ST_SPLIT;
  const int ngrp = ST_THREADS / ST_SPLIT;
  for (int c0 = 0; c0 < ncol; c0 += ngrp) {
    const int c = c0 + threadIdx.x / ST_SPLIT;
    const bool on = c < ncol;
    for (int i = 0; i < r; ++i) {             // L y = b
      cd s = {0.0, 0.0};
      if (on)
        for (int q = part; q < i; q += ST_SPLIT) s = cadd(s, cmul(L[i * ld + q], X[(int64_t)q * ldx + c]));
#pragma unroll
      for (int o = ST_SPLIT / 2; o > 0; o >>= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
      }
      if (on && part == 0) X[(int64_t)i * ldx + c] = cscale(csub(X[(int64_t)i * ldx + c], s), 1.0 / L[i * ld + i].x);
      __syncwarp();
    }
    for (int i = r - 1; i >= 0; --i) {        // L^H z = y
      cd s = {0.0, 0.0};
      if (on)
        for (int q = i + 1 + part; q < r; q += ST_SPLIT) s = cadd(s, cconjmul(L[q * ld + i], X[(int64_t)q * ldx + c]));
#pragma unroll
      for (int o = ST_SPLIT / 2; o > 0; o >>= 1) {
        s.x += __shfl_xor_sync(0xffffffffu, s.x, o);
        s.y += __shfl_xor_sync(0xffffffffu, s.y, o);
      }
      if (on && part == 0) X[(int64_t)i * ldx + c] = cscale(csub(X[(int64_t)i * ldx + c], s), 1.0 / L[i * ld + i].x);
      __syncwarp();
    }
  }
  __syncthreads();
}

struct StepParams {
  const double* S;      // [K][3][2r][2r]
  const double* R;      // [K][2][2r][N]
  const double* hn2;    // [K]
  int K, r, N;
  cd* alpha;            // [K][r] accepted
  cd* alpha_try;        // [K][r] candidate just evaluated -> next candidate
  double* rho;          // [K] accepted objective (+inf before the first evaluation)
  double* lam;          // [K]
  cd* JhJ;              // [K][r][r] of the accepted point
  cd* rhs;              // [K][r]
  cd* Bout;             // [K][r][N] amplitudes / modes of the accepted point
  int* done;            // [K]
  cd* scratch;          // per trial: X = [C | F] -> [B | P] (r x (N + r), row-major), F (r x r), W (r x r), E (r x N)[, L]
  int64_t scratch_stride;   // complex elements per trial
  double nu, tol;
  int use_smem;
};

__global__ void __launch_bounds__(ST_THREADS, 1)
bop_step_kernel(const StepParams q) {
  extern __shared__ double st_smem[];
  __shared__ double red[ST_THREADS / 32];
  __shared__ int ok, accept, finished;
  const int k = blockIdx.x;
  if (q.done[k]) return;
  const int r = q.r, N = q.N, r2 = 2 * r, M = N + r;
  const int ld = r + 1;
  cd* X = q.scratch + (int64_t)k * q.scratch_stride;    // [B | P] after the solve
  cd* Fm = X + (int64_t)r * M;                           // F = Phi^H T Phi, row-major
  cd* Wk = Fm + (int64_t)r * r;                          // B B^H, then F^H P
  cd* Ek = Wk + (int64_t)r * r;                          // F^H B  (r x N)
  cd* rowF = reinterpret_cast<cd*>(st_smem);             // staged row m of F (r) and of [B | P] (M)
  cd* rowX = rowF + r;
  cd* L = q.use_smem ? rowX + M : Ek + (int64_t)r * N;
  const double* S0 = q.S + (int64_t)k * 3 * r2 * r2;
  const double* S1 = S0 + (int64_t)r2 * r2;
  const double* S2 = S1 + (int64_t)r2 * r2;
  const double* R0 = q.R + (int64_t)k * 2 * r2 * N;
  const double* R1 = R0 + (int64_t)r2 * N;
  auto herm = [&](const double* S, int j, int l) -> cd {     // (Phi^H T^w Phi)[j][l] from the real 2r x 2r matrix
    return {S[(int64_t)j * r2 + l] + S[(int64_t)(r + j) * r2 + r + l], S[(int64_t)j * r2 + r + l] - S[(int64_t)(r + j) * r2 + l]};
  };
  auto rect = [&](const double* R, int j, int n) -> cd {     // (Phi^H T^w H)[j][n]
    return {R[(int64_t)j * N + n], -R[(int64_t)(r + j) * N + n]};
  };
  if (threadIdx.x == 0) { ok = 1; accept = 0; finished = 0; }
  for (int e = threadIdx.x; e < r * r; e += ST_THREADS) {
    const int i = e / r, j = e % r;
    L[i * ld + j] = herm(S0, i, j);
    const cd f = herm(S1, i, j);
    Fm[e] = f;
    X[(int64_t)i * M + N + j] = f;
  }
  for (int e = threadIdx.x; e < r * N; e += ST_THREADS) X[(int64_t)(e / N) * M + e % N] = rect(R0, e / N, e % N);
  __syncthreads();
  chol_lower(L, r, ld, &ok);
  double rho_try = __longlong_as_double(0x7ff0000000000000LL);
  if (ok) {
    chol_solve_rows(L, r, ld, X, M, M);
    double s = 0.0;                                       // rho = ||H||^2 - Re tr(C^H B)
    for (int e = threadIdx.x; e < r * N; e += ST_THREADS) {
      const cd c = rect(R0, e / N, e % N), b = X[(int64_t)(e / N) * M + e % N];
      s += c.x * b.x + c.y * b.y;
    }
    rho_try = q.hn2[k] - block_sum_st(s, red);
  }
  const double rho_old = q.rho[k];
  if (threadIdx.x == 0) {
    if (ok && rho_try < rho_old) {
      accept = 1;
      finished = (rho_old < 1e300) && (rho_old - rho_try <= q.tol * rho_old);
      q.rho[k] = rho_try;
      q.lam[k] = fmax(q.lam[k] / q.nu, 1e-12);
    } else {
      const double l = q.lam[k] * q.nu;
      q.lam[k] = l;
      // no accepted point yet (singular start), runaway damping, or stagnation (the candidate is within tol of the
      // accepted objective: converged): stop this trial
      finished = (l > 1e12) || !(rho_old < 1e300) || (ok && rho_try - rho_old <= q.tol * rho_old);
    }
  }
  __syncthreads();
  if (accept) {
    // accepted: alpha <- alpha_try, store B, J^H J and rhs of this point
    for (int j = threadIdx.x; j < r; j += ST_THREADS) q.alpha[(int64_t)k * r + j] = q.alpha_try[(int64_t)k * r + j];
    cd* Bk = q.Bout + (int64_t)k * r * N;
    for (int e = threadIdx.x; e < r * N; e += ST_THREADS) Bk[e] = X[(int64_t)(e / N) * M + e % N];
    // Wk = B B^H: column n of B staged in shared memory per step; 8 outputs per thread and pass (registers)
    constexpr int CH = 8;
    for (int e0 = 0; e0 < r * r; e0 += CH * ST_THREADS) {
      cd acc[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u) acc[u] = {0.0, 0.0};
      for (int n = 0; n < N; ++n) {
        __syncthreads();
        for (int j = threadIdx.x; j < r; j += ST_THREADS) rowF[j] = X[(int64_t)j * M + n];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const int e = e0 + threadIdx.x + u * ST_THREADS;
          if (e < r * r) acc[u] = cadd(acc[u], cmulc(rowF[e / r], rowF[e % r]));
        }
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int e = e0 + threadIdx.x + u * ST_THREADS;
        if (e < r * r) Wk[e] = acc[u];
      }
    }
    __syncthreads();
    // [F^H B | F^H P]: rows m of F and of [B | P] staged per step; thread owns outputs (j, c), c fastest.
    // J^H J = (E2 - F^H P) o conj(B B^H);   E = Ct - F^H B
    cd* Jk = q.JhJ + (int64_t)k * r * r;
    for (int e0 = 0; e0 < r * M; e0 += CH * ST_THREADS) {
      cd acc[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u) acc[u] = {0.0, 0.0};
      for (int m = 0; m < r; ++m) {
        __syncthreads();
        for (int j = threadIdx.x; j < r; j += ST_THREADS) rowF[j] = Fm[(int64_t)m * r + j];
        for (int c = threadIdx.x; c < M; c += ST_THREADS) rowX[c] = X[(int64_t)m * M + c];
        __syncthreads();
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const int e = e0 + threadIdx.x + u * ST_THREADS;
          if (e < r * M) acc[u] = cadd(acc[u], cconjmul(rowF[e / M], rowX[e % M]));
        }
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int e = e0 + threadIdx.x + u * ST_THREADS;
        if (e < r * M) {
          const int j = e / M, c = e % M;
          if (c < N) {
            Ek[(int64_t)j * N + c] = csub(rect(R1, j, c), acc[u]);
          } else {
            const int l = c - N;
            const cd a1 = csub(herm(S2, j, l), acc[u]);
            const cd w = Wk[(int64_t)j * r + l];
            Jk[(int64_t)j * r + l] = cmul(a1, {w.x, -w.y});
          }
        }
      }
    }
    __syncthreads();
    // rhs_j = sum_n conj(B[j][n]) E[j][n]: one warp per row j
    for (int j = threadIdx.x / 32; j < r; j += ST_THREADS / 32) {
      cd s = {0.0, 0.0};
      for (int n = threadIdx.x % 32; n < N; n += 32) s = cadd(s, cconjmul(X[(int64_t)j * M + n], Ek[(int64_t)j * N + n]));
      s.x = warp_sum(s.x);
      s.y = warp_sum(s.y);
      if (threadIdx.x % 32 == 0) q.rhs[(int64_t)k * r + j] = s;
    }
    __syncthreads();
  }
  if (finished) {
    if (threadIdx.x == 0) q.done[k] = 1;
    return;
  }
  // next candidate: (J^H J + lam diag) delta = rhs at the accepted point
  const double lam = q.lam[k];
  const cd* Jk = q.JhJ + (int64_t)k * r * r;
  for (int e = threadIdx.x; e < r * r; e += ST_THREADS) {
    const int i = e / r, j = e % r;
    cd v = Jk[e];
    if (i == j) v.x += lam * v.x;
    L[i * ld + j] = v;
  }
  cd* dl = X;                                           // delta as a one-column row-major matrix (ldx = 1)
  for (int j = threadIdx.x; j < r; j += ST_THREADS) dl[j] = q.rhs[(int64_t)k * r + j];
  if (threadIdx.x == 0) ok = 1;
  __syncthreads();
  chol_lower(L, r, ld, &ok);
  if (!ok) {                                            // numerically singular model: stop this trial at the accepted point
    if (threadIdx.x == 0) q.done[k] = 1;
    return;
  }
  chol_solve_rows(L, r, ld, dl, 1, 1);
  for (int j = threadIdx.x; j < r; j += ST_THREADS) q.alpha_try[(int64_t)k * r + j] = cadd(q.alpha[(int64_t)k * r + j], dl[j]);
}

}  // namespace
}  // namespace era5svd

extern "C" {

static size_t bop_scratch_elems(int64_t r, int64_t N) {
  // [B | P] r x (N + r), F r x r, W r x r, E r x N, and L r x (r + 1) when it does not fit shared memory
  return (size_t)(r * (N + r) + 2 * r * r + r * N + r * (r + 1));
}

size_t era5svd_bop_workspace_bytes(int64_t K, int64_t p, int64_t r, int64_t N) {
  if (K <= 0 || p <= 0 || r <= 0 || N <= 0) return 0;
  const size_t psi = (size_t)K * p * 2 * r * 8;
  const size_t S = (size_t)K * 3 * 4 * r * r * 8, R = (size_t)K * 2 * 2 * r * N * 8;
  return psi + S + R + (size_t)K * bop_scratch_elems(r, N) * 16 + (size_t)K * 8 + 256;
}

// One Levenberg-Marquardt iteration for every trial that is not done (see the file header).
int era5svd_bop_iterate_f64(const double* H, int64_t n_time, int64_t N, int64_t ldh, const double* t, const int* idx,
                            int64_t K, int64_t p, int64_t r, double* alpha, double* alpha_try, double* rho,
                            double* lam, double* JhJ, double* rhs, double* Bout, int* done, double nu, double tol,
                            int first, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(H && t && idx && alpha && alpha_try && rho && lam && JhJ && rhs && Bout && done, "bop_iterate: null pointer");
  ERA5SVD_REQUIRE(K > 0 && p > 0 && r > 0 && N > 0 && ldh >= N && p <= n_time && K <= 65535 && r <= 128 && N <= 128,
                  "bop_iterate: bad shape K=%lld p=%lld r=%lld N=%lld", (long long)K, (long long)p, (long long)r, (long long)N);
  const size_t need = era5svd_bop_workspace_bytes(K, p, r, N);
  if (!workspace || workspace_bytes < need) {
    set_error("bop_iterate: workspace too small (%zu < %zu)", workspace_bytes, need);
    return ERA5SVD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  double* Psi = (double*)workspace;
  double* S = Psi + (size_t)K * p * 2 * r;
  double* R = S + (size_t)K * 3 * 4 * r * r;
  double* hn2 = R + (size_t)K * 2 * 2 * r * N;
  cd* scratch = (cd*)(((uintptr_t)(hn2 + K) + 15) & ~(uintptr_t)15);
  int rc;
  if (first) {
    bop_hnorm_kernel<<<(unsigned)K, 256, 0, st>>>(H, ldh, idx, (int)p, (int)N, hn2);
    if ((rc = check_launch("bop_hnorm_kernel"))) return rc;
  }
  dim3 gp((unsigned)ceil_div(p * r, 256), (unsigned)K);
  bop_phi_kernel<<<gp, 256, 0, st>>>((const cd*)alpha_try, t, idx, (int)K, (int)p, (int)r, Psi, done);
  if ((rc = check_launch("bop_phi_kernel"))) return rc;
  const int na = (int)(2 * r);
  dim3 gs((unsigned)ceil_div(na, 64), (unsigned)ceil_div(na, 64), (unsigned)K);
  bop_gram_kernel<3, false><<<gs, 256, 0, st>>>(Psi, t, idx, H, ldh, (int)K, (int)p, na, na, S, done);
  if ((rc = check_launch("bop_gram_kernel<S>"))) return rc;
  dim3 gr((unsigned)ceil_div(N, 64), (unsigned)ceil_div(na, 64), (unsigned)K);
  bop_gram_kernel<2, true><<<gr, 256, 0, st>>>(Psi, t, idx, H, ldh, (int)K, (int)p, na, (int)N, R, done);
  if ((rc = check_launch("bop_gram_kernel<R>"))) return rc;
  StepParams q;
  q.S = S; q.R = R; q.hn2 = hn2; q.K = (int)K; q.r = (int)r; q.N = (int)N;
  q.alpha = (cd*)alpha; q.alpha_try = (cd*)alpha_try; q.rho = rho; q.lam = lam; q.JhJ = (cd*)JhJ; q.rhs = (cd*)rhs;
  q.Bout = (cd*)Bout; q.done = done; q.scratch = scratch; q.nu = nu; q.tol = tol;
  q.scratch_stride = (int64_t)bop_scratch_elems(r, N);
  const size_t rows = (size_t)(2 * r + N) * 16;                 // staged rows of F and [B | P]
  const size_t lbytes = (size_t)r * (r + 1) * 16;
  q.use_smem = rows + lbytes <= 200 * 1024;
  const size_t smem = rows + (q.use_smem ? lbytes : 0);
  ERA5SVD_CUDA(cudaFuncSetAttribute(bop_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  bop_step_kernel<<<(unsigned)K, ST_THREADS, smem, st>>>(q);
  return check_launch("bop_step_kernel");
}

}  // extern "C"
