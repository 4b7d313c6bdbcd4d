// (c)/(d) Small float64 factor kernels: everything that is time-sized (n) or sketch-sized (l).
// These replace the LAPACK calls behind sklearn's normalisers and small SVD
// (scipy.linalg.lu / qr / svd, sklearn/utils/extmath.py:371-383, :615) and the eigensolve of the
// Gram-route standard SVD (np.linalg.svd, src/dmd_era5/era5_svd/era5_svd.py:251).
// They are latency-bound, replicated on every GPU, and run as one CTA (the solvers) or a handful of
// CTAs (the small GEMM).
#include <cstdlib>

#include "common.cuh"

namespace era5svd {

// ---------------------------------------------------------------------------------------------
// C = alpha * op(A) * op(B) + beta * C     (32 x 32 tile, 16 x 16 threads, 2 x 2 per thread)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gemm_f64_kernel(int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha,
                const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                double beta, double* __restrict__ C, int64_t ldc) {
  constexpr int TS = 32, KS = 32;
  __shared__ double As[KS][TS + 1];
  __shared__ double Bs[KS][TS + 1];
  const int t = threadIdx.x;
  const int ty = t / 16, tx = t % 16;
  const int64_t i0 = (int64_t)blockIdx.y * TS, j0 = (int64_t)blockIdx.x * TS;
  double acc[2][2] = {{0, 0}, {0, 0}};
  for (int64_t k0 = 0; k0 < K; k0 += KS) {
    // 32 x 32 elements each, 4 per thread
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = t + 256 * e;
      {
        // A element (i, k): choose the mapping that is contiguous in memory
        int ii, kk;
        if (transA) { ii = idx % TS; kk = idx / TS; } else { kk = idx % KS; ii = idx / KS; }
        int64_t gi = i0 + ii, gk = k0 + kk;
        double v = 0.0;
        if (gi < M && gk < K) v = transA ? A[gk * lda + gi] : A[gi * lda + gk];
        As[kk][ii] = v;
      }
      {
        int jj, kk;
        if (transB) { kk = idx % KS; jj = idx / KS; } else { jj = idx % TS; kk = idx / TS; }
        int64_t gj = j0 + jj, gk = k0 + kk;
        double v = 0.0;
        if (gj < N && gk < K) v = transB ? B[gj * ldb + gk] : B[gk * ldb + gj];
        Bs[kk][jj] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
      double a0 = As[kk][ty], a1 = As[kk][ty + 16];
      double b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
      acc[0][0] = fma(a0, b0, acc[0][0]);
      acc[0][1] = fma(a0, b1, acc[0][1]);
      acc[1][0] = fma(a1, b0, acc[1][0]);
      acc[1][1] = fma(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      int64_t gi = i0 + ty + 16 * a, gj = j0 + tx + 16 * b;
      if (gi < M && gj < N) {
        double* c = C + gi * ldc + gj;
        double v = alpha * acc[a][b];
        if (beta != 0.0) v += beta * (*c);
        *c = v;
      }
    }
}

// ---------------------------------------------------------------------------------------------
// Symmetric eigensolver: parallel cyclic two-sided Jacobi, one CTA of 1024 threads.
// Round-robin ordering: n_pad - 1 steps per sweep, n_pad / 2 disjoint (p, q) pairs per step.
// A group of 16 lanes owns one pair per step: it computes the rotation from the 2 x 2 block,
// applies it to columns p, q of A and V (phase 1), and after a block barrier to rows p, q of A
// (phase 2).  Working copies live in shared memory with an ODD leading dimension (conflict-free
// column walks for doubles) when they fit, else in the global workspace (one CTA: block barriers
// order the accesses; L1/L2 resident).  FP64-pipe bound: ~4 DFMA per updated element pair.
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Register-tiled right-looking Cholesky of an l x l (l <= 128) symmetric matrix by one CTA of 32 x 32 threads, row
// scaling deferred: thread (ty, tx) holds the entries rows ty + 32 a, columns tx + 32 b (a, b < 4) of the upper triangle in
// r[a][b] for the whole factorisation; on return row i holds the UNSCALED final row (R = D^-1/2 r).  Two images of the
// pivot row cross shared memory per step - prow_c = the row itself (column factors), prow_r = the row divided by its pivot
// for the columns > j and ZERO elsewhere (row factors; rows <= j are final) - both formed by the warp that owns the row
// (the pivot reaches its lanes by shuffle), so the other 31 warps execute no compare / select / multiply: 8 shared loads
// and <= 10 DFMAs per step, one barrier.  (The first version, with the pivot test and the scaling in every thread, issued
// 79 instructions per warp and step - ncu, profiles/r02_small_kernels.md.)  A dependent or non-positive pivot drops its
// direction (zero row factors, *dropped = 1), never a NaN.  prow_c / prow_r: [2][128], zeroed here.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void chol_register_tiled(double (&r)[4][4], int l, const double* __restrict__ gdiag,
                                                    double rel_tol, double (*prow_c)[128], double (*prow_r)[128],
                                                    int* dropped, int tx, int ty) {
  const int t = ty * 32 + tx;
  if (t < 256) { (&prow_c[0][0])[t] = 0.0; (&prow_r[0][0])[t] = 0.0; }
  if (t == 0) *dropped = 0;
  __syncthreads();
  auto publish = [&](int jn) {               // executed by the warp that owns row jn (ty == jn % 32): row jn is final
    const int an = jn >> 5, ln = jn & 31;
    double v[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) v[b] = an == 0 ? r[0][b] : (an == 1 ? r[1][b] : (an == 2 ? r[2][b] : r[3][b]));
    const double dsel = an == 0 ? v[0] : (an == 1 ? v[1] : (an == 2 ? v[2] : v[3]));
    const double d = __shfl_sync(0xffffffffu, dsel, ln);
    const double inv_d = (d > rel_tol * gdiag[jn] && d > 0.0) ? 1.0 / d : 0.0;
    if (inv_d == 0.0 && tx == 0) *dropped = 1;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int c = tx + 32 * b;
      if (c < l) {
        prow_c[jn & 1][c] = v[b];
        prow_r[jn & 1][c] = c > jn ? v[b] * inv_d : 0.0;
      }
    }
  };
  if (ty == 0) publish(0);
  for (int j = 0; j < l; ++j) {
    __syncthreads();
    const double* pwc = prow_c[j & 1];
    const double* pwr = prow_r[j & 1];
    double pc[4];
#pragma unroll
    for (int b = 0; b < 4; ++b) pc[b] = pwc[tx + 32 * b];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (ty + 32 * a > j) {                                   // warp-uniform: rows of this slot still open
        const double pra = pwr[ty + 32 * a];
#pragma unroll
        for (int b = a + 1; b < 4; ++b) r[a][b] = fma(-pra, pc[b], r[a][b]);
        if (tx >= ty) r[a][a] = fma(-pra, pc[a], r[a][a]);     // diagonal block: upper part only
      }
    }
    const int jn = j + 1;
    if (jn < l && ty == (jn & 31)) publish(jn);                // the next pivot row is final now
  }
  __syncthreads();                                             // *dropped is complete
}

constexpr int JAC_THREADS = 1024;
constexpr int JAC_GROUP = 16;
constexpr int JAC_GROUPS = JAC_THREADS / JAC_GROUP;

__device__ __forceinline__ void jacobi_pair(int step, int i, int n_pad, int& p, int& q) {
  const int r = n_pad - 1;
  if (i == 0) {
    p = r;
    q = step;
  } else {
    p = step + i;
    if (p >= r) p -= r;
    q = step + r - i;
    if (q >= r) q -= r;
  }
  if (p > q) { int tmp = p; p = q; q = tmp; }
}

// SMEM is a template parameter so that the shared-memory instantiation compiles to LDS / STS: with a run-time
// `use_smem ? smem : ws` pointer every access was a GENERIC load / store (LD.E / ST.E in the SASS).
template <bool SMEM>
__global__ void __launch_bounds__(JAC_THREADS, 1)
syevj_kernel(const double* __restrict__ A, int n, int64_t lda, double* __restrict__ W,
             double* __restrict__ V, int64_t ldv, int max_sweeps, double tol_in, double* __restrict__ ws,
             const int* __restrict__ run_if) {
  // run_if != NULL: this launch is the fallback of syev_chol_jacobi_kernel and runs only if that reported a
  // dropped pivot (*run_if == 1)
  if (run_if && *run_if == 0) return;
  extern __shared__ double smem[];
  const int t = threadIdx.x;
  const int g = t / JAC_GROUP, gl = t % JAC_GROUP;
  const unsigned gmask = 0xFFFFu << (16 * (g & 1));   // the 16 lanes of this group within its warp
  const int n_pad = n + (n & 1);
  const int npairs = n_pad / 2;
  const int ldw = n | 1;
  // layout: [cs: 2 * npairs doubles][rank: n ints][Aw][Vw]
  double* cs = smem;
  int* rank = reinterpret_cast<int*>(smem + 2 * npairs);
  double* Aw;
  if constexpr (SMEM) Aw = smem + 2 * npairs + (n + 1) / 2;
  else Aw = ws;
  double* Vw = Aw + n * ldw;
  for (int r = t / 32; r < n; r += JAC_THREADS / 32)
    for (int c = t % 32; c < n; c += 32) {
      Aw[r * ldw + c] = 0.5 * (A[(int64_t)r * lda + c] + A[(int64_t)c * lda + r]);  // enforce symmetry
      Vw[r * ldw + c] = (r == c) ? 1.0 : 0.0;
    }
  __syncthreads();

  // relative off-diagonal threshold |a_pq| <= tol * sqrt(|a_pp a_qq|), tol = eps * sqrt(n) (cf. xGESVJ)
  const double tol = tol_in > 0.0 ? tol_in : 2.220446049250313e-16 * sqrt((double)n);
  const double tol2 = tol * tol;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    int did = 0;
    for (int step = 0; step < n_pad - 1; ++step) {
      // phase 1: rotation parameters + right multiplication (columns p, q of A and V)
      for (int i = g; i < npairs; i += JAC_GROUPS) {
        int p, q;
        jacobi_pair(step, i, n_pad, p, q);
        double c = 1.0, s = 0.0;
        if (q < n) {  // q == n only for the padding index of odd n
          const double app = Aw[p * ldw + p], aqq = Aw[q * ldw + q], apq = Aw[p * ldw + q];
          __syncwarp(gmask);  // every lane of the group has read the 2 x 2 block before any lane overwrites it
          if (apq * apq > tol2 * fabs(app * aqq) && fabs(apq) > 1e-300) {
            // tan(theta) only steers convergence: any t gives an exactly orthogonal rotation as long as
            // c = 1 / sqrt(1 + t^2), s = t c are formed in float64.  So t comes from fast float32 ops
            // (relative error ~1e-7: the rotated a_pq drops by 1e-7 instead of to zero).
            const float dq = (float)(aqq - app), ap2 = 2.0f * (float)apq;
            double t64;
            if (fabsf(ap2) > 1e-30f && fabsf(dq) < 1e30f) {
              const float tau = __fdividef(dq, ap2);
              const float tf = __fdividef(copysignf(1.0f, tau), fabsf(tau) + sqrtf(fmaf(tau, tau, 1.0f)));
              t64 = (double)tf;
            } else {
              const double tau = (aqq - app) / (2.0 * apq);
              t64 = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            }
            c = rsqrt(fma(t64, t64, 1.0));
            s = t64 * c;
            did = 1;
            for (int r = gl; r < n; r += JAC_GROUP) {
              double x = Aw[r * ldw + p], y = Aw[r * ldw + q];
              Aw[r * ldw + p] = c * x - s * y;
              Aw[r * ldw + q] = s * x + c * y;
              x = Vw[r * ldw + p]; y = Vw[r * ldw + q];
              Vw[r * ldw + p] = c * x - s * y;
              Vw[r * ldw + q] = s * x + c * y;
            }
          }
        }
        if (gl == 0) { cs[2 * i] = c; cs[2 * i + 1] = s; }
      }
      __syncthreads();
      // phase 2: left multiplication (rows p, q of A)
      for (int i = g; i < npairs; i += JAC_GROUPS) {
        const double c = cs[2 * i], s = cs[2 * i + 1];
        if (s != 0.0) {
          int p, q;
          jacobi_pair(step, i, n_pad, p, q);
          for (int col = gl; col < n; col += JAC_GROUP) {
            const double x = Aw[p * ldw + col], y = Aw[q * ldw + col];
            Aw[p * ldw + col] = c * x - s * y;
            Aw[q * ldw + col] = s * x + c * y;
          }
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(did)) break;
  }
  // eigenvalues = diag(Aw); rank them in descending order (stable on ties) and scatter.
  for (int j = t; j < n; j += JAC_THREADS) {
    const double wj = Aw[j * ldw + j];
    int rk = 0;
    for (int i = 0; i < n; ++i) {
      const double wi = Aw[i * ldw + i];
      rk += (wi > wj) || (wi == wj && i < j);
    }
    W[rk] = wj;
    rank[j] = rk;
  }
  __syncthreads();
  for (int r = t / 32; r < n; r += JAC_THREADS / 32)
    for (int c = t % 32; c < n; c += 32) V[(int64_t)r * ldv + rank[c]] = Vw[r * ldw + c];
}

// ---------------------------------------------------------------------------------------------
// Symmetric POSITIVE DEFINITE eigensolver for the sketch-sized factors (n <= 128): Cholesky G = R^T R followed by
// ONE-SIDED Jacobi on the rows of R (Veselic / Hari).  Row rotations R <- J^T R leave R^T R = G unchanged and
// drive the rows to mutual orthogonality; at convergence R = Lambda^(1/2) U^T, i.e. eigenvalue j = |row j|^2 and
// eigenvector j = row j / |row j|.  Compared with the two-sided kernel above (columns of A, columns of V, rows of
// A: 580 KB of shared-memory traffic per step, two barriers) a step reads and writes each row once (194 KB, one
// barrier) and needs no eigenvector accumulation; the two-sided kernel is shared-memory-bandwidth bound (ncu: 79 %
// of the LSU wavefront peak).  One 16-lane group owns one (p, q) pair per step: its lanes hold the two rows in
// registers, form the three dot products (butterfly shuffles), rotate and store.
// A dropped Cholesky pivot (G numerically singular, e.g. a sketch wider than the rank) sets *status = 1 and leaves
// the outputs untouched: the caller then runs the two-sided kernel, which needs no definiteness.
// ---------------------------------------------------------------------------------------------
constexpr int J1_MAXN = 128;
constexpr int J1_PER_LANE = J1_MAXN / JAC_GROUP;     // 8 row elements per lane

template <typename TS>
__device__ __forceinline__ TS group16_sum(TS v, unsigned mask) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, 16);
  return v;
}

// TS = the type the SWEEPS run in.  double: the eigen-decomposition to working precision.  float (callers that ask for a
// relative decoupling >= 1e-5, i.e. the Rayleigh-Ritz rotation of the randomized driver, which is a choice of basis and not
// a result): the Cholesky factor is still formed in float64 (cond(G) ~ 1e7 is beyond float32) and then rotated as float32
// rows (cond(R) ~ 3e3) - FFMA instead of the 2-cycle DFMA, single instead of paired shuffles, half the shared-memory
// bytes; the dependent chain of a step, which bounds the sweeps, is about half as long.  Eigenvalues / vectors then carry
// float32 accuracy (~1e-6).
template <typename TS>
__global__ void __launch_bounds__(JAC_THREADS, 1)
syev_chol_jacobi_kernel(const double* __restrict__ A, int n, int64_t lda, double* __restrict__ W,
                        double* __restrict__ V, int64_t ldv, int max_sweeps, double tol_in, double pivot_tol,
                        int* __restrict__ status) {
  extern __shared__ double smem[];
  __shared__ double prow_c[2][128];
  __shared__ double prow_r[2][128];
  __shared__ int s_dropped;
  __shared__ double gdiag[J1_MAXN];
  __shared__ double lam[J1_MAXN];
  __shared__ TS nrm2[J1_MAXN];
  __shared__ int rank[J1_MAXN];
  const int t = threadIdx.x, tx = t % 32, ty = t / 32;
  const int nper = (n + JAC_GROUP - 1) / JAC_GROUP;     // row elements per lane of a 16-lane group (uniform)
  const int ldw = nper * JAC_GROUP;             // rows padded with zeros to whole groups: the sweeps load / store unguarded
  TS* Rw = reinterpret_cast<TS*>(smem);         // [n][ldw]
  for (int j = t; j < n; j += JAC_THREADS) gdiag[j] = A[(int64_t)j * lda + j];
  // ---- register-tiled right-looking Cholesky (chol_register_tiled), straight from global memory ----
  {
    double r[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = ty + 32 * a, c = tx + 32 * b;
        r[a][b] = (i < n && c < n && c >= i) ? 0.5 * (A[(int64_t)i * lda + c] + A[(int64_t)c * lda + i]) : 0.0;
      }
    __syncthreads();
    chol_register_tiled(r, n, gdiag, pivot_tol, prow_c, prow_r, &s_dropped, tx, ty);
    if (s_dropped) {                            // uniform over the CTA
      if (t == 0) *status = 1;
      return;
    }
    // R = D^(-1/2) * (unscaled rows); the diagonal owner of row i publishes the scale
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (ty == tx && ty + 32 * a < n) lam[ty + 32 * a] = rsqrt(r[a][a]);
    __syncthreads();
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = ty + 32 * a, c = tx + 32 * b;
        if (i < n && c < ldw) Rw[i * ldw + c] = (c < n && c >= i) ? (TS)(r[a][b] * lam[i]) : (TS)0;
      }
  }
  if (t == 0) *status = 0;
  __syncthreads();

  // ---- one-sided Jacobi on the rows of R ----
  const int g = t / JAC_GROUP, gl = t % JAC_GROUP;
  const unsigned gmask = 0xFFFFu << (16 * (g & 1));
  const int n_pad = n + (n & 1);
  const int npairs = n_pad / 2;
  const TS tiny = sizeof(TS) == 8 ? (TS)1e-300 : (TS)1e-37;
  const double tol = tol_in > 0.0 ? tol_in : 2.220446049250313e-16 * sqrt((double)n);
  const TS tol2 = (TS)(tol * tol);
  // Early stop for a caller-given tolerance (float32 data paths): the iteration converges quadratically, so a sweep
  // that met no pair above sqrt(tol / 100) leaves every pair below tol (the factor 100 covers relative gaps down
  // to ~1 %); the usual extra sweep that only verifies convergence is then skipped.  With tol_in <= 0 (float64
  // parity) the sweep-without-rotation rule stays.
  const bool early = tol_in > 0.0;
  const TS tol_early = (TS)(tol * 0.01);
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    int did = 0;
    // squared row norms, recomputed once per sweep and then carried through the rotations (a'_pp = a_pp - t a_pq,
    // a'_qq = a_qq + t a_pq): they only steer the angle and the convergence test, so a step needs ONE dot product
    // (a_pq) instead of three.  The eigenvalues come from fresh norms below.
    for (int i = g; i < n; i += JAC_GROUPS) {
      TS sq = 0;
      for (int c = gl; c < n; c += JAC_GROUP) sq = fma(Rw[i * ldw + c], Rw[i * ldw + c], sq);
      sq = group16_sum(sq, gmask);
      if (gl == 0) nrm2[i] = sq;
    }
    __syncthreads();
    for (int step = 0; step < n_pad - 1; ++step) {
      if (g < npairs) {
        int p, q;
        jacobi_pair(step, g, n_pad, p, q);
        if (q < n) {
          TS x[J1_PER_LANE], y[J1_PER_LANE];
          const TS app = nrm2[p], aqq = nrm2[q];
          TS apq = 0;
          TS* px = Rw + p * ldw + gl;
          TS* py = Rw + q * ldw + gl;
#pragma unroll
          for (int k = 0; k < J1_PER_LANE; ++k) {
            if (k < nper) {                    // uniform; the pad columns hold zeros and stay zero under rotations
              x[k] = px[JAC_GROUP * k];
              y[k] = py[JAC_GROUP * k];
              apq = fma(x[k], y[k], apq);
            }
          }
          apq = group16_sum(apq, gmask);
          const bool big = apq * apq > tol2 * app * aqq && fabs(apq) > tiny;
          if (early ? (apq * apq > tol_early * app * aqq) : big) did = 1;
          if (big) {
            // rotation that orthogonalises rows p, q (same angle as the two-sided rotation of the 2 x 2 Gram block);
            // tan(theta) from fast float32 ops, c and s completed in the sweep type (exactly orthogonal rotation)
            const float dq = (float)(aqq - app), ap2 = 2.0f * (float)apq;
            TS tt;
            if (fabsf(ap2) > 1e-30f && fabsf(dq) < 1e30f) {
              const float tau = __fdividef(dq, ap2);
              const float tf = __fdividef(copysignf(1.0f, tau), fabsf(tau) + sqrtf(fmaf(tau, tau, 1.0f)));
              tt = (TS)tf;
            } else {
              const double tau = ((double)aqq - (double)app) / (2.0 * (double)apq);
              tt = (TS)((tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau)));
            }
            const TS cs = rsqrt(fma(tt, tt, (TS)1));
            const TS sn = tt * cs;
            if (gl == 0) { nrm2[p] = app - tt * apq; nrm2[q] = aqq + tt * apq; }
#pragma unroll
            for (int k = 0; k < J1_PER_LANE; ++k) {
              if (k < nper) {
                px[JAC_GROUP * k] = cs * x[k] - sn * y[k];
                py[JAC_GROUP * k] = sn * x[k] + cs * y[k];
              }
            }
          }
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(did)) break;
  }
  // ---- eigenvalues = squared row norms, eigenvectors = normalised rows; descending order ----
  for (int i = g; i < n; i += JAC_GROUPS) {
    double s = 0.0;
    for (int c = gl; c < n; c += JAC_GROUP) s = fma((double)Rw[i * ldw + c], (double)Rw[i * ldw + c], s);
    s = group16_sum(s, gmask);
    if (gl == 0) lam[i] = s;
  }
  __syncthreads();
  for (int j = t; j < n; j += JAC_THREADS) {
    const double wj = lam[j];
    int rk = 0;
    for (int i = 0; i < n; ++i) {
      const double wi = lam[i];
      rk += (wi > wj) || (wi == wj && i < j);
    }
    W[rk] = wj;
    rank[j] = rk;
  }
  __syncthreads();
  for (int r = ty; r < n; r += 32) {           // r = component, c = eigenvector (row of R)
    for (int c = tx; c < n; c += 32) V[(int64_t)r * ldv + rank[c]] = (double)Rw[c * ldw + r] * rsqrt(lam[c]);
  }
}

// ---------------------------------------------------------------------------------------------
// Cholesky G = R^T R (upper R) + explicit inverse, one CTA (32 x 32 threads).  Working copies in
// shared memory when they fit (l <= 118), else in the caller's R / Rinv buffers.
// ---------------------------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(1024, 1)
chol_inv_kernel(const double* __restrict__ G, int l, int64_t ldg, double* __restrict__ R,
                int64_t ldr, double* __restrict__ Rinv, int64_t ldri, double rel_tol) {
  constexpr bool use_smem = SMEM;
  extern __shared__ double smem[];
  const int t = threadIdx.x, tx = t % 32, ty = t / 32;
  // smem: [gdiag: l][scale: l][Rw][Iw]  (Rw, Iw only when use_smem)
  double* gdiag = smem;
  double* scale = smem + l;
  const int ldw = use_smem ? (l | 1) : 0;
  double* Rw;
  double* Iw;
  if constexpr (SMEM) {
    Rw = smem + 2 * l;
    Iw = Rw + l * ldw;
  } else {
    Rw = R;
    Iw = Rinv;
  }
  // 32-bit index arithmetic: l <= 8192, so l * ld fits comfortably
  const int ld_r = use_smem ? ldw : (int)ldr, ld_i = use_smem ? ldw : (int)ldri;
  // copy the upper triangle (symmetrised), zero the strictly lower part
  for (int r = ty; r < l; r += 32)
    for (int c = tx; c < l; c += 32) {
      Rw[r * ld_r + c] = (c >= r) ? 0.5 * (G[(int64_t)r * ldg + c] + G[(int64_t)c * ldg + r]) : 0.0;
      Iw[r * ld_i + c] = 0.0;
    }
  for (int j = t; j < l; j += 1024) gdiag[j] = G[(int64_t)j * ldg + j];
  // Right-looking Cholesky with the row scaling deferred: one barrier per step.  At step j the
  // (unscaled) pivot row is final; trailing rows get  R[i][c] -= R[j][i] R[j][c] / d_j.
  const bool small = l <= 128;
  if (small) {
    // Small factor (the sketch width): REGISTER-TILED right-looking Cholesky.  Thread (ty, tx) keeps the entries
    // rows ty + 32 a, columns tx + 32 b (a, b < 4) of the trailing matrix in registers for the whole factorisation.
    // Per step only the pivot row travels through shared memory (published by its owners right after their update,
    // double buffered, one barrier per step), and the reciprocal pivot is computed once, by the thread that owns
    // it.  ~45 instructions per thread and step; the smem read-modify-write version was issue bound
    // (ncu: 915 k warp instructions, 69 % issue utilisation for l = 110).
    __shared__ double prow_c[2][128];
    __shared__ double prow_r[2][128];
    __shared__ int s_dropped;
    __syncthreads();                         // Rw (upper triangle, symmetrised) and gdiag are in place
    double r[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = ty + 32 * a, c = tx + 32 * b;
        r[a][b] = (i < l && c < l && c >= i) ? Rw[i * ld_r + c] : 0.0;
      }
    chol_register_tiled(r, l, gdiag, rel_tol, prow_c, prow_r, &s_dropped, tx, ty);
    // back to the working copy (the triangular inverse below reads columns of R)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int i = ty + 32 * a, c = tx + 32 * b;
        if (i < l && c < l && c >= i) Rw[i * ld_r + c] = r[a][b];
      }
  } else {
    for (int j = 0; j < l; ++j) {
      __syncthreads();
      const double d = Rw[j * ld_r + j];
      // dependent (or non-positive) pivot: the direction is dropped (no update, zero row), never NaN
      const double inv_d = (d > rel_tol * gdiag[j] && d > 0.0) ? 1.0 / d : 0.0;
      const double* prow = Rw + j * ld_r;
      for (int i = j + 1 + ty; i < l; i += 32) {
        const double f = prow[i] * inv_d;
        double* row = Rw + i * ld_r;
        // columns c >= i only, lanes strided: first column handled by lane ((i - j - 1) % 32)
        for (int c = j + 1 + tx; c < l; c += 32)
          if (c >= i) row[c] = fma(-f, prow[c], row[c]);
      }
    }
  }
  __syncthreads();
  for (int j = t; j < l; j += 1024) {
    const double d = Rw[j * ld_r + j];
    scale[j] = (d > rel_tol * gdiag[j] && d > 0.0) ? rsqrt(d) : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < l; r += 32) {
    const double sc = scale[r];
    for (int c = r + tx; c < l; c += 32) {
      const double v = Rw[r * ld_r + c];
      Rw[r * ld_r + c] = (c == r) ? (sc > 0.0 ? v * sc : 1e150) : v * sc;
    }
  }
  __syncthreads();
  // inverse, one warp per column c of Rinv: solve R x = e_c by column-oriented back substitution:
  // step i finalises x_i = x_i / r_ii (reciprocal precomputed) and the lanes subtract r_ki x_i from the
  // entries k < i: no reductions, no divisions in the loop.
  for (int j = t; j < l; j += 1024) scale[j] = 1.0 / Rw[j * ld_r + j];      // reuse: reciprocal diagonal
  __syncthreads();
  // x lives in column c of Iw itself (zero-initialised above; stride ld_i is odd in shared memory, so the
  // column walk is conflict free)
  if (l <= 128) {
    // register-resident variant: the warp solves its (up to four) columns ty, ty + 32, ... TOGETHER; lane tx holds
    // rows tx + 32 m of each.  One step = four conflict-free column loads of R shared by the four solves, one
    // shuffle broadcast and four DFMAs per column; the four dependency chains interleave.
    double x[4][4];
    int c_hi = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ck = ty + 32 * k;
      if (ck < l) c_hi = ck;
#pragma unroll
      for (int m = 0; m < 4; ++m) x[k][m] = (tx + 32 * m == ck && ck < l) ? 1.0 : 0.0;
    }
    // The solved entry x_i stays UNSCALED in its owner lane (the factor 1 / r_ii is applied once at the end), and the
    // strictly lower part of Rw is zero, so one step is: (slot + 1) column loads, the reciprocal diagonal, one shuffle
    // broadcast + one multiply per solve and (slot + 1) DFMAs per solve - no per-row compare / select except for the owner
    // lane's own diagonal entry.  Lanes whose row lies beyond l read a clamped (valid) row and hold values nobody uses.
    const double* rowp[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) rowp[m] = Rw + (tx + 32 * m < l ? tx + 32 * m : l - 1) * ld_r;
    // rows are visited from the bottom; the 32-row slot of the current row is a compile-time constant inside each
    // of the four blocks below, so x[k][slot] needs no dynamic register indexing and rows of higher slots (> i,
    // already final) are not touched
#pragma unroll
    for (int slot = 3; slot >= 0; --slot) {
      const int i_top = c_hi < 32 * slot + 31 ? c_hi : 32 * slot + 31;
      for (int i = i_top; i >= 32 * slot; --i) {
        const int owner = i & 31;
        const double sc = scale[i];
        double rk[4];
#pragma unroll
        for (int m = 0; m <= slot; ++m) rk[m] = rowp[m][i];
        if (tx >= owner) rk[slot] = 0.0;       // own diagonal entry, and (rows beyond l only) the clamped row's values
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const double xi = __shfl_sync(0xffffffffu, x[k][slot], owner) * sc;
#pragma unroll
          for (int m = 0; m <= slot; ++m) x[k][m] = fma(-rk[m], xi, x[k][m]);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double sc = tx + 32 * m < l ? scale[tx + 32 * m] : 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) x[k][m] *= sc;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ck = ty + 32 * k;
      if (ck < l)
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int row = tx + 32 * m;
          if (row <= ck) Iw[row * ld_i + ck] = x[k][m];
        }
    }
  } else {
    for (int c = ty; c < l; c += 32) {
      if (tx == 0) Iw[c * ld_i + c] = 1.0;
      __syncwarp();
      for (int i = c; i >= 0; --i) {
        const double xi = Iw[i * ld_i + c] * scale[i];
        __syncwarp();
        if (tx == 0) Iw[i * ld_i + c] = xi;
        for (int k = tx; k < i; k += 32) Iw[k * ld_i + c] = fma(-Rw[k * ld_r + i], xi, Iw[k * ld_i + c]);
        __syncwarp();
      }
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int r = ty; r < l; r += 32)
      for (int c = tx; c < l; c += 32) {
        R[(int64_t)r * ldr + c] = Rw[r * ldw + c];
        Rinv[(int64_t)r * ldri + c] = Iw[r * ldw + c];
      }
  }
}

__global__ void __launch_bounds__(128)
col_normalize_kernel(double* __restrict__ P, int64_t n, int64_t l, int64_t ldp,
                     double* __restrict__ norms_out) {
  const int64_t j = blockIdx.x;
  __shared__ double red[4];
  __shared__ double s_norm;
  double s = 0.0;
  for (int64_t r = threadIdx.x; r < n; r += blockDim.x) {
    double v = P[r * ldp + j];
    s = fma(v, v, s);
  }
  s = warp_sum(s);
  if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double nrm = sqrt(red[0] + red[1] + red[2] + red[3]);
    s_norm = nrm;
    if (norms_out) norms_out[j] = nrm;
  }
  __syncthreads();
  const double nrm = s_norm;
  if (nrm > 0.0)
    for (int64_t r = threadIdx.x; r < n; r += blockDim.x) P[r * ldp + j] /= nrm;
}

template <typename Ts, typename Td>
__global__ void __launch_bounds__(256)
convert_kernel(const Ts* __restrict__ src, int64_t lds, Td* __restrict__ dst, int64_t ldd,
               int64_t rows, int64_t cols) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  int64_t r = idx / cols, c = idx % cols;
  dst[r * ldd + c] = (Td)src[r * lds + c];
}

__global__ void __launch_bounds__(256)
scale_rows_kernel(double* __restrict__ V, int64_t k, int64_t n, int64_t ldv,
                  const double* __restrict__ scale) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= k * n) return;
  int64_t r = idx / n, c = idx % n;
  V[r * ldv + c] *= scale[r];
}

// s = sqrt(max(w, 0)), inv_s = 1 / s (0 where s == 0)
__global__ void __launch_bounds__(128)
sigma_from_eig_kernel(const double* __restrict__ w, int64_t l, double* __restrict__ s,
                      double* __restrict__ inv_s) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= l) return;
  double v = w[i];
  double sv = v > 0.0 ? sqrt(v) : 0.0;
  s[i] = sv;
  if (inv_s) inv_s[i] = sv > 0.0 ? 1.0 / sv : 0.0;
}

}  // namespace era5svd

extern "C" {

int era5svd_sigma_from_eig_f64(const double* w, int64_t l, double* s, double* inv_s, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(w && s && l > 0, "sigma_from_eig: bad argument");
  sigma_from_eig_kernel<<<(unsigned)ceil_div(l, 128), 128, 0, as_stream(stream)>>>(w, l, s, inv_s);
  return check_launch("sigma_from_eig_kernel");
}

int era5svd_gemm_f64(int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                     double* C, int64_t ldc, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(A && B && C, "gemm_f64: null pointer");
  ERA5SVD_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_f64: bad shape");
  ERA5SVD_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "gemm_f64: bad leading dimension");
  ERA5SVD_REQUIRE(ceil_div(M, 32) <= 65535, "gemm_f64: M too large");
  dim3 grid((unsigned)ceil_div(N, 32), (unsigned)ceil_div(M, 32));
  gemm_f64_kernel<<<grid, 256, 0, as_stream(stream)>>>(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  return check_launch("gemm_f64_kernel");
}

// bit 0: force the two-sided kernel (ERA5SVD_SYEVJ_TWOSIDED=1; diagnostics and tests)
static unsigned syevj_flags() {
  const char* e = getenv("ERA5SVD_SYEVJ_TWOSIDED");
  return (e && e[0] == '1') ? 1u : 0u;
}

static size_t syevj_small_smem_doubles(int64_t n) {
  int64_t npairs = (n + (n & 1)) / 2;
  return (size_t)(2 * npairs + (n + 1) / 2);
}
static size_t syevj_smem_bytes(int64_t n) {
  return (syevj_small_smem_doubles(n) + (size_t)(2 * n * (n | 1))) * sizeof(double);
}

size_t era5svd_syevj_workspace_bytes(int64_t n) {
  if (n <= 0) return 0;
  return (size_t)(2 * n * (n | 1)) * sizeof(double) + 16;     // + the status word of the positive-definite fast path
}

int era5svd_syevj_f64(double* A, int64_t n, int64_t lda, double* W, double* V, int64_t ldv,
                      int max_sweeps, double tol, void* workspace, size_t workspace_bytes, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(A && W && V, "syevj: null pointer");
  ERA5SVD_REQUIRE(n > 0 && n <= 16384 && lda >= n && ldv >= n, "syevj: bad shape n=%lld", (long long)n);
  if (max_sweeps <= 0) max_sweeps = 30;
  // Sketch-sized factors (Gram matrices: positive definite unless the sketch is wider than the rank): Cholesky +
  // one-sided Jacobi.  It reports a dropped pivot through a status word; the two-sided kernel below then runs
  // (it returns at once when the fast path has succeeded).
  int* status = nullptr;
  if (n <= J1_MAXN && n >= 2 && workspace && workspace_bytes >= era5svd_syevj_workspace_bytes(n) &&
      !(syevj_flags() & 1)) {
    status = reinterpret_cast<int*>(static_cast<char*>(workspace) + (size_t)(2 * n * (n | 1)) * sizeof(double));
    const size_t elems = (size_t)n * (size_t)((n + JAC_GROUP - 1) / JAC_GROUP * JAC_GROUP);   // rows padded to whole lane groups
    int rc;
    if (tol >= 1e-5) {        // a basis rotation, not a result (the randomized driver's Rayleigh-Ritz step): float32 sweeps
      ERA5SVD_CUDA(ensure_dynamic_smem((const void*)syev_chol_jacobi_kernel<float>, elems * sizeof(float)));
      syev_chol_jacobi_kernel<float><<<1, JAC_THREADS, elems * sizeof(float), as_stream(stream)>>>(A, (int)n, lda, W, V, ldv, max_sweeps, tol, 1e-13, status);
      rc = check_launch("syev_chol_jacobi_kernel<float>");
    } else {
      ERA5SVD_CUDA(ensure_dynamic_smem((const void*)syev_chol_jacobi_kernel<double>, elems * sizeof(double)));
      syev_chol_jacobi_kernel<double><<<1, JAC_THREADS, elems * sizeof(double), as_stream(stream)>>>(A, (int)n, lda, W, V, ldv, max_sweeps, tol, 1e-13, status);
      rc = check_launch("syev_chol_jacobi_kernel<double>");
    }
    if (rc) return rc;
  }
  const size_t full = syevj_smem_bytes(n);
  const size_t limit = 227 * 1024;
  int use_smem = full <= limit;
  size_t smem = use_smem ? full : syevj_small_smem_doubles(n) * sizeof(double);
  if (!use_smem) {
    size_t need = era5svd_syevj_workspace_bytes(n);
    if (!workspace || workspace_bytes < need) {
      set_error("syevj: workspace too small (%zu < %zu)", workspace_bytes, need);
      return ERA5SVD_ERR_WORKSPACE;
    }
  }
  ERA5SVD_CUDA(ensure_dynamic_smem((const void*)syevj_kernel<true>, limit));
  ERA5SVD_CUDA(ensure_dynamic_smem((const void*)syevj_kernel<false>, limit));
  if (use_smem)
    syevj_kernel<true><<<1, JAC_THREADS, smem, as_stream(stream)>>>(A, (int)n, lda, W, V, ldv, max_sweeps, tol, (double*)workspace, status);
  else
    syevj_kernel<false><<<1, JAC_THREADS, smem, as_stream(stream)>>>(A, (int)n, lda, W, V, ldv, max_sweeps, tol, (double*)workspace, status);
  return check_launch("syevj_kernel");
}

int era5svd_chol_inv_f64(const double* G, int64_t l, int64_t ldg, double* R, int64_t ldr,
                         double* Rinv, int64_t ldri, double rel_tol, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(G && R && Rinv, "chol_inv: null pointer");
  ERA5SVD_REQUIRE(l > 0 && l <= 8192 && ldg >= l && ldr >= l && ldri >= l, "chol_inv: bad shape l=%lld", (long long)l);
  const size_t small = (size_t)(2 * l) * sizeof(double);
  const size_t full = small + (size_t)(2 * l * (l | 1)) * sizeof(double);
  const int use_smem = full <= 220 * 1024;
  const size_t smem = use_smem ? full : small;
  if (use_smem) {
    ERA5SVD_CUDA(ensure_dynamic_smem((const void*)chol_inv_kernel<true>, 220 * 1024));
    chol_inv_kernel<true><<<1, 1024, smem, as_stream(stream)>>>(G, (int)l, ldg, R, ldr, Rinv, ldri, rel_tol);
  } else {
    chol_inv_kernel<false><<<1, 1024, smem, as_stream(stream)>>>(G, (int)l, ldg, R, ldr, Rinv, ldri, rel_tol);
  }
  return check_launch("chol_inv_kernel");
}

int era5svd_col_normalize_f64(double* P, int64_t n, int64_t l, int64_t ldp, double* norms_out,
                              void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(P, "col_normalize: null pointer");
  ERA5SVD_REQUIRE(n > 0 && l > 0 && ldp >= l, "col_normalize: bad shape");
  col_normalize_kernel<<<(unsigned)l, 128, 0, as_stream(stream)>>>(P, n, l, ldp, norms_out);
  return check_launch("col_normalize_kernel");
}

int era5svd_convert(const void* src, int dtype_src, int64_t lds, void* dst, int dtype_dst,
                    int64_t ldd, int64_t rows, int64_t cols, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(src && dst, "convert: null pointer");
  ERA5SVD_REQUIRE(valid_dtype(dtype_src) && valid_dtype(dtype_dst), "convert: bad dtype");
  ERA5SVD_REQUIRE(rows > 0 && cols > 0 && lds >= cols && ldd >= cols, "convert: bad shape");
  cudaStream_t st = as_stream(stream);
  unsigned blocks = (unsigned)ceil_div(rows * cols, 256);
  if (dtype_src == ERA5SVD_F64 && dtype_dst == ERA5SVD_F32)
    convert_kernel<double, float><<<blocks, 256, 0, st>>>((const double*)src, lds, (float*)dst, ldd, rows, cols);
  else if (dtype_src == ERA5SVD_F32 && dtype_dst == ERA5SVD_F64)
    convert_kernel<float, double><<<blocks, 256, 0, st>>>((const float*)src, lds, (double*)dst, ldd, rows, cols);
  else if (dtype_src == ERA5SVD_F32)
    convert_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, rows, cols);
  else
    convert_kernel<double, double><<<blocks, 256, 0, st>>>((const double*)src, lds, (double*)dst, ldd, rows, cols);
  return check_launch("convert_kernel");
}

int era5svd_scale_rows_f64(double* V, int64_t k, int64_t n, int64_t ldv, const double* scale,
                           void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(V && scale, "scale_rows: null pointer");
  ERA5SVD_REQUIRE(k > 0 && n > 0 && ldv >= n, "scale_rows: bad shape");
  scale_rows_kernel<<<(unsigned)ceil_div(k * n, 256), 256, 0, as_stream(stream)>>>(V, k, n, ldv, scale);
  return check_launch("scale_rows_kernel");
}

}  // extern "C"
