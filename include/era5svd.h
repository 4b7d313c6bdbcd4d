/*
 * era5svd.h - C ABI of the B200-native DMD-ERA5 SVD stage (libera5svd.so, sm_100a).
 *
 * The reference (ClimeTrend/DMD-ERA5) is pure Python and has NO FFI of its own
 * (SURVEY.md section 8b); its seam for this path is
 *     svd_on_era5(da, parsed_config) -> (U, s, V)      src/dmd_era5/era5_svd/era5_svd.py:230-263
 * plus the matrix-build calls of era5_svd.main          era5_svd.py:384-414
 * which bottom out in NumPy / SciPy / scikit-learn BLAS+LAPACK calls.  Each entry point
 * below replaces one of those library call sites (cited per function); INTEGRATION.md
 * shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative era5svd_status otherwise;
 *     era5svd_last_error() returns a thread-local message for the last failure.
 *   - the caller owns ALL memory (device pointers, e.g. torch tensors' data_ptr());
 *     the library allocates nothing persistent; ops that need scratch take a workspace
 *     pointer + size and have a *_workspace_bytes() query.
 *   - all ops are asynchronous on the cudaStream_t passed as `void* stream`
 *     (0 = legacy default stream); no hidden synchronisation, re-entrant across streams.
 *   - all matrices are ROW-MAJOR with an explicit leading dimension in ELEMENTS.
 *   - dtype: ERA5SVD_F32 / ERA5SVD_F64 for the tall (space-sized) operands; every small
 *     (time- or sketch-sized) factor is float64.
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     ERA5SVD_ERR_CUDA.
 */
#ifndef ERA5SVD_H
#define ERA5SVD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERA5SVD_VERSION 100 /* 0.1.0 */

typedef enum {
  ERA5SVD_OK = 0,
  ERA5SVD_ERR_ARG = -1,       /* invalid argument (shape, dtype, alignment, null) */
  ERA5SVD_ERR_CUDA = -2,      /* CUDA runtime / launch error, or no device       */
  ERA5SVD_ERR_WORKSPACE = -3, /* workspace too small                             */
  ERA5SVD_ERR_UNSUPPORTED = -4
} era5svd_status;

typedef enum { ERA5SVD_F32 = 0, ERA5SVD_F64 = 1 } era5svd_dtype;

/* arithmetic used by the tall GEMM passes */
typedef enum {
  ERA5SVD_PREC_NATIVE = 0, /* FMA in the storage dtype (FP64 DFMA for f64, FP32 FFMA for f32) */
  ERA5SVD_PREC_TF32X3 = 1  /* f32 storage, tcgen05 kind::tf32 with 3-term hi/lo split, fp32 accumulate in TMEM */
} era5svd_precision;

/* build flags */
#define ERA5SVD_BUILD_MEAN_CENTER 1u
#define ERA5SVD_BUILD_SCALE 2u
#define ERA5SVD_BUILD_CHECK_FINITE 4u
#define ERA5SVD_BUILD_NO_TMA 8u /* diagnostics: float32 sources take the register-staged kernel instead of the TMA one */

int era5svd_version(void);
const char* era5svd_last_error(void);
/* number of CUDA kernels launched by this library in the calling process (bench's gpu_launches) */
unsigned long long era5svd_launch_count(void);

/* Measured tensor peaks for the roofline denominators (bench.py; not part of the data path; synchronous).
 * era5svd_probe_tf32_tflops : dense kind::tf32 TFLOP/s of the whole chip with every SM issuing tcgen05.mma M = 128,
 *                             K = 8, width N back to back for ~`seconds` (form 0: A from shared memory, 1: A from TMEM).
 * era5svd_probe_dmma_tflops : FP64 tensor TFLOP/s (mma.sync.m8n8k4.f64, the path of the float64 tall kernels). */
int era5svd_probe_tf32_tflops(int form, int N, double seconds, double* tflops);
int era5svd_probe_dmma_tflops(double seconds, double* tflops);

/* ------------------------------------------------------------------------------------------
 * (a) matrix build.  Replaces standardize_data (slice_tools.py:171-177: xarray mean / subtract /
 * std(ddof=0) / divide), flatten_era5_variables' transpose+concat copy (slice_tools.py:323-336)
 * and the sklearn check_array finiteness pass (sklearn/utils/extmath.py:546).
 *
 * src : one (variable, level) block in the NATIVE ERA5 layout, time-major:
 *       element (t, p) at src[t * src_ld + p], t < T, p < P  (P = lat*lon points, or a shard of them)
 * X   : destination rows [P] x columns [T], row r = p, X[p * ldx + t]   (space x time, time fastest)
 *       value = (src - mean_p) / std_p * weight_p   following the flags; dtype_x may differ from
 *       dtype_src (explicit cast, opt-in extension; default equal = reference behaviour)
 * mean_out / std_out : length-P arrays in dtype_x, nullable; std is of the CENTRED data, ddof = 0,
 *       NaN-skipping like xarray; no epsilon guard (std == 0 -> inf/nan exactly like the reference)
 * weights : nullable length-P array (dtype_x), e.g. sqrt(cos(lat)) - opt-in extension
 * nonfinite_flag : nullable device int, set to 1 if any non-finite input was seen (CHECK_FINITE)
 */
int era5svd_build_rows(const void* src, int dtype_src, int64_t T, int64_t src_ld, int64_t P,
                       void* X, int dtype_x, int64_t ldx, void* mean_out, void* std_out,
                       const void* weights, unsigned flags, int* nonfinite_flag, void* stream);

/* The finiteness pass alone, for a matrix that did not come through the build (svd_on_era5's array input): sets *flag
 * (device int, zeroed by the caller) to 1 if any element of X[rows x cols] (row pitch ld) is NaN or +-Inf.  Replaces
 * sklearn's check_array inside randomized_svd (sklearn/utils/extmath.py:546), which raises "Input contains NaN". */
int era5svd_check_finite(const void* X, int dtype, int64_t rows, int64_t cols, int64_t ld, int* flag, void* stream);

/* Same build, float32 matrix, writing the tf32 hi / lo images the tensor-core passes consume
 * (Xhi + Xlo == X exactly) in the same pass; X itself is optional (nullable). */
int era5svd_build_rows_split(const void* src, int dtype_src, int64_t T, int64_t src_ld, int64_t P,
                             float* X, float* Xhi, float* Xlo, int64_t ldx, float* mean_out,
                             float* std_out, const float* weights, unsigned flags,
                             int* nonfinite_flag, void* stream);

/* ------------------------------------------------------------------------------------------
 * (b) tall GEMM passes of the randomized range finder.
 *
 * era5svd_sketch:  Y[m x l] = X[m x n] * Om[n x l]           replaces `A @ Q`   extmath.py:378, 383
 *                  (also U = Q * Uhat, extmath.py:619, with X := Y, Om := small factor)
 * era5svd_project: Z[n x l] (float64) (+)= X[m x n]^T * Y[m x l]   replaces `A.T @ Q` extmath.py:379
 *                  and `Q.T @ M` extmath.py:606 (transposed), and the Gram matrices Y^T Y / X^T X
 *                  (pass Y := X).  Partial sums over row ranges go to the workspace in the accumulate
 *                  type and are reduced in float64 in a fixed order (deterministic).
 * X, Om, Y share `dtype`.  A delay-embedded block j is addressed by passing X + j with the same ldx
 * (slice_tools.py:207-211 is never materialised).
 */
int era5svd_sketch(const void* X, int dtype, int64_t m, int64_t n, int64_t ldx, const void* Om,
                   int64_t l, int64_t ldo, void* Y, int64_t ldy, int precision, void* stream);

size_t era5svd_project_workspace_bytes(int dtype, int64_t m, int64_t n, int64_t l, int precision);
int era5svd_project(const void* X, int dtype, int64_t m, int64_t n, int64_t ldx, const void* Y,
                    int64_t l, int64_t ldy, double* Z, int64_t ldz, int accumulate, int precision,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * (b') the same passes on the tensor cores: tcgen05.mma kind::tf32 fed by TMA, fp32 accumulators in
 * TMEM, 3-term split x = hi + lo (hi = tf32(x), lo = x - hi, both stored as float32) so that
 * a*b ~ a_hi*b_hi + a_lo*b_hi + a_hi*b_lo keeps fp32-level accuracy.  float32 storage only.
 *
 * era5svd_split_tf32      : hi / lo images of a tall float32 matrix (done once per matrix).
 * era5svd_sketch_tf32x3   : Y = X * Om with X given as (Xhi, Xlo) - or with Xlo == NULL and Xhi the plain
 *                           float32 matrix, split on chip (X then crosses HBM once; l <= 128);
 *                           Om is the float64 small factor
 *                           (n x l, split on the fly into the workspace).  Writes any of Y (plain
 *                           float32) and the pair (Yhi, Ylo) for a following project; all share ldy,
 *                           which must be >= round_up(l, 16) and a multiple of 4 (pad columns get 0).
 * era5svd_project_tf32x3  : Z (float64, n x l) (+)= X^T Y from (Xhi, Xlo), (Yhi, Ylo); l <= 128.  Xlo == NULL:
 *                           Xhi is the plain float32 matrix, split on chip; then Ylo == NULL is allowed as well: Yhi is
 *                           the plain float32 Y (the sketch's Y output), split on chip too, so that the tall factor
 *                           crosses HBM once per pass instead of as two images.
 * Pointers may carry a column offset (delay-embedded window X + j): only 4-byte alignment of the
 * tall operands is required, row pitches must be multiples of 4 floats.
 */
int era5svd_split_tf32(const float* X, int64_t rows, int64_t cols, int64_t ldx, float* hi, float* lo,
                       int64_t ld_out, void* stream);
size_t era5svd_sketch_tf32x3_workspace_bytes(int64_t n, int64_t l);
int era5svd_sketch_tf32x3(const float* Xhi, const float* Xlo, int64_t m, int64_t n, int64_t ldx,
                          const double* Om, int64_t l, int64_t ldo, float* Y, float* Yhi, float* Ylo,
                          int64_t ldy, void* workspace, size_t workspace_bytes, void* stream);
/* The small factor of the range finder is OURS to choose (any well-conditioned basis of the same span serves the
 * power iteration, extmath.py:371-383 normalises only for stability), so the driver keeps it in tf32-representable
 * values: its lo image is then zero and the sketch needs two tensor-core products per k-step instead of three, and
 * half the Om^T tile traffic.
 * era5svd_round_tf32_f64 : A <- tf32(A) in place (float64 storage, round to nearest, 10 explicit mantissa bits).
 * era5svd_sketch_tf32x2  : Y = X * tf32(Om), X the plain float32 matrix (split hi/lo on chip), l <= 128; same
 *                          outputs, pitches and workspace as era5svd_sketch_tf32x3.  Exact (to the fp32 accumulate)
 *                          for a factor that went through era5svd_round_tf32_f64. */
int era5svd_round_tf32_f64(double* A, int64_t rows, int64_t cols, int64_t lda, void* stream);
int era5svd_sketch_tf32x2(const float* X, int64_t m, int64_t n, int64_t ldx, const double* Om, int64_t l,
                          int64_t ldo, float* Y, float* Yhi, float* Ylo, int64_t ldy, void* workspace,
                          size_t workspace_bytes, void* stream);
/* Single-product TF32 passes for the EARLY power iterations (extmath.py:377-379): subspace iteration is self-correcting,
 * so only the last iteration and the final range / projection passes (extmath.py:383, :606) need fp32-level products.
 * The raw float32 tiles go TMA -> shared memory -> tcgen05.mma directly (the tensor core truncates fp32 operands to
 * tf32), one product per k-step, no hi / lo images, no transform warps: the pass is HBM bound.
 * era5svd_sketch_tf32x1  : Y (plain float32, pitch ldy as above) = X * tf32(Om); workspace as era5svd_sketch_tf32x3.
 * era5svd_project_tf32x1 : Z (float64, n x l) (+)= X^T Y with both operands truncated to tf32; workspace as
 *                          era5svd_project_tf32x3.  l <= 128. */
int era5svd_sketch_tf32x1(const float* X, int64_t m, int64_t n, int64_t ldx, const double* Om, int64_t l,
                          int64_t ldo, float* Y, int64_t ldy, void* workspace, size_t workspace_bytes, void* stream);
int era5svd_project_tf32x1(const float* X, int64_t m, int64_t n, int64_t ldx, const float* Y, int64_t l,
                           int64_t ldy, double* Z, int64_t ldz, int accumulate, void* workspace,
                           size_t workspace_bytes, void* stream);
/* era5svd_project_tf32x2 : Z (+)= X^T tf32(Y): X split hi / lo on chip (exact), the plain float32 Y taken truncated to tf32 -
 *                          two products per k-step instead of three, no lo image of Y.  For the projection of the LAST power
 *                          iteration: Z = X^T (Y + dY) = X^T Y + X^T dY, and an error in Y is filtered by X^T like the
 *                          iteration itself filters the sketch (it lands in the dominant subspace), whereas the final
 *                          projection B = Q^T X (extmath.py:606) keeps all three products.  Workspace as project_tf32x3. */
int era5svd_project_tf32x2(const float* X, int64_t m, int64_t n, int64_t ldx, const float* Y, int64_t l,
                           int64_t ldy, double* Z, int64_t ldz, int accumulate, void* workspace,
                           size_t workspace_bytes, void* stream);
size_t era5svd_project_tf32x3_workspace_bytes(int64_t m, int64_t n, int64_t l);
int era5svd_project_tf32x3(const float* Xhi, const float* Xlo, int64_t m, int64_t n, int64_t ldx,
                           const float* Yhi, const float* Ylo, int64_t l, int64_t ldy, double* Z,
                           int64_t ldz, int accumulate, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * (c)/(d) small float64 factor kernels (replicated on every GPU; per-CTA solvers).
 */
/* C[M x N] = alpha * op(A) * op(B) + beta * C, row-major; transX != 0 means the operand is stored
 * transposed (op(A) = A^T with A stored [K x M]). */
int era5svd_gemm_f64(int transA, int transB, int64_t M, int64_t N, int64_t K, double alpha,
                     const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                     double* C, int64_t ldc, void* stream);

/* Symmetric eigen-decomposition by parallel cyclic Jacobi (one CTA).  A[n x n] is destroyed;
 * W[n] = eigenvalues in DESCENDING order, V[n x n] = eigenvectors in columns (same order).
 * Replaces the small dense solves of scipy.linalg.svd(B, gesdd) (extmath.py:615) and the
 * eigensolve behind the Gram-route standard SVD (np.linalg.svd, era5_svd.py:251).
 * max_sweeps <= 0 selects the default (30); iteration stops early at convergence, i.e. when a whole
 * sweep finds every |a_pq| <= tol * sqrt(|a_pp a_qq|); tol <= 0 selects eps * sqrt(n).
 * n <= 128 (the sketch-sized Gram matrices): Cholesky A = R^T R followed by ONE-SIDED Jacobi on the rows of R
 * (eigenvalue = squared row norm, eigenvector = normalised row; a third of the shared-memory traffic of the
 * two-sided iteration).  A dropped pivot (A numerically singular or indefinite) falls back to the two-sided
 * iteration inside the same call, so the contract above holds for any symmetric A.  The workspace
 * (era5svd_syevj_workspace_bytes, required for n <= 128) also carries the 4-byte status word of that hand-over.
 * tol >= 1e-5 on that path means the caller wants a basis ROTATION, not a result (the Rayleigh-Ritz step of the randomized
 * driver: any well-conditioned W serves, as long as the same W is used afterwards): the Cholesky factor is formed in
 * float64 and then rotated as FLOAT32 rows; W / V come back with float32 accuracy (~1e-6). */
size_t era5svd_syevj_workspace_bytes(int64_t n);
int era5svd_syevj_f64(double* A, int64_t n, int64_t lda, double* W, double* V, int64_t ldv,
                      int max_sweeps, double tol, void* workspace, size_t workspace_bytes, void* stream);

/* Symmetric eigensolver for time-sized matrices, k largest eigenpairs (the eigensolve behind the Gram-route
 * standard SVD: np.linalg.svd(X, full_matrices=False) followed by the [:k] truncation, era5_svd.py:249-254).
 *   tridiag_reduce        : A (n x n symmetric, both triangles, row-major) = Q T Q^T.  On return d[n], e[n] (e[n-1] = 0)
 *                           hold T, tau[n] the reflector scalars and row j of A (columns j+1..) the Householder vector
 *                           of column j (v[0] = 1 stored); the rest of A is destroyed.  Multi-CTA, memory bound.
 *   tridiag_eig_topk      : the k largest eigenvalues of T (W, descending; bisection on Sturm counts) and vectors
 *                           Z (n x k, columns; inverse iteration).  Vectors of close eigenvalues are independent but
 *                           not orthogonal: the caller orthonormalises and applies one Rayleigh-Ritz step (Y = T Z
 *                           via tridiag_apply).
 *   tridiag_backtransform : V (n x k) = Q Z. */
size_t era5svd_tridiag_reduce_workspace_bytes(int64_t n);
int era5svd_tridiag_reduce_f64(double* A, int64_t n, int64_t lda, double* d, double* e, double* tau,
                               void* workspace, size_t workspace_bytes, void* stream);
size_t era5svd_tridiag_eig_topk_workspace_bytes(int64_t n, int64_t k);
int era5svd_tridiag_eig_topk_f64(const double* d, const double* e, int64_t n, int64_t k, double* W, double* Z,
                                 int64_t ldz, void* workspace, size_t workspace_bytes, void* stream);
int era5svd_tridiag_apply_f64(const double* d, const double* e, int64_t n, int64_t k, const double* Z, int64_t ldz,
                              double* Y, int64_t ldy, void* stream);
int era5svd_tridiag_backtransform_f64(const double* A, int64_t n, int64_t lda, const double* tau, int64_t k,
                                      const double* Z, int64_t ldz, double* V, int64_t ldv, void* stream);

/* Cholesky G = R^T R (R upper triangular) of a symmetric positive (semi-)definite l x l matrix and
 * the explicit inverse Rinv = R^{-1} (upper triangular), one CTA.  Replaces scipy qr / lu
 * normalisers (extmath.py:371-383) in CholeskyQR form.  A pivot that falls below
 * rel_tol * G[j][j] (column j numerically dependent on the previous ones) is replaced by a huge
 * value so that row/column j of Rinv vanish: rank-deficient directions give zero columns, not NaN. */
int era5svd_chol_inv_f64(const double* G, int64_t l, int64_t ldg, double* R, int64_t ldr,
                         double* Rinv, int64_t ldri, double rel_tol, void* stream);

/* P[:, j] /= ||P[:, j]||_2 for an n x l float64 matrix; norms_out (nullable) receives the norms. */
int era5svd_col_normalize_f64(double* P, int64_t n, int64_t l, int64_t ldp, double* norms_out,
                              void* stream);

/* s[i] = sqrt(max(w[i], 0)), inv_s[i] = 1 / s[i] (0 where s[i] == 0; nullable): singular values from
 * the eigenvalues of B B^T. */
int era5svd_sigma_from_eig_f64(const double* w, int64_t l, double* s, double* inv_s, void* stream);

/* dst[r x c] (dtype_dst) = src[r x c] (dtype_src), row-major with leading dimensions. */
int era5svd_convert(const void* src, int dtype_src, int64_t lds, void* dst, int dtype_dst,
                    int64_t ldd, int64_t rows, int64_t cols, void* stream);

/* ------------------------------------------------------------------------------------------
 * (d') Optimized DMD / BOP-DMD on the SVD-projected coefficients, batched over bagging trials.  NOT in the reference
 * (it only cites the method, README.md:85, :139); named by BASELINE.json (north_star (d), configs[4]).  Variable
 * projection + Levenberg-Marquardt (Askham & Kutz 2018), restated in oracle/bopdmd_np.py.
 *
 * H (n_time x N, row-major, ldh) : projected coefficients, row i = snapshot at time t[i]  (= (diag(s) V)^T of the SVD)
 * idx (K x p int32)              : time-ordered snapshot subset of every trial (K = 1 with idx = 0..n_time-1: full fit)
 * alpha, alpha_try (K x r complex128, interleaved re/im): accepted eigenvalues / candidate evaluated by this call;
 *                                  before the first call alpha_try holds the initial guess, rho = +inf, lam = lam0, done = 0
 * rho, lam (K), JhJ (K x r x r complex), rhs (K x r complex), Bout (K x r x N complex: B of the accepted point,
 * mode j = Bout[j, :] in the coordinates of H's columns), done (K int32: 1 once converged / failed).
 * One call = one LM iteration of every trial with done == 0:  Psi = [Re Phi | Im Phi] -> batched FP64 Gram GEMMs ->
 * one CTA per trial (complex Cholesky, B, rho, accept / reject, J^H J, rhs, next candidate).  first != 0 also
 * computes the per-trial ||H[idx]||_F^2. */
size_t era5svd_bop_workspace_bytes(int64_t K, int64_t p, int64_t r, int64_t N);
int era5svd_bop_iterate_f64(const double* H, int64_t n_time, int64_t N, int64_t ldh, const double* t, const int* idx,
                            int64_t K, int64_t p, int64_t r, double* alpha, double* alpha_try, double* rho,
                            double* lam, double* JhJ, double* rhs, double* Bout, int* done, double nu, double tol,
                            int first, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * svd_flip (extmath.py:964-972, u-based): per column of U the FIRST row holding max |U[:, j]|.
 * col_absmax writes, per column j < k: absmax[j], row[j] = row_offset + local row (lowest on ties),
 * sign[j] = sign of that entry (+1 / -1 / 0).  combine reduces R stacked candidate sets
 * ([R x k] each, e.g. all-gathered over ranks) with the same tie rule.  scale_cols applies
 * U[:, j] *= sign[j]; scale_rows applies V[j, :] *= sign[j] (float64).
 */
size_t era5svd_col_absmax_workspace_bytes(int64_t m, int64_t k);
int era5svd_col_absmax(const void* U, int dtype, int64_t m, int64_t k, int64_t ldu,
                       int64_t row_offset, double* absmax, int64_t* row, double* sign,
                       void* workspace, size_t workspace_bytes, void* stream);
int era5svd_maxloc_combine(const double* absmax, const int64_t* row, const double* sign, int64_t R,
                           int64_t k, double* sign_out, void* stream);
int era5svd_scale_cols(void* U, int dtype, int64_t m, int64_t k, int64_t ldu, const double* scale,
                       void* stream);
int era5svd_scale_rows_f64(double* V, int64_t k, int64_t n, int64_t ldv, const double* scale,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * era5svd_comm_* : the collectives of the row-sharded path as kernels over peer memory (NVLink / NVSwitch P2P through
 * CUDA IPC), SURVEY.md 8b / 8e.  The reference has no counterpart (single process, era5_svd.py:246-259); what is
 * reduced is what its GEMMs sum over ALL rows: Z = X^T Y (extmath.py:379, 606), the l x l Gram matrix, and the
 * first-maximum rule of svd_flip (extmath.py:964-972).  One process per GPU of ONE node:
 *   create   : allocates this rank's window (two data slots of slot_bytes + flags) and exports its IPC handle
 *              (era5svd_comm_handle_bytes() bytes) - the caller exchanges the handles (torch.distributed all_gather);
 *   connect  : maps every peer's window; `handles` = nranks consecutive handles in rank order;
 *   allreduce_f64 / allgather_f64 : one kernel on `stream` each; sums are formed in rank order 0..R-1 on every rank,
 *              so the replicated small factors stay bit-identical;  count <= era5svd_comm_capacity() doubles;
 *   fuse_next_project(n, l) : the NEXT era5svd_project* call on this host thread (n x l result) finishes with ONE kernel
 *              that sums the partial tiles of the projection AND the ranks (reduce_partials_allreduce_kernel): Z holds
 *              the all-reduced result.  ERA5SVD_ERR_UNSUPPORTED when n * l does not fit a slot.
 * Every rank must issue the same sequence of collectives.  Not thread safe per communicator. */
size_t era5svd_comm_handle_bytes(void);
int era5svd_comm_create(int nranks, int rank, int64_t slot_bytes, void** comm_out, void* handle_out);
int era5svd_comm_connect(void* comm, const void* handles);
int era5svd_comm_destroy(void* comm);
int64_t era5svd_comm_capacity(void* comm);
unsigned long long era5svd_comm_fused_count(void* comm);
int era5svd_comm_allreduce_f64(void* comm, double* buf, int64_t count, void* stream);
int era5svd_comm_allgather_f64(void* comm, const double* src, int64_t count, double* dst, void* stream);
int era5svd_comm_fuse_next_project(void* comm, int64_t n, int64_t l);

#ifdef __cplusplus
}
#endif
#endif /* ERA5SVD_H */
