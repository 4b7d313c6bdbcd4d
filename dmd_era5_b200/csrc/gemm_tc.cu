// (b) Tall GEMM passes on the 5th-generation tensor cores: tcgen05.mma kind::tf32 fed by TMA, fp32
// accumulators in TMEM, 3-term TF32 split for fp32-level accuracy ("3xTF32"):
//
//     x = x_hi + x_lo,  x_hi = tf32(x),  x_lo = x - x_hi         (both exactly representable operands)
//     a * b  ~=  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo                (drops a_lo*b_lo ~ 2^-22 |a b|)
//
// The tall operands are stored pre-split in HBM (X_hi / X_lo written once by era5svd_split_tf32,
// Y_hi / Y_lo written by the sketch epilogue), so both kernels are pure TMA -> UMMA pipelines with
// no register-path transform.
//
//   sketch :  Y[m x l]   = X[m x n] * Om[n x l]     A = X tile   (K-major, K = time)
//                                                   B = Om^T     (K-major)        D: 128 rows x Npad
//   project:  Z^T[l x n] = Y^T[l x m] * X[m x n]    A = Y tile   (MN-major, K = space rows)
//                                                   B = X tile   (MN-major)       D: 128 (l) x time chunk
//
// Both kernels read X tiles as [rows][32 time values] boxes, so X is never transposed in HBM: sketch
// consumes them K-major (16 B swizzle granules), project MN-major (32-bit MN-major operands use the
// 32 B-granule 128 B swizzle, TMA mode SWIZZLE_128B_ATOM_32B).
// Replaces `A @ Q` / `A.T @ Q` / `Q.T @ M` (sklearn/utils/extmath.py:378-383, :606) for float32 data.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace era5svd {

// from gemm_tc2.cu: on-chip split variants (X given as one plain float32 matrix)
int sketch_tf32x3_raw(const float* X, int64_t m, int64_t n, int64_t ldx, const double* Om, int64_t l,
                      int64_t ldo, float* Y, float* Yhi, float* Ylo, int64_t ldy, void* workspace,
                      cudaStream_t st, int om_tf32);

int project_tf32x3_raw(const float* X, int64_t m, int64_t n, int64_t ldx, const float* Yhi, const float* Ylo,
                       int64_t l, int64_t ldy, double* Z, int64_t ldz, int accumulate, void* workspace,
                       size_t workspace_bytes, cudaStream_t st, int y_tf32);
size_t project_tf32x3_raw_workspace_bytes(int64_t m, int64_t n, int64_t l);

// from gemm_simt.cu
int launch_reduce_partials_f32(const float* part, int64_t splits, int64_t n, int64_t l, int64_t lp, double* Z,
                               int64_t ldz, int accumulate, cudaStream_t st);

namespace tc {

constexpr int BM = 128;          // sketch: rows per tile (UMMA M)
constexpr int BK = 32;           // tf32 elements per 128-byte swizzle row
constexpr int UMMA_K = 8;        // tf32: 32 bytes of K per instruction
constexpr int SWIZZLE_ATOM = 1024;   // 8 rows x 128 B

// ---------------------------------------------------------------------------------------------
// host: TMA descriptors
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D float32 tensor [outer rows][inner elements], row pitch ld elements, box = box_inner x box_outer,
// 128 B swizzle, out-of-bounds elements read as zero.  `base` may be any 4-byte aligned address: the
// map is anchored at the enclosing 16-byte boundary and *col_shift returns the element offset to add
// to every inner coordinate (this is how delay-embedded column windows X[:, j:] are addressed).
int make_tmap(CUtensorMap* map, const float* base, int64_t inner, int64_t outer, int64_t ld,
              uint32_t box_inner, uint32_t box_outer, int* col_shift, CUtensorMapSwizzle swizzle);
int make_tmap(CUtensorMap* map, const float* base, int64_t inner, int64_t outer, int64_t ld,
              uint32_t box_inner, uint32_t box_outer, int* col_shift) {
  return make_tmap(map, base, inner, outer, ld, box_inner, box_outer, col_shift, CU_TENSOR_MAP_SWIZZLE_128B);
}
int make_tmap(CUtensorMap* map, const float* base, int64_t inner, int64_t outer, int64_t ld,
              uint32_t box_inner, uint32_t box_outer, int* col_shift, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the driver");
    return ERA5SVD_ERR_CUDA;
  }
  uintptr_t addr = reinterpret_cast<uintptr_t>(base);
  int shift = (int)((addr & 15u) / 4u);
  addr &= ~(uintptr_t)15u;
  if ((ld * 4) % 16 != 0) {
    set_error("tf32x3: row pitch (%lld floats) must be a multiple of 4", (long long)ld);
    return ERA5SVD_ERR_ARG;
  }
  // The encoding is a pure function of (address, shape, pitch, box, swizzle): a small per-thread cache saves the driver
  // call for the descriptors that recur in every pass of a step (X, the workspace images of Om^T, Y).
  struct Key { uintptr_t addr; int64_t inner, outer, ld; uint32_t bi, bo; int sw; };
  struct Slot { Key k; CUtensorMap m; bool valid; };
  static thread_local Slot cache[64];
  const Key key{addr, inner + shift, outer, ld, box_inner, box_outer, (int)swizzle};
  const size_t h = (size_t)((addr >> 4) * 0x9E3779B97F4A7C15ull + (uint64_t)outer * 31 + (uint64_t)inner * 7 + box_outer + (uint64_t)swizzle * 131) % 64;
  Slot& sl = cache[h];
  if (sl.valid && sl.k.addr == key.addr && sl.k.inner == key.inner && sl.k.outer == key.outer && sl.k.ld == key.ld &&
      sl.k.bi == key.bi && sl.k.bo == key.bo && sl.k.sw == key.sw) {
    *map = sl.m;
    *col_shift = shift;
    return ERA5SVD_OK;
  }
  cuuint64_t dims[2] = {(cuuint64_t)(inner + shift), (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, reinterpret_cast<void*>(addr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) { sl.k = key; sl.m = *map; sl.valid = true; }
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld box=%ux%u", (int)r, (long long)inner,
              (long long)outer, (long long)ld, box_inner, box_outer);
    return ERA5SVD_ERR_CUDA;
  }
  *col_shift = shift;
  return ERA5SVD_OK;
}

// ---------------------------------------------------------------------------------------------
// split kernels
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ X, int64_t rows, int64_t cols, int64_t ldx,
                  float* __restrict__ hi, float* __restrict__ lo, int64_t ldo) {
  const int64_t total = rows * cols;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    int64_t r = idx / cols, c = idx - r * cols;
    float x = X[r * ldx + c];
    float h = tf32_hi(x);
    hi[r * ldo + c] = h;
    lo[r * ldo + c] = x - h;
  }
}

// Om (float64, n x l) -> Om^T hi / lo (float32, [npad][ldt]), shifted right by `shift` columns:
// out[j][t] = Om[t - shift][j] for shift <= t < shift + n, zero elsewhere (rows >= l are zero too).
// The shift is how a column window X[:, c : c + n] whose start is not 16-byte aligned is addressed:
// TMA boxes must start on 16-byte boundaries under the 128 B swizzle, so the box starts at the
// aligned column below c and the zero rows of Om cancel the extra columns.
// The float64 source lets lo carry bits beyond fp32: lo = fp32(om - hi).
__global__ void __launch_bounds__(256)
split_omega_t_kernel(const double* __restrict__ Om, int64_t n, int64_t l, int64_t ldo,
                     float* __restrict__ hi, float* __restrict__ lo, int64_t npad, int64_t ldt, int shift) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npad * ldt) return;
  int64_t j = idx / ldt, t = idx - j * ldt - shift;   // j = sketch column, t = time
  float h = 0.f, w = 0.f;
  if (j < l && t >= 0 && t < n) {
    double v = Om[t * ldo + j];
    h = tf32_hi((float)v);
    w = (float)(v - (double)h);
  }
  hi[idx] = h;
  lo[idx] = w;
}

// ---------------------------------------------------------------------------------------------
// sketch kernel (persistent, warp specialised)
// ---------------------------------------------------------------------------------------------
struct SketchParams {
  int64_t m;
  int64_t num_tiles;
  int num_k;          // ceil(n / 32)
  int npad;           // UMMA N (multiple of 16, <= 256)
  int stages;
  int nprod;          // 3: hi/lo images of both operands (3xTF32); 1: ONE product on the raw float32 X tile (the tensor
                      //    core truncates it to tf32) and the tf32-exact Om^T image - the low-precision power iterations
  float* Y;           // nullable
  float* Yhi;         // nullable
  float* Ylo;         // nullable
  int64_t ldy;
};

constexpr int SK_THREADS = 192;   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

__global__ void __launch_bounds__(SK_THREADS, 1)
sketch_tc_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
                 const __grid_constant__ CUtensorMap tm_ohi, const __grid_constant__ CUtensorMap tm_olo,
                 const SketchParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;   // warp-uniform
  const uint32_t a_bytes = BM * BK * 4;                 // 16 KB
  const uint32_t b_bytes = (uint32_t)p.npad * BK * 4;   // npad x 128 B
  const bool one = p.nprod == 1;                        // kernel parameter: uniform
  const uint32_t stage_bytes = one ? a_bytes + b_bytes : 2 * a_bytes + 2 * b_bytes;
  const uint32_t off_bhi = one ? a_bytes : 2 * a_bytes; // stage layout: A_hi [A_lo] B_hi [B_lo]
  // barriers + tmem pointer live after the stages
  const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);
  const uint32_t acc_cols = p.npad <= 128 ? 128u : 256u;    // columns per accumulator buffer

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_xhi); tma_prefetch_desc(&tm_xlo);
    tma_prefetch_desc(&tm_ohi); tma_prefetch_desc(&tm_olo);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * acc_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int32_t row0 = (int32_t)(tile * BM);
        for (int kc = 0; kc < p.num_k; ++kc) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
          mbar_arrive_expect_tx(full_bar(s), stage_bytes);
          tma_load_2d(st, &tm_xhi, kc * BK, row0, full_bar(s));
          tma_load_2d(st + off_bhi, &tm_ohi, kc * BK, 0, full_bar(s));
          if (!one) {
            tma_load_2d(st + a_bytes, &tm_xlo, kc * BK, row0, full_bar(s));
            tma_load_2d(st + off_bhi + b_bytes, &tm_olo, kc * BK, 0, full_bar(s));
          }
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    const uint32_t idesc = make_idesc_tf32(BM, p.npad, 0, 0);
    int s = 0; uint32_t ph = 0;
    int buf = 0; uint32_t aph = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(buf), aph ^ 1u);      // epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)buf * acc_cols;
      for (int kc = 0; kc < p.num_k; ++kc) {
        mbar_wait(full_bar(s), ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint32_t koff = kk * UMMA_K * 4;   // 32 bytes per k-step inside the swizzle row
            const uint64_t a_hi = make_smem_desc(st + koff, 16, SWIZZLE_ATOM);
            const uint64_t b_hi = make_smem_desc(st + off_bhi + koff, 16, SWIZZLE_ATOM);
            if (one) {
              umma_tf32_ss(d_tmem, a_hi, b_hi, idesc, (kc | kk) != 0);
            } else {
              const uint64_t a_lo = make_smem_desc(st + a_bytes + koff, 16, SWIZZLE_ATOM);
              const uint64_t b_lo = make_smem_desc(st + off_bhi + b_bytes + koff, 16, SWIZZLE_ATOM);
              umma_tf32_ss(d_tmem, a_lo, b_hi, idesc, (kc | kk) != 0);   // small terms first
              umma_tf32_ss(d_tmem, a_hi, b_lo, idesc, 1);
              umma_tf32_ss(d_tmem, a_hi, b_hi, idesc, 1);
            }
          }
          umma_commit(empty_bar(s));               // smem stage free once these MMAs retire
          if (kc == p.num_k - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (++buf == 2) { buf = 0; aph ^= 1u; }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global (Y, Y_hi, Y_lo) =====
    const int q = warp % 4;                        // TMEM lane quadrant this warp may access
    int buf = 0; uint32_t aph = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tfull_bar(buf), aph);
      tcgen05_fence_after();
      const int64_t row = tile * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + (uint32_t)buf * acc_cols + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.npad; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_wait_ld();
        if (row < p.m) {
          const int64_t off = row * p.ldy + c0;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float4 y = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                   __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
            if (p.Y) *reinterpret_cast<float4*>(p.Y + off + 4 * g) = y;
            if (p.Yhi) {
              float4 h = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
              float4 w = make_float4(y.x - h.x, y.y - h.y, y.z - h.z, y.w - h.w);
              *reinterpret_cast<float4*>(p.Yhi + off + 4 * g) = h;
              *reinterpret_cast<float4*>(p.Ylo + off + 4 * g) = w;
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
      if (++buf == 2) { buf = 0; aph ^= 1u; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * acc_cols);
}

// ---------------------------------------------------------------------------------------------
// single-product sketch, TWO row tiles per Om^T tile (persistent, warp specialised)
//
// sketch_tc_kernel re-streams the Om^T tile (npad x 32 floats, L2 resident) for every 128-row tile: per HBM byte of X
// it moves 1.875 bytes through the L2 fabric and 3.75 bytes through the shared-memory port (TMA writes + SS-form
// operand reads), and measured it is those two - not HBM - that bound it once the SM clock sits at the power cap
// (c3: 7.27 ms in the step against 5.8 ms alone; profiles/r02_pitch_sensitivity.txt).  Here one stage holds a
// [256 rows x 32] X box (one TMA call, two consecutive 128-row operand tiles) and ONE Om^T tile used by both:
// 1.44 bytes of L2 traffic and 2.9 bytes of shared-memory traffic per HBM byte.  TMEM: 2 tiles x 2 buffers x 128
// columns, so the epilogue of one 256-row group overlaps the MMAs of the next.
// ---------------------------------------------------------------------------------------------
struct SketchX1Params {
  int64_t m;
  int64_t num_groups;   // ceil(m / 256)
  int num_k;            // ceil(kspan / 32)
  int npad;             // UMMA N (multiple of 16, <= 128)
  int stages;
  float* Y;
  int64_t ldy;
};

constexpr int SX1_THREADS = 192;   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue

__global__ void __launch_bounds__(SX1_THREADS, 1)
sketch_x1_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_o,
                 const SketchX1Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;   // warp-uniform
  const uint32_t a_bytes = BM * BK * 4;                 // one 128-row operand tile: 16 KB
  const uint32_t b_bytes = (uint32_t)p.npad * BK * 4;
  const uint32_t stage_bytes = 2 * a_bytes + b_bytes;   // [A0 | A1 | B]
  const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 4);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);   // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_o); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  // TMEM map: buffer b, tile t at column (2 b + t) * 128

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int64_t grp = blockIdx.x; grp < p.num_groups; grp += gridDim.x) {
        const int32_t row0 = (int32_t)(grp * 2 * BM);
        for (int kc = 0; kc < p.num_k; ++kc) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
          mbar_arrive_expect_tx(full_bar(s), stage_bytes);
          tma_load_2d(st, &tm_x, kc * BK, row0, full_bar(s));                  // 256 rows: tiles A0, A1 back to back
          tma_load_2d(st + 2 * a_bytes, &tm_o, kc * BK, 0, full_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    const uint32_t idesc = make_idesc_tf32(BM, p.npad, 0, 0);
    int s = 0; uint32_t ph = 0;
    int buf = 0; uint32_t aph = 0;
    for (int64_t grp = blockIdx.x; grp < p.num_groups; grp += gridDim.x) {
      mbar_wait(tempty_bar(buf), aph ^ 1u);      // epilogue has drained this accumulator pair
      tcgen05_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)buf * 256u, d1 = d0 + 128u;
      for (int kc = 0; kc < p.num_k; ++kc) {
        mbar_wait(full_bar(s), ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
#pragma unroll
          for (int kk = 0; kk < BK / UMMA_K; ++kk) {
            const uint32_t koff = kk * UMMA_K * 4;
            const uint64_t b = make_smem_desc(st + 2 * a_bytes + koff, 16, SWIZZLE_ATOM);
            umma_tf32_ss(d0, make_smem_desc(st + koff, 16, SWIZZLE_ATOM), b, idesc, (kc | kk) != 0);
            umma_tf32_ss(d1, make_smem_desc(st + a_bytes + koff, 16, SWIZZLE_ATOM), b, idesc, (kc | kk) != 0);
          }
          umma_commit(empty_bar(s));
          if (kc == p.num_k - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
      if (++buf == 2) { buf = 0; aph ^= 1u; }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global Y =====
    const int q = warp % 4;
    int buf = 0; uint32_t aph = 0;
    for (int64_t grp = blockIdx.x; grp < p.num_groups; grp += gridDim.x) {
      mbar_wait_hint(tfull_bar(buf), aph, 20000u);
      tcgen05_fence_after();
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        const int64_t row = grp * 2 * BM + t * BM + q * 32 + lane;
        const uint32_t taddr = tmem_base + (uint32_t)(2 * buf + t) * 128u + ((uint32_t)(q * 32) << 16);
        for (int c0 = 0; c0 < p.npad; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_wait_ld();
          if (row < p.m) {
            float4* dst = reinterpret_cast<float4*>(p.Y + row * p.ldy + c0);
#pragma unroll
            for (int g = 0; g < 4; ++g)
              dst[g] = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]),
                                   __uint_as_float(v[4 * g + 2]), __uint_as_float(v[4 * g + 3]));
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(buf));
      if (++buf == 2) { buf = 0; aph ^= 1u; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------
// project kernel: one CTA = (time chunk, row split); accumulates Z^T chunk over its rows in TMEM
// ---------------------------------------------------------------------------------------------
struct ProjectParams {
  int64_t m, n;
  int l;
  int ks;             // rows per stage (16)
  int ncc;            // 32-wide time chunks per CTA
  int nmma;           // UMMA pieces per k-step (N = ncc*32 / nmma each)
  int stages;
  int nprod;          // 3: hi/lo images of X and Y; 1: one product on the raw float32 tiles (truncated by the tensor core)
  int nya;            // 128-column tiles of Y (M tiles of the MMA): 1, or 2 for 128 < l <= 256 (single product only: the
                      // Gram blocks of the standard route - twice the columns per read of X)
  int xshift;         // the window starts xshift (0..3) columns right of the 16-byte aligned map origin
  int64_t rows_per_split;
  float* part;        // [splits][n][l]
};

constexpr int PJ_THREADS = 192;

__global__ void __launch_bounds__(PJ_THREADS, 1)
project_tc_kernel(const __grid_constant__ CUtensorMap tm_xhi, const __grid_constant__ CUtensorMap tm_xlo,
                  const __grid_constant__ CUtensorMap tm_yhi, const __grid_constant__ CUtensorMap tm_ylo,
                  const ProjectParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;   // warp-uniform
  const uint32_t box_bytes = (uint32_t)p.ks * BK * 4;       // one [ks rows x 32] box (2 KB for ks = 16)
  const uint32_t y_bytes = 4u * (uint32_t)p.nya * box_bytes;   // 128 sketch columns per Y tile
  const uint32_t x_bytes = (uint32_t)p.ncc * box_bytes;
  const bool one = p.nprod == 1;                            // kernel parameter: uniform
  const uint32_t stage_bytes = one ? y_bytes + x_bytes : 2 * y_bytes + 2 * x_bytes;
  const uint32_t off_xhi = one ? y_bytes : 2 * y_bytes;     // stage layout: Y_hi [Y_lo] X_hi [X_lo]
  const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  const uint32_t tfull_bar = bar_base + 8u * (2 * p.stages);
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 1);
  const uint32_t nc = (uint32_t)p.ncc * BK;                 // time columns of this CTA
  const uint32_t acc_cols = nc * (uint32_t)p.nya;           // accumulator of Y tile a: columns [a * nc, (a + 1) * nc)
  const uint32_t tmem_cols = acc_cols <= 32 ? 32u : acc_cols <= 64 ? 64u : acc_cols <= 128 ? 128u : acc_cols <= 256 ? 256u : 512u;

  const int64_t t0 = (int64_t)blockIdx.x * nc;              // first time column
  const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_split;
  const int64_t r_end = min(p.m, r_begin + p.rows_per_split);
  const int num_k = (int)((r_end - r_begin + p.ks - 1) / p.ks);

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_xhi); tma_prefetch_desc(&tm_xlo);
    tma_prefetch_desc(&tm_yhi); tma_prefetch_desc(&tm_ylo);
  }
  if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (elect_one()) {
      int s = 0; uint32_t ph = 0;
      for (int kc = 0; kc < num_k; ++kc) {
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
        const int32_t row0 = (int32_t)(r_begin + (int64_t)kc * p.ks);
        mbar_arrive_expect_tx(full_bar(s), stage_bytes);
        for (int c = 0; c < 4 * p.nya; ++c) {
          tma_load_2d(st + c * box_bytes, &tm_yhi, c * BK, row0, full_bar(s));
          if (!one) tma_load_2d(st + y_bytes + c * box_bytes, &tm_ylo, c * BK, row0, full_bar(s));
        }
        for (int c = 0; c < p.ncc; ++c) {
          const int32_t tc0 = (int32_t)(t0 + (int64_t)c * BK);
          tma_load_2d(st + off_xhi + c * box_bytes, &tm_xhi, tc0, row0, full_bar(s));
          if (!one) tma_load_2d(st + off_xhi + x_bytes + c * box_bytes, &tm_xlo, tc0, row0, full_bar(s));
        }
        if (++s == p.stages) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t npiece = nc / (uint32_t)p.nmma;                 // UMMA N
    const uint32_t idesc = make_idesc_tf32(128, (int)npiece, 1, 1);
    int s = 0; uint32_t ph = 0;
    for (int kc = 0; kc < num_k; ++kc) {
      mbar_wait(full_bar(s), ph);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
        for (int ks = 0; ks < p.ks / UMMA_K; ++ks) {
          const uint32_t koff = (uint32_t)ks * UMMA_K * 128;       // next 8 K rows (two 4-row swizzle atoms)
          const uint64_t a_hi = make_smem_desc(st + koff, box_bytes, 512, LAYOUT_SW128_BASE32B);
          const uint64_t a_lo = make_smem_desc(st + y_bytes + koff, box_bytes, 512, LAYOUT_SW128_BASE32B);
          for (int pc = 0; pc < p.nmma; ++pc) {
            const uint32_t xoff = (uint32_t)pc * (npiece / BK) * box_bytes + koff;
            const uint64_t b_hi = make_smem_desc(st + off_xhi + xoff, box_bytes, 512, LAYOUT_SW128_BASE32B);
            const uint32_t d_tmem = tmem_base + (uint32_t)pc * npiece;
            if (one) {
              umma_tf32_ss(d_tmem, a_hi, b_hi, idesc, (kc | ks) != 0);
              if (p.nya == 2) {               // second 128-column tile of Y against the same X boxes
                const uint64_t a2 = make_smem_desc(st + 4 * box_bytes + koff, box_bytes, 512, LAYOUT_SW128_BASE32B);
                umma_tf32_ss(d_tmem + nc, a2, b_hi, idesc, (kc | ks) != 0);
              }
            } else {
              const uint64_t b_lo = make_smem_desc(st + off_xhi + x_bytes + xoff, box_bytes, 512, LAYOUT_SW128_BASE32B);
              umma_tf32_ss(d_tmem, a_lo, b_hi, idesc, (kc | ks) != 0);
              umma_tf32_ss(d_tmem, a_hi, b_lo, idesc, 1);
              umma_tf32_ss(d_tmem, a_hi, b_hi, idesc, 1);
            }
          }
        }
        umma_commit(empty_bar(s));
        if (kc == num_k - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (++s == p.stages) { s = 0; ph ^= 1u; }
    }
  } else {
    // epilogue: lane i of the accumulator = sketch column i; columns = time offsets
    const int q = warp % 4;
    const int i = q * 32 + lane;
    float* out = p.part + (int64_t)blockIdx.y * p.n * p.l;
    if (num_k > 0) {
      mbar_wait(tfull_bar, 0);
      tcgen05_fence_after();
    }
    for (int a = 0; a < p.nya; ++a) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)a * nc;
      const int col = i + 128 * a;                       // sketch column held by this lane in Y tile a
      for (uint32_t c0 = 0; c0 < nc; c0 += 16) {
        uint32_t v[16];
        if (num_k > 0) {
          tmem_ld16(taddr + c0, v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (col < p.l) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int64_t t = t0 + c0 + j - p.xshift;     // window-relative time index
            if (t >= 0 && t < p.n) out[t * p.l + col] = __uint_as_float(v[j]);
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace tc

// ---------------------------------------------------------------------------------------------
// host entry points
// ---------------------------------------------------------------------------------------------
static int round_up(int64_t a, int64_t b) { return (int)(ceil_div(a, b) * b); }

struct PjPlan {
  int nchunks, ncc, nmma;
  int64_t splits, rows_per_split;
  size_t bytes;
};

// max_rows: rows summed in one TMEM accumulator before the partial tile goes to float64.  The tensor core's fp32
// accumulate truncates; for the 3xTF32 passes that bias must stay below the 1e-6 level (4096 rows), the single-product
// power-iteration passes carry 2^-11 per operand anyway (16384 rows: a quarter of the partial-tile traffic).
static PjPlan pj_plan(int64_t m, int64_t n, int64_t l, int64_t max_rows = 4096) {
  PjPlan pl;
  // l > 128: two Y tiles share the 512 TMEM columns, so a CTA takes at most 256 time columns
  pl.nchunks = (int)ceil_div(n, l > 128 ? 256 : 512);
  pl.ncc = (int)ceil_div(ceil_div(n, pl.nchunks), 32);          // 32-wide chunks per CTA (<= 16)
  // UMMA N = ncc * 32 / nmma must be <= 256 and a multiple of 32 (whole boxes per piece)
  pl.nmma = 1;
  while (pl.ncc * 32 / pl.nmma > 256 || pl.ncc % pl.nmma != 0) {
    ++pl.nmma;
    if (pl.nmma > pl.ncc) { pl.ncc += 1; pl.nmma = 1; }
  }
  // rows per split <= 4096 (bounds the fp32 running sums), CTA count a multiple of the SM count
  const int sms = sm_count();
  int64_t splits = ceil_div(m, max_rows);
  int64_t ctas = splits * pl.nchunks;
  ctas = ceil_div(ctas, sms) * sms;
  splits = max_rows > 4096 ? ctas / pl.nchunks : ceil_div(ctas, pl.nchunks);   // x1: never more CTAs than whole waves
  if (splits < 1) splits = 1;
  const int64_t cap = ((int64_t)512 << 20) / (n * l * 4 > 0 ? n * l * 4 : 1);
  if (splits > cap) splits = cap > 0 ? cap : 1;
  if (splits > 65535) splits = 65535;
  int64_t rps = ceil_div(ceil_div(m, splits), 16) * 16;
  splits = ceil_div(m, rps);
  pl.splits = splits;
  pl.rows_per_split = rps;
  pl.bytes = (size_t)(splits * n * l * 4);
  return pl;
}

}  // namespace era5svd

extern "C" {

int era5svd_split_tf32(const float* X, int64_t rows, int64_t cols, int64_t ldx, float* hi, float* lo,
                       int64_t ld_out, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(X && hi && lo, "split_tf32: null pointer");
  ERA5SVD_REQUIRE(rows > 0 && cols > 0 && ldx >= cols && ld_out >= cols, "split_tf32: bad shape");
  int64_t blocks = ceil_div(rows * cols, 256 * 8);
  int64_t cap = (int64_t)sm_count() * 32;
  if (blocks > cap) blocks = cap;
  tc::split_tf32_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(X, rows, cols, ldx, hi, lo, ld_out);
  return check_launch("split_tf32_kernel");
}

size_t era5svd_sketch_tf32x3_workspace_bytes(int64_t n, int64_t l) {
  using namespace era5svd;
  if (n <= 0 || l <= 0) return 0;
  return (size_t)2 * round_up(l, 16) * round_up(n + 3, 4) * sizeof(float);
}

static int sketch_tf32_impl(const float* Xhi, const float* Xlo, int64_t m, int64_t n, int64_t ldx,
                           const double* Om, int64_t l, int64_t ldo, float* Y, float* Yhi, float* Ylo,
                           int64_t ldy, void* workspace, size_t workspace_bytes, void* stream, int om_tf32,
                           int nprod = 3) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(Xhi && Om, "sketch_tf32x3: null pointer");
  ERA5SVD_REQUIRE(Y || Yhi, "sketch_tf32x3: no output requested");
  ERA5SVD_REQUIRE((Yhi == nullptr) == (Ylo == nullptr), "sketch_tf32x3: Yhi and Ylo go together");
  ERA5SVD_REQUIRE(m > 0 && n > 0 && l > 0 && ldx >= n && ldo >= l, "sketch_tf32x3: bad shape");
  const int npad = round_up(l, 16);
  if (npad > 256) {
    set_error("sketch_tf32x3: l = %lld > 256 is not supported by the tensor-core path", (long long)l);
    return ERA5SVD_ERR_UNSUPPORTED;
  }
  ERA5SVD_REQUIRE(ldy >= npad && ldy % 4 == 0, "sketch_tf32x3: ldy must be >= round_up(l, 16) = %d and a multiple of 4", npad);
  ERA5SVD_REQUIRE(m < ((int64_t)1 << 31), "sketch_tf32x3: m too large for TMA coordinates");
  for (const float* ptr : {Y, Yhi, Ylo})
    ERA5SVD_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "sketch_tf32x3: outputs must be 16-byte aligned");
  const size_t need = era5svd_sketch_tf32x3_workspace_bytes(n, l);
  if (!workspace || workspace_bytes < need) {
    set_error("sketch_tf32x3: workspace too small (%zu < %zu)", workspace_bytes, need);
    return ERA5SVD_ERR_WORKSPACE;
  }
  ERA5SVD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0, "sketch_tf32x3: workspace must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (!Xlo && nprod != 1)   // Xhi is the plain float32 matrix: split on chip (gemm_tc2.cu), X read from HBM once
    return sketch_tf32x3_raw(Xhi, m, n, ldx, Om, l, ldo, Y, Yhi, Ylo, ldy, workspace, st, om_tf32);
  ERA5SVD_REQUIRE(!om_tf32, "sketch_tf32x2: needs the plain float32 matrix (Xlo == NULL, on-chip split)");
  if (nprod == 1) Xlo = Xhi;   // single product: the raw tile is its own hi image (the tensor core truncates), no lo image
  CUtensorMap tm_xhi, tm_xlo, tm_ohi, tm_olo;
  int xs = 0, xs2 = 0, os = 0, os2 = 0, rc;
  if ((rc = tc::make_tmap(&tm_xhi, Xhi, n, m, ldx, tc::BK, tc::BM, &xs))) return rc;
  if ((rc = tc::make_tmap(&tm_xlo, Xlo, n, m, ldx, tc::BK, tc::BM, &xs2))) return rc;
  ERA5SVD_REQUIRE(xs == xs2, "sketch_tf32x3: Xhi and Xlo must have the same 16-byte phase");
  const int64_t kspan = n + xs;                  // K extent seen from the aligned map origin
  const int64_t ldt = round_up(kspan, 4);
  float* ohi = (float*)workspace;
  float* olo = ohi + (int64_t)npad * ldt;
  tc::split_omega_t_kernel<<<(unsigned)ceil_div((int64_t)npad * ldt, 256), 256, 0, st>>>(Om, n, l, ldo, ohi, olo, npad, ldt, xs);
  if ((rc = check_launch("split_omega_t_kernel"))) return rc;
  if ((rc = tc::make_tmap(&tm_ohi, ohi, kspan, npad, ldt, tc::BK, (uint32_t)npad, &os))) return rc;
  if ((rc = tc::make_tmap(&tm_olo, olo, kspan, npad, ldt, tc::BK, (uint32_t)npad, &os2))) return rc;
  ERA5SVD_REQUIRE(os == 0 && os2 == 0, "sketch_tf32x3: workspace must be 16-byte aligned");

  if (nprod == 1 && npad <= 128 && m > 2 * tc::BM && !getenv("ERA5SVD_SKETCH_X1_SINGLE")) {
    // two 128-row tiles per Om^T tile (sketch_x1_kernel)
    CUtensorMap tm_x2;
    int xs3 = 0;
    if ((rc = tc::make_tmap(&tm_x2, Xhi, n, m, ldx, tc::BK, 2 * tc::BM, &xs3))) return rc;
    tc::SketchX1Params q;
    q.m = m;
    q.num_groups = ceil_div(m, 2 * tc::BM);
    q.num_k = (int)ceil_div(kspan, tc::BK);
    q.npad = npad;
    q.Y = Y;
    q.ldy = ldy;
    const size_t stage = 2 * (size_t)tc::BM * tc::BK * 4 + (size_t)npad * tc::BK * 4;
    q.stages = (int)((227 * 1024 - 1024 - 256) / stage);
    if (q.stages > 6) q.stages = 6;
    const size_t smem = q.stages * stage + 1024 + 256;
    ERA5SVD_CUDA(ensure_dynamic_smem((const void*)tc::sketch_x1_kernel, smem));
    const int64_t grid = q.num_groups < sm_count() ? q.num_groups : sm_count();
    tc::sketch_x1_kernel<<<(unsigned)grid, tc::SX1_THREADS, smem, st>>>(tm_x2, tm_ohi, q);
    return check_launch("sketch_x1_kernel");
  }
  tc::SketchParams p;
  p.m = m;
  p.num_tiles = ceil_div(m, tc::BM);
  p.num_k = (int)ceil_div(kspan, tc::BK);
  p.npad = npad;
  p.nprod = nprod;
  p.Y = Y; p.Yhi = Yhi; p.Ylo = Ylo;
  p.ldy = ldy;
  const size_t stage_bytes = (nprod == 1 ? 1 : 2) * ((size_t)tc::BM * tc::BK * 4 + (size_t)npad * tc::BK * 4);
  const size_t budget = 227 * 1024 - 1024 - 256;
  p.stages = (int)(budget / stage_bytes);
  if (p.stages > (nprod == 1 ? 7 : 6)) p.stages = nprod == 1 ? 7 : 6;
  ERA5SVD_REQUIRE(p.stages >= 2, "sketch_tf32x3: not enough shared memory for two stages");
  const size_t smem = p.stages * stage_bytes + 1024 + 256;
  ERA5SVD_CUDA(ensure_dynamic_smem((const void*)tc::sketch_tc_kernel, smem));
  int64_t grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  tc::sketch_tc_kernel<<<(unsigned)grid, tc::SK_THREADS, smem, st>>>(tm_xhi, tm_xlo, tm_ohi, tm_olo, p);
  return check_launch("sketch_tc_kernel");
}

int era5svd_sketch_tf32x3(const float* Xhi, const float* Xlo, int64_t m, int64_t n, int64_t ldx,
                          const double* Om, int64_t l, int64_t ldo, float* Y, float* Yhi, float* Ylo,
                          int64_t ldy, void* workspace, size_t workspace_bytes, void* stream) {
  return sketch_tf32_impl(Xhi, Xlo, m, n, ldx, Om, l, ldo, Y, Yhi, Ylo, ldy, workspace, workspace_bytes, stream, 0);
}

int era5svd_sketch_tf32x2(const float* X, int64_t m, int64_t n, int64_t ldx, const double* Om, int64_t l,
                          int64_t ldo, float* Y, float* Yhi, float* Ylo, int64_t ldy, void* workspace,
                          size_t workspace_bytes, void* stream) {
  return sketch_tf32_impl(X, nullptr, m, n, ldx, Om, l, ldo, Y, Yhi, Ylo, ldy, workspace, workspace_bytes, stream, 1);
}

namespace era5svd {
namespace tc {
__global__ void __launch_bounds__(256)
round_tf32_f64_kernel(double* __restrict__ A, int64_t rows, int64_t cols, int64_t lda) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * cols) return;
  const int64_t r = idx / cols, c = idx - r * cols;
  A[r * lda + c] = (double)tf32_hi((float)A[r * lda + c]);
}
}  // namespace tc
}  // namespace era5svd

int era5svd_round_tf32_f64(double* A, int64_t rows, int64_t cols, int64_t lda, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(A, "round_tf32: null pointer");
  ERA5SVD_REQUIRE(rows > 0 && cols > 0 && lda >= cols, "round_tf32: bad shape");
  tc::round_tf32_f64_kernel<<<(unsigned)ceil_div(rows * cols, (int64_t)256), 256, 0, as_stream(stream)>>>(A, rows, cols, lda);
  return check_launch("round_tf32_f64_kernel");
}

int era5svd_sketch_tf32x1(const float* X, int64_t m, int64_t n, int64_t ldx, const double* Om, int64_t l,
                          int64_t ldo, float* Y, int64_t ldy, void* workspace, size_t workspace_bytes, void* stream) {
  return sketch_tf32_impl(X, nullptr, m, n, ldx, Om, l, ldo, Y, nullptr, nullptr, ldy, workspace, workspace_bytes,
                          stream, 0, 1);
}

size_t era5svd_project_tf32x3_workspace_bytes(int64_t m, int64_t n, int64_t l) {
  using namespace era5svd;
  if (m <= 0 || n <= 0 || l <= 0) return 0;
  if (l > 128) return pj_plan(m, n + 3, l, 16384).bytes;      // single-product projection with two Y tiles (l <= 256)
  size_t a = pj_plan(m, n + 3, l).bytes;
  const size_t b = project_tf32x3_raw_workspace_bytes(m, n, l), c = pj_plan(m, n + 3, l, 16384).bytes;
  if (b > a) a = b;
  return c > a ? c : a;
}

static int project_tf32_impl(const float* Xhi, const float* Xlo, int64_t m, int64_t n, int64_t ldx,
                             const float* Yhi, const float* Ylo, int64_t l, int64_t ldy, double* Z,
                             int64_t ldz, int accumulate, void* workspace, size_t workspace_bytes,
                             void* stream, int nprod) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(Xhi && Yhi && Z, "project_tf32x3: null pointer");
  ERA5SVD_REQUIRE(Ylo || !Xlo, "project_tf32x3: a plain Y (Ylo == NULL) needs the on-chip split path (Xlo == NULL too)");
  ERA5SVD_REQUIRE(m > 0 && n > 0 && l > 0 && ldx >= n && ldy >= l && ldz >= l, "project_tf32x3: bad shape");
  if (l > 128 && !(nprod == 1 && l <= 256)) {
    set_error("project_tf32x3: l = %lld is not supported by the tensor-core path (<= 128; <= 256 single product)",
              (long long)l);
    return ERA5SVD_ERR_UNSUPPORTED;
  }
  ERA5SVD_REQUIRE(m < ((int64_t)1 << 31), "project_tf32x3: m too large for TMA coordinates");
  if (!Xlo && nprod != 1)   // Xhi is the plain float32 matrix: split on chip (gemm_tc2.cu)
    return project_tf32x3_raw(Xhi, m, n, ldx, Yhi, Ylo, l, ldy, Z, ldz, accumulate, workspace, workspace_bytes,
                              as_stream(stream), nprod == 2);
  if (nprod == 1) { Xlo = Xhi; Ylo = Yhi; }   // single product on the raw tiles: no lo images are loaded
  // plan for the widest window (xshift <= 3) so that the workspace query needs no pointer
  const PjPlan pl = pj_plan(m, n + 3, l, nprod == 1 ? 16384 : 4096);
  if (!workspace || workspace_bytes < pl.bytes) {
    set_error("project_tf32x3: workspace too small (%zu < %zu)", workspace_bytes, pl.bytes);
    return ERA5SVD_ERR_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  constexpr int KS = 16;
  CUtensorMap tm_xhi, tm_xlo, tm_yhi, tm_ylo;
  int xs = 0, xs2 = 0, ys = 0, ys2 = 0, rc;
  if ((rc = tc::make_tmap(&tm_xhi, Xhi, n, m, ldx, tc::BK, KS, &xs, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = tc::make_tmap(&tm_xlo, Xlo, n, m, ldx, tc::BK, KS, &xs2, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = tc::make_tmap(&tm_yhi, Yhi, l, m, ldy, tc::BK, KS, &ys, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = tc::make_tmap(&tm_ylo, Ylo, l, m, ldy, tc::BK, KS, &ys2, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  ERA5SVD_REQUIRE(xs == xs2, "project_tf32x3: Xhi and Xlo must have the same 16-byte phase");
  ERA5SVD_REQUIRE(ys == 0 && ys2 == 0, "project_tf32x3: Yhi / Ylo must be 16-byte aligned");

  tc::ProjectParams p;
  p.m = m; p.n = n; p.l = (int)l;
  p.ks = KS;
  p.ncc = pl.ncc;
  p.nmma = pl.nmma;
  p.nprod = nprod;
  p.nya = l > 128 ? 2 : 1;
  p.xshift = xs;
  p.rows_per_split = pl.rows_per_split;
  p.part = (float*)workspace;
  const size_t box = (size_t)KS * tc::BK * 4;
  const size_t stage_bytes = (nprod == 1 ? 1 : 2) * (4 * (size_t)p.nya * box + (size_t)pl.ncc * box);
  const size_t budget = 227 * 1024 - 1024 - 256;
  p.stages = (int)(budget / stage_bytes);
  if (p.stages > 8) p.stages = 8;
  ERA5SVD_REQUIRE(p.stages >= 2, "project_tf32x3: not enough shared memory for two stages");
  const size_t smem = p.stages * stage_bytes + 1024 + 256;
  ERA5SVD_CUDA(ensure_dynamic_smem((const void*)tc::project_tc_kernel, smem));
  dim3 grid((unsigned)pl.nchunks, (unsigned)pl.splits);
  tc::project_tc_kernel<<<grid, tc::PJ_THREADS, smem, st>>>(tm_xhi, tm_xlo, tm_yhi, tm_ylo, p);
  if ((rc = check_launch("project_tc_kernel"))) return rc;
  return launch_reduce_partials_f32(p.part, pl.splits, n, l, l, Z, ldz, accumulate, st);
}

int era5svd_project_tf32x3(const float* Xhi, const float* Xlo, int64_t m, int64_t n, int64_t ldx,
                           const float* Yhi, const float* Ylo, int64_t l, int64_t ldy, double* Z,
                           int64_t ldz, int accumulate, void* workspace, size_t workspace_bytes,
                           void* stream) {
  return project_tf32_impl(Xhi, Xlo, m, n, ldx, Yhi, Ylo, l, ldy, Z, ldz, accumulate, workspace, workspace_bytes,
                           stream, 3);
}

int era5svd_project_tf32x2(const float* X, int64_t m, int64_t n, int64_t ldx, const float* Y, int64_t l,
                           int64_t ldy, double* Z, int64_t ldz, int accumulate, void* workspace,
                           size_t workspace_bytes, void* stream) {
  return project_tf32_impl(X, nullptr, m, n, ldx, Y, nullptr, l, ldy, Z, ldz, accumulate, workspace, workspace_bytes,
                           stream, 2);
}

int era5svd_project_tf32x1(const float* X, int64_t m, int64_t n, int64_t ldx, const float* Y, int64_t l,
                           int64_t ldy, double* Z, int64_t ldz, int accumulate, void* workspace,
                           size_t workspace_bytes, void* stream) {
  return project_tf32_impl(X, nullptr, m, n, ldx, Y, nullptr, l, ldy, Z, ldz, accumulate, workspace, workspace_bytes,
                           stream, 1);
}

}  // extern "C"
