// (b) Tall GEMM passes, tcgen05 generation 2: X is read from HBM ONCE as plain float32 and split on
// chip.  The raw tile lands in shared memory by TMA; four "transform" warps read it, form
// hi = tf32(x), lo = x - hi in registers and store both into TENSOR MEMORY with tcgen05.st; the MMAs
// then take A from TMEM (tcgen05.mma "TS" form) and only the small operand (Om^T resp. Y, pre-split,
// L2 resident) from shared memory.  Compared with gemm_tc.cu (hi / lo images of X in HBM) this halves
// the HBM traffic of every pass and removes the A-operand reads from the shared-memory port, which
// is what bounds the SS form (ncu: profiles/r01_ncu_full_tc_v1.md).
//
//   sketch :  Y = X Om        A = X tile   (TMEM: lane = row,  columns = time)     B = Om^T (K-major smem)
//   project:  Z = X^T Y       A = X^T tile (TMEM: lane = time, columns = rows)     B = Y    (N-major smem)
#include "common.cuh"
#include "tc_common.cuh"

namespace era5svd {

void launch_reduce_partials_f32(const float* part, int64_t splits, int64_t n, int64_t l, double* Z,
                                int64_t ldz, int accumulate, cudaStream_t st);

namespace tc {

int make_tmap(CUtensorMap* map, const float* base, int64_t inner, int64_t outer, int64_t ld,
              uint32_t box_inner, uint32_t box_outer, int* col_shift, CUtensorMapSwizzle swizzle);
__global__ void split_omega_t_kernel(const double* __restrict__ Om, int64_t n, int64_t l, int64_t ldo,
                                     float* __restrict__ hi, float* __restrict__ lo, int64_t npad, int64_t ldt,
                                     int shift);

constexpr int BM2 = 128;
constexpr int BK2 = 32;
constexpr int UK2 = 8;

struct Sketch2Params {
  int64_t m;
  int64_t num_tiles;
  int num_k;
  int npad;
  int stages;      // smem ring (raw A tile + B hi/lo)
  float* Y;
  float* Yhi;
  float* Ylo;
  int64_t ldy;
};

// warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: transform (smem -> hi/lo -> TMEM; two groups of four
// warps taking alternate k-chunks), warps 10-13: epilogue
constexpr int SK2_THREADS = 448;
constexpr int SK2_EPI_WARP0 = 10;
constexpr int A_RING = 4;          // TMEM A buffers: 64 columns each (32 hi + 32 lo)
constexpr uint32_t A_COLS = 64;

__global__ void __launch_bounds__(SK2_THREADS, 1)
sketch_tc2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_ohi,
                  const __grid_constant__ CUtensorMap tm_olo, const Sketch2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t a_bytes = BM2 * BK2 * 4;
  const uint32_t b_bytes = (uint32_t)p.npad * BK2 * 4;
  const uint32_t stage_bytes = a_bytes + 2 * b_bytes;
  const uint32_t bar_base = smem_base + (uint32_t)p.stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (p.stages + s); };
  auto aready_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + b); };
  auto aempty_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + A_RING + b); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + 2 * A_RING + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * p.stages + 2 * A_RING + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * p.stages + 2 * A_RING + 4);
  // TMEM map (512 columns): [0,128) acc 0, [128,256) acc 1, [256,512) A ring
  const uint32_t acc_cols = 128;
  const uint32_t a_col0 = 256;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < A_RING; ++b) {
      mbar_init(aready_bar(b), 4);    // one arrive per transform warp
      mbar_init(aempty_bar(b), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_ohi); tma_prefetch_desc(&tm_olo);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int32_t row0 = (int32_t)(tile * BM2);
        for (int kc = 0; kc < p.num_k; ++kc) {
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
          mbar_arrive_expect_tx(full_bar(s), stage_bytes);
          tma_load_2d(st, &tm_x, kc * BK2, row0, full_bar(s));
          tma_load_2d(st + a_bytes, &tm_ohi, kc * BK2, 0, full_bar(s));
          tma_load_2d(st + a_bytes + b_bytes, &tm_olo, kc * BK2, 0, full_bar(s));
          if (++s == p.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A (hi / lo) from TMEM, B (Om^T hi / lo) from smem =====
    const uint32_t idesc = make_idesc_tf32(BM2, p.npad, 0, 0);
    int s = 0; uint32_t ph = 0;
    int ab = 0; uint32_t aph = 0;
    int buf = 0; uint32_t tph = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(buf), tph ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)buf * acc_cols;
      for (int kc = 0; kc < p.num_k; ++kc) {
        mbar_wait(full_bar(s), ph);          // B tiles landed (and the raw A tile)
        mbar_wait(aready_bar(ab), aph);      // A hi / lo stored to TMEM by the transform warps
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t st = smem_base + (uint32_t)s * stage_bytes;
          const uint32_t a_hi = tmem_base + a_col0 + (uint32_t)ab * A_COLS;
          const uint32_t a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < BK2 / UK2; ++kk) {
            const uint32_t koff = kk * UK2 * 4;
            const uint64_t b_hi = make_smem_desc(st + a_bytes + koff, 16, 1024);
            const uint64_t b_lo = make_smem_desc(st + a_bytes + b_bytes + koff, 16, 1024);
            umma_tf32_ts(d_tmem, a_lo + kk * UK2, b_hi, idesc, (kc | kk) != 0);
            umma_tf32_ts(d_tmem, a_hi + kk * UK2, b_lo, idesc, 1);
            umma_tf32_ts(d_tmem, a_hi + kk * UK2, b_hi, idesc, 1);
          }
          umma_commit(empty_bar(s));
          umma_commit(aempty_bar(ab));
          if (kc == p.num_k - 1) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++s == p.stages) { s = 0; ph ^= 1u; }
        if (++ab == A_RING) { ab = 0; aph ^= 1u; }
      }
      if (++buf == 2) { buf = 0; tph ^= 1u; }
    }
  } else if (warp < SK2_EPI_WARP0) {
    // ===== transform: raw X tile (smem, 128 B swizzle) -> hi / lo -> TMEM (lane = row) =====
    const int q = warp % 4;
    const int grp = (warp - 2) / 4;                    // this group handles chunks with (count & 1) == grp
    const int r = q * 32 + lane;                       // row of the tile == TMEM lane
    int s = 0; uint32_t ph = 0;
    int ab = 0; uint32_t aph = 0;
    uint32_t cnt = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < p.num_k; ++kc, ++cnt) {
        if ((cnt & 1u) != (uint32_t)grp) {
          if (++s == p.stages) { s = 0; ph ^= 1u; }
          if (++ab == A_RING) { ab = 0; aph ^= 1u; }
          continue;
        }
        mbar_wait(full_bar(s), ph);
        mbar_wait(aempty_bar(ab), aph ^ 1u);           // MMAs that read this TMEM buffer have retired
        tcgen05_fence_after();
        const uint32_t row_addr = smem_base + (uint32_t)s * stage_bytes + (uint32_t)r * 128u;
        const uint32_t t_hi = tmem_base + a_col0 + (uint32_t)ab * A_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int chunk = half * 4 + c;            // 16-byte chunk of the 128-byte row
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(row_addr + (uint32_t)((chunk ^ (r & 7)) << 4)));
            // truncation split (1 LOP + 1 FADD per element): hi = x with the low 13 mantissa bits cleared,
            // lo = x - hi exact, |lo| < 2^-10 |x|
            const float h0 = tf32_trunc(v.x), h1 = tf32_trunc(v.y), h2 = tf32_trunc(v.z), h3 = tf32_trunc(v.w);
            hi[4 * c + 0] = __float_as_uint(h0); lo[4 * c + 0] = __float_as_uint(v.x - h0);
            hi[4 * c + 1] = __float_as_uint(h1); lo[4 * c + 1] = __float_as_uint(v.y - h1);
            hi[4 * c + 2] = __float_as_uint(h2); lo[4 * c + 2] = __float_as_uint(v.z - h2);
            hi[4 * c + 3] = __float_as_uint(h3); lo[4 * c + 3] = __float_as_uint(v.w - h3);
          }
          tmem_st16(t_hi + half * 16, hi);
          tmem_st16(t_hi + 32 + half * 16, lo);
        }
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(aready_bar(ab));
        if (++s == p.stages) { s = 0; ph ^= 1u; }
        if (++ab == A_RING) { ab = 0; aph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global (Y, Y_hi, Y_lo) =====
    const int q = warp % 4;
    int buf = 0; uint32_t tph = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tfull_bar(buf), tph);
      tcgen05_fence_after();
      const int64_t row = tile * BM2 + q * 32 + lane;
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
      if (++buf == 2) { buf = 0; tph ^= 1u; }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

static int round_up2(int64_t a, int64_t b) { return (int)(ceil_div(a, b) * b); }

// Y = X Om with X a plain float32 matrix (split on chip).  Called by era5svd_sketch_tf32x3 when Xlo == NULL.
int sketch_tf32x3_raw(const float* X, int64_t m, int64_t n, int64_t ldx, const double* Om, int64_t l,
                      int64_t ldo, float* Y, float* Yhi, float* Ylo, int64_t ldy, void* workspace,
                      cudaStream_t st) {
  const int npad = round_up2(l, 16);
  if (npad > 128) {
    set_error("sketch_tf32x3 (on-chip split): l = %lld > 128 is not supported", (long long)l);
    return ERA5SVD_ERR_UNSUPPORTED;
  }
  CUtensorMap tm_x, tm_ohi, tm_olo;
  int xs = 0, os = 0, os2 = 0, rc;
  if ((rc = tc::make_tmap(&tm_x, X, n, m, ldx, tc::BK2, tc::BM2, &xs, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  const int64_t kspan = n + xs;
  const int64_t ldt = round_up2(kspan, 4);
  float* ohi = (float*)workspace;
  float* olo = ohi + (int64_t)npad * ldt;
  tc::split_omega_t_kernel<<<(unsigned)ceil_div((int64_t)npad * ldt, 256), 256, 0, st>>>(Om, n, l, ldo, ohi, olo, npad, ldt, xs);
  if ((rc = check_launch("split_omega_t_kernel"))) return rc;
  if ((rc = tc::make_tmap(&tm_ohi, ohi, kspan, npad, ldt, tc::BK2, (uint32_t)npad, &os, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = tc::make_tmap(&tm_olo, olo, kspan, npad, ldt, tc::BK2, (uint32_t)npad, &os2, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  ERA5SVD_REQUIRE(os == 0 && os2 == 0, "sketch_tf32x3: workspace must be 16-byte aligned");

  tc::Sketch2Params p;
  p.m = m;
  p.num_tiles = ceil_div(m, tc::BM2);
  p.num_k = (int)ceil_div(kspan, tc::BK2);
  p.npad = npad;
  p.Y = Y; p.Yhi = Yhi; p.Ylo = Ylo;
  p.ldy = ldy;
  const size_t stage_bytes = (size_t)tc::BM2 * tc::BK2 * 4 + 2 * (size_t)npad * tc::BK2 * 4;
  const size_t budget = 227 * 1024 - 1024 - 512;
  p.stages = (int)(budget / stage_bytes);
  if (p.stages > 6) p.stages = 6;
  ERA5SVD_REQUIRE(p.stages >= 2, "sketch_tf32x3: not enough shared memory for two stages");
  const size_t smem = p.stages * stage_bytes + 1024 + 512;
  ERA5SVD_CUDA(cudaFuncSetAttribute(tc::sketch_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  tc::sketch_tc2_kernel<<<(unsigned)grid, tc::SK2_THREADS, smem, st>>>(tm_x, tm_ohi, tm_olo, p);
  return check_launch("sketch_tc2_kernel");
}

}  // namespace era5svd
