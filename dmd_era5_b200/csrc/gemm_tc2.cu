// (b) Tall GEMM passes, tcgen05 generation 2: X is read from HBM ONCE as plain float32 and split on
// chip.  The raw tile lands in shared memory by TMA; "transform" warps read it, form
// hi = trunc_tf32(x), lo = x - hi in registers and store both into TENSOR MEMORY with tcgen05.st; the
// MMAs then take A from TMEM (tcgen05.mma "TS" form) and only the small operand (Om^T, pre-split,
// L2 resident) from shared memory.  Compared with gemm_tc.cu (hi / lo images of X in HBM) this halves
// the HBM traffic of the pass and removes the A-operand reads from the shared-memory port, which is
// what bounds the SS form (ncu: profiles/r01_ncu_full_tc_v1.md).
//
//   sketch :  Y = X Om        A = X tile (TMEM: lane = row, columns = time)     B = Om^T (K-major smem)
//
// The MMA issuer must live entirely in the uniform datapath: the warp index comes from __shfl_sync (so that
// the role branches are warp-uniform for the compiler) and the single issuing thread is picked with
// elect.sync.  With `threadIdx.x / 32` and `lane == 0` ptxas wraps EVERY tcgen05.mma into an R2UR /
// BRA.U.ANY loop (~100 clk per instruction), which bounded generation 2 of these kernels at half the
// tensor rate.  Measured instruction cost (profiles/r01_microbench_mma_probe.txt): ~9 + N/2 clk for
// M = 128, K = 8 with A in TMEM, i.e. a 3xTF32 k-step at N = 112 takes ~180 clk however it is split.
//
// Three decoupled rings, so that the HBM prefetch distance does not depend on MMA completion:
//   raw-A ring (smem, 16 KB/slot) : TMA  -> transform warps          freed as soon as it has been read
//   B ring     (smem, 28 KB/slot) : TMA  -> MMA                      freed by tcgen05.commit
//   A ring     (TMEM, 64 cols)    : transform -> MMA                 freed by tcgen05.commit
#include "common.cuh"
#include "tc_common.cuh"

namespace era5svd {

int launch_reduce_partials_f32(const float* part, int64_t splits, int64_t n, int64_t l, int64_t lp, double* Z,
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
  int ra;          // raw-A ring depth
  int rb;          // B ring depth
  int use_lo;      // 1: Om = hi + lo (three MMAs per k-step); 0: Om is tf32-exact, lo tile neither loaded nor multiplied
  float* Y;
  float* Yhi;
  float* Ylo;
  int64_t ldy;
};

// warp 0: TMA producer for raw A, warp 1: MMA issuer + TMEM alloc, warps 2-9: transform (two groups of four
// warps taking alternate k-chunks), warps 10-13: epilogue, warp 14: TMA producer for B
constexpr int SK2_THREADS = 480;
constexpr int SK2_EPI_WARP0 = 10;
constexpr int SK2_BPROD_WARP = 14;
constexpr int AT_RING = 4;         // TMEM A buffers: 64 columns each (32 hi + 32 lo)
constexpr uint32_t AT_COLS = 64;

struct Ring {
  int i = 0;
  uint32_t ph = 0;
  int depth;
  __device__ explicit Ring(int d) : depth(d) {}
  __device__ void next() {
    if (++i == depth) { i = 0; ph ^= 1u; }
  }
};

__global__ void __launch_bounds__(SK2_THREADS, 1)
sketch_tc2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_ohi,
                  const __grid_constant__ CUtensorMap tm_olo, const Sketch2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;   // warp-uniform
  const uint32_t a_bytes = BM2 * BK2 * 4;                   // raw A slot
  const uint32_t b_half = (uint32_t)p.npad * BK2 * 4;       // Om^T hi (or lo) tile
  const uint32_t b_bytes = p.use_lo ? 2 * b_half : b_half;  // B slot
  const uint32_t b_base = smem_base + (uint32_t)p.ra * a_bytes;
  const uint32_t bar_base = b_base + (uint32_t)p.rb * b_bytes;
  int nb = 0;
  const uint32_t araw_full = bar_base + 8u * nb;  nb += p.ra;
  const uint32_t araw_empty = bar_base + 8u * nb; nb += p.ra;
  const uint32_t b_full = bar_base + 8u * nb;     nb += p.rb;
  const uint32_t b_empty = bar_base + 8u * nb;    nb += p.rb;
  const uint32_t at_ready = bar_base + 8u * nb;   nb += AT_RING;
  const uint32_t at_empty = bar_base + 8u * nb;   nb += AT_RING;
  const uint32_t tfull = bar_base + 8u * nb;      nb += 2;
  const uint32_t tempty = bar_base + 8u * nb;     nb += 2;
  const uint32_t tmem_slot = bar_base + 8u * nb;
  // TMEM map (512 columns): [0,128) acc 0, [128,256) acc 1, [256,512) A ring
  const uint32_t acc_cols = 128;
  const uint32_t at_col0 = 256;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ra; ++i) {
      mbar_init(araw_full + 8u * i, 1);
      mbar_init(araw_empty + 8u * i, 4);   // the four transform warps that read the slot
    }
    for (int i = 0; i < p.rb; ++i) {
      mbar_init(b_full + 8u * i, 1);
      mbar_init(b_empty + 8u * i, 1);
    }
    for (int i = 0; i < AT_RING; ++i) {
      mbar_init(at_ready + 8u * i, 4);
      mbar_init(at_empty + 8u * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull + 8u * i, 1);
      mbar_init(tempty + 8u * i, 4);
    }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_x);
  if (warp == SK2_BPROD_WARP && lane == 0) { tma_prefetch_desc(&tm_ohi); tma_prefetch_desc(&tm_olo); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== TMA producer: raw X tiles (HBM stream) =====
    if (elect_one()) {
      Ring ra(p.ra);
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int32_t row0 = (int32_t)(tile * BM2);
        for (int kc = 0; kc < p.num_k; ++kc) {
          mbar_wait_hint(araw_empty + 8u * ra.i, ra.ph ^ 1u, 2000u);
          mbar_arrive_expect_tx(araw_full + 8u * ra.i, a_bytes);
          tma_load_2d(smem_base + (uint32_t)ra.i * a_bytes, &tm_x, kc * BK2, row0, araw_full + 8u * ra.i);
          ra.next();
        }
      }
    }
  } else if (warp == SK2_BPROD_WARP) {
    // ===== TMA producer: Om^T hi / lo tiles (L2 resident, same for every row tile) =====
    if (elect_one()) {
      Ring rb(p.rb);
      for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int kc = 0; kc < p.num_k; ++kc) {
          mbar_wait(b_empty + 8u * rb.i, rb.ph ^ 1u);
          const uint32_t dst = b_base + (uint32_t)rb.i * b_bytes;
          mbar_arrive_expect_tx(b_full + 8u * rb.i, b_bytes);
          tma_load_2d(dst, &tm_ohi, kc * BK2, 0, b_full + 8u * rb.i);
          if (p.use_lo) tma_load_2d(dst + b_half, &tm_olo, kc * BK2, 0, b_full + 8u * rb.i);
          rb.next();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A (hi / lo) from TMEM, B (Om^T hi / lo) from smem =====
    const uint32_t idesc = make_idesc_tf32(BM2, p.npad, 0, 0);
    const bool use_lo = p.use_lo != 0;            // kernel parameter: uniform, so the issue loop stays branch-free per chunk
    Ring rb(p.rb), at(AT_RING), acc(2);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(tempty + 8u * acc.i, acc.ph ^ 1u);
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc.i * acc_cols;
      for (int kc = 0; kc < p.num_k; ++kc) {
        mbar_wait(b_full + 8u * rb.i, rb.ph);
        mbar_wait(at_ready + 8u * at.i, at.ph);
        tcgen05_fence_after();
        if (elect_one()) {
          const uint32_t bs = b_base + (uint32_t)rb.i * b_bytes;
          const uint32_t a_hi = tmem_base + at_col0 + (uint32_t)at.i * AT_COLS;
          const uint32_t a_lo = a_hi + 32;
#pragma unroll
          for (int kk = 0; kk < BK2 / UK2; ++kk) {
            const uint32_t koff = kk * UK2 * 4;
            const uint64_t b_hi = make_smem_desc(bs + koff, 16, 1024);
            const uint64_t b_lo = make_smem_desc(bs + b_half + koff, 16, 1024);
            umma_tf32_ts(d_tmem, a_lo + kk * UK2, b_hi, idesc, (kc | kk) != 0);
            if (use_lo) umma_tf32_ts(d_tmem, a_hi + kk * UK2, b_lo, idesc, 1);
            umma_tf32_ts(d_tmem, a_hi + kk * UK2, b_hi, idesc, 1);
          }
          umma_commit(b_empty + 8u * rb.i);
          umma_commit(at_empty + 8u * at.i);
          if (kc == p.num_k - 1) umma_commit(tfull + 8u * acc.i);
        }
        __syncwarp();
        rb.next();
        at.next();
      }
      acc.next();
    }
  } else if (warp < SK2_EPI_WARP0) {
    // ===== transform: raw X tile (smem, 128 B swizzle) -> hi / lo -> TMEM (lane = row) =====
    const int q = warp % 4;
    const int grp = (warp - 2) / 4;                    // this group handles chunks with (count & 1) == grp
    const int r = q * 32 + lane;                       // row of the tile == TMEM lane
    Ring ra(p.ra), at(AT_RING);
    uint32_t cnt = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      for (int kc = 0; kc < p.num_k; ++kc, ++cnt) {
        if ((cnt & 1u) == (uint32_t)grp) {
          mbar_wait_hint(araw_full + 8u * ra.i, ra.ph, 2000u);
          const uint32_t row_addr = smem_base + (uint32_t)ra.i * a_bytes + (uint32_t)r * 128u;
          float4 v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c)                  // 16-byte chunk c of the 128-byte row, un-swizzled
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[c].x), "=f"(v[c].y), "=f"(v[c].z), "=f"(v[c].w)
                         : "r"(row_addr + (uint32_t)((c ^ (r & 7)) << 4)));
          mbar_wait_hint(at_empty + 8u * at.i, at.ph ^ 1u, 2000u);   // MMAs that read this TMEM buffer have retired
          tcgen05_fence_after();
          const uint32_t t_hi = tmem_base + at_col0 + (uint32_t)at.i * AT_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              // truncation split (1 LOP + 1 FADD per element): hi = x with the low 13 mantissa bits
              // cleared, lo = x - hi exact, |lo| < 2^-10 |x|
              const float4 x = v[half * 4 + c];
              const float h0 = tf32_trunc(x.x), h1 = tf32_trunc(x.y), h2 = tf32_trunc(x.z), h3 = tf32_trunc(x.w);
              hi[4 * c + 0] = __float_as_uint(h0); lo[4 * c + 0] = __float_as_uint(x.x - h0);
              hi[4 * c + 1] = __float_as_uint(h1); lo[4 * c + 1] = __float_as_uint(x.y - h1);
              hi[4 * c + 2] = __float_as_uint(h2); lo[4 * c + 2] = __float_as_uint(x.z - h2);
              hi[4 * c + 3] = __float_as_uint(h3); lo[4 * c + 3] = __float_as_uint(x.w - h3);
            }
            tmem_st16(t_hi + half * 16, hi);
            tmem_st16(t_hi + 32 + half * 16, lo);
          }
          tmem_wait_st();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) {
            // The raw slot is released only HERE, after the loaded values have been consumed.  An arrive placed
            // right after the ld.shared instructions can become visible before those loads have read shared
            // memory; the TMA refill then races with them (seen as corrupted low rows of a tile once the MMA
            // issue got fast enough for the producer to run just one slot ahead).
            mbar_arrive(araw_empty + 8u * ra.i);
            mbar_arrive(at_ready + 8u * at.i);
          }
        }
        ra.next();
        at.next();
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global (Y, Y_hi, Y_lo) =====
    const int q = warp % 4;
    Ring acc(2);
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait_hint(tfull + 8u * acc.i, acc.ph, 20000u);
      tcgen05_fence_after();
      const int64_t row = tile * BM2 + q * 32 + lane;
      const uint32_t taddr = tmem_base + (uint32_t)acc.i * acc_cols + ((uint32_t)(q * 32) << 16);
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
      if (lane == 0) mbar_arrive(tempty + 8u * acc.i);
      acc.next();
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


// ---------------------------------------------------------------------------------------------
// project v2:  Z[n x l] += X^T Y over a row split, X split on chip.
//   A = X^T tile in TMEM (lane = time within a 128-wide tile, columns = rows of the slot, K-major)
//   B = Y tile (hi / lo, pre-split by the sketch epilogue) from smem, N-major (32 B-granule swizzle)
//   D = [128 time lanes x 128 sketch columns] per time tile, two time tiles (256 time columns) per CTA,
//       accumulated in TMEM over the CTA's whole row split, then written as a float32 partial tile.
// grid = (ceil(n / 256), splits); one slot = 16 rows x 256 time values (16 KB of X).
// ---------------------------------------------------------------------------------------------
struct Project2Params {
  int64_t m, n;
  int l;
  int xshift;
  int ra, rb;
  int64_t rows_per_split;
  int split_y;        // 1: tm_yhi maps the PLAIN float32 Y; its lo image is formed on chip by the (otherwise idle) epilogue warps
                      // 2: plain Y taken as tf32(Y) (the tensor core truncates it): NO lo image, two products per k-step
  int lp;             // column pitch of a partial tile: l rounded up to 16 (pad columns hold the zeros the MMA produced)
  float* part;        // [splits][n][lp]
};

constexpr int PJ2_THREADS = 480;
constexpr int PJ2_KS = 16;                 // rows per slot
constexpr int PJ2_TT = 2;                  // time tiles per CTA
constexpr int PJ2_NC = PJ2_TT * 128;       // time columns per CTA
constexpr uint32_t PJ2_AT_COLS = PJ2_TT * 2 * PJ2_KS;   // 64 TMEM columns per A slot

__global__ void __launch_bounds__(PJ2_THREADS, 1)
project_tc2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_yhi,
                   const __grid_constant__ CUtensorMap tm_ylo, const Project2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;   // warp-uniform
  const uint32_t box_bytes = PJ2_KS * BK2 * 4;                  // [16 rows x 32 floats] = 2 KB
  const uint32_t a_bytes = (PJ2_NC / BK2) * box_bytes;          // 8 boxes = 16 KB
  const uint32_t y_half = 4 * box_bytes;                        // 128 sketch columns, hi (or lo)
  const uint32_t b_bytes = 2 * y_half;
  const uint32_t b_base = smem_base + (uint32_t)p.ra * a_bytes;
  const uint32_t bar_base = b_base + (uint32_t)p.rb * b_bytes;
  int nb = 0;
  const uint32_t araw_full = bar_base + 8u * nb;  nb += p.ra;
  const uint32_t araw_empty = bar_base + 8u * nb; nb += p.ra;
  const uint32_t b_full = bar_base + 8u * nb;     nb += p.rb;
  const uint32_t b_empty = bar_base + 8u * nb;    nb += p.rb;
  const uint32_t at_ready = bar_base + 8u * nb;   nb += AT_RING;
  const uint32_t at_empty = bar_base + 8u * nb;   nb += AT_RING;
  const uint32_t tfull = bar_base + 8u * nb;      nb += 1;
  const uint32_t b_raw = bar_base + 8u * nb;      nb += p.rb;    // split_y: raw Y tile landed (TMA) -> epilogue warps
  const uint32_t tmem_slot = bar_base + 8u * nb;
  const uint32_t at_col0 = 256;                                  // [0,256): two accumulators, [256,512): A ring

  const int64_t t0 = (int64_t)blockIdx.x * PJ2_NC;
  const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_split;
  const int64_t r_end = min(p.m, r_begin + p.rows_per_split);
  const int num_k = (int)((r_end - r_begin + PJ2_KS - 1) / PJ2_KS);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ra; ++i) {
      mbar_init(araw_full + 8u * i, 1);
      mbar_init(araw_empty + 8u * i, 8);      // all eight transform warps read every slot
    }
    for (int i = 0; i < p.rb; ++i) {
      mbar_init(b_full + 8u * i, p.split_y == 1 ? 4 : 1);   // split_y == 1: the four warps that wrote the lo image arrive
      mbar_init(b_empty + 8u * i, 1);
      mbar_init(b_raw + 8u * i, 1);
    }
    for (int i = 0; i < AT_RING; ++i) {
      mbar_init(at_ready + 8u * i, 8);
      mbar_init(at_empty + 8u * i, 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&tm_x);
  if (warp == SK2_BPROD_WARP && lane == 0) { tma_prefetch_desc(&tm_yhi); tma_prefetch_desc(&tm_ylo); }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // ===== TMA producer: raw X slots (HBM stream), 8 boxes of [16 rows x 32 time] =====
    if (elect_one()) {
      Ring ra(p.ra);
      for (int kc = 0; kc < num_k; ++kc) {
        mbar_wait(araw_empty + 8u * ra.i, ra.ph ^ 1u);
        const uint32_t dst = smem_base + (uint32_t)ra.i * a_bytes;
        const int32_t row0 = (int32_t)(r_begin + (int64_t)kc * PJ2_KS);
        mbar_arrive_expect_tx(araw_full + 8u * ra.i, a_bytes);
        for (int c = 0; c < PJ2_NC / BK2; ++c)
          tma_load_2d(dst + c * box_bytes, &tm_x, (int32_t)(t0 + c * BK2), row0, araw_full + 8u * ra.i);
        ra.next();
      }
    }
  } else if (warp == SK2_BPROD_WARP) {
    // ===== TMA producer: Y hi / lo slots, 4 + 4 boxes of [16 rows x 32 columns] =====
    if (elect_one()) {
      Ring rb(p.rb);
      for (int kc = 0; kc < num_k; ++kc) {
        mbar_wait(b_empty + 8u * rb.i, rb.ph ^ 1u);
        const uint32_t dst = b_base + (uint32_t)rb.i * b_bytes;
        const int32_t row0 = (int32_t)(r_begin + (int64_t)kc * PJ2_KS);
        if (p.split_y == 1) {
          mbar_arrive_expect_tx(b_raw + 8u * rb.i, y_half);
          for (int c = 0; c < 4; ++c) tma_load_2d(dst + c * box_bytes, &tm_yhi, c * BK2, row0, b_raw + 8u * rb.i);
        } else if (p.split_y == 2) {
          mbar_arrive_expect_tx(b_full + 8u * rb.i, y_half);
          for (int c = 0; c < 4; ++c) tma_load_2d(dst + c * box_bytes, &tm_yhi, c * BK2, row0, b_full + 8u * rb.i);
        } else {
          mbar_arrive_expect_tx(b_full + 8u * rb.i, b_bytes);
          for (int c = 0; c < 4; ++c) {
            tma_load_2d(dst + c * box_bytes, &tm_yhi, c * BK2, row0, b_full + 8u * rb.i);
            tma_load_2d(dst + y_half + c * box_bytes, &tm_ylo, c * BK2, row0, b_full + 8u * rb.i);
          }
        }
        rb.next();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc_tf32(128, (p.l + 15) / 16 * 16, 0, 1);   // A K-major (TMEM), B MN-major, N = l padded to 16
    const bool with_ylo = p.split_y != 2;          // kernel parameter: uniform
    Ring rb(p.rb), at(AT_RING);
    for (int kc = 0; kc < num_k; ++kc) {
      mbar_wait(b_full + 8u * rb.i, rb.ph);
      mbar_wait(at_ready + 8u * at.i, at.ph);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t bs = b_base + (uint32_t)rb.i * b_bytes;
        const uint32_t a0 = tmem_base + at_col0 + (uint32_t)at.i * PJ2_AT_COLS;
#pragma unroll
        for (int tt = 0; tt < PJ2_TT; ++tt) {
          const uint32_t d_tmem = tmem_base + (uint32_t)tt * 128;
#pragma unroll
          for (int ks = 0; ks < PJ2_KS / UK2; ++ks) {
            const uint32_t a_hi = a0 + (uint32_t)tt * 2 * PJ2_KS + ks * UK2;
            const uint32_t a_lo = a_hi + PJ2_KS;
            const uint32_t koff = (uint32_t)ks * UK2 * 128;      // 8 K rows = two 4-row swizzle atoms
            const uint64_t b_hi = make_smem_desc(bs + koff, box_bytes, 512, LAYOUT_SW128_BASE32B);
            const uint64_t b_lo = make_smem_desc(bs + y_half + koff, box_bytes, 512, LAYOUT_SW128_BASE32B);
            umma_tf32_ts(d_tmem, a_lo, b_hi, idesc, (kc | ks) != 0);
            if (with_ylo) umma_tf32_ts(d_tmem, a_hi, b_lo, idesc, 1);
            umma_tf32_ts(d_tmem, a_hi, b_hi, idesc, 1);
          }
        }
        umma_commit(b_empty + 8u * rb.i);
        umma_commit(at_empty + 8u * at.i);
        if (kc == num_k - 1) umma_commit(tfull);
      }
      __syncwarp();
      rb.next();
      at.next();
    }
  } else if (warp < SK2_EPI_WARP0) {
    // ===== transform: X slot (smem [row][time]) -> hi / lo of X^T -> TMEM (lane = time, column = row) =====
    const int q = warp % 4;
    const int tt = (warp - 2) / 4;                       // time tile handled by this group of four warps
    const int tl = q * 32 + lane;                        // time index within the tile == TMEM lane
    const int box = (tt * 128 + tl) / BK2;               // which [16 x 32] box holds this time column
    const int col = tl % BK2;                            // float index inside the 128-byte box row
    Ring ra(p.ra), at(AT_RING);
    for (int kc = 0; kc < num_k; ++kc) {
      mbar_wait_hint(araw_full + 8u * ra.i, ra.ph, 2000u);
      const uint32_t base = smem_base + (uint32_t)ra.i * a_bytes + (uint32_t)box * box_bytes;
      float x[PJ2_KS];
#pragma unroll
      for (int k = 0; k < PJ2_KS; ++k) {                 // row k of the box: 128 B, 16-byte chunks XOR-swizzled by (k & 7)
        const uint32_t addr = base + (uint32_t)k * 128u + (uint32_t)((((col >> 2) ^ (k & 7)) << 4) | ((col & 3) << 2));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[k]) : "r"(addr));
      }
      mbar_wait_hint(at_empty + 8u * at.i, at.ph ^ 1u, 2000u);
      tcgen05_fence_after();
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int k = 0; k < PJ2_KS; ++k) {
        const float h = tf32_trunc(x[k]);
        hi[k] = __float_as_uint(h);
        lo[k] = __float_as_uint(x[k] - h);
      }
      const uint32_t t_hi = tmem_base + at_col0 + (uint32_t)at.i * PJ2_AT_COLS + (uint32_t)tt * 2 * PJ2_KS +
                            ((uint32_t)(q * 32) << 16);
      tmem_st16(t_hi, hi);
      tmem_st16(t_hi + PJ2_KS, lo);
      tmem_wait_st();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(araw_empty + 8u * ra.i);        // released after the loaded values were consumed (see sketch)
        mbar_arrive(at_ready + 8u * at.i);
      }
      ra.next();
      at.next();
    }
  } else {
    // ===== epilogue: accumulators -> float32 partial tile part[split][time][column] =====
    const int q = warp % 4;
    float* out = p.part + (int64_t)blockIdx.y * p.n * p.lp;
    if (p.split_y == 1) {
      // Y arrives as ONE plain float32 image (the sketch then writes, and this pass reads, m*l*4 bytes instead of
      // twice that).  The tensor core truncates its operands to tf32, so the raw tile already serves as the hi
      // operand; these warps (idle until the accumulators are complete) form lo = y - trunc(y) next to it.  The
      // tile layout (32-byte-granule swizzle) does not matter: the split is elementwise, same offsets in both tiles.
      Ring rb(p.rb);
      const uint32_t e = (uint32_t)(q * 32 + lane);            // 128 threads x 4 chunks of 16 bytes = y_half (8 KB)
      for (int kc = 0; kc < num_k; ++kc) {
        mbar_wait_hint(b_raw + 8u * rb.i, rb.ph, 2000u);
        const uint32_t src = b_base + (uint32_t)rb.i * b_bytes;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t off = (e + 128u * c) << 4;
          float4 y;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w) : "r"(src + off));
          const float4 w = make_float4(y.x - tf32_trunc(y.x), y.y - tf32_trunc(y.y), y.z - tf32_trunc(y.z), y.w - tf32_trunc(y.w));
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(src + y_half + off), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(b_full + 8u * rb.i);
        rb.next();
      }
    }
    if (num_k > 0) {
      mbar_wait_hint(tfull, 0, 50000u);
      tcgen05_fence_after();
    }
    for (int tt = 0; tt < PJ2_TT; ++tt) {
      const int64_t t = t0 + tt * 128 + q * 32 + lane - p.xshift;     // window-relative time index of this lane
      const uint32_t taddr = tmem_base + (uint32_t)tt * 128 + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.l; c0 += 16) {
        uint32_t v[16];
        if (num_k > 0) {
          tmem_ld16(taddr + (uint32_t)c0, v);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0u;
        }
        if (t >= 0 && t < p.n) {
          // 16-byte aligned rows of the partial tile: four 16-byte stores per lane instead of sixteen scalar ones
          float4* dst = reinterpret_cast<float4*>(out + t * p.lp + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                                 __uint_as_float(v[4 * j + 3]));
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc

static int round_up2(int64_t a, int64_t b) { return (int)(ceil_div(a, b) * b); }

// Y = X Om with X a plain float32 matrix (split on chip).  Called by era5svd_sketch_tf32x3 when Xlo == NULL.
// om_tf32 != 0: Y = X tf32(Om) - the small factor is taken rounded to tf32 (for a factor that already holds
// tf32-representable values, era5svd_round_tf32_f64, this is exact), so its lo image vanishes: two MMAs per k-step
// instead of three and half the Om^T tile traffic.
int sketch_tf32x3_raw(const float* X, int64_t m, int64_t n, int64_t ldx, const double* Om, int64_t l,
                      int64_t ldo, float* Y, float* Yhi, float* Ylo, int64_t ldy, void* workspace,
                      cudaStream_t st, int om_tf32) {
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
  // shared memory: B ring of 4 slots (L2 latency), the rest for the raw-A ring (HBM latency), up to 8 slots.
  // The raw-A depth must be EVEN: the two transform groups take alternate chunks, so with an even depth every slot
  // always belongs to the same group.  With an odd depth a group meets a slot only every other use and its parity wait
  // can be satisfied by the use it skipped (seen as a race and, eventually, a hung barrier at depth 5).
  p.use_lo = om_tf32 ? 0 : 1;
  const size_t a_bytes = (size_t)tc::BM2 * tc::BK2 * 4, b_bytes = (p.use_lo ? 2 : 1) * (size_t)npad * tc::BK2 * 4;
  const size_t budget = 227 * 1024 - 1024 - 512;
  p.rb = p.use_lo ? 4 : 6;
  p.ra = (int)((budget - p.rb * b_bytes) / a_bytes);
  if (p.ra > 8) p.ra = 8;
  p.ra &= ~1;
  ERA5SVD_REQUIRE(p.ra >= 2, "sketch_tf32x3: not enough shared memory");
  const size_t smem = p.ra * a_bytes + p.rb * b_bytes + 1024 + 512;
  ERA5SVD_CUDA(ensure_dynamic_smem((const void*)tc::sketch_tc2_kernel, smem));
  int64_t grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  tc::sketch_tc2_kernel<<<(unsigned)grid, tc::SK2_THREADS, smem, st>>>(tm_x, tm_ohi, tm_olo, p);
  return check_launch("sketch_tc2_kernel");
}


struct Pj2Plan {
  int nchunks;
  int64_t splits, rows_per_split;
  size_t bytes;
};

// One plan for the workspace query and the launch: it depends on (m, n, l) only (the column shift of an
// unaligned window adds at most one time chunk, which only matters for the wave rounding).
static Pj2Plan pj2_plan(int64_t m, int64_t n, int64_t l, int64_t rows_cap = 4096) {
  Pj2Plan pl;
  pl.nchunks = (int)ceil_div(n + 3, tc::PJ2_NC);
  const int sms = sm_count();
  // <= ~4096 rows per TMEM accumulation: the tensor core's fp32 accumulate truncates, and the resulting bias on
  // sigma grows with the rows summed on chip (measured on the c2 bench: 2e-6 at 4096, 2e-5 at 8192, 4e-5 at 16384);
  // the partial tiles are then added in float64.  This also holds for the two-product projection of the LAST power iteration
  // (era5svd_project_tf32x2): its Z feeds the Rayleigh-Ritz matrix T = Omega^T Z, whose trailing eigenvalues sit at
  // ~1e-7 of the leading one - with 16384 rows per accumulation (tried, round 2) the bias makes T numerically indefinite,
  // the positive-definite eigensolver drops a pivot and the two-sided fallback costs 2 ms per step.
  int64_t splits = ceil_div(m, rows_cap);
  int64_t ctas = ceil_div(splits * pl.nchunks, sms) * sms;      // whole waves
  splits = ctas / pl.nchunks;                                    // never more CTAs than whole waves
  if (splits < 1) splits = 1;
  const int64_t lp = ceil_div(l, (int64_t)16) * 16;            // column pitch of a partial tile (16-byte aligned rows)
  const int64_t cap = ((int64_t)512 << 20) / (n * lp * 4 > 0 ? n * lp * 4 : 1);
  if (splits > cap) splits = cap > 0 ? cap : 1;
  if (splits > 65535) splits = 65535;
  int64_t rps = ceil_div(ceil_div(m, splits), tc::PJ2_KS) * tc::PJ2_KS;
  pl.splits = ceil_div(m, rps);
  pl.rows_per_split = rps;
  pl.bytes = (size_t)(pl.splits * n * lp * 4);
  return pl;
}

size_t project_tf32x3_raw_workspace_bytes(int64_t m, int64_t n, int64_t l) { return pj2_plan(m, n, l).bytes; }

// Z (+)= X^T Y with X a plain float32 matrix (split on chip), Y given as (Yhi, Ylo).
int project_tf32x3_raw(const float* X, int64_t m, int64_t n, int64_t ldx, const float* Yhi, const float* Ylo,
                       int64_t l, int64_t ldy, double* Z, int64_t ldz, int accumulate, void* workspace,
                       size_t workspace_bytes, cudaStream_t st, int y_tf32) {
  CUtensorMap tm_x, tm_yhi, tm_ylo;
  int xs = 0, ys = 0, ys2 = 0, rc;
  if ((rc = tc::make_tmap(&tm_x, X, n, m, ldx, tc::BK2, tc::PJ2_KS, &xs, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = tc::make_tmap(&tm_yhi, Yhi, l, m, ldy, tc::BK2, tc::PJ2_KS, &ys, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  // Ylo == NULL: Yhi is the plain float32 Y, split on chip
  if ((rc = tc::make_tmap(&tm_ylo, Ylo ? Ylo : Yhi, l, m, ldy, tc::BK2, tc::PJ2_KS, &ys2, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  ERA5SVD_REQUIRE(ys == 0 && ys2 == 0, "project_tf32x3: Yhi / Ylo must be 16-byte aligned");
  const Pj2Plan pl = pj2_plan(m, n, l);
  if (!workspace || workspace_bytes < pl.bytes) {
    set_error("project_tf32x3: workspace too small (%zu < %zu)", workspace_bytes, pl.bytes);
    return ERA5SVD_ERR_WORKSPACE;
  }
  tc::Project2Params p;
  p.m = m; p.n = n; p.l = (int)l; p.lp = round_up2(l, 16);
  p.split_y = Ylo ? 0 : (y_tf32 ? 2 : 1);
  p.xshift = xs;
  p.rows_per_split = pl.rows_per_split;
  p.part = (float*)workspace;
  const size_t a_bytes = (size_t)(tc::PJ2_NC / tc::BK2) * tc::PJ2_KS * tc::BK2 * 4;   // 16 KB
  const size_t b_bytes = 2 * 4 * (size_t)tc::PJ2_KS * tc::BK2 * 4;                    // 16 KB
  p.rb = 4;
  p.ra = 8;
  const size_t smem = p.ra * a_bytes + p.rb * b_bytes + 1024 + 512;     // 512 B hold the (2 ra + 3 rb + 2 AT_RING + 2) barriers
  ERA5SVD_CUDA(ensure_dynamic_smem((const void*)tc::project_tc2_kernel, smem));
  dim3 grid((unsigned)ceil_div(n + xs, tc::PJ2_NC), (unsigned)pl.splits);
  tc::project_tc2_kernel<<<grid, tc::PJ2_THREADS, smem, st>>>(tm_x, tm_yhi, tm_ylo, p);
  if ((rc = check_launch("project_tc2_kernel"))) return rc;
  return launch_reduce_partials_f32(p.part, pl.splits, n, l, round_up2(l, 16), Z, ldz, accumulate, st);
}

}  // namespace era5svd
