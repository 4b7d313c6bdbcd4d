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
// Three decoupled rings, so that the HBM prefetch distance does not depend on MMA completion:
//   raw-A ring (smem, 16 KB/slot) : TMA  -> transform warps          freed as soon as it has been read
//   B ring     (smem, 28 KB/slot) : TMA  -> MMA                      freed by tcgen05.commit
//   A ring     (TMEM, 64 cols)    : transform -> MMA                 freed by tcgen05.commit
#include <cstdlib>
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
  int ra;          // raw-A ring depth
  int rb;          // B ring depth
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

// sketch, "merged N": two MMAs per k-step.
//   One tcgen05.mma of M = 128, K = 8 (tf32) costs ~105-116 clk for any N <= 192 and ~136 clk at N = 256
//   (profiles/r01_microbench_mma_probe.txt), so the two products that share A_hi are issued as ONE
//   instruction against the concatenated B tile (the hi and lo tiles of Om^T are adjacent in shared memory,
//   one K-major descriptor spans both):
//        D[:, 0:npad)      (+)= A_hi * Om_hi^T   (N = 2 npad)   and   (+)= A_lo * Om_hi^T   (N = npad)
//        D[:, npad:2npad)  (+)= A_hi * Om_lo^T
//   The epilogue adds the two halves.  TMEM: [0, 2 npad) one accumulator, [256, 512) A ring of 4 chunk slots
//   (32 hi + 32 lo columns); the deep A ring hides the MMA completion latency, which a double-buffered
//   accumulator with a 2-slot ring did not (1.06 ms vs 1.02 ms for three MMAs per k-step).
__global__ void __launch_bounds__(SK2_THREADS, 1)
sketch_tc3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_ohi,
                  const __grid_constant__ CUtensorMap tm_olo, const Sketch2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;   // warp-uniform
  const uint32_t a_bytes = BM2 * BK2 * 4;                   // raw A slot
  const uint32_t b_half = (uint32_t)p.npad * BK2 * 4;       // Om^T hi (or lo) tile, a multiple of 1024
  const uint32_t b_bytes = 2 * b_half;                      // B slot: [hi rows | lo rows]
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
  const uint32_t tempty = bar_base + 8u * nb;     nb += 1;
  const uint32_t tmem_slot = bar_base + 8u * nb;
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
    mbar_init(tfull, 1);
    mbar_init(tempty, 4);
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
          mbar_wait(araw_empty + 8u * ra.i, ra.ph ^ 1u);
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
          tma_load_2d(dst + b_half, &tm_olo, kc * BK2, 0, b_full + 8u * rb.i);
          rb.next();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: A (hi / lo) from TMEM, B = [Om^T hi | Om^T lo] from smem =====
    const uint32_t idesc_hl = make_idesc_tf32(BM2, 2 * p.npad, 0, 0);   // A_hi x [hi | lo]
    const uint32_t idesc_h = make_idesc_tf32(BM2, p.npad, 0, 0);        // A_lo x hi
    Ring rb(p.rb), at(AT_RING);
    uint32_t tph = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, tph ^= 1u) {
      mbar_wait(tempty, tph ^ 1u);              // the epilogue has drained the previous tile
      tcgen05_fence_after();
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
            const uint64_t bd = make_smem_desc(bs + (uint32_t)kk * UK2 * 4, 16, 1024);
            umma_tf32_ts(tmem_base, a_hi + kk * UK2, bd, idesc_hl, (kc | kk) != 0);
            umma_tf32_ts(tmem_base, a_lo + kk * UK2, bd, idesc_h, 1);
          }
          umma_commit(b_empty + 8u * rb.i);
          umma_commit(at_empty + 8u * at.i);
          if (kc == p.num_k - 1) umma_commit(tfull);
        }
        __syncwarp();
        rb.next();
        at.next();
      }
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
          mbar_wait(araw_full + 8u * ra.i, ra.ph);
          const uint32_t row_addr = smem_base + (uint32_t)ra.i * a_bytes + (uint32_t)r * 128u;
          float4 v[8];
#pragma unroll
          for (int c = 0; c < 8; ++c)                  // 16-byte chunk c of the 128-byte row, un-swizzled
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v[c].x), "=f"(v[c].y), "=f"(v[c].z), "=f"(v[c].w)
                         : "r"(row_addr + (uint32_t)((c ^ (r & 7)) << 4)));
          mbar_wait(at_empty + 8u * at.i, at.ph ^ 1u);           // MMAs that read this TMEM buffer have retired
          tcgen05_fence_after();
          const uint32_t t_hi = tmem_base + at_col0 + (uint32_t)at.i * AT_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              // truncation split (1 LOP + 1 FADD per element): hi = x with the low 13 mantissa bits
              // cleared (what the tensor core would read anyway), lo = x - hi exact, |lo| < 2^-10 |x|
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
            // The raw slot is released only here, after the loaded values have been consumed: an arrive placed
            // right after the ld.shared instructions can become visible before the loads have read shared
            // memory, and the refill by TMA then races with them (observed: low rows of a tile corrupted).
            mbar_arrive(araw_empty + 8u * ra.i);
            mbar_arrive(at_ready + 8u * at.i);
          }
        }
        ra.next();
        at.next();
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers (sum of the two halves) -> global (Y, Y_hi, Y_lo) =====
    const int q = warp % 4;
    uint32_t tph = 0;
    for (int64_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, tph ^= 1u) {
      mbar_wait(tfull, tph);
      tcgen05_fence_after();
      const int64_t row = tile * BM2 + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.npad; c0 += 32) {
        uint32_t v[32], w[32];
        tmem_ld16(taddr + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
        tmem_ld16(taddr + (uint32_t)(p.npad + c0), *reinterpret_cast<uint32_t(*)[16]>(&w[0]));
        const bool second = c0 + 16 < p.npad;
        if (second) {
          tmem_ld16(taddr + (uint32_t)(c0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
          tmem_ld16(taddr + (uint32_t)(p.npad + c0 + 16), *reinterpret_cast<uint32_t(*)[16]>(&w[16]));
        }
        tmem_wait_ld();
        if (c0 + 32 >= p.npad) {
          // last columns are in registers: release the accumulator before the global stores
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty);
        }
        if (row < p.m) {
          const int64_t off = row * p.ldy + c0;
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            if (g < 4 || second) {
              float4 y = make_float4(__uint_as_float(v[4 * g]) + __uint_as_float(w[4 * g]),
                                     __uint_as_float(v[4 * g + 1]) + __uint_as_float(w[4 * g + 1]),
                                     __uint_as_float(v[4 * g + 2]) + __uint_as_float(w[4 * g + 2]),
                                     __uint_as_float(v[4 * g + 3]) + __uint_as_float(w[4 * g + 3]));
              if (p.Y) *reinterpret_cast<float4*>(p.Y + off + 4 * g) = y;
              if (p.Yhi) {
                float4 hh = make_float4(tf32_hi(y.x), tf32_hi(y.y), tf32_hi(y.z), tf32_hi(y.w));
                float4 ll = make_float4(y.x - hh.x, y.y - hh.y, y.z - hh.z, y.w - hh.w);
                *reinterpret_cast<float4*>(p.Yhi + off + 4 * g) = hh;
                *reinterpret_cast<float4*>(p.Ylo + off + 4 * g) = ll;
              }
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}


struct Project2Params {
  int64_t m, n;
  int l;
  int xshift;
  int ra, rb;
  int64_t rows_per_split;
  float* part;        // [splits][n][l]
};

constexpr int PJ2_THREADS = 480;

// ---------------------------------------------------------------------------------------------
// project v3 ("merged N"):  Z[n x l] += X^T Y over a row split, X split on chip, two MMAs per k-step.
//   A = X^T tile in TMEM (lane = time within ONE 128-wide time tile, columns = rows of the slot, K-major)
//   B = [Y_hi | Y_lo] (pre-split by the sketch epilogue) from smem, N-major: the hi boxes and the lo boxes
//       are adjacent 32-column groups of one descriptor, so A_hi x [Y_hi | Y_lo] is ONE N = 256 instruction
//   D = [128 time lanes x 256]: columns [0,128) = A_hi Y_hi + A_lo Y_hi, [128,256) = A_hi Y_lo; the
//       epilogue adds the halves and writes a float32 partial tile.
// grid = (ceil(n / 128), splits); one slot = 32 rows x 128 time values (16 KB of X, 32 KB of Y hi/lo).
// TMEM: [0,256) accumulator, [256,512) A ring of 4 slots x (32 hi + 32 lo) columns.
// ---------------------------------------------------------------------------------------------
constexpr int PJ3_KS = 32;                 // rows per slot
constexpr int PJ3_NC = 128;                // time columns per CTA
constexpr int PJ3_AT_RING = 4;
constexpr uint32_t PJ3_AT_COLS = 2 * PJ3_KS;

__global__ void __launch_bounds__(PJ2_THREADS, 1)
project_tc3_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_yhi,
                   const __grid_constant__ CUtensorMap tm_ylo, const Project2Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x / 32, 0), lane = threadIdx.x % 32;   // warp-uniform
  const uint32_t box_bytes = PJ3_KS * BK2 * 4;                  // [32 rows x 32 floats] = 4 KB
  const uint32_t a_bytes = (PJ3_NC / BK2) * box_bytes;          // 4 boxes = 16 KB
  const uint32_t y_half = 4 * box_bytes;                        // 128 sketch columns, hi (or lo)
  const uint32_t b_bytes = 2 * y_half;                          // 32 KB
  const uint32_t b_base = smem_base + (uint32_t)p.ra * a_bytes;
  const uint32_t bar_base = b_base + (uint32_t)p.rb * b_bytes;
  int nb = 0;
  const uint32_t araw_full = bar_base + 8u * nb;  nb += p.ra;
  const uint32_t araw_empty = bar_base + 8u * nb; nb += p.ra;
  const uint32_t b_full = bar_base + 8u * nb;     nb += p.rb;
  const uint32_t b_empty = bar_base + 8u * nb;    nb += p.rb;
  const uint32_t at_ready = bar_base + 8u * nb;   nb += PJ3_AT_RING;
  const uint32_t at_empty = bar_base + 8u * nb;   nb += PJ3_AT_RING;
  const uint32_t tfull = bar_base + 8u * nb;      nb += 1;
  const uint32_t tmem_slot = bar_base + 8u * nb;
  const uint32_t at_col0 = 256;

  const int64_t t0 = (int64_t)blockIdx.x * PJ3_NC;
  const int64_t r_begin = (int64_t)blockIdx.y * p.rows_per_split;
  const int64_t r_end = min(p.m, r_begin + p.rows_per_split);
  const int num_k = (int)((r_end - r_begin + PJ3_KS - 1) / PJ3_KS);

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ra; ++i) {
      mbar_init(araw_full + 8u * i, 1);
      mbar_init(araw_empty + 8u * i, 4);      // the four transform warps of the group that owns the chunk
    }
    for (int i = 0; i < p.rb; ++i) {
      mbar_init(b_full + 8u * i, 1);
      mbar_init(b_empty + 8u * i, 1);
    }
    for (int i = 0; i < PJ3_AT_RING; ++i) {
      mbar_init(at_ready + 8u * i, 4);
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
    // ===== TMA producer: raw X slots (HBM stream), 4 boxes of [32 rows x 32 time] =====
    if (elect_one()) {
      Ring ra(p.ra);
      for (int kc = 0; kc < num_k; ++kc) {
        mbar_wait(araw_empty + 8u * ra.i, ra.ph ^ 1u);
        const uint32_t dst = smem_base + (uint32_t)ra.i * a_bytes;
        const int32_t row0 = (int32_t)(r_begin + (int64_t)kc * PJ3_KS);
        mbar_arrive_expect_tx(araw_full + 8u * ra.i, a_bytes);
        for (int c = 0; c < PJ3_NC / BK2; ++c)
          tma_load_2d(dst + c * box_bytes, &tm_x, (int32_t)(t0 + c * BK2), row0, araw_full + 8u * ra.i);
        ra.next();
      }
    }
  } else if (warp == SK2_BPROD_WARP) {
    // ===== TMA producer: Y hi / lo slots, 4 + 4 boxes of [32 rows x 32 columns] =====
    if (elect_one()) {
      Ring rb(p.rb);
      for (int kc = 0; kc < num_k; ++kc) {
        mbar_wait(b_empty + 8u * rb.i, rb.ph ^ 1u);
        const uint32_t dst = b_base + (uint32_t)rb.i * b_bytes;
        const int32_t row0 = (int32_t)(r_begin + (int64_t)kc * PJ3_KS);
        mbar_arrive_expect_tx(b_full + 8u * rb.i, b_bytes);
        for (int c = 0; c < 4; ++c) {
          tma_load_2d(dst + c * box_bytes, &tm_yhi, c * BK2, row0, b_full + 8u * rb.i);
          tma_load_2d(dst + y_half + c * box_bytes, &tm_ylo, c * BK2, row0, b_full + 8u * rb.i);
        }
        rb.next();
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc_hl = make_idesc_tf32(128, 256, 0, 1);   // A K-major (TMEM), B MN-major, [Y_hi | Y_lo]
    const uint32_t idesc_h = make_idesc_tf32(128, 128, 0, 1);    // A_lo x Y_hi
    Ring rb(p.rb), at(PJ3_AT_RING);
    for (int kc = 0; kc < num_k; ++kc) {
      mbar_wait(b_full + 8u * rb.i, rb.ph);
      mbar_wait(at_ready + 8u * at.i, at.ph);
      tcgen05_fence_after();
      if (elect_one()) {
        const uint32_t bs = b_base + (uint32_t)rb.i * b_bytes;
        const uint32_t a_hi = tmem_base + at_col0 + (uint32_t)at.i * PJ3_AT_COLS;
        const uint32_t a_lo = a_hi + PJ3_KS;
#pragma unroll
        for (int ks = 0; ks < PJ3_KS / UK2; ++ks) {
          // 8 K rows = two 4-row swizzle atoms = 1024 B inside every 32-column box
          const uint64_t bd = make_smem_desc(bs + (uint32_t)ks * UK2 * 128, box_bytes, 512, LAYOUT_SW128_BASE32B);
          umma_tf32_ts(tmem_base, a_hi + ks * UK2, bd, idesc_hl, (kc | ks) != 0);
          umma_tf32_ts(tmem_base, a_lo + ks * UK2, bd, idesc_h, 1);
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
    const int grp = (warp - 2) / 4;                      // this group handles slots with (kc & 1) == grp
    const int tl = q * 32 + lane;                        // time index within the tile == TMEM lane
    const int box = tl / BK2;                            // which [32 x 32] box holds this time column
    const int col = tl % BK2;                            // float index inside the 128-byte box row
    Ring ra(p.ra), at(PJ3_AT_RING);
    for (int kc = 0; kc < num_k; ++kc) {
      if ((kc & 1) == grp) {
        mbar_wait(araw_full + 8u * ra.i, ra.ph);
        const uint32_t base = smem_base + (uint32_t)ra.i * a_bytes + (uint32_t)box * box_bytes;
        float x[PJ3_KS];
#pragma unroll
        for (int k = 0; k < PJ3_KS; ++k) {               // row k of the box: 128 B, 16-byte chunks XOR-swizzled by (k & 7)
          const uint32_t addr = base + (uint32_t)k * 128u + (uint32_t)((((col >> 2) ^ (k & 7)) << 4) | ((col & 3) << 2));
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[k]) : "r"(addr));
        }
        mbar_wait(at_empty + 8u * at.i, at.ph ^ 1u);
        tcgen05_fence_after();
        const uint32_t t_hi = tmem_base + at_col0 + (uint32_t)at.i * PJ3_AT_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float v = x[half * 16 + k];
            const float h = tf32_trunc(v);
            hi[k] = __float_as_uint(h);
            lo[k] = __float_as_uint(v - h);
          }
          tmem_st16(t_hi + half * 16, hi);
          tmem_st16(t_hi + PJ3_KS + half * 16, lo);
        }
        tmem_wait_st();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(araw_empty + 8u * ra.i);      // released after the loaded values were consumed (see sketch)
          mbar_arrive(at_ready + 8u * at.i);
        }
      }
      ra.next();
      at.next();
    }
  } else {
    // ===== epilogue: accumulator halves summed -> float32 partial tile part[split][time][column] =====
    const int q = warp % 4;
    float* out = p.part + (int64_t)blockIdx.y * p.n * p.l;
    if (num_k > 0) {
      mbar_wait(tfull, 0);
      tcgen05_fence_after();
    }
    const int64_t t = t0 + q * 32 + lane - p.xshift;      // window-relative time index of this lane
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < p.l; c0 += 16) {
      uint32_t v[16], w[16];
      if (num_k > 0) {
        tmem_ld16(taddr + (uint32_t)c0, v);
        tmem_ld16(taddr + 128u + (uint32_t)c0, w);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { v[j] = 0u; w[j] = 0u; }
      }
      if (t >= 0 && t < p.n) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < p.l) out[t * p.l + c0 + j] = __uint_as_float(v[j]) + __uint_as_float(w[j]);
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
  // shared memory: B ring of 4 slots (L2 latency), the rest for the raw-A ring (HBM latency), up to 8 slots
  const size_t a_bytes = (size_t)tc::BM2 * tc::BK2 * 4, b_bytes = 2 * (size_t)npad * tc::BK2 * 4;
  const size_t budget = 227 * 1024 - 1024 - 512;
  p.rb = 4;
  p.ra = (int)((budget - p.rb * b_bytes) / a_bytes);
  if (p.ra > 8) p.ra = 8;
  ERA5SVD_REQUIRE(p.ra >= 2, "sketch_tf32x3: not enough shared memory");
  const size_t smem = p.ra * a_bytes + p.rb * b_bytes + 1024 + 512;
  ERA5SVD_CUDA(cudaFuncSetAttribute(tc::sketch_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
  tc::sketch_tc3_kernel<<<(unsigned)grid, tc::SK2_THREADS, smem, st>>>(tm_x, tm_ohi, tm_olo, p);
  return check_launch("sketch_tc3_kernel");
}


struct Pj2Plan {
  int nchunks;
  int64_t splits, rows_per_split;
  size_t bytes;
};

// One plan for the workspace query and the launch: it depends on (m, n, l) only (the column shift of an
// unaligned window adds at most one time tile, which only matters for the wave rounding).
static Pj2Plan pj2_plan(int64_t m, int64_t n, int64_t l) {
  Pj2Plan pl;
  pl.nchunks = (int)ceil_div(n + 3, tc::PJ3_NC);
  const int sms = sm_count();
  int64_t splits = ceil_div(m, 8192);
  int64_t ctas = ceil_div(splits * pl.nchunks, sms) * sms;      // whole waves
  splits = ctas / pl.nchunks;                                    // never more CTAs than whole waves
  if (splits < 1) splits = 1;
  const int64_t cap = ((int64_t)512 << 20) / (n * l * 4 > 0 ? n * l * 4 : 1);
  if (splits > cap) splits = cap > 0 ? cap : 1;
  if (splits > 65535) splits = 65535;
  int64_t rps = ceil_div(ceil_div(m, splits), tc::PJ3_KS) * tc::PJ3_KS;
  pl.splits = ceil_div(m, rps);
  pl.rows_per_split = rps;
  pl.bytes = (size_t)(pl.splits * n * l * 4);
  return pl;
}

size_t project_tf32x3_raw_workspace_bytes(int64_t m, int64_t n, int64_t l) { return pj2_plan(m, n, l).bytes; }

// Z (+)= X^T Y with X a plain float32 matrix (split on chip), Y given as (Yhi, Ylo).
int project_tf32x3_raw(const float* X, int64_t m, int64_t n, int64_t ldx, const float* Yhi, const float* Ylo,
                       int64_t l, int64_t ldy, double* Z, int64_t ldz, int accumulate, void* workspace,
                       size_t workspace_bytes, cudaStream_t st) {
  CUtensorMap tm_x, tm_yhi, tm_ylo;
  int xs = 0, ys = 0, ys2 = 0, rc;
  if ((rc = tc::make_tmap(&tm_x, X, n, m, ldx, tc::BK2, tc::PJ3_KS, &xs, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = tc::make_tmap(&tm_yhi, Yhi, l, m, ldy, tc::BK2, tc::PJ3_KS, &ys, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = tc::make_tmap(&tm_ylo, Ylo, l, m, ldy, tc::BK2, tc::PJ3_KS, &ys2, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  ERA5SVD_REQUIRE(ys == 0 && ys2 == 0, "project_tf32x3: Yhi / Ylo must be 16-byte aligned");
  const Pj2Plan pl = pj2_plan(m, n, l);
  if (!workspace || workspace_bytes < pl.bytes) {
    set_error("project_tf32x3: workspace too small (%zu < %zu)", workspace_bytes, pl.bytes);
    return ERA5SVD_ERR_WORKSPACE;
  }
  tc::Project2Params p;
  p.m = m; p.n = n; p.l = (int)l;
  p.xshift = xs;
  p.rows_per_split = pl.rows_per_split;
  p.part = (float*)workspace;
  const size_t a_bytes = (size_t)(tc::PJ3_NC / tc::BK2) * tc::PJ3_KS * tc::BK2 * 4;   // 16 KB
  const size_t b_bytes = 2 * 4 * (size_t)tc::PJ3_KS * tc::BK2 * 4;                    // 32 KB
  p.rb = 4;
  p.ra = 5;
  const size_t smem = p.ra * a_bytes + p.rb * b_bytes + 1024 + 512;
  ERA5SVD_CUDA(cudaFuncSetAttribute(tc::project_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div(n + xs, tc::PJ3_NC), (unsigned)pl.splits);
  tc::project_tc3_kernel<<<grid, tc::PJ2_THREADS, smem, st>>>(tm_x, tm_yhi, tm_ylo, p);
  if ((rc = check_launch("project_tc3_kernel"))) return rc;
  launch_reduce_partials_f32(p.part, pl.splits, n, l, Z, ldz, accumulate, st);
  return check_launch("reduce_partials_kernel");
}

}  // namespace era5svd
