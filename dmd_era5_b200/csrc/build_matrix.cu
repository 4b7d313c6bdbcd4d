// (a) Matrix build: native ERA5 layout (time, point) -> snapshot matrix rows (point, time),
// fused with time-mean removal, optional unit-variance scaling, optional area weights, cast.
//
// Reference semantics restated (not copied):
//   standardize_data            src/dmd_era5/slice_tools/slice_tools.py:171-177
//   flatten_era5_variables      src/dmd_era5/slice_tools/slice_tools.py:323-336
//   check_array finiteness      sklearn/utils/extmath.py:546
//
// HBM-bound.  Two kernels per (variable, level) block:
//   row_stats_kernel  : one thread per grid point walks the T snapshots (a warp reads 32
//                       consecutive points = one 128 B line per snapshot, fully coalesced);
//                       float64 accumulation, NaN skipping.
//   transpose_kernel  : 32x32 shared-memory tile transpose; reads coalesced along points,
//                       writes coalesced along time, applies (x - mean) / std * w and the cast.
// Algorithmic bytes per element: read 4 (stats) [+4 second stats pass when scaling] + read 4 +
// write 4 (f32).  See DESIGN.md for the single-read fused variant.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace era5svd {

// Centre in the wider of (source, matrix) types, then round to the matrix type.  With equal types
// (the reference's behaviour) this is exactly `x - mean` in that type.
template <typename Ts, typename Tx>
__device__ __forceinline__ Tx centre(Ts raw, Tx mean) {
  if (sizeof(Ts) > sizeof(Tx)) return (Tx)((double)raw - (double)mean);
  return (Tx)raw - mean;
}

template <typename Ts, typename Tx>
__global__ void __launch_bounds__(128)
row_stats_kernel(const Ts* __restrict__ src, int64_t T, int64_t src_ld, int64_t P,
                 Tx* __restrict__ mean_out, Tx* __restrict__ std_out, int do_scale) {
  int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const Ts* col = src + p;
  // pass 1: NaN-skipping mean (xarray's mean(dim) default), float64 accumulation.
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int64_t cnt = 0;
  int64_t t = 0;
  for (; t + 4 <= T; t += 4) {
    double v0 = (double)col[(t + 0) * src_ld];
    double v1 = (double)col[(t + 1) * src_ld];
    double v2 = (double)col[(t + 2) * src_ld];
    double v3 = (double)col[(t + 3) * src_ld];
    if (v0 == v0) { s0 += v0; ++cnt; }
    if (v1 == v1) { s1 += v1; ++cnt; }
    if (v2 == v2) { s2 += v2; ++cnt; }
    if (v3 == v3) { s3 += v3; ++cnt; }
  }
  for (; t < T; ++t) {
    double v = (double)col[t * src_ld];
    if (v == v) { s0 += v; ++cnt; }
  }
  double mean = cnt > 0 ? ((s0 + s1) + (s2 + s3)) / (double)cnt : __longlong_as_double(0x7ff8000000000000LL);
  Tx mean_x = (Tx)mean;
  mean_out[p] = mean_x;
  if (!do_scale) return;
  // pass 2: std (ddof = 0) of the CENTRED data, as the reference computes it: the centred values
  // are formed in the matrix dtype (x - mean), their own (tiny) mean is removed again by nanstd.
  double a = 0.0, q = 0.0;
  for (t = 0; t < T; ++t) {
    Ts raw = col[t * src_ld];
    if (raw == raw) {
      double xc = (double)centre<Ts, Tx>(raw, mean_x);
      a += xc;
      q += xc * xc;
    }
  }
  double m2 = cnt > 0 ? a / (double)cnt : 0.0;
  double var = cnt > 0 ? q / (double)cnt - m2 * m2 : __longlong_as_double(0x7ff8000000000000LL);
  if (var < 0.0) var = 0.0;
  std_out[p] = (Tx)sqrt(var);
}

template <typename Ts, typename Tx>
__global__ void __launch_bounds__(256)
transpose_kernel(const Ts* __restrict__ src, int64_t T, int64_t src_ld, int64_t P,
                 Tx* __restrict__ X, int64_t ldx, const Tx* __restrict__ mean,
                 const Tx* __restrict__ stdv, const Tx* __restrict__ weights,
                 int check_finite, int* __restrict__ nonfinite_flag,
                 float* __restrict__ Xhi, float* __restrict__ Xlo) {
  __shared__ Ts tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int64_t t0 = (int64_t)blockIdx.y * 32;
  int bad = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t t = t0 + ty + 8 * i, p = p0 + tx;
    Ts v = Ts(0);
    if (t < T && p < P) {
      v = src[t * src_ld + p];
    }
    tile[ty + 8 * i][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t p = p0 + ty + 8 * i, t = t0 + tx;
    if (p < P && t < T) {
      Ts raw = tile[tx][ty + 8 * i];
      Tx v = mean ? centre<Ts, Tx>(raw, mean[p]) : (Tx)raw;
      if (stdv) v = v / stdv[p];
      if (weights) v = v * weights[p];
      // the flag reports what is WRITTEN: a NaN / Inf source value, but also 0 / 0 from a time-constant point under
      // scale=True (std = 0; the reference then raises "Input contains NaN" in sklearn's check_array)
      if (check_finite && !isfinite((double)v)) bad = 1;
      if (X) X[p * ldx + t] = v;
      if (Xhi) {   // tf32 hi / lo images for the tensor-core passes (float32 matrices only)
        const float h = tc::tf32_hi((float)v);
        Xhi[p * ldx + t] = h;
        Xlo[p * ldx + t] = (float)v - h;
      }
    }
  }
  if (check_finite) {
    int any = __syncthreads_or(bad);
    if (any && tx == 0 && ty == 0) atomicExch(nonfinite_flag, 1);
  }
}

// Single-read variant: one CTA keeps a [T x PB] tile of the native array (PB = 32 points, every
// snapshot) in shared memory, computes the per-point statistics from it and writes the PB finished
// rows of X.  HBM traffic = read m*n*b + write m*n*b: the algorithmic minimum of the build.
//   load : warp w reads snapshots w, w + 8, ... ; a warp-load is 32 consecutive points (128 B for fp32)
//   stats: warp w owns points w, w + 8, ... ; lanes stride over time, float64 warp reduction
//   store: same ownership; a warp writes one row of X with lanes along time (fully coalesced)
// The tile row pitch PB + 1 keeps both the column walks (stats / store) and the row fills conflict free.
constexpr int FB_PB = 32;
constexpr int FB_THREADS = 512;
constexpr int FB_UNROLL = 8;     // independent 128-byte loads in flight per warp

template <typename Ts, typename Tx>
__global__ void __launch_bounds__(FB_THREADS)
fused_build_kernel(const Ts* __restrict__ src, int64_t T, int64_t src_ld, int64_t P,
                   Tx* __restrict__ X, int64_t ldx, Tx* __restrict__ mean_out, Tx* __restrict__ std_out,
                   const Tx* __restrict__ weights, int center, int do_scale, int check_finite,
                   int* __restrict__ nonfinite_flag, float* __restrict__ Xhi, float* __restrict__ Xlo) {
  extern __shared__ unsigned char fb_smem[];
  Ts* tile = reinterpret_cast<Ts*>(fb_smem);                  // [T][PB + 1]
  constexpr int LD = FB_PB + 1;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32, nwarps = FB_THREADS / 32;
  const int64_t p0 = (int64_t)blockIdx.x * FB_PB;
  const int64_t p = p0 + lane;
  int bad = 0;
  for (int64_t tb = warp; tb < T; tb += (int64_t)nwarps * FB_UNROLL) {
    Ts v[FB_UNROLL];
#pragma unroll
    for (int i = 0; i < FB_UNROLL; ++i) {
      const int64_t t = tb + (int64_t)i * nwarps;
      v[i] = (t < T && p < P) ? src[t * src_ld + p] : Ts(0);
    }
#pragma unroll
    for (int i = 0; i < FB_UNROLL; ++i) {
      const int64_t t = tb + (int64_t)i * nwarps;
      if (t < T) tile[t * LD + lane] = v[i];
    }
  }
  __syncthreads();
  for (int pp = warp; pp < FB_PB; pp += nwarps) {
    const int64_t gp = p0 + pp;
    if (gp >= P) break;
    Tx mean_x = Tx(0), std_x = Tx(1);
    if (center) {
      double s = 0.0;
      int cnt = 0;
      for (int64_t t = lane; t < T; t += 32) {
        const double v = (double)tile[t * LD + pp];
        if (v == v) { s += v; ++cnt; }
      }
      s = warp_sum(s);
      cnt = warp_sum(cnt);
      const double mean = cnt > 0 ? s / (double)cnt : __longlong_as_double(0x7ff8000000000000LL);
      mean_x = (Tx)mean;
      if (do_scale) {
        double a = 0.0, q = 0.0;
        for (int64_t t = lane; t < T; t += 32) {
          const Ts raw = tile[t * LD + pp];
          if (raw == raw) {
            const double xc = (double)centre<Ts, Tx>(raw, mean_x);
            a += xc;
            q += xc * xc;
          }
        }
        a = warp_sum(a);
        q = warp_sum(q);
        const double m2 = cnt > 0 ? a / (double)cnt : 0.0;
        double var = cnt > 0 ? q / (double)cnt - m2 * m2 : __longlong_as_double(0x7ff8000000000000LL);
        if (var < 0.0) var = 0.0;
        std_x = (Tx)sqrt(var);
      }
      if (lane == 0) {
        mean_out[gp] = mean_x;
        if (do_scale) std_out[gp] = std_x;
      }
    }
    const Tx w = weights ? weights[gp] : Tx(1);
    for (int64_t t = lane; t < T; t += 32) {
      const Ts raw = tile[t * LD + pp];
      Tx v = center ? centre<Ts, Tx>(raw, mean_x) : (Tx)raw;
      if (do_scale) v = v / std_x;
      if (weights) v = v * w;
      if (check_finite && !isfinite((double)v)) bad = 1;        // the value written (see transpose_kernel)
      if (X) X[gp * ldx + t] = v;
      if (Xhi) {
        const float h = tc::tf32_hi((float)v);
        Xhi[gp * ldx + t] = h;
        Xlo[gp * ldx + t] = (float)v - h;
      }
    }
  }
  if (check_finite) {
    int any = __syncthreads_or(bad);
    if (any && threadIdx.x == 0) atomicExch(nonfinite_flag, 1);
  }
}

// TMA variant of the single-read build for float32 sources (the real ERA5 dtype).  The register-staged loads of
// fused_build_kernel keep only 16 x 8 x 128 B = 16 KB per CTA in flight, which is what bounds it (c2: 3.2 TB/s,
// c3 with one resident CTA: 2.0 TB/s).  Here ONE thread issues the whole [T x 32 points] tile as 2-D TMA boxes
// (32 points x 64 snapshots, 8 KB each), so the complete tile (95 KB at T = 744, 187 KB at T = 1460) is in flight
// at once and no registers are spent on staging.
//   smem tile : [Tpad][32] float32, 128-byte rows, TMA SWIZZLE_128B: 16-byte chunk c of row t sits at chunk
//               c ^ (t & 7), so a quarter-warp reading chunk c of 8 consecutive snapshots hits 8 distinct
//               16-byte bank groups -> the transposing column reads are conflict free WITHOUT padding.
//   warp w    : points 4 (w % 8) .. + 3 (one 16-byte chunk), time half w / 8; lane = snapshot.
//   stats     : float64 partial sums per (warp, point), combined across the two time halves through smem.
//   store     : four 128-byte coalesced row stores per warp iteration (one per point of the chunk).
constexpr int TB_THREADS = 512;
constexpr int TB_BOX_T = 64;

// PB = 32: 128-byte rows under SWIZZLE_128B (chunk c of row t at c ^ (t & 7)); PB = 16: 64-byte rows under SWIZZLE_64B
// (address bits [4,6) XOR bits [7,9): chunk c of row t at c ^ ((t >> 1) & 3)).  Either way a quarter-warp reading one
// chunk of eight consecutive snapshots touches eight distinct 16-byte bank groups.
template <int PB>
__device__ __forceinline__ float4 tb_ld_chunk(uint32_t tile, int t, int c) {
  float4 v;
  const uint32_t sw = PB == 32 ? (uint32_t)(c ^ (t & 7)) : (uint32_t)(c ^ ((t >> 1) & 3));
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(tile + (uint32_t)t * (uint32_t)(PB * 4) + (sw << 4)));
  return v;
}

// PB = points per tile: 32 (128-byte source segments) while two [T x 32] tiles fit one SM; 16 (64-byte segments) for
// longer series (T = 1460: 94 KB per tile), so that TWO CTAs stay resident and one CTA's loads overlap the other's
// statistics / stores - with a single resident CTA the load and store phases alternate and HBM idles half the time.
// Warp w: chunk w % (PB / 4), time part w / (PB / 4) (2 parts for PB = 32, 4 for PB = 16).
// thread-block-cluster helpers (the CL > 1 instantiations: a long series is split over CL CTAs along time)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_peer_f64(const double* local, uint32_t rank) {
  uint32_t a;
  double v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(tc::smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ int ld_peer_s32(const int* local, uint32_t rank) {
  uint32_t a;
  int v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(tc::smem_u32(local)), "r"(rank));
  asm volatile("ld.shared::cluster.s32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

// CL = CTAs per point tile (thread-block cluster along TIME).  A series too long for two resident [T x 32] tiles
// (T > ~850) used to fall back to 16-point tiles: 64-byte source segments, one DRAM activate per 64 bytes, 0.59 of
// the copy rate at T = 1460.  With CL = 2 (4) each CTA holds [T / CL x 32 points] (128-byte segments, two resident
// CTAs as at T = 744), the per-point sums cross the cluster through distributed shared memory (32 doubles per CTA)
// and every CTA writes its own time range of the 32 rows (>= 2.9 KB contiguous per row at T = 1460).
template <typename Tx, int PB, int CL>
__global__ void __launch_bounds__(TB_THREADS)
fused_build_tma_kernel(const __grid_constant__ CUtensorMap tm_src, int T, int64_t P, int col_shift,
                       Tx* __restrict__ X, int64_t ldx, Tx* __restrict__ mean_out, Tx* __restrict__ std_out,
                       const Tx* __restrict__ weights, int center, int do_scale, int check_finite,
                       int* __restrict__ nonfinite_flag, float* __restrict__ Xhi, float* __restrict__ Xlo) {
  extern __shared__ unsigned char fb_smem[];
  constexpr int NCH = PB / 4, NP = (TB_THREADS / 32) / NCH, ROWB = PB * 4;
  __shared__ double part_a[NP][PB], part_q[NP][PB];
  __shared__ int part_n[NP][PB];
  __shared__ __align__(8) uint64_t bar_storage;
  const uint32_t tile = (tc::smem_u32(fb_smem) + 1023u) & ~1023u;
  const uint32_t bar = tc::smem_u32(&bar_storage);
  __shared__ double cl_a[PB], cl_q[PB];     // this CTA's per-point sums, read by its cluster peers (CL > 1)
  __shared__ int cl_n[PB];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int c = warp % NCH, h = warp / NCH;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  const int64_t p0 = (int64_t)(blockIdx.x / CL) * PB;
  // snapshots [tg0, tg0 + Tl) of the series belong to this CTA (whole 64-snapshot boxes per CTA); tile row = t - tg0
  const int Tc = CL > 1 ? ((T + CL - 1) / CL + TB_BOX_T - 1) / TB_BOX_T * TB_BOX_T : T;
  const int tg0 = (int)crank * Tc;
  const int Tl = tg0 >= T ? 0 : (T - tg0 < Tc ? T - tg0 : Tc);
  const int nbox = (Tl + TB_BOX_T - 1) / TB_BOX_T;
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::fence_barrier_init();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc::mbar_arrive_expect_tx(bar, (uint32_t)nbox * (uint32_t)(TB_BOX_T * ROWB));
    for (int b = 0; b < nbox; ++b)
      tc::tma_load_2d(tile + (uint32_t)b * (uint32_t)(TB_BOX_T * ROWB), &tm_src, (int32_t)p0, tg0 + b * TB_BOX_T, bar);
  }
  __syncthreads();                 // barrier initialised before anyone polls it
  // time range of this warp: NP parts split on multiples of 32 snapshots (tile rows, i.e. relative to tg0)
  const int t_part = ((Tl + 32 * NP - 1) / (32 * NP)) * 32;
  const int t_begin = h * t_part < Tl ? h * t_part : Tl;
  const int t_end = (h + 1) * t_part < Tl ? (h + 1) * t_part : Tl;
  // a source whose base is not 16-byte aligned is mapped from the aligned address below it (boxes under the 128-byte
  // swizzle must start on 16-byte boundaries): tile column x holds point p0 + x - col_shift
  const int64_t gp0 = p0 + 4 * c - col_shift;
  bool valid[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) valid[e] = gp0 + e >= 0 && gp0 + e < P;
  Tx w4[4] = {Tx(1), Tx(1), Tx(1), Tx(1)};
  if (weights)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      if (valid[e]) w4[e] = weights[gp0 + e];
  tc::mbar_wait(bar, 0);

  int bad = 0;
  Tx mean_x[4] = {Tx(0), Tx(0), Tx(0), Tx(0)}, std_x[4] = {Tx(1), Tx(1), Tx(1), Tx(1)};
  if (center) {
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    int cnt[4] = {0, 0, 0, 0};
    for (int t = t_begin + lane; t < t_end; t += 32) {
      const float4 v4 = tb_ld_chunk<PB>(tile, t, c);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (v[e] == v[e]) { s[e] += (double)v[e]; ++cnt[e]; }
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s[e] = warp_sum(s[e]);
      cnt[e] = warp_sum(cnt[e]);
    }
    if (lane == 0)
#pragma unroll
      for (int e = 0; e < 4; ++e) { part_a[h][4 * c + e] = s[e]; part_n[h][4 * c + e] = cnt[e]; }
    __syncthreads();
    if constexpr (CL > 1) {
      // this CTA's totals -> cl_a / cl_n, visible to the peers after the cluster barrier
      if (threadIdx.x < PB) {
        double tot = 0.0;
        int nv = 0;
#pragma unroll
        for (int hh = 0; hh < NP; ++hh) { nv += part_n[hh][threadIdx.x]; tot += part_a[hh][threadIdx.x]; }
        cl_a[threadIdx.x] = tot;
        cl_n[threadIdx.x] = nv;
      }
      cluster_sync_all();
    }
    int n_valid[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      n_valid[e] = 0;
      double tot = 0.0;
      if constexpr (CL > 1) {
        for (uint32_t rk = 0; rk < (uint32_t)CL; ++rk) {       // fixed rank order: every CTA forms the same mean
          tot += ld_peer_f64(&cl_a[4 * c + e], rk);
          n_valid[e] += ld_peer_s32(&cl_n[4 * c + e], rk);
        }
      } else {
#pragma unroll
        for (int hh = 0; hh < NP; ++hh) { n_valid[e] += part_n[hh][4 * c + e]; tot += part_a[hh][4 * c + e]; }
      }
      const double mean = n_valid[e] > 0 ? tot / (double)n_valid[e] : __longlong_as_double(0x7ff8000000000000LL);
      mean_x[e] = (Tx)mean;
    }
    if (do_scale) {
      __syncthreads();             // part_a is reused below
      double a[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
      for (int t = t_begin + lane; t < t_end; t += 32) {
        const float4 v4 = tb_ld_chunk<PB>(tile, t, c);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (v[e] == v[e]) {
            const double xc = (double)centre<float, Tx>(v[e], mean_x[e]);
            a[e] += xc;
            q[e] += xc * xc;
          }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        a[e] = warp_sum(a[e]);
        q[e] = warp_sum(q[e]);
      }
      if (lane == 0)
#pragma unroll
        for (int e = 0; e < 4; ++e) { part_a[h][4 * c + e] = a[e]; part_q[h][4 * c + e] = q[e]; }
      __syncthreads();
      if constexpr (CL > 1) {
        cluster_sync_all();        // every peer has finished reading cl_a of the mean pass
        if (threadIdx.x < PB) {
          double at = 0.0, qt = 0.0;
#pragma unroll
          for (int hh = 0; hh < NP; ++hh) { at += part_a[hh][threadIdx.x]; qt += part_q[hh][threadIdx.x]; }
          cl_a[threadIdx.x] = at;
          cl_q[threadIdx.x] = qt;
        }
        cluster_sync_all();
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        double at = 0.0, qt = 0.0;
        if constexpr (CL > 1) {
          for (uint32_t rk = 0; rk < (uint32_t)CL; ++rk) {
            at += ld_peer_f64(&cl_a[4 * c + e], rk);
            qt += ld_peer_f64(&cl_q[4 * c + e], rk);
          }
        } else {
#pragma unroll
          for (int hh = 0; hh < NP; ++hh) { at += part_a[hh][4 * c + e]; qt += part_q[hh][4 * c + e]; }
        }
        const double m2 = n_valid[e] > 0 ? at / (double)n_valid[e] : 0.0;
        double var = n_valid[e] > 0 ? qt / (double)n_valid[e] - m2 * m2 : __longlong_as_double(0x7ff8000000000000LL);
        if (var < 0.0) var = 0.0;
        std_x[e] = (Tx)sqrt(var);
      }
    }
    if (crank == 0 && h == 0 && lane < 4 && gp0 + lane >= 0 && gp0 + lane < P) {
      // lane e publishes point e (static register indexing kept by the unrolled select)
      Tx mv = mean_x[0], sv = std_x[0];
#pragma unroll
      for (int e = 1; e < 4; ++e)
        if (lane == e) { mv = mean_x[e]; sv = std_x[e]; }
      mean_out[gp0 + lane] = mv;
      if (do_scale) std_out[gp0 + lane] = sv;
    }
  }
  for (int t = t_begin + lane; t < t_end; t += 32) {
    const float4 v4 = tb_ld_chunk<PB>(tile, t, c);
    const float raw[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (valid[e]) {
        Tx v = center ? centre<float, Tx>(raw[e], mean_x[e]) : (Tx)raw[e];
        if (do_scale) v = v / std_x[e];
        if (weights) v = v * w4[e];
        if (check_finite && !isfinite((double)v)) bad = 1;      // the value written (see transpose_kernel)
        const int64_t off = (gp0 + e) * ldx + tg0 + t;
        if (X) X[off] = v;
        if (Xhi) {
          const float hh = tc::tf32_hi((float)v);
          Xhi[off] = hh;
          Xlo[off] = (float)v - hh;
        }
      }
    }
  }
  if (check_finite) {
    int any = __syncthreads_or(bad);
    if (any && threadIdx.x == 0) atomicExch(nonfinite_flag, 1);
  }
  if constexpr (CL > 1) {
    if (center) cluster_sync_all();      // a CTA's shared memory must outlive its peers' reads of cl_a / cl_q
  }
}

namespace tc {
int make_tmap(CUtensorMap* map, const float* base, int64_t inner, int64_t outer, int64_t ld,
              uint32_t box_inner, uint32_t box_outer, int* col_shift, CUtensorMapSwizzle swizzle);
}

// float32 source, tile fits in shared memory, 16-byte row pitch: the TMA kernel.  Returns 1 when it ran,
// 0 when the shape does not qualify (caller falls back), < 0 on error.
// bit 0: ERA5SVD_BUILD_PB32=1 keeps 32-point tiles for long series too (diagnostics: scripts/time_build.py)
static unsigned build_debug_flags() {
  const char* e = getenv("ERA5SVD_BUILD_PB32");
  const char* f = getenv("ERA5SVD_BUILD_PB16");       // bit 1: 16-point tiles for short series as well
  const char* g = getenv("ERA5SVD_BUILD_CLUSTER");    // bit 2: long series as 32-point tiles split over a cluster
  return ((e && e[0] == '1') ? 1u : 0u) | ((f && f[0] == '1') ? 2u : 0u) | ((g && g[0] == '1') ? 4u : 0u);
}

template <typename Tx, int PB, int CL>
static int launch_build_tma(const CUtensorMap& tm, unsigned tiles, size_t tile_bytes, int64_t T, int64_t P, int shift, Tx* X,
                            int64_t ldx, Tx* mean_out, Tx* std_out, const Tx* weights, bool center, bool scale, bool check,
                            int* nonfinite_flag, cudaStream_t st, float* Xhi, float* Xlo) {
  auto kern = fused_build_tma_kernel<Tx, PB, CL>;
  ERA5SVD_CUDA(ensure_dynamic_smem((const void*)kern, tile_bytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(tiles * CL);
  cfg.blockDim = dim3(TB_THREADS);
  cfg.dynamicSmemBytes = tile_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  ERA5SVD_CUDA(cudaLaunchKernelEx(&cfg, kern, tm, (int)T, P, shift, X, ldx, mean_out, std_out, weights, center ? 1 : 0,
                                  (center && scale) ? 1 : 0, check ? 1 : 0, nonfinite_flag, Xhi, Xlo));
  return ERA5SVD_OK;
}

template <typename Tx>
int try_build_rows_tma(const float* src, int64_t T, int64_t src_ld, int64_t P, Tx* X, int64_t ldx, Tx* mean_out,
                       Tx* std_out, const Tx* weights, bool center, bool scale, bool check, int* nonfinite_flag,
                       cudaStream_t st, float* Xhi, float* Xlo) {
  if ((src_ld * 4) % 16 != 0 || P + 4 >= (int64_t)1 << 31) return 0;
  // 32-point tiles (128-byte source segments) while two [T x 32] tiles fit one SM (T <= ~850); longer series take
  // 16-point tiles (64-byte segments), still two resident CTAs (T <= ~1750).  The alternative for long series -
  // 32-point tiles split along time over a cluster of 2 / 4 CTAs, statistics through distributed shared memory
  // (ERA5SVD_BUILD_CLUSTER=1) - was measured and is NOT the default: with centring it is slower (T = 1460: 3.91 vs
  // 3.00 ms, the cluster barriers couple the CTAs' load and store phases), without statistics slightly faster
  // (2.83 vs 2.99 ms); profiles/r02_build_cluster.txt.  Series too long for a 16-point tile (3500 < T <= ~7000) do take
  // it with 4 CTAs per tile; with 8 (c4's 8760 hourly snapshots: 2.60 TB/s) it is no better than the register-staged
  // fallback (2.51 TB/s, and faster with scaling), which therefore keeps those.
  const size_t two_resident = 112 * 1024;
  const unsigned dbg = build_debug_flags();
  auto part_bytes = [&](int cl, int rowb) {
    const int64_t tc_ = cl > 1 ? ceil_div(ceil_div(T, (int64_t)cl), (int64_t)TB_BOX_T) * TB_BOX_T
                               : ceil_div(T, (int64_t)TB_BOX_T) * TB_BOX_T;
    return (size_t)tc_ * rowb + 1024;
  };
  int cl = 1, pb = 32;
  const size_t one_resident = 225 * 1024;
  if (dbg & 2) {
    pb = 16;
  } else if (!(dbg & 1) && part_bytes(1, 128) > two_resident) {
    if (dbg & 4) {                                               // diagnostics: cluster variant wherever it fits
      cl = part_bytes(2, 128) <= two_resident ? 2 : (part_bytes(4, 128) <= one_resident ? 4 : 8);
    } else if (part_bytes(1, 64) <= one_resident) {
      pb = 16;                                                   // T <= ~3500 (two resident CTAs up to ~1750)
    } else if (part_bytes(4, 128) <= one_resident) {
      cl = 4;                                                    // T <= ~7000
    } else {
      return 0;       // longer still (c4: 8760): 8 CTAs per tile measured no better than the staged fallback (r02_build_cluster.txt)
    }
  }
  const size_t tile_bytes = part_bytes(cl, pb * 4);
  if (tile_bytes > 225 * 1024) return 0;
  CUtensorMap tm;
  int shift = 0;
  int rc = tc::make_tmap(&tm, src, P, T, src_ld, (uint32_t)pb, TB_BOX_T, &shift,
                         pb == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  const unsigned tiles = (unsigned)ceil_div(P + shift, (int64_t)pb);
#define ERA5SVD_BUILD_ARGS tm, tiles, tile_bytes, T, P, shift, X, ldx, mean_out, std_out, weights, center, scale, check, \
                           nonfinite_flag, st, Xhi, Xlo
  if (pb == 16) rc = launch_build_tma<Tx, 16, 1>(ERA5SVD_BUILD_ARGS);
  else if (cl == 1) rc = launch_build_tma<Tx, 32, 1>(ERA5SVD_BUILD_ARGS);
  else if (cl == 2) rc = launch_build_tma<Tx, 32, 2>(ERA5SVD_BUILD_ARGS);
  else if (cl == 4) rc = launch_build_tma<Tx, 32, 4>(ERA5SVD_BUILD_ARGS);
  else rc = launch_build_tma<Tx, 32, 8>(ERA5SVD_BUILD_ARGS);
#undef ERA5SVD_BUILD_ARGS
  if (rc) return rc;
  rc = check_launch("fused_build_tma_kernel");
  return rc ? rc : 1;
}

template <typename Ts, typename Tx>
int build_rows_impl(const void* src, int64_t T, int64_t src_ld, int64_t P, void* X, int64_t ldx,
                    void* mean_out, void* std_out, const void* weights, unsigned flags,
                    int* nonfinite_flag, cudaStream_t st, float* Xhi = nullptr, float* Xlo = nullptr) {
  const bool center = flags & ERA5SVD_BUILD_MEAN_CENTER;
  const bool scale = flags & ERA5SVD_BUILD_SCALE;
  const bool check = (flags & ERA5SVD_BUILD_CHECK_FINITE) && nonfinite_flag;
  if constexpr (sizeof(Ts) == 4) {
    if (!(flags & ERA5SVD_BUILD_NO_TMA)) {
      int rc = try_build_rows_tma<Tx>((const float*)src, T, src_ld, P, (Tx*)X, ldx, (Tx*)mean_out, (Tx*)std_out,
                                      (const Tx*)weights, center, scale, check, nonfinite_flag, st, Xhi, Xlo);
      if (rc != 0) return rc < 0 ? rc : ERA5SVD_OK;
    }
  }
  // single-read fused kernel whenever the [T x 32] tile fits in shared memory (T <= ~1700 for fp32)
  const size_t tile_bytes = (size_t)T * (FB_PB + 1) * sizeof(Ts);
  if (tile_bytes <= 220 * 1024) {
    auto kern = fused_build_kernel<Ts, Tx>;
    ERA5SVD_CUDA(ensure_dynamic_smem((const void*)kern, tile_bytes));
    kern<<<(unsigned)ceil_div(P, FB_PB), FB_THREADS, tile_bytes, st>>>(
        (const Ts*)src, T, src_ld, P, (Tx*)X, ldx, (Tx*)mean_out, (Tx*)std_out, (const Tx*)weights, center ? 1 : 0,
        (center && scale) ? 1 : 0, check ? 1 : 0, nonfinite_flag, Xhi, Xlo);
    return check_launch("fused_build_kernel");
  }
  if (center) {
    int threads = 128;
    int64_t blocks = ceil_div(P, threads);
    row_stats_kernel<Ts, Tx><<<(unsigned)blocks, threads, 0, st>>>(
        (const Ts*)src, T, src_ld, P, (Tx*)mean_out, (Tx*)std_out, scale ? 1 : 0);
    int rc = check_launch("row_stats_kernel");
    if (rc) return rc;
  }
  dim3 block(32, 8);
  dim3 grid((unsigned)ceil_div(P, 32), (unsigned)ceil_div(T, 32));
  transpose_kernel<Ts, Tx><<<grid, block, 0, st>>>(
      (const Ts*)src, T, src_ld, P, (Tx*)X, ldx, center ? (const Tx*)mean_out : nullptr,
      (center && scale) ? (const Tx*)std_out : nullptr, (const Tx*)weights, check ? 1 : 0,
      nonfinite_flag, Xhi, Xlo);
  return check_launch("transpose_kernel");
}

// finiteness pass over a matrix that did not come through the build kernel (svd_on_era5's array input): one CTA per
// group of rows, lanes along the row (coalesced), 16-byte loads when the row pitch allows
template <typename T>
__global__ void __launch_bounds__(256)
check_finite_kernel(const T* __restrict__ X, int64_t rows, int64_t cols, int64_t ld, int* __restrict__ flag) {
  int bad = 0;
  for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const T* row = X + r * ld;
    for (int64_t c = threadIdx.x; c < cols; c += blockDim.x) {
      const T v = row[c];
      if (!(v - v == T(0))) bad = 1;           // NaN and +-Inf both fail
    }
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) atomicExch(flag, 1);
}

}  // namespace era5svd

extern "C" int era5svd_check_finite(const void* X, int dtype, int64_t rows, int64_t cols, int64_t ld, int* flag,
                                    void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(X && flag, "check_finite: null pointer");
  ERA5SVD_REQUIRE(valid_dtype(dtype), "check_finite: bad dtype");
  ERA5SVD_REQUIRE(rows > 0 && cols > 0 && ld >= cols, "check_finite: bad shape");
  int64_t blocks = rows < (int64_t)sm_count() * 16 ? rows : (int64_t)sm_count() * 16;
  if (dtype == ERA5SVD_F32)
    check_finite_kernel<float><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const float*)X, rows, cols, ld, flag);
  else
    check_finite_kernel<double><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>((const double*)X, rows, cols, ld, flag);
  return check_launch("check_finite_kernel");
}

extern "C" int era5svd_build_rows(const void* src, int dtype_src, int64_t T, int64_t src_ld,
                                  int64_t P, void* X, int dtype_x, int64_t ldx, void* mean_out,
                                  void* std_out, const void* weights, unsigned flags,
                                  int* nonfinite_flag, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(src && X, "build_rows: null src/X");
  ERA5SVD_REQUIRE(valid_dtype(dtype_src) && valid_dtype(dtype_x), "build_rows: bad dtype");
  ERA5SVD_REQUIRE(T > 0 && P > 0 && src_ld >= P && ldx >= T, "build_rows: bad shape T=%lld P=%lld src_ld=%lld ldx=%lld",
                  (long long)T, (long long)P, (long long)src_ld, (long long)ldx);
  ERA5SVD_REQUIRE(ceil_div(T, 32) <= 65535, "build_rows: T too large for one launch");
  if (flags & ERA5SVD_BUILD_SCALE)
    ERA5SVD_REQUIRE(flags & ERA5SVD_BUILD_MEAN_CENTER, "build_rows: SCALE requires MEAN_CENTER (reference quirk Q4: scale alone is ignored by the caller)");
  if (flags & ERA5SVD_BUILD_MEAN_CENTER) ERA5SVD_REQUIRE(mean_out, "build_rows: mean_out required with MEAN_CENTER");
  if (flags & ERA5SVD_BUILD_SCALE) ERA5SVD_REQUIRE(std_out, "build_rows: std_out required with SCALE");
  cudaStream_t st = as_stream(stream);
  if (dtype_src == ERA5SVD_F32 && dtype_x == ERA5SVD_F32)
    return build_rows_impl<float, float>(src, T, src_ld, P, X, ldx, mean_out, std_out, weights, flags, nonfinite_flag, st);
  if (dtype_src == ERA5SVD_F64 && dtype_x == ERA5SVD_F64)
    return build_rows_impl<double, double>(src, T, src_ld, P, X, ldx, mean_out, std_out, weights, flags, nonfinite_flag, st);
  if (dtype_src == ERA5SVD_F32 && dtype_x == ERA5SVD_F64)
    return build_rows_impl<float, double>(src, T, src_ld, P, X, ldx, mean_out, std_out, weights, flags, nonfinite_flag, st);
  return build_rows_impl<double, float>(src, T, src_ld, P, X, ldx, mean_out, std_out, weights, flags, nonfinite_flag, st);
}

extern "C" int era5svd_build_rows_split(const void* src, int dtype_src, int64_t T, int64_t src_ld,
                                        int64_t P, float* X, float* Xhi, float* Xlo, int64_t ldx,
                                        float* mean_out, float* std_out, const float* weights,
                                        unsigned flags, int* nonfinite_flag, void* stream) {
  using namespace era5svd;
  ERA5SVD_REQUIRE(src && Xhi && Xlo, "build_rows_split: null src/Xhi/Xlo");
  ERA5SVD_REQUIRE(valid_dtype(dtype_src), "build_rows_split: bad dtype");
  ERA5SVD_REQUIRE(T > 0 && P > 0 && src_ld >= P && ldx >= T, "build_rows_split: bad shape T=%lld P=%lld src_ld=%lld ldx=%lld",
                  (long long)T, (long long)P, (long long)src_ld, (long long)ldx);
  ERA5SVD_REQUIRE(ceil_div(T, 32) <= 65535, "build_rows_split: T too large for one launch");
  if (flags & ERA5SVD_BUILD_SCALE)
    ERA5SVD_REQUIRE(flags & ERA5SVD_BUILD_MEAN_CENTER, "build_rows_split: SCALE requires MEAN_CENTER");
  if (flags & ERA5SVD_BUILD_MEAN_CENTER) ERA5SVD_REQUIRE(mean_out, "build_rows_split: mean_out required with MEAN_CENTER");
  if (flags & ERA5SVD_BUILD_SCALE) ERA5SVD_REQUIRE(std_out, "build_rows_split: std_out required with SCALE");
  cudaStream_t st = as_stream(stream);
  if (dtype_src == ERA5SVD_F32)
    return build_rows_impl<float, float>(src, T, src_ld, P, X, ldx, mean_out, std_out, weights, flags, nonfinite_flag, st, Xhi, Xlo);
  return build_rows_impl<double, float>(src, T, src_ld, P, X, ldx, mean_out, std_out, weights, flags, nonfinite_flag, st, Xhi, Xlo);
}
