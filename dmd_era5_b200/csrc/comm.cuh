// Intra-node peer-memory communicator (CUDA IPC windows over NVLink / NVSwitch).  Not part of the public ABI.
//
// Every rank owns one window  [slot 0 | slot 1 | flag pad]  allocated with cudaMalloc and mapped into every peer through
// cudaIpcOpenMemHandle.  A collective writes this rank's contribution into the slot of the current epoch's parity, raises
// one flag per (block, peer) with a system-scope release store, waits for the peers' flags with system-scope acquire loads,
// and then READS the peers' slots directly over NVLink - one kernel, no host round trip, no stream hop.
//
// Slot reuse is safe with two slots: a rank can only enter epoch e + 2 (which overwrites the slot of epoch e) after it left
// epoch e + 1, i.e. after every peer raised its e + 1 flags - and a peer raises them from a kernel that runs, in stream
// order, after its epoch-e kernel (the one that read our slot) has finished.
#pragma once

#include <stdint.h>

namespace era5svd {

constexpr int COMM_MAX_RANKS = 16;
constexpr int COMM_MAX_BLOCKS = 320;      // upper bound of the grids that take part (>= 2 x 148 SMs)

struct CommDev {                          // passed to kernels by value
  int nranks, rank;
  uint32_t epoch;                         // value the flags of this collective carry (monotonic, wraps)
  double* slot[COMM_MAX_RANKS];           // this epoch's data slot of every rank (own and peer-mapped)
  uint32_t* flags[COMM_MAX_RANKS];        // flag pad of every rank: [COMM_MAX_BLOCKS][COMM_MAX_RANKS]
};

struct Comm;

// The communicator bound by era5svd_comm_fuse_next_project on this thread (consumed by the next reduction of partial tiles),
// or nullptr.  take_bound_comm() clears the binding; *n / *l are the shape the caller promised.
Comm* take_bound_comm(int64_t* n, int64_t* l);
// Advance the epoch and fill the device view.  Returns false (and sets the error) when `count` doubles do not fit a slot.
bool comm_next(Comm* c, int64_t count, CommDev* out);
void comm_note_fused(Comm* c);
int comm_max_grid();                      // largest grid a collective kernel may use

#ifdef __CUDACC__
// Block-level exchange: every thread of the block has written its part of the local slot; on return the same block's part
// of EVERY rank's slot is visible.  Block b of one rank pairs with block b of the others (same grid on every rank).
__device__ __forceinline__ void comm_block_exchange(const CommDev& c) {
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < c.nranks) {
    const int t = threadIdx.x;
    uint32_t* remote = c.flags[t] + (size_t)blockIdx.x * COMM_MAX_RANKS + c.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(c.epoch) : "memory");
    const uint32_t* mine = c.flags[c.rank] + (size_t)blockIdx.x * COMM_MAX_RANKS + t;
    uint32_t v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    } while ((int32_t)(v - c.epoch) < 0);
  }
  __syncthreads();
}

// Coherent (system-scope) load of a peer's value: never served from a stale L1 line of an earlier epoch.
__device__ __forceinline__ double comm_load(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// Sum over the ranks in rank order 0, 1, ..., R-1: the same association on every rank, so the replicas stay bit-identical.
__device__ __forceinline__ double comm_sum_ranks(const CommDev& c, int64_t idx) {
  double v[COMM_MAX_RANKS];
#pragma unroll
  for (int r = 0; r < COMM_MAX_RANKS; ++r)
    if (r < c.nranks) v[r] = comm_load(c.slot[r] + idx);
  double s = v[0];
#pragma unroll
  for (int r = 1; r < COMM_MAX_RANKS; ++r)
    if (r < c.nranks) s += v[r];
  return s;
}
#endif

}  // namespace era5svd
