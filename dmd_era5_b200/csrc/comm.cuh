// Intra-node peer-memory communicator (CUDA IPC windows over NVLink / NVSwitch).  Not part of the public ABI.
//
// Every rank owns one window  [slot 0 | slot 1 | flag pad]  allocated with cudaMalloc and mapped into every peer through
// cudaIpcOpenMemHandle.  A collective writes this rank's contribution into the slot of the current epoch's parity, raises
// ONE flag per peer with a system-scope release store once all of its blocks have written, waits for the peers' flags with
// system-scope acquire loads, and then READS the peers' slots directly over NVLink - one kernel, no host round trip, no
// stream hop.
//
// Slot reuse is safe with two slots: a rank can only enter epoch e + 2 (which overwrites the slot of epoch e) after it left
// epoch e + 1, i.e. after every peer raised its e + 1 flags - and a peer raises them from a kernel that runs, in stream
// order, after its epoch-e kernel (the one that read our slot) has finished.
#pragma once

#include <stdint.h>
#include <stdio.h>

namespace era5svd {

constexpr int COMM_MAX_RANKS = 16;
constexpr long long COMM_TIMEOUT_CLOCKS = 40000000000ll;   // ~20 s at 2 GHz

struct CommDev {                          // passed to kernels by value
  int nranks, rank;
  uint32_t epoch;                         // value the flags of this collective carry (monotonic, wraps)
  double* slot[COMM_MAX_RANKS];           // this epoch's data slot of every rank (own and peer-mapped)
  uint32_t* flags[COMM_MAX_RANKS];        // flag pad of every rank: [COMM_MAX_RANKS] arrival words, one per sender
  uint32_t* ctr;                          // LOCAL: blocks of this grid that have written their part (this epoch's parity)
  uint32_t* go;                           // LOCAL: epoch whose peer data may be read
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
__device__ __forceinline__ uint32_t comm_ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-level exchange.  Every block has written its part of the local slot; on return EVERY rank's whole slot is visible
// to every block.  Two levels, so that what crosses NVLink for the synchronisation is one flag per peer and collective
// (a flag round per block - the first version - cost 7 system-scope releases per block and made the large messages
// slower than NCCL): blocks count in on a local word; block 0 waits for the count, raises this rank's flag at every
// peer (release, system scope), waits for the peers' flags (acquire, system scope) and publishes a local go word that
// the other blocks poll (gpu scope).  The grid must be co-resident (comm_max_grid()).
__device__ __forceinline__ void comm_grid_exchange(const CommDev& c) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();                            // this block's slot writes before its count
    atomicAdd(c.ctr, 1u);
  }
  if (blockIdx.x == 0) {
    const long long t0 = clock64();
    if (threadIdx.x == 0) {
      while (comm_ld_acquire_gpu(c.ctr) < gridDim.x) {
        if (clock64() - t0 > COMM_TIMEOUT_CLOCKS) __trap();
      }
      *c.ctr = 0;                               // next used two epochs from now, after everybody has passed this point
    }
    __syncthreads();
    if ((int)threadIdx.x < c.nranks) {
      const int t = threadIdx.x;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(c.flags[t] + c.rank), "r"(c.epoch) : "memory");
      const uint32_t* mine = c.flags[c.rank] + t;
      uint32_t v;
      for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - c.epoch) >= 0) break;
        // a peer that never arrives (it failed, or the ranks issued different collectives) must not hang the GPU: after
        // ~20 s of polling the kernel aborts, the stream reports a launch failure and the process exits
        if (clock64() - t0 > COMM_TIMEOUT_CLOCKS) {
          printf("era5svd comm: rank %d waited too long for rank %d (epoch %u): aborting\n", c.rank, t, c.epoch);
          __trap();
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(c.go), "r"(c.epoch) : "memory");
  } else if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while ((int32_t)(comm_ld_acquire_gpu(c.go) - c.epoch) < 0) {
      if (clock64() - t0 > 2 * COMM_TIMEOUT_CLOCKS) __trap();
    }
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
