// era5svd_comm_* : the collectives of the row-sharded SVD as OUR kernels over peer memory.
//
// What crosses NVLink in this path is tiny and latency bound (Z = X^T Y: n x l float64 = 0.65 - 1.3 MB per tall pass; the
// l x l Gram matrix; k flip candidates), and every rank needs the full sum: a one-shot exchange - write the local
// contribution into a peer-mapped slot, flag, read all peers' slots and add them in rank order - is one kernel on the
// compute stream.  The sum over the partial tiles of the projection and the sum over the ranks are ONE kernel
// (reduce_partials_allreduce_kernel, gemm_simt.cu), requested with era5svd_comm_fuse_next_project.
#include "comm.cuh"

#include <string.h>

#include "common.cuh"

namespace era5svd {

struct Comm {
  int nranks = 0, rank = 0, dev = 0;
  size_t slot_doubles = 0;        // capacity of one slot
  size_t slot_bytes = 0;
  uint32_t epoch = 0;
  char* base = nullptr;           // own window
  char* peer[COMM_MAX_RANKS] = {};
  bool connected = false;
  unsigned long long fused = 0, calls = 0;
};

// flag pad: [COMM_MAX_RANKS] arrival words (written by the peers) + the local words ctr[2], go
static size_t flag_bytes() { return (size_t)(COMM_MAX_RANKS + 4) * sizeof(uint32_t); }

static thread_local Comm* g_bound = nullptr;
static thread_local int64_t g_bound_n = 0, g_bound_l = 0;

Comm* take_bound_comm(int64_t* n, int64_t* l) {
  Comm* c = g_bound;
  if (c) { *n = g_bound_n; *l = g_bound_l; }
  g_bound = nullptr;
  return c;
}

int comm_max_grid() {
  return 4 * sm_count();       // 256-thread kernels of <= 64 registers: four CTAs per SM are resident together
}

bool comm_next(Comm* c, int64_t count, CommDev* out) {
  if (!c || !c->connected) { set_error("comm: communicator is not connected"); return false; }
  if (count <= 0 || (size_t)count > c->slot_doubles) {
    set_error("comm: %lld doubles do not fit a slot of %zu", (long long)count, c->slot_doubles);
    return false;
  }
  c->epoch += 1;
  if (c->epoch == 0) c->epoch = 1;          // flags start at 0: never hand out epoch 0
  const size_t off = (c->epoch & 1u) ? c->slot_bytes : 0;
  out->nranks = c->nranks;
  out->rank = c->rank;
  out->epoch = c->epoch;
  for (int r = 0; r < COMM_MAX_RANKS; ++r) {
    out->slot[r] = r < c->nranks ? reinterpret_cast<double*>(c->peer[r] + off) : nullptr;
    out->flags[r] = r < c->nranks ? reinterpret_cast<uint32_t*>(c->peer[r] + 2 * c->slot_bytes) : nullptr;
  }
  uint32_t* local = reinterpret_cast<uint32_t*>(c->base + 2 * c->slot_bytes) + COMM_MAX_RANKS;
  out->ctr = local + (c->epoch & 1u);
  out->go = local + 2;
  c->calls += 1;
  return true;
}

// In-place all-reduce (sum) of `count` doubles.
__global__ void __launch_bounds__(256) comm_allreduce_f64_kernel(double* __restrict__ buf, int64_t count, CommDev c) {
  const int64_t per = (count + gridDim.x - 1) / gridDim.x;
  const int64_t a = (int64_t)blockIdx.x * per, b = a + per < count ? a + per : count;
  double* mine = c.slot[c.rank];
  for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) mine[i] = buf[i];
  comm_grid_exchange(c);
  for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) buf[i] = comm_sum_ranks(c, i);
}

// All-gather: dst[r * count + i] = rank r's src[i].
__global__ void __launch_bounds__(256) comm_allgather_f64_kernel(const double* __restrict__ src, int64_t count,
                                                                 double* __restrict__ dst, CommDev c) {
  const int64_t per = (count + gridDim.x - 1) / gridDim.x;
  const int64_t a = (int64_t)blockIdx.x * per, b = a + per < count ? a + per : count;
  double* mine = c.slot[c.rank];
  for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) mine[i] = src[i];
  comm_grid_exchange(c);
  for (int r = 0; r < c.nranks; ++r)
    for (int64_t i = a + threadIdx.x; i < b; i += blockDim.x) dst[(int64_t)r * count + i] = comm_load(c.slot[r] + i);
}

void comm_note_fused(Comm* c) { if (c) c->fused += 1; }

static unsigned grid_for(int64_t count) {
  // one value per thread while the grid allows: a thread's peer loads are one NVLink round trip however many ranks it
  // reads, but consecutive values of the same thread would pay one round trip EACH (measured: 4 values per thread 20 us,
  // against 9.4 us for a single block's exchange)
  int64_t g = ceil_div(count, 256);
  const int mx = comm_max_grid();
  if (g > mx) g = mx;
  if (g < 1) g = 1;
  return (unsigned)g;
}

}  // namespace era5svd

extern "C" {

using namespace era5svd;

size_t era5svd_comm_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

int era5svd_comm_create(int nranks, int rank, int64_t slot_bytes, void** comm_out, void* handle_out) {
  ERA5SVD_REQUIRE(comm_out && handle_out, "comm_create: null pointer");
  ERA5SVD_REQUIRE(nranks >= 1 && nranks <= COMM_MAX_RANKS && rank >= 0 && rank < nranks,
                  "comm_create: %d ranks (rank %d) not supported (max %d)", nranks, rank, COMM_MAX_RANKS);
  ERA5SVD_REQUIRE(slot_bytes >= 8, "comm_create: slot too small");
  Comm* c = new Comm();
  c->nranks = nranks;
  c->rank = rank;
  c->slot_bytes = ((size_t)slot_bytes + 255) / 256 * 256;
  c->slot_doubles = c->slot_bytes / 8;
  cudaError_t e = cudaGetDevice(&c->dev);
  const size_t total = 2 * c->slot_bytes + flag_bytes();
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->base, total);
  if (e == cudaSuccess) e = cudaMemset(c->base, 0, total);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->base);
  if (e != cudaSuccess) {
    set_error("comm_create: %s", cudaGetErrorString(e));
    if (c->base) cudaFree(c->base);
    delete c;
    return ERA5SVD_ERR_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  c->peer[rank] = c->base;
  *comm_out = c;
  return ERA5SVD_OK;
}

int era5svd_comm_connect(void* comm, const void* handles) {
  Comm* c = static_cast<Comm*>(comm);
  ERA5SVD_REQUIRE(c && handles, "comm_connect: null pointer");
  ERA5SVD_REQUIRE(!c->connected, "comm_connect: already connected");
  for (int r = 0; r < c->nranks; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("comm_connect: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      (void)cudaGetLastError();
      for (int q = 0; q < r; ++q)
        if (q != c->rank && c->peer[q]) { cudaIpcCloseMemHandle(c->peer[q]); c->peer[q] = nullptr; }
      return ERA5SVD_ERR_CUDA;
    }
    c->peer[r] = static_cast<char*>(p);
  }
  c->connected = true;
  return ERA5SVD_OK;
}

int era5svd_comm_destroy(void* comm) {
  Comm* c = static_cast<Comm*>(comm);
  if (!c) return ERA5SVD_OK;
  if (g_bound == c) g_bound = nullptr;
  cudaDeviceSynchronize();
  for (int r = 0; r < c->nranks; ++r)
    if (r != c->rank && c->peer[r]) cudaIpcCloseMemHandle(c->peer[r]);
  if (c->base) cudaFree(c->base);
  delete c;
  return ERA5SVD_OK;
}

int64_t era5svd_comm_capacity(void* comm) { return comm ? (int64_t)static_cast<Comm*>(comm)->slot_doubles : 0; }

unsigned long long era5svd_comm_fused_count(void* comm) { return comm ? static_cast<Comm*>(comm)->fused : 0; }

int era5svd_comm_allreduce_f64(void* comm, double* buf, int64_t count, void* stream) {
  Comm* c = static_cast<Comm*>(comm);
  ERA5SVD_REQUIRE(c && buf, "comm_allreduce: null pointer");
  CommDev d;
  if (!comm_next(c, count, &d)) return ERA5SVD_ERR_ARG;
  comm_allreduce_f64_kernel<<<grid_for(count), 256, 0, as_stream(stream)>>>(buf, count, d);
  return check_launch("comm_allreduce_f64_kernel");
}

int era5svd_comm_allgather_f64(void* comm, const double* src, int64_t count, double* dst, void* stream) {
  Comm* c = static_cast<Comm*>(comm);
  ERA5SVD_REQUIRE(c && src && dst, "comm_allgather: null pointer");
  CommDev d;
  if (!comm_next(c, count, &d)) return ERA5SVD_ERR_ARG;
  comm_allgather_f64_kernel<<<grid_for(count), 256, 0, as_stream(stream)>>>(src, count, dst, d);
  return check_launch("comm_allgather_f64_kernel");
}

int era5svd_comm_fuse_next_project(void* comm, int64_t n, int64_t l) {
  Comm* c = static_cast<Comm*>(comm);
  ERA5SVD_REQUIRE(c && c->connected, "comm_fuse_next_project: communicator is not connected");
  ERA5SVD_REQUIRE(n > 0 && l > 0, "comm_fuse_next_project: bad shape");
  if ((size_t)(n * l) > c->slot_doubles) {
    set_error("comm_fuse_next_project: %lld x %lld doubles do not fit a slot of %zu", (long long)n, (long long)l,
              c->slot_doubles);
    return ERA5SVD_ERR_UNSUPPORTED;
  }
  g_bound = c;
  g_bound_n = n;
  g_bound_l = l;
  return ERA5SVD_OK;
}

}  // extern "C"
