// Error string, version, launch counter, device info.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace era5svd {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

cudaError_t ensure_dynamic_smem(const void* func, size_t bytes) {
  struct Entry { const void* f; int dev; size_t bytes; };
  static thread_local Entry cache[64];
  static thread_local int used = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  for (int i = 0; i < used; ++i)
    if (cache[i].f == func && cache[i].dev == dev) {
      if (cache[i].bytes >= bytes) return cudaSuccess;
      cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
      if (e == cudaSuccess) cache[i].bytes = bytes;
      return e;
    }
  cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess && used < 64) cache[used++] = Entry{func, dev, bytes};
  return e;
}

}  // namespace era5svd

extern "C" {

int era5svd_version(void) { return ERA5SVD_VERSION; }

const char* era5svd_last_error(void) { return era5svd::g_err; }

unsigned long long era5svd_launch_count(void) {
  return era5svd::g_launches.load(std::memory_order_relaxed);
}

}  // extern "C"
