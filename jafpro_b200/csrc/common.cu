// Error reporting, launch accounting and version for the jafpro_b200 C ABI.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace {
thread_local char g_err[512] = "";
thread_local char g_kernel[160] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace

namespace jaf {

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void note_kernel(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_kernel, sizeof(g_kernel), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return JAF_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return JAF_ERR_CUDA;
}

int current_device() {
  int dev = -1;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  return dev;
}

int sm_count(int dev) {
  static volatile int cached[kMaxDevices] = {};
  if (dev < 0 || dev >= kMaxDevices) return 0;
  int v = cached[dev];
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = v;
  }
  return v;
}

int finish_launch(const char* what, int launches) {
  g_launches.fetch_add((uint64_t)launches, std::memory_order_relaxed);
  return cuda_status(cudaGetLastError(), what);
}

}  // namespace jaf

extern "C" {

int jaf_version(void) { return 100; }
const char* jaf_last_error(void) { return g_err; }
const char* jaf_last_kernel(void) { return g_kernel; }
uint64_t jaf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
