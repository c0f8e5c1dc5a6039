// Shared host/device helpers for the jafpro_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "jafpro_b200.h"

namespace jaf {

// Thread-local error message behind jaf_last_error().
void set_error(const char* fmt, ...);
// cudaGetLastError() -> status; bumps the process-wide launch counter by `launches`.
int finish_launch(const char* what, int launches = 1);
int cuda_status(cudaError_t e, const char* what);

// Name of the kernel variant the calling thread launched last (behind jaf_last_kernel()): measurement tools
// report the kernel that actually ran instead of guessing from the shapes.
void note_kernel(const char* fmt, ...);

// Per-device state.  Function attributes (dynamic shared memory), __constant__ symbols and SM counts belong to a
// DEVICE, not to the process: a process that touches a second GPU (nn.DataParallel, test/conv_pro_test.py:114-141)
// must set them again there.  current_device() < 0 on error; sm_count() is cached per device.
constexpr int kMaxDevices = 64;
int current_device();
int sm_count(int dev);
// `if (!once.done(dev)) { ...set attributes on dev...; once.mark(dev); }` — idempotent work, so a race between two
// threads on the same device only repeats it.
struct PerDeviceOnce {
  volatile unsigned char flags[kMaxDevices] = {};
  bool done(int dev) const { return dev >= 0 && dev < kMaxDevices && flags[dev] != 0; }
  void mark(int dev) { if (dev >= 0 && dev < kMaxDevices) flags[dev] = 1; }
};

// raster.cu: z-buffer keys [B,S,S] u64 of the projected poses into `workspace` (jaf_raster_workspace_bytes)
// keys_clean: the first B*is*is keys are already empty (the caller vouches for it): skip the clear
int raster_keys_from_poses(const float* cam, const float* verts, const int* fidx, int B, int V, int F, int is,
                           float eye_z, float near_, float far_, void* workspace, cudaStream_t st, int* launches,
                           bool keys_clean = false);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace jaf

#define JAF_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) {                                               \
      jaf::set_error("%s: %s (%s)", __func__, msg, #cond);       \
      return JAF_ERR_INVALID;                                    \
    }                                                            \
  } while (0)

#define JAF_CUDA(call)                                           \
  do {                                                           \
    int st_ = jaf::cuda_status((call), #call);                   \
    if (st_ != JAF_OK) return st_;                               \
  } while (0)

// ---------------------------------------------------------------------------------
// Device-side load/store flavours.
//   ld_stream*: read-once data (flows, logits, masks): bypass L1 so the gather
//               footprint of the reference images keeps the cache.
//   st_stream*: write-once outputs: evict-first in L2.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_stream_f32x2(const float* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_s32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// L1-bypassing loads that ask L2 to retain the line (the data is read once more, later, by the same CTA)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float ld_stream_keep_f32(const float* p, uint64_t pol) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float2 ld_stream_keep_f32x2(const float* p, uint64_t pol) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;"
               : "=f"(v.x), "=f"(v.y)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint4 ld_gather_u128(const uint4* p) {
  // read-only path, L1-allocating: neighbouring output pixels re-use the same taps
  return __ldg(p);
}
__device__ __forceinline__ void st_stream_u128(uint4* p, uint4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream_f32(float* p, float v) { __stcs(p, v); }

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
