// Shared helpers for the sm_100a kernels behind the C ABI in include/simple_rf_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define SRF_API extern "C" __attribute__((visibility("default")))

namespace srf {

// Last error text, readable through srf_last_error().  One slot per host thread.
extern thread_local char g_last_error[512];

inline int fail(const char* where, const char* what) {
  snprintf(g_last_error, sizeof(g_last_error), "%s: %s", where, what);
  return 1;
}

inline int check_launch(const char* where) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(where, cudaGetErrorString(e));
  return 0;
}

#define SRF_REQUIRE(cond, where, msg) \
  do { if (!(cond)) return ::srf::fail(where, msg); } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Streaming (read-once / write-once) global accesses: keep them out of L1.
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float* p, float v) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// Saved activation tiles of the fused MLP: [tile][act_slots data images + mask images][16 KB].  The mask images follow the
// data images; the 1 KB record of data image s (one 32-bit word per (column group, row): bit i = pre-activation value of
// column 32 * group + i is positive, i.e. the ReLU output is non-zero) sits in mask image s / 16 at byte (s % 16) * 1024 + group * 512 + row * 4.
__host__ __device__ inline int act_tile_images(int act_slots) { return act_slots + (act_slots + 15) / 16; }
__host__ __device__ inline size_t act_mask_offset(int act_slots, int slot) {
  return (size_t)(act_slots + slot / 16) * 16384 + (size_t)(slot % 16) * 1024;
}

// cudaFuncSetAttribute is a PER-DEVICE setting: true the first time `mask` is consulted on the current device (the caller then
// configures the kernel).  Racing host threads may both configure - the call is idempotent.
inline bool first_use_on_this_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace srf
