// Shared helpers for libgnc.so (sm_100a).  Error reporting, launch accounting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/gnc.h"

namespace gnc {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, const char* a = "", long long b = 0, long long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// After every launch: catch configuration errors immediately (no sync).
inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return GNC_ECUDA;
  }
  count_launch();
  return GNC_OK;
}

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE attribute: remembered per (call site, device
// ordinal) so that a second GPU used by the same process is configured too; safe to call from several threads.
struct SmemAttrOnce {
  std::atomic<unsigned long long> done{0};
  template <typename Kernel>
  int ensure(Kernel kernel, int bytes, const char* what) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return GNC_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) {
      snprintf(g_err, sizeof(g_err), "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
      return GNC_ECUDA;
    }
    done.fetch_or(bit, std::memory_order_release);
    return GNC_OK;
  }
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// Streaming 128-bit accesses: data touched once should not displace L1 lines.
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

}  // namespace gnc

#define GNC_REQUIRE(cond, msg)                                   \
  do {                                                           \
    if (!(cond)) return gnc::fail(GNC_EINVAL, "%s", msg);        \
  } while (0)
