// Host-side helpers shared by the translation units of libctta: error reporting and the launch counter.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include "../../include/ctta.h"

namespace ctta {
int set_error(int code, const char* fmt, ...);
extern std::atomic<long long> g_launch_count;
inline void count_launch(int n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }
constexpr int kMaxDevices = 64;
int current_device();                                    // cudaGetDevice (0 on error)
int sm_count();                                          // SM count of the CURRENT device (cached per device)
int ensure_dynamic_smem(const void* fn, int bytes);      // cudaFuncAttributeMaxDynamicSharedMemorySize once per (device, fn)
}  // namespace ctta

#define CTTA_REQUIRE(cond, ...)                                                  \
  do {                                                                           \
    if (!(cond)) return ::ctta::set_error(CTTA_ERR_INVALID, __VA_ARGS__);        \
  } while (0)

#define CTTA_CUDA(call)                                                                                 \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess)                                                                             \
      return ::ctta::set_error(CTTA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                               __FILE__, __LINE__);                                                     \
  } while (0)

#define CTTA_LAUNCH_CHECK()                                                                             \
  do {                                                                                                  \
    cudaError_t e__ = cudaGetLastError();                                                               \
    if (e__ != cudaSuccess)                                                                             \
      return ::ctta::set_error(CTTA_ERR_CUDA, "kernel launch failed: %s (%s:%d)",                       \
                               cudaGetErrorString(e__), __FILE__, __LINE__);                            \
    ::ctta::count_launch();                                                                             \
  } while (0)
