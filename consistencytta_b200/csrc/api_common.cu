// libctta: error state, version, launch counter.
#include "ctta_internal.h"
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <set>
#include <utility>

namespace ctta {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launch_count{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  return dev;
}

// per-DEVICE caches: a process may drive several GPUs (one engine per device), and function attributes / SM counts
// belong to the device that is current at launch time
int sm_count() {
  static std::atomic<int> n[kMaxDevices];
  const int dev = current_device();
  if (dev < 0 || dev >= kMaxDevices) return 148;
  int v = n[dev].load(std::memory_order_relaxed);
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev].store(v, std::memory_order_relaxed);
  }
  return v;
}

int ensure_dynamic_smem(const void* fn, int bytes) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> configured;
  const std::pair<int, const void*> key(current_device(), fn);
  std::lock_guard<std::mutex> lock(mu);
  if (configured.count(key)) return 0;
  CTTA_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  configured.insert(key);
  return 0;
}
}  // namespace ctta

extern "C" {
const char* ctta_last_error(void) { return ctta::g_err; }
int ctta_version(void) { return 100; }
long long ctta_launch_count(void) { return ctta::g_launch_count.load(); }
}
