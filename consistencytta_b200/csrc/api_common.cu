// libctta: error state, version, launch counter.
#include "ctta_internal.h"
#include <cstdarg>
#include <cstdio>

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

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
}  // namespace ctta

extern "C" {
const char* ctta_last_error(void) { return ctta::g_err; }
int ctta_version(void) { return 100; }
long long ctta_launch_count(void) { return ctta::g_launch_count.load(); }
}
