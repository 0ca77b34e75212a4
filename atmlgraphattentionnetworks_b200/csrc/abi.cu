// ABI plumbing: version, thread-local error message, device attribute cache.
#include "common.cuh"
#include <string.h>

namespace b200gat {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;   // kernels launched by this library (not atomic: informational)

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) return 0;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return static_cast<int>(e);
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace b200gat

extern "C" int b200gat_abi_version(void) { return B200GAT_ABI_VERSION; }

extern "C" uint64_t b200gat_launch_count(void) { return b200gat::g_launches; }

extern "C" int b200gat_last_error(char* buf, size_t buf_len) {
  size_t n = strlen(b200gat::g_err);
  if (buf && buf_len) {
    size_t m = n < buf_len - 1 ? n : buf_len - 1;
    memcpy(buf, b200gat::g_err, m);
    buf[m] = 0;
  }
  return static_cast<int>(n);
}
