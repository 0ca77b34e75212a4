// ABI plumbing: version, thread-local error message, device attribute cache.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>

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

int64_t l2_bytes() {
  static thread_local int cached_dev = -1;
  static thread_local int64_t cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return int64_t(126) << 20;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrL2CacheSize, dev) != cudaSuccess || n <= 0) n = 126 << 20;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

bool edge_schedule_streaming(int64_t span, int64_t row_bytes) {
  static int forced = -1;   // 0: by span, 1: cached, 2: stream
  if (forced < 0) {
    const char* e = getenv("B200GAT_EDGE_SCHEDULE");
    forced = (e && e[0] == 'c') ? 1 : ((e && e[0] == 's') ? 2 : 0);
  }
  if (forced) return forced == 2;
  if (span < 0) return false;                       // unknown locality: the occupancy-first schedule
  return (span + 1) * row_bytes > l2_bytes() / 2;
}

int make_dropout(const float* mask, const b200gat_dropout& d, DropoutSpec* out) {
  out->mask = mask; out->seed = nullptr; out->threshold = 0u; out->scale = 1.f;
  if (mask) return 0;                                   // the tensor wins
  B200GAT_REQUIRE(d.p >= 0.f && d.p <= 1.f, B200GAT_E_SHAPE, "dropout: p = %g outside [0, 1]", double(d.p));
  if (d.p <= 0.f) return 0;
  B200GAT_REQUIRE(d.seed != nullptr, B200GAT_E_NULL, "dropout: p > 0 needs the device seed words");
  out->seed = d.seed;
  if (d.p >= 1.f) { out->threshold = 0xFFFFFFFFu; out->scale = 0.f; return 0; }   // everything dropped
  const double t = double(d.p) * 4294967296.0;
  out->threshold = t >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(t);
  out->scale = 1.f / (1.f - d.p);
  return 0;
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
