// capi.cu -- library-level entry points: version, thread-local error text, launch counter, device queries.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace pea {

static thread_local char tls_error[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;  // B200; not cached so a later call can still query the device
  }
  return cached;
}

}  // namespace pea

extern "C" int pea_version(void) { return PEA_B200_VERSION; }
extern "C" const char* pea_last_error(void) { return pea::tls_error; }
extern "C" long long pea_launch_count(void) { return pea::g_launches.load(); }
