// C-ABI plumbing: thread-local error message, version, device probe.
#include "common.cuh"
#include "v3s.h"
#include <stdarg.h>

namespace v3s {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
  return 2;
}

}  // namespace v3s

extern "C" const char* v3s_last_error(void) { return v3s::g_err; }

extern "C" int v3s_version(void) { return 100; }

// 0 when a compute-capability-10.x device is current; the product path refuses anything else.
extern "C" int v3s_check_device(void) {
  int dev = 0;
  V3S_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  V3S_CUDA(cudaGetDeviceProperties(&prop, dev));
  V3S_REQUIRE(prop.major == 10, "v3s needs an sm_100a device (B200); found sm_%d%d (%s)", prop.major, prop.minor, prop.name);
  return 0;
}
