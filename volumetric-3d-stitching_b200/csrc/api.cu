// C-ABI plumbing: thread-local error message, version, device probe.
#include "common.cuh"
#include "v3s.h"
#include <stdarg.h>

namespace v3s {

static thread_local char g_err[512] = "";
long long g_kernel_launches = 0;

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

// number of CUDA kernels this library has launched so far (bench.py's gpu_launches)
extern "C" long long v3s_kernel_launches(void) { return v3s::g_kernel_launches; }

// 0 when a compute-capability-10.x device is current; the product path refuses anything else.
extern "C" int v3s_check_device(void) {
  int dev = 0;
  V3S_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  V3S_CUDA(cudaGetDeviceProperties(&prop, dev));
  V3S_REQUIRE(prop.major == 10, "v3s needs an sm_100a device (B200); found sm_%d%d (%s)", prop.major, prop.minor, prop.name);
  return 0;
}

// Fill n floats at dst (which may be a peer-mapped address of another GPU) -- used to validate the
// NVLink peer mapping before the sharded symmetric Gram stores through it.
namespace v3s {
__global__ void peer_fill_kernel(float* dst, float value, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = value;
}
}  // namespace v3s
extern "C" int v3s_peer_fill(float* dst, float value, long long n, void* stream) {
  V3S_REQUIRE(dst != nullptr && n >= 0, "v3s_peer_fill: bad arguments");
  if (n == 0) return 0;
  v3s::peer_fill_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(dst, value, n);
  V3S_CHECK_LAUNCH();
  return 0;
}
