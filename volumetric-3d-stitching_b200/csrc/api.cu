// C-ABI plumbing: thread-local error message, version, device probe.
#include "common.cuh"
#include <cstring>
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

// ---- small device -> host reads that bypass the copy engines --------------------------------------------------
// A cudaMemcpy D2H of a few bytes queues on the device-to-host copy engine BEHIND whatever bulk transfer is running
// there (the 13 GB pattern matrix of OrientationRecoverySolver.solve travels while stages 2-4 run): every Ritz check
// of the eigen-solver, every validation scalar then waits for hundreds of megabytes.  These reads go through a
// pinned, device-mapped mailbox instead: a one-block kernel stores the bytes over PCIe, the stream is synchronised,
// the host copies them out.
namespace v3s {
constexpr size_t kMailboxBytes = 256 * 1024;
static unsigned char* g_mailbox = nullptr;
__global__ void store_to_host_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst, long long bytes) {
  const long long n16 = ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) ? bytes / 16 : 0;
  for (long long i = threadIdx.x; i < n16; i += blockDim.x)
    reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
  for (long long i = n16 * 16 + threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}
int read_back(void* host_dst, const void* dev_src, size_t bytes, cudaStream_t st) {
  V3S_REQUIRE(bytes <= kMailboxBytes, "v3s read_back: %zu bytes exceed the mailbox", bytes);
  if (!g_mailbox) V3S_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&g_mailbox), kMailboxBytes, cudaHostAllocMapped | cudaHostAllocPortable));
  if (bytes == 0) return 0;
  store_to_host_kernel<<<1, 256, 0, st>>>(reinterpret_cast<const unsigned char*>(dev_src), g_mailbox, (long long)bytes);
  V3S_CHECK_LAUNCH();
  V3S_CUDA(cudaStreamSynchronize(st));
  memcpy(host_dst, g_mailbox, bytes);
  return 0;
}
}  // namespace v3s
extern "C" int v3s_read_small(const void* dev_src, void* host_dst, long long bytes, void* stream) {
  V3S_REQUIRE(dev_src && host_dst && bytes >= 0, "v3s_read_small: bad arguments");
  return v3s::read_back(host_dst, dev_src, (size_t)bytes, (cudaStream_t)stream);
}
