// Shared helpers for the v3s CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "v3s kernels are written for sm_100a (B200) only"
#endif

namespace v3s {

// ---- error plumbing (C-ABI returns int, message via v3s_last_error) -----------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define V3S_CUDA(call)                                                             \
  do {                                                                             \
    cudaError_t _e = (call);                                                       \
    if (_e != cudaSuccess) return v3s::cuda_fail(_e, #call, __FILE__, __LINE__);   \
  } while (0)

extern long long g_kernel_launches;   // kernels launched by this library since it was loaded (api.cu)
}  // namespace v3s
#define V3S_CHECK_LAUNCH()                      \
  do {                                          \
    ++v3s::g_kernel_launches;                   \
    V3S_CUDA(cudaGetLastError());               \
  } while (0)
namespace v3s {

#define V3S_REQUIRE(cond, ...)                                                     \
  do {                                                                             \
    if (!(cond)) { v3s::set_error(__VA_ARGS__); return 1; }                        \
  } while (0)

constexpr int kNumSMs = 148;

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- warp / block reductions ------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w < v ? w : v; }
  return v;
}

// Block-wide sum; `scratch` needs >= 32 elements of T in shared memory. Result valid in all threads.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  T r = (lane < nw) ? scratch[lane] : T(0);
  r = warp_sum(r);
  return r;
}

// FP32 operand block X [N, 8] of the block mat-vec, stored in "strip layout": columns of P in
// strips of 128, each strip = 1024 floats laid out as float4 [half = j/4][cc = col%4][lane = (col%128)/4],
// so that one 4 KB bulk copy per strip lands in shared memory ready for conflict-free 128-bit reads
// by the lane that owns columns lane*4 .. lane*4+3.  Buffer length: xf_strip_floats(N).
__host__ __device__ inline long long xf_strip_index(long long col, int j) {
  const long long strip = col >> 7;
  const int cl = (int)(col & 127);
  return strip * 1024 + (long long)(((((j >> 2) << 2) + (cl & 3)) * 32 + (cl >> 2)) * 4 + (j & 3));
}
__host__ __device__ inline long long xf_strip_floats(long long N) { return ((N + 127) >> 7) * 1024; }

// small device -> host read through a pinned mapped mailbox (api.cu): synchronises the stream, uses no copy engine
int read_back(void* host_dst, const void* dev_src, size_t bytes, cudaStream_t st);

// order-preserving float <-> uint key (larger float => larger key)
__device__ __forceinline__ uint32_t float_to_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
  uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

}  // namespace v3s
