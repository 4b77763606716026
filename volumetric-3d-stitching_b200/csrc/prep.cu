// K2a -- centre + unit-normalise the pattern rows and emit the split-precision Gram operands.
//
// Restates distance.py:70-71 (remove_offset: subtract the per-pixel mean over patterns) and
// distance.py:66-67 (set_norm_to_one: rows / ||row||_2).  Outputs per row:
//   a_f32 [n_rows, P]        centred unit row, FP32 (used by the exact distance refinement)
//   a_hi, a_lo [n_rows, Ppad] BF16 split a ~= hi + lo (16 mantissa bits), zero padded to Ppad
// HBM-bound: reads the pattern matrix twice (column sums, then rows), writes 4+2+2 B/element.
#include "common.cuh"
#include "v3s.h"
#include <cuda_fp16.h>

namespace v3s {

// column sums in FP64: thread per column (coalesced over columns), CTA-y over row chunks
__global__ void colsum_kernel(const float* __restrict__ img, long long n_rows, int P, long long stride,
                              int rows_per_cta, double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= P) return;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > n_rows) r1 = n_rows;
  double acc = 0.0;
  const float* p = img + r0 * stride + c;
  long long r = r0;
  for (; r + 4 <= r1; r += 4) {   // 4 independent loads in flight
    const float a = p[0], b = p[stride], d = p[2 * stride], e = p[3 * stride];
    acc += (double)a + (double)b + ((double)d + (double)e);
    p += 4 * stride;
  }
  for (; r < r1; ++r) { acc += (double)p[0]; p += stride; }
  atomicAdd(&sums[c], acc);
}
// the same with four columns per thread (128-bit loads; P and the row stride multiples of 4, 16-byte aligned rows):
// four times the bytes in flight per thread and a quarter of the column blocks, so more row chunks run in parallel
__global__ void colsum4_kernel(const float* __restrict__ img, long long n_rows, int P4, long long stride,
                               int rows_per_cta, double* __restrict__ sums) {
  const int c4 = blockIdx.x * blockDim.x + threadIdx.x;
  if (c4 >= P4) return;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  long long r1 = r0 + rows_per_cta;
  if (r1 > n_rows) r1 = n_rows;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  const float* p = img + r0 * stride + 4 * (long long)c4;
  long long r = r0;
  for (; r + 4 <= r1; r += 4) {
    const float4 u = *reinterpret_cast<const float4*>(p), v = *reinterpret_cast<const float4*>(p + stride),
                 w = *reinterpret_cast<const float4*>(p + 2 * stride), x = *reinterpret_cast<const float4*>(p + 3 * stride);
    a0 += (double)u.x + (double)v.x + ((double)w.x + (double)x.x);
    a1 += (double)u.y + (double)v.y + ((double)w.y + (double)x.y);
    a2 += (double)u.z + (double)v.z + ((double)w.z + (double)x.z);
    a3 += (double)u.w + (double)v.w + ((double)w.w + (double)x.w);
    p += 4 * stride;
  }
  for (; r < r1; ++r) {
    const float4 u = *reinterpret_cast<const float4*>(p);
    a0 += (double)u.x; a1 += (double)u.y; a2 += (double)u.z; a3 += (double)u.w;
    p += stride;
  }
  atomicAdd(&sums[4 * c4 + 0], a0);
  atomicAdd(&sums[4 * c4 + 1], a1);
  atomicAdd(&sums[4 * c4 + 2], a2);
  atomicAdd(&sums[4 * c4 + 3], a3);
}

constexpr int kPrepThreads = 256;

__global__ void __launch_bounds__(kPrepThreads)
center_normalize_kernel(const float* __restrict__ img, long long n_rows, int P, long long stride,
                        const double* __restrict__ mean, float* __restrict__ a_f32, __nv_bfloat16* __restrict__ a_hi,
                        __nv_bfloat16* __restrict__ a_lo, int Ppad, double* __restrict__ norms,
                        int* __restrict__ zero_flag, __half* __restrict__ a_f16 = nullptr) {
  __shared__ double scratch[32];
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const float* src = img + row * stride;
    double ss = 0.0;
    for (int c = threadIdx.x; c < P; c += kPrepThreads) {
      const double x = (double)src[c] - mean[c];
      ss += x * x;
    }
    ss = block_sum(ss, scratch);
    const double nrm = sqrt(ss);
    const bool zero = !(nrm > 0.0);
    if (threadIdx.x == 0) {
      if (norms) norms[row] = nrm;
      if (zero_flag) zero_flag[row] = zero ? 1 : 0;
    }
    const double inv = zero ? 0.0 : 1.0 / nrm;
    float* d32 = a_f32 ? a_f32 + row * (long long)P : nullptr;
    __nv_bfloat16* dh = a_hi ? a_hi + row * (long long)Ppad : nullptr;
    __nv_bfloat16* dl = a_lo ? a_lo + row * (long long)Ppad : nullptr;
    __half* d16 = a_f16 ? a_f16 + row * (long long)Ppad : nullptr;
    for (int c = threadIdx.x; c < Ppad; c += kPrepThreads) {
      float a = 0.f;
      if (c < P) {
        a = (float)(((double)src[c] - mean[c]) * inv);
        if (d32) d32[c] = a;
      }
      if (d16) d16[c] = __float2half_rn(a * 256.0f);       // the fused Gram kernel's operand fl16(2^8 a) (= v3s_split_f16)
      if (dh) {
        const __nv_bfloat16 h = __float2bfloat16_rn(a);
        dh[c] = h;
        dl[c] = __float2bfloat16_rn(a - __bfloat162float(h));
      }
    }
  }
}

// The same for the fused path's outputs (FP32 rows + optional FP16 operand), four pixels per thread and iteration:
// 128-bit loads of the pattern row, 128-bit stores of the FP32 row, 64-bit stores of the FP16 operand.  P, Ppad and the
// row stride are multiples of 4 and the rows 16-byte aligned.  Same arithmetic per element as the scalar kernel.
__global__ void __launch_bounds__(kPrepThreads)
center_normalize4_kernel(const float* __restrict__ img, long long n_rows, int P, long long stride,
                         const double* __restrict__ mean, float* __restrict__ a_f32, __half* __restrict__ a_f16, int Ppad,
                         int* __restrict__ zero_flag) {
  __shared__ double scratch[32];
  const int P4 = P / 4, Ppad4 = Ppad / 4;
  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(img + row * stride);
    double ss = 0.0;
    for (int c = threadIdx.x; c < P4; c += kPrepThreads) {
      const float4 v = src[c];
      const double2 m0 = reinterpret_cast<const double2*>(mean)[2 * c], m1 = reinterpret_cast<const double2*>(mean)[2 * c + 1];
      const double x0 = (double)v.x - m0.x, x1 = (double)v.y - m0.y, x2 = (double)v.z - m1.x, x3 = (double)v.w - m1.y;
      ss += x0 * x0;
      ss += x1 * x1;
      ss += x2 * x2;
      ss += x3 * x3;
    }
    ss = block_sum(ss, scratch);
    const double nrm = sqrt(ss);
    const bool zero = !(nrm > 0.0);
    if (threadIdx.x == 0 && zero_flag) zero_flag[row] = zero ? 1 : 0;
    const double inv = zero ? 0.0 : 1.0 / nrm;
    float4* d32 = a_f32 ? reinterpret_cast<float4*>(a_f32 + row * (long long)P) : nullptr;
    uint2* d16 = a_f16 ? reinterpret_cast<uint2*>(a_f16 + row * (long long)Ppad) : nullptr;
    for (int c = threadIdx.x; c < Ppad4; c += kPrepThreads) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < P4) {
        const float4 v = src[c];
        const double2 m0 = reinterpret_cast<const double2*>(mean)[2 * c], m1 = reinterpret_cast<const double2*>(mean)[2 * c + 1];
        a.x = (float)(((double)v.x - m0.x) * inv);
        a.y = (float)(((double)v.y - m0.y) * inv);
        a.z = (float)(((double)v.z - m1.x) * inv);
        a.w = (float)(((double)v.w - m1.y) * inv);
        if (d32) d32[c] = a;
      }
      if (d16) {
        const __half2 h01 = __halves2half2(__float2half_rn(a.x * 256.0f), __float2half_rn(a.y * 256.0f));
        const __half2 h23 = __halves2half2(__float2half_rn(a.z * 256.0f), __float2half_rn(a.w * 256.0f));
        uint2 o;
        o.x = *reinterpret_cast<const unsigned*>(&h01);
        o.y = *reinterpret_cast<const unsigned*>(&h23);
        d16[c] = o;
      }
    }
  }
}

// BF16 hi/lo split of already-normalised rows (used after an all-gather of the FP32 rows: shipping
// 4 B/element and splitting locally is cheaper than shipping the 4 + 2 + 2 B/element of all three)
__global__ void split_bf16_kernel(const float* __restrict__ a, long long n_rows, int P, int Ppad,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const long long total = n_rows * (long long)Ppad;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / Ppad;
    const int c = (int)(e - r * Ppad);
    const float v = c < P ? a[r * P + c] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[e] = h;
    lo[e] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

}  // namespace v3s

using namespace v3s;

extern "C" int v3s_split_bf16(const float* a_f32, long long n_rows, int P, int Ppad, void* a_hi, void* a_lo, void* stream) {
  V3S_REQUIRE(n_rows >= 0 && P > 0 && Ppad >= P && Ppad % 8 == 0 && a_hi && a_lo, "v3s_split_bf16: bad arguments");
  if (n_rows == 0) return 0;
  long long g = ceil_div64(n_rows * (long long)Ppad, 256);
  if (g > 16LL * kNumSMs) g = 16LL * kNumSMs;
  split_bf16_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(a_f32, n_rows, P, Ppad, (__nv_bfloat16*)a_hi, (__nv_bfloat16*)a_lo);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_column_sums(const float* img, long long n_rows, int P, long long stride, double* sums,
                               void* stream) {
  V3S_REQUIRE(P > 0 && n_rows >= 0 && stride >= P, "v3s_column_sums: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  V3S_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * P, st));
  if (n_rows == 0) return 0;
  if (P % 4 == 0 && stride % 4 == 0 && ((uintptr_t)img & 15) == 0) {
    const int P4 = P / 4;
    const int bx4 = (P4 + 127) / 128;
    long long want = (8LL * kNumSMs + bx4 - 1) / bx4;
    if (want < 1) want = 1;
    long long rows4 = ceil_div64(n_rows, want);
    if (rows4 < 16) rows4 = 16;
    const int by4 = (int)ceil_div64(n_rows, rows4);
    colsum4_kernel<<<dim3(bx4, by4), 128, 0, st>>>(img, n_rows, P4, stride, (int)rows4, sums);
    V3S_CHECK_LAUNCH();
    return 0;
  }
  const int bx = (P + 127) / 128;
  long long want_y = (4LL * kNumSMs + bx - 1) / bx;
  if (want_y < 1) want_y = 1;
  long long rows_per = ceil_div64(n_rows, want_y);
  if (rows_per < 16) rows_per = 16;
  const int by = (int)ceil_div64(n_rows, rows_per);
  colsum_kernel<<<dim3(bx, by), 128, 0, st>>>(img, n_rows, P, stride, (int)rows_per, sums);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_center_normalize(const float* img, long long n_rows, int P, long long stride, const double* mean,
                                    float* a_f32, void* a_hi, void* a_lo, int Ppad, double* norms, int* zero_flag,
                                    void* stream) {
  V3S_REQUIRE(P > 0 && n_rows >= 0 && stride >= P, "v3s_center_normalize: bad shape");
  V3S_REQUIRE((a_hi == nullptr) == (a_lo == nullptr), "v3s_center_normalize: a_hi and a_lo go together");
  V3S_REQUIRE(a_hi == nullptr || (Ppad >= P && Ppad % 8 == 0), "v3s_center_normalize: Ppad must be >= P and a multiple of 8");
  if (n_rows == 0) return 0;
  const int grid = (int)(n_rows < 8LL * kNumSMs ? n_rows : 8LL * kNumSMs);
  center_normalize_kernel<<<grid, kPrepThreads, 0, (cudaStream_t)stream>>>(
      img, n_rows, P, stride, mean, a_f32, (__nv_bfloat16*)a_hi, (__nv_bfloat16*)a_lo, a_hi ? Ppad : P, norms, zero_flag);
  V3S_CHECK_LAUNCH();
  return 0;
}

// v3s_center_normalize with the FP16 operand of the fused Gram / top-k kernel (v3s_split_f16's output: fl16(2^8 a),
// zero padded to Ppad columns) written in the same pass -- one read of the patterns less than centring + splitting.
extern "C" int v3s_center_normalize_f16(const float* img, long long n_rows, int P, long long stride, const double* mean,
                                        float* a_f32, void* a_f16, int Ppad, int* zero_flag, void* stream) {
  V3S_REQUIRE(P > 0 && n_rows >= 0 && stride >= P, "v3s_center_normalize_f16: bad shape");
  V3S_REQUIRE(a_f16 != nullptr && Ppad >= P && Ppad % 8 == 0, "v3s_center_normalize_f16: a_f16 with Ppad >= P, a multiple of 8");
  if (n_rows == 0) return 0;
  const int grid = (int)(n_rows < 8LL * kNumSMs ? n_rows : 8LL * kNumSMs);
  if (P % 4 == 0 && Ppad % 4 == 0 && stride % 4 == 0 && (((uintptr_t)img | (uintptr_t)a_f32 | (uintptr_t)mean) & 15) == 0 &&
      ((uintptr_t)a_f16 & 7) == 0) {
    center_normalize4_kernel<<<grid, kPrepThreads, 0, (cudaStream_t)stream>>>(img, n_rows, P, stride, mean, a_f32, (__half*)a_f16,
                                                                              Ppad, zero_flag);
    V3S_CHECK_LAUNCH();
    return 0;
  }
  center_normalize_kernel<<<grid, kPrepThreads, 0, (cudaStream_t)stream>>>(img, n_rows, P, stride, mean, a_f32, nullptr, nullptr,
                                                                           Ppad, nullptr, zero_flag, (__half*)a_f16);
  V3S_CHECK_LAUNCH();
  return 0;
}
