// K2a -- centre + unit-normalise the pattern rows and emit the split-precision Gram operands.
//
// Restates distance.py:70-71 (remove_offset: subtract the per-pixel mean over patterns) and
// distance.py:66-67 (set_norm_to_one: rows / ||row||_2).  Outputs per row:
//   a_f32 [n_rows, P]        centred unit row, FP32 (used by the exact distance refinement)
//   a_hi, a_lo [n_rows, Ppad] BF16 split a ~= hi + lo (16 mantissa bits), zero padded to Ppad
// HBM-bound: reads the pattern matrix twice (column sums, then rows), writes 4+2+2 B/element.
#include "common.cuh"
#include "v3s.h"

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

constexpr int kPrepThreads = 256;

__global__ void __launch_bounds__(kPrepThreads)
center_normalize_kernel(const float* __restrict__ img, long long n_rows, int P, long long stride,
                        const double* __restrict__ mean, float* __restrict__ a_f32, __nv_bfloat16* __restrict__ a_hi,
                        __nv_bfloat16* __restrict__ a_lo, int Ppad, double* __restrict__ norms,
                        int* __restrict__ zero_flag) {
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
    for (int c = threadIdx.x; c < Ppad; c += kPrepThreads) {
      float a = 0.f;
      if (c < P) {
        a = (float)(((double)src[c] - mean[c]) * inv);
        if (d32) d32[c] = a;
      }
      if (dh) {
        const __nv_bfloat16 h = __float2bfloat16_rn(a);
        dh[c] = h;
        dl[c] = __float2bfloat16_rn(a - __bfloat162float(h));
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
