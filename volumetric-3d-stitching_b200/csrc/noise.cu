// Poisson photon noise on the synthesised patterns (BASELINE config 5).  Not part of the reference
// (it has no noise model); a builder-defined input transform applied to the pattern
// matrix before stage 2, so that both pipelines can be fed the same noisy patterns.
//   each pattern is scaled to `photons` expected counts in total, every pixel is replaced by a
//   Poisson draw of its expectation (the result is the photon count, as float).
// Counter-based RNG (Philox-4x32-10) keyed by (seed; global pattern index, pixel, draw) so the
// noise does not depend on how the patterns are sharded across GPUs or on launch geometry.
#include "common.cuh"
#include "v3s.h"

namespace v3s {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

struct PhiloxStream {            // uniform floats in (0,1), 4 per Philox call
  uint2 key; uint32_t c0, c1, draw; uint4 buf; int have;
  __device__ PhiloxStream(unsigned long long seed, unsigned long long row, uint32_t pix)
      : key(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(row >> 32))), c0((uint32_t)row), c1(pix), draw(0),
        have(0) {}
  __device__ float next() {
    if (have == 0) { buf = philox4x32_10(make_uint4(c0, c1, draw++, 0x5eedu), key); have = 4; }
    const uint32_t u = have == 4 ? buf.x : have == 3 ? buf.y : have == 2 ? buf.z : buf.w;
    --have;
    return ((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f);
  }
};

__device__ float poisson_draw(PhiloxStream& s, float lam) {
  if (!(lam > 0.f)) return 0.f;
  if (lam < 10.f) {                       // Knuth: multiply uniforms until below exp(-lambda)
    const float L = __expf(-lam);
    float p = 1.f;
    int k = -1;
    do { ++k; p *= s.next(); } while (p > L && k < 200);
    return (float)k;
  }
  // PTRS (Hoermann 1993): transformed rejection with squeeze
  const float slam = sqrtf(lam), loglam = logf(lam);
  const float b = 0.931f + 2.53f * slam, a = -0.059f + 0.02483f * b;
  const float invalpha = 1.1239f + 1.1328f / (b - 3.4f), vr = 0.9277f - 3.6224f / (b - 2.f);
  for (int it = 0; it < 64; ++it) {
    const float U = s.next() - 0.5f, V = s.next();
    const float us = 0.5f - fabsf(U);
    const float k = floorf((2.f * a / us + b) * U + lam + 0.43f);
    if (us >= 0.07f && V <= vr) return k;
    if (k < 0.f || (us < 0.013f && V > us)) continue;
    if (logf(V) + logf(invalpha) - logf(a / (us * us) + b) <= -lam + k * loglam - lgammaf(k + 1.f)) return k;
  }
  return floorf(lam + 0.5f);
}

__global__ void __launch_bounds__(256)
poisson_noise_kernel(float* __restrict__ img, long long n_rows, int P, long long stride, long long row0_global,
                     double photons, unsigned long long seed) {
  __shared__ double scratch[32];
  for (long long r = blockIdx.x; r < n_rows; r += gridDim.x) {
    float* row = img + r * stride;
    double s = 0.0;
    for (int c = threadIdx.x; c < P; c += 256) s += (double)row[c];
    s = block_sum(s, scratch);
    const float scale = s > 0.0 ? (float)(photons / s) : 0.f;
    for (int c = threadIdx.x; c < P; c += 256) {
      PhiloxStream st(seed, (unsigned long long)(row0_global + r), (uint32_t)c);
      row[c] = poisson_draw(st, row[c] * scale);
    }
    __syncthreads();
  }
}

}  // namespace v3s

using namespace v3s;

extern "C" int v3s_poisson_noise(float* img, long long n_rows, int P, long long stride, long long row0_global,
                                 double photons_per_pattern, unsigned long long seed, void* stream) {
  V3S_REQUIRE(n_rows >= 0 && P > 0 && stride >= P && photons_per_pattern > 0, "v3s_poisson_noise: bad arguments");
  if (n_rows == 0) return 0;
  const int grid = (int)(n_rows < 8LL * kNumSMs ? n_rows : 8LL * kNumSMs);
  poisson_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(img, n_rows, P, stride, row0_global, photons_per_pattern, seed);
  V3S_CHECK_LAUNCH();
  return 0;
}
