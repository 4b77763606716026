// K1 -- pattern synthesis, reference-exact mode.
//
// Restates projection.py:55-71 (generate_single_image) + rotation.py:143-162
// (rotate_structure_index) per pattern, fused in one persistent CTA per SM:
//   A. ED_rot[i,j,k] = Tri(F)(Rm (U_j, U_k, U_i))           (trilinear gather from smem)
//      fused with the axis-2 DFT restricted to the few k_z planes the Ewald sphere touches
//   B. 2-D DFT over axes 0,1 of those planes, magnitude, fftshift
//   C. trilinear sampling at the (rotation-independent) detector q-grid, circular mask
// The object F stays resident in shared memory across all patterns of the CTA.
// Coordinates / interpolation weights are computed in FP64, values in FP32.
#include "common.cuh"
#include "v3s.h"
#include <math.h>
#include <vector>

namespace v3s {

struct __align__(16) PixelTap {   // 16 bytes, one 128-bit load
  int   idx;        // i0 | i1<<8 | zrel<<16 | valid<<24
  float t0, t1, t2;
};

constexpr int kSynthThreads = 512;
constexpr int kMaxPlanes = 4;
constexpr int kSynthMaxN = 256;       // voxels per side (the cell index travels in 8 bits of PixelTap::idx)

// f in [0, n-1] (FP64 grid coordinate) -> cell index i in [0, n-2] and weight t = f - i as FP32 (t = 1 at the upper
// edge, like the reference's clipped searchsorted cell).  The nearest integer of f comes from the magic-number add
// (its low word), the remainder is exact in FP64 and rounded to FP32 once; a negative remainder moves one cell down.
__device__ __forceinline__ void split_cell(double f, int n, int& i, float& t) {
  const double magic = 6755399441055744.0;             // 2^52 + 2^51
  const double fr = f + magic;
  i = __double2loint(fr);
  double rem = f - (fr - magic);                        // in [-0.5, 0.5], exact
  if (rem < 0.0) { rem += 1.0; --i; }
  t = (float)rem;
  if (i < 0) { i = 0; t = 0.f; }                        // (f = -0 / rounding at the lower edge)
  if (i > n - 2) { t += (float)(i - (n - 2)); i = n - 2; }
}

struct SynthParams {
  int n;            // voxels per side
  int nz;           // number of axis-2 planes needed
  int z0;           // first (shifted) plane index
  int P;            // pixels per pattern
  long long N;      // patterns
};

__global__ void __launch_bounds__(kSynthThreads, 1)
synth_kernel(const float* __restrict__ ed, const double* __restrict__ rot, const PixelTap* __restrict__ taps,
             float* __restrict__ out, long long out_stride, SynthParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = prm.n, n2 = n * n, n3 = n2 * n, nz = prm.nz;
  float*  F    = reinterpret_cast<float*>(smem_raw);                         // n^3
  float2* slab = reinterpret_cast<float2*>(F + ((n3 + 3) & ~3));             // nz*n^2
  float2* tmp  = slab + nz * n2;                                              // nz*n^2
  float*  gmag = reinterpret_cast<float*>(tmp + nz * n2);                     // nz*n^2
  float2* tw   = reinterpret_cast<float2*>(gmag + ((nz * n2 + 1) & ~1));      // n
  float2* twk  = tw + n;                                                       // kMaxPlanes * n: twiddles of the axis-2 DFT
  __shared__ double Rm[9];
  __shared__ double Uax[kSynthMaxN];                                          // U_i, rotation.py:146

  const int tid = threadIdx.x;
  for (int i = tid; i < n3; i += kSynthThreads) F[i] = ed[i];
  for (int i = tid; i < n; i += kSynthThreads) {
    double s, c;
    sincospi(-2.0 * (double)i / (double)n, &s, &c);
    tw[i] = make_float2((float)c, (float)s);
  }
  // frequency index (mod n) of each needed shifted plane: f = c - n/2  (np.fft.fftshift)
  int fz[kMaxPlanes];
#pragma unroll
  for (int z = 0; z < kMaxPlanes; ++z) fz[z] = (((prm.z0 + z) - n / 2) % n + n) % n;

  __syncthreads();
  // axis-2 DFT of plane z at position k: exp(-2 pi i k f_z / n) = tw[(k f_z) mod n] -- the same for every column, so
  // tabulated once per CTA instead of a running phase per thread
  for (int i = tid; i < kMaxPlanes * n; i += kSynthThreads) {
    const int z = i / n, k = i - z * n;
    twk[i] = tw[(k * fz[z]) % n];
  }
  const double half_n = 0.5 * (double)n;
  const double U0 = (1.0 - (n + 1) * 0.5) / half_n;   // U_i = (i + 1 - (n+1)/2) / (n/2), rotation.py:146
  const double Ulast = ((double)n - (n + 1) * 0.5) / half_n;
  const double mU0h = -U0 * half_n;
  for (int i = tid; i < n; i += kSynthThreads) Uax[i] = (i + 1 - (n + 1) * 0.5) / half_n;

  for (long long pat = blockIdx.x; pat < prm.N; pat += gridDim.x) {
    __syncthreads();
    if (tid < 9) Rm[tid] = rot[pat * 9 + tid];
    __syncthreads();
    const double r00 = Rm[0], r01 = Rm[1], r02 = Rm[2], r10 = Rm[3], r11 = Rm[4], r12 = Rm[5], r20 = Rm[6],
                 r21 = Rm[7], r22 = Rm[8];

    // ---- phase A: rotate + DFT along axis 2 ------------------------------------------
    for (int col = tid; col < n2; col += kSynthThreads) {
      const int i = col / n, j = col - i * n;
      // p = (x, y, z) = (U_j, U_k, U_i): array axes are (z, x, y), projection.py:79
      const double px = Uax[j], pz = Uax[i];
      const double b0 = r00 * px + r02 * pz, b1 = r10 * px + r12 * pz, b2 = r20 * px + r22 * pz;
      float2 acc[kMaxPlanes];
#pragma unroll
      for (int z = 0; z < kMaxPlanes; ++z) acc[z] = make_float2(0.f, 0.f);
      // one voxel of the rotated object: trilinear sample of F at R p, p = (U_j, U_k, U_i); 0 outside the grid
      // (rotation.py:152-160).  Branch-free (clamped cell, value selected at the end) so that two voxels per
      // iteration interleave their dependent chains: FP64 coordinates -> cell split -> 8 gathers -> 7 lerps.
      auto sample = [&](int k) -> float {
        const double py = Uax[k];                      // (warp-wide broadcast from shared memory)
        const double q0 = b0 + r01 * py, q1 = b1 + r11 * py, q2 = b2 + r21 * py;
        const bool inb = !(q0 < U0 || q0 > Ulast || q1 < U0 || q1 > Ulast || q2 < U0 || q2 > Ulast);
        // grid coordinate f = (q - U0) n/2 in [0, n-1]; cell index and FP32 weight without double <-> int
        // conversions (quarter-rate pipe): adding 2^52 + 2^51 rounds f to the nearest integer in the low word
        int i0, i1, i2;
        float t0, t1, t2;
        split_cell(inb ? fma(q0, half_n, mU0h) : 0.0, n, i0, t0);
        split_cell(inb ? fma(q1, half_n, mU0h) : 0.0, n, i1, t1);
        split_cell(inb ? fma(q2, half_n, mU0h) : 0.0, n, i2, t2);
        const float* b = F + (i0 * n + i1) * n + i2;
        const float c000 = b[0], c001 = b[1], c010 = b[n], c011 = b[n + 1];
        const float c100 = b[n2], c101 = b[n2 + 1], c110 = b[n2 + n], c111 = b[n2 + n + 1];
        const float c00 = c000 + t2 * (c001 - c000), c01 = c010 + t2 * (c011 - c010);
        const float c10 = c100 + t2 * (c101 - c100), c11 = c110 + t2 * (c111 - c110);
        const float c0 = c00 + t1 * (c01 - c00), c1 = c10 + t1 * (c11 - c10);
        return inb ? c0 + t0 * (c1 - c0) : 0.f;
      };
      int k = 0;
      for (; k + 1 < n; k += 2) {
        const float v0 = sample(k), v1 = sample(k + 1);
#pragma unroll
        for (int z = 0; z < kMaxPlanes; ++z) {
          if (z < nz) {
            const float2 w0 = twk[z * n + k], w1 = twk[z * n + k + 1];
            acc[z].x = fmaf(v0, w0.x, acc[z].x);
            acc[z].y = fmaf(v0, w0.y, acc[z].y);
            acc[z].x = fmaf(v1, w1.x, acc[z].x);
            acc[z].y = fmaf(v1, w1.y, acc[z].y);
          }
        }
      }
      if (k < n) {
        const float v0 = sample(k);
#pragma unroll
        for (int z = 0; z < kMaxPlanes; ++z) {
          if (z < nz) {
            const float2 w0 = twk[z * n + k];
            acc[z].x = fmaf(v0, w0.x, acc[z].x);
            acc[z].y = fmaf(v0, w0.y, acc[z].y);
          }
        }
      }
#pragma unroll
      for (int z = 0; z < kMaxPlanes; ++z)
        if (z < nz) slab[z * n2 + col] = acc[z];
    }
    __syncthreads();

    // ---- phase B1: DFT along axis 0, output index in fftshift order ------------------
    for (int o = tid; o < nz * n2; o += kSynthThreads) {
      const int z = o / n2, r = o - z * n2, c0 = r / n, j = r - c0 * n;
      const int f = ((c0 - n / 2) % n + n) % n;
      float2 a = make_float2(0.f, 0.f);
      int p = 0;
      const float2* s = slab + z * n2 + j;
      for (int i = 0; i < n; ++i) {
        const float2 x = s[i * n], w = tw[p];
        a.x += x.x * w.x - x.y * w.y;
        a.y += x.x * w.y + x.y * w.x;
        p += f; if (p >= n) p -= n;
      }
      tmp[o] = a;
    }
    __syncthreads();
    // ---- phase B2: DFT along axis 1, magnitude ---------------------------------------
    for (int o = tid; o < nz * n2; o += kSynthThreads) {
      const int z = o / n2, r = o - z * n2, c0 = r / n, c1 = r - c0 * n;
      const int f = ((c1 - n / 2) % n + n) % n;
      float2 a = make_float2(0.f, 0.f);
      int p = 0;
      const float2* s = tmp + z * n2 + c0 * n;
      for (int j = 0; j < n; ++j) {
        const float2 x = s[j], w = tw[p];
        a.x += x.x * w.x - x.y * w.y;
        a.y += x.x * w.y + x.y * w.x;
        p += f; if (p >= n) p -= n;
      }
      gmag[o] = sqrtf(a.x * a.x + a.y * a.y);
    }
    __syncthreads();

    // ---- phase C: sample on the detector grid -----------------------------------------
    float* dst = out + pat * out_stride;
    // four pixels per iteration: their tap records (global / L2) are requested together, so one memory latency
    // covers four interpolations instead of one
    for (int p0 = tid; p0 < prm.P; p0 += 4 * kSynthThreads) {
      PixelTap t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * kSynthThreads;
        t[u] = p < prm.P ? taps[p] : PixelTap{0, 0.f, 0.f, 0.f};
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = p0 + u * kSynthThreads;
        float v = 0.f;
        if (t[u].idx >> 24) {
          const int i0 = t[u].idx & 0xff, i1 = (t[u].idx >> 8) & 0xff, z = (t[u].idx >> 16) & 0xff;
          const float* g0 = gmag + z * n2 + i0 * n + i1;
          const float* g1 = g0 + n2;
          const float a00 = g0[0] + t[u].t2 * (g1[0] - g0[0]);
          const float a01 = g0[1] + t[u].t2 * (g1[1] - g0[1]);
          const float a10 = g0[n] + t[u].t2 * (g1[n] - g0[n]);
          const float a11 = g0[n + 1] + t[u].t2 * (g1[n + 1] - g0[n + 1]);
          const float a0 = a00 + t[u].t1 * (a01 - a00), a1 = a10 + t[u].t1 * (a11 - a10);
          v = a0 + t[u].t0 * (a1 - a0);
        }
        if (p < prm.P) dst[p] = v;
      }
    }
  }
}

// Host-side geometry (projection.py:95-117, 121-160, 164-190), FP64, once per call.
static int build_taps(int n, int n_p, const v3s_geometry* g, std::vector<PixelTap>& taps, int* z0, int* nz) {
  const double width = g->pixel * ((double)g->n_p_nobin / (double)n_p) * (double)n_p;  // SuperPixel * N_p
  // k axis: projection.py:121-126 with Length = max - min of 2e-7 * U
  std::vector<double> U(n), kax(n);
  for (int i = 0; i < n; ++i) U[i] = ((double)(i + 1) - (n + 1) / 2.0) / (n / 2.0);
  const double length = g->grid_scale * U[n - 1] - g->grid_scale * U[0];
  const double d = length / (n - 1);
  const double nyq = 0.5 / d;
  const double N1 = (n - 1) / 2.0;
  const double scale = 2 * nyq / n;
  for (int c = 0; c < n; ++c) kax[c] = (-N1 + c) * scale;
  auto locate = [&](double x, int* idx, double* t) -> bool {
    if (x < kax[0] || x > kax[n - 1] || x != x) return false;
    int cnt = 0;                       // searchsorted(..., side='left')
    while (cnt < n && kax[cnt] < x) ++cnt;
    int i = cnt - 1;
    if (i < 0) i = 0;
    if (i > n - 2) i = n - 2;
    *idx = i;
    *t = (x - kax[i]) / (kax[i + 1] - kax[i]);
    return true;
  };
  const int P = n_p * n_p;
  taps.assign(P, PixelTap{0, 0.f, 0.f, 0.f});
  std::vector<int> iz(P, -1);
  int zmin = 1 << 30, zmax = -1;
  for (int u = 0; u < n_p; ++u) {
    for (int v = 0; v < n_p; ++v) {
      // camera_x_y: U = (Width/(N-1)) * (linspace(1,N,N) - (N+1)/2); meshgrid 'xy'
      const double cx = (width / (n_p - 1)) * ((double)(v + 1) - (n_p + 1) / 2.0);
      const double cy = (width / (n_p - 1)) * ((double)(u + 1) - (n_p + 1) / 2.0);
      const double T = g->lambda * sqrt(cx * cx + cy * cy + g->zd * g->zd);
      const double qx = cx / T, qy = cy / T, qz = g->zd / T - 1.0 / g->lambda;
      const bool masked = (cx * cx + cy * cy) > (width / 2) * (width / 2);
      int i0, i1, i2;
      double t0, t1, t2;
      if (masked) continue;
      if (!locate(qx, &i0, &t0) || !locate(qy, &i1, &t1) || !locate(qz, &i2, &t2)) continue;
      const int p = u * n_p + v;
      iz[p] = i2;
      zmin = i2 < zmin ? i2 : zmin;
      zmax = i2 > zmax ? i2 : zmax;
      taps[p].idx = i0 | (i1 << 8) | (1 << 24);
      taps[p].t0 = (float)t0; taps[p].t1 = (float)t1; taps[p].t2 = (float)t2;
    }
  }
  if (zmax < 0) { *z0 = 0; *nz = 2; return 0; }   // every pixel masked / out of range: all-zero patterns
  *z0 = zmin;
  *nz = zmax + 2 - zmin;   // planes zmin .. zmax+1
  for (int p = 0; p < P; ++p)
    if (iz[p] >= 0) taps[p].idx |= ((iz[p] - zmin) << 16);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Direct mode: the analytical scattering amplitude on each orientation's rotated Ewald-sphere
// q-grid (BASELINE north_star kernel 1; closed form verified in SURVEY Appendix A.1):
//   I(u,v) = | sum_{ijk} F[i,j,k] exp(-2 pi i (M kappa) . (a_i, a_j, a_k)) |,  M = R Pm,
//   Pm: (c0,c1,c2) -> (c1,c2,c0), a_i = grid_scale * U_i, kappa = (Q_x,Q_y,Q_z)[u,v].
// It is the same physical image as the reference's rotate -> FFT -> interpolate chain without its
// two trilinear interpolations (corr 0.998, not the same numbers), so it is an option, not the
// default.  One CTA per (pattern, 512-pixel tile), one pixel per thread; the object sits in
// shared memory and is read as warp-wide broadcasts, the triple sum is evaluated separably
// (n^3 real x complex FMAs per pixel: FP32-pipe bound), phases by SFU sincospi.
constexpr int kDirectThreads = 512;
constexpr int kDirectMaxN = 32;

struct DirectPixel { float kx, ky, kz, valid; };   // kappa * 2 * grid_scale: phase / pi = k . U

__global__ void __launch_bounds__(kDirectThreads)
synth_direct_kernel(const float* __restrict__ ed, const double* __restrict__ rot, const DirectPixel* __restrict__ pix,
                    float* __restrict__ out, long long out_stride, int n, int P, long long N) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n2 = n * n, n3 = n2 * n;
  const int kpad = (n + 3) & ~3;                                   // rows of F padded to float4
  float* F = reinterpret_cast<float*>(smem_raw);                    // n*n*kpad
  __shared__ float Mrot[9];
  const int tid = threadIdx.x;
  for (int e = tid; e < n2 * kpad; e += kDirectThreads) {
    const int ij = e / kpad, k = e - ij * kpad;
    F[e] = k < n ? ed[ij * n + k] : 0.f;
  }
  (void)n3;
  const long long pat = blockIdx.y;
  if (tid < 9) {
    // M = R Pm: column c of M is column (c+2)%3 ... (M x)_r = sum_c R[r][c] (Pm x)_c with Pm x = (x1, x2, x0)
    const int r = tid / 3, c = tid % 3;
    const int src = (c + 2) % 3;            // M[r][c] multiplies x_c; (Pm x)_src = x_c  <=>  src = (c + 2) % 3
    Mrot[tid] = (float)rot[pat * 9 + r * 3 + src];
  }
  __syncthreads();
  const int p = blockIdx.x * kDirectThreads + tid;
  float result = 0.f;
  const bool active = p < P && pix[p < P ? p : 0].valid != 0.f;
  if (active) {
    const DirectPixel q = pix[p];
    const float k0 = Mrot[0] * q.kx + Mrot[1] * q.ky + Mrot[2] * q.kz;
    const float k1 = Mrot[3] * q.kx + Mrot[4] * q.ky + Mrot[5] * q.kz;
    const float k2 = Mrot[6] * q.kx + Mrot[7] * q.ky + Mrot[8] * q.kz;
    const float half_n = 0.5f * (float)n;
    float2 e2[kDirectMaxN];
#pragma unroll
    for (int k = 0; k < kDirectMaxN; ++k) {
      if (k < n) {
        const float U = ((float)(k + 1) - 0.5f * (float)(n + 1)) / half_n;
        float sn, cs;
        sincospif(-k2 * U, &sn, &cs);
        e2[k] = make_float2(cs, sn);
      } else {
        e2[k] = make_float2(0.f, 0.f);
      }
    }
    // e1(j) = exp(-i pi k1 U_j) by rotation: U_{j+1} - U_j = 2/n
    const float U_first = (1.f - 0.5f * (float)(n + 1)) / half_n;
    float e1r0, e1i0, w1r, w1i;
    sincospif(-k1 * U_first, &e1i0, &e1r0);
    sincospif(-k1 * (2.f / (float)n), &w1i, &w1r);
    float ar = 0.f, ai = 0.f;
    for (int i = 0; i < n; ++i) {
      float br = 0.f, bi = 0.f;
      float e1r = e1r0, e1i = e1i0;
      for (int j = 0; j < n; ++j) {
        const float4* f4 = reinterpret_cast<const float4*>(F + (i * n + j) * kpad);
        float cr = 0.f, ci = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < kDirectMaxN / 4; ++k4) {
          if (k4 * 4 < n) {
            const float4 f = f4[k4];                                // warp-wide broadcast
            cr = fmaf(f.x, e2[4 * k4].x, cr);     ci = fmaf(f.x, e2[4 * k4].y, ci);
            cr = fmaf(f.y, e2[4 * k4 + 1].x, cr); ci = fmaf(f.y, e2[4 * k4 + 1].y, ci);
            cr = fmaf(f.z, e2[4 * k4 + 2].x, cr); ci = fmaf(f.z, e2[4 * k4 + 2].y, ci);
            cr = fmaf(f.w, e2[4 * k4 + 3].x, cr); ci = fmaf(f.w, e2[4 * k4 + 3].y, ci);
          }
        }
        br += cr * e1r - ci * e1i;
        bi += cr * e1i + ci * e1r;
        const float t = e1r * w1r - e1i * w1i;
        e1i = e1r * w1i + e1i * w1r;
        e1r = t;
      }
      const float U = ((float)(i + 1) - 0.5f * (float)(n + 1)) / half_n;
      float sn, cs;
      sincospif(-k0 * U, &sn, &cs);
      ar += br * cs - bi * sn;
      ai += br * sn + bi * cs;
    }
    result = sqrtf(ar * ar + ai * ai);
  }
  if (p < P) out[pat * out_stride + p] = result;
}

}  // namespace v3s

using namespace v3s;

extern "C" void v3s_default_geometry(v3s_geometry* g) {
  // projection.py:96-107
  g->pixel = 75e-6; g->zd = 0.5; g->lambda = 2 * 1e-9; g->n_p_nobin = 1024; g->grid_scale = 2e-7;
}

extern "C" long long v3s_synth_workspace_bytes(int n_p) { return (long long)sizeof(PixelTap) * n_p * n_p; }

// Geometry set-up, once per (n, n_p, geometry): fills the caller's device workspace with the
// per-pixel interpolation taps and returns the k_z plane range in meta_host = {z0, nz}.
extern "C" int v3s_synth_prepare(int n, int n_p, const v3s_geometry* geom, void* ws_dev, int* meta_host, void* stream) {
  V3S_REQUIRE(n >= 2 && n <= 255, "v3s_synth_prepare: n_voxel_1d=%d outside [2,255]", n);
  V3S_REQUIRE(n_p >= 2, "v3s_synth_prepare: bad n_p=%d", n_p);
  v3s_geometry g;
  if (geom) g = *geom; else v3s_default_geometry(&g);
  std::vector<PixelTap> taps;
  int z0, nz;
  build_taps(n, n_p, &g, taps, &z0, &nz);
  V3S_REQUIRE(nz <= kMaxPlanes, "v3s_synth_prepare: Ewald sphere spans %d k_z planes (max %d)", nz, kMaxPlanes);
  // pageable source: the copy is staged before the call returns, so `taps` may be released
  V3S_CUDA(cudaMemcpyAsync(ws_dev, taps.data(), sizeof(PixelTap) * n_p * n_p, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  meta_host[0] = z0;
  meta_host[1] = nz;
  return 0;
}

extern "C" int v3s_synth_patterns(const float* ed, int n, const double* rot9, long long N, int n_p, const void* ws_dev,
                                  int z0, int nz, float* out, long long out_stride, void* stream) {
  V3S_REQUIRE(n >= 2 && n <= 255 && n_p >= 2 && N >= 0, "v3s_synth_patterns: bad n=%d n_p=%d N=%lld", n, n_p, N);
  V3S_REQUIRE(nz >= 1 && nz <= kMaxPlanes && ws_dev != nullptr, "v3s_synth_patterns: call v3s_synth_prepare first");
  if (N == 0) return 0;
  const int P = n_p * n_p;
  const int n2 = n * n, n3 = n2 * n;
  size_t smem = sizeof(float) * ((n3 + 3) & ~3) + 2 * sizeof(float2) * nz * n2 + sizeof(float) * ((nz * n2 + 1) & ~1) +
                sizeof(float2) * n + sizeof(float2) * kMaxPlanes * n;
  V3S_REQUIRE(smem <= 227 * 1024, "v3s_synth_patterns: object n=%d needs %zu B shared memory (max 232448)", n, smem);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    V3S_CUDA(cudaFuncSetAttribute(synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  SynthParams prm{n, nz, z0, P, N};
  int grid = (int)(N < (long long)kNumSMs ? N : kNumSMs);
  synth_kernel<<<grid, kSynthThreads, smem, (cudaStream_t)stream>>>(ed, rot9, (const PixelTap*)ws_dev, out, out_stride, prm);
  V3S_CHECK_LAUNCH();
  return 0;
}

// Direct mode (see synth_direct_kernel).  ws_dev: n_p*n_p*16 bytes (same size as the exact mode's).
extern "C" int v3s_synth_patterns_direct(const float* ed, int n, const double* rot9, long long N, int n_p,
                                         const v3s_geometry* geom, void* ws_dev, float* out, long long out_stride,
                                         void* stream) {
  V3S_REQUIRE(n >= 2 && n <= kDirectMaxN, "v3s_synth_patterns_direct: n_voxel_1d=%d outside [2,%d]", n, kDirectMaxN);
  V3S_REQUIRE(n_p >= 2 && N >= 0 && N < 65536LL * 65535LL && ws_dev != nullptr, "v3s_synth_patterns_direct: bad arguments");
  v3s_geometry g;
  if (geom) g = *geom; else v3s_default_geometry(&g);
  const int P = n_p * n_p;
  std::vector<DirectPixel> pix(P);
  const double width = g.pixel * ((double)g.n_p_nobin / (double)n_p) * (double)n_p;
  for (int u = 0; u < n_p; ++u)
    for (int v = 0; v < n_p; ++v) {
      const double cx = (width / (n_p - 1)) * ((double)(v + 1) - (n_p + 1) / 2.0);
      const double cy = (width / (n_p - 1)) * ((double)(u + 1) - (n_p + 1) / 2.0);
      const double T = g.lambda * sqrt(cx * cx + cy * cy + g.zd * g.zd);
      const bool masked = (cx * cx + cy * cy) > (width / 2) * (width / 2);
      const double sc = 2.0 * g.grid_scale;                       // phase / pi = (2 a kappa) with a = grid_scale * U
      pix[u * n_p + v] = DirectPixel{(float)(sc * cx / T), (float)(sc * cy / T), (float)(sc * (g.zd / T - 1.0 / g.lambda)),
                                     masked ? 0.f : 1.f};
    }
  cudaStream_t st = (cudaStream_t)stream;
  V3S_CUDA(cudaMemcpyAsync(ws_dev, pix.data(), sizeof(DirectPixel) * P, cudaMemcpyHostToDevice, st));
  if (N == 0) return 0;
  const int kpad = (n + 3) & ~3;
  const size_t smem = sizeof(float) * n * n * kpad;
  V3S_REQUIRE(smem <= 227 * 1024, "v3s_synth_patterns_direct: shared memory %zu B too large", smem);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    V3S_CUDA(cudaFuncSetAttribute(synth_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  for (long long p0 = 0; p0 < N; p0 += 65535) {     // grid.y limit
    const long long np = (N - p0 < 65535) ? (N - p0) : 65535;
    dim3 grid((unsigned)((P + kDirectThreads - 1) / kDirectThreads), (unsigned)np);
    synth_direct_kernel<<<grid, kDirectThreads, smem, st>>>(ed, rot9 + p0 * 9, (const DirectPixel*)ws_dev,
                                                            out + p0 * out_stride, out_stride, n, P, np);
    V3S_CHECK_LAUNCH();
  }
  return 0;
}
