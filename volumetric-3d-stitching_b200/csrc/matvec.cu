// K3 -- block mat-vec  Y = P_rows . X  for the block-Lanczos eigen-solver (HBM-bound).
//
// Replaces the dense `P_ep @ v` that ARPACK drives inside scipy.sparse.linalg.eigs
// (diffusion.py:61).  P_rows is the FP32 row block [n_rows, N] of the Markov matrix, X is the
// FP64 Lanczos block [N, b] (b <= 8).  Algorithmic traffic per call: 4 n_rows N bytes of P,
// read exactly once with 128-bit coalesced loads; X is staged per 128-column strip in shared
// memory as FP32; products accumulate in FP32 per lane (8 rows x 8 vectors = 64 accumulators)
// and are combined across the warp with shuffles in FP64 ("warp-shuffle-reduced").
// Work item = (8-row group per warp, column chunk); column chunks give enough CTAs to fill
// 148 SMs at small N.  Per-chunk partial sums are written separately and reduced in a fixed
// order by a second tiny kernel, so results are run-to-run deterministic.
#include "common.cuh"
#include "v3s.h"

namespace v3s {

constexpr int MV_B = 8;            // vectors per block
constexpr int MV_ROWS = 8;         // rows per warp
constexpr int MV_WARPS = 4;        // warps per CTA
constexpr int MV_THREADS = MV_WARPS * 32;
constexpr int MV_STRIP = 128;      // columns per strip (32 lanes x float4)

__global__ void x_to_f32_kernel(const double* __restrict__ X, long long N, int b, float* __restrict__ Xf) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < N * MV_B; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / MV_B;
    const int j = (int)(e - i * MV_B);
    Xf[e] = j < b ? (float)X[i * b + j] : 0.f;
  }
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// grid.x = row groups of MV_WARPS*MV_ROWS rows, grid.y = column chunks
__global__ void __launch_bounds__(MV_THREADS)
block_matvec_kernel(const float* __restrict__ P, long long ldp, long long n_rows, long long N,
                    const float* __restrict__ Xf, long long chunk_cols, double* __restrict__ Ypart) {
  // X strip as float4 [half][cc][lane]: lane owns columns lane*4+cc; conflict-free 128-bit reads
  __shared__ float4 xs[2][2 * 4 * 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long rbase = ((long long)blockIdx.x * MV_WARPS + warp) * MV_ROWS;
  const long long c_begin = (long long)blockIdx.y * chunk_cols;
  long long c_end = c_begin + chunk_cols;
  if (c_end > N) c_end = N;
  const int nstrips = (int)((c_end - c_begin + MV_STRIP - 1) / MV_STRIP);

  float acc[MV_ROWS][MV_B];
#pragma unroll
  for (int r = 0; r < MV_ROWS; ++r)
#pragma unroll
    for (int j = 0; j < MV_B; ++j) acc[r][j] = 0.f;

  const float* prow[MV_ROWS];
#pragma unroll
  for (int r = 0; r < MV_ROWS; ++r) {
    long long rr = rbase + r;
    if (rr >= n_rows) rr = n_rows - 1;          // clamp: duplicates are discarded on store
    prow[r] = P + rr * ldp;
  }

  auto stage_x = [&](int buf, int s) {
    // MV_STRIP*MV_B floats = 1024 floats = 256 float4; 128 threads x 2
    const long long c0 = c_begin + (long long)s * MV_STRIP;
    for (int i = tid; i < MV_STRIP * MV_B / 4; i += MV_THREADS) {
      const int cl = i >> 1, half = i & 1;           // column within the strip, which half of the 8 vectors
      const long long col = c0 + cl;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col < c_end) v = *reinterpret_cast<const float4*>(Xf + col * MV_B + half * 4);
      xs[buf][(half * 4 + (cl & 3)) * 32 + (cl >> 2)] = v;
    }
  };

  if (nstrips > 0) stage_x(0, 0);
  __syncthreads();
  for (int s = 0; s < nstrips; ++s) {
    const int buf = s & 1;
    if (s + 1 < nstrips) stage_x(buf ^ 1, s + 1);
    const long long c = c_begin + (long long)s * MV_STRIP + lane * 4;
    float4 p[MV_ROWS];
    if (c + 3 < c_end) {
#pragma unroll
      for (int r = 0; r < MV_ROWS; ++r) p[r] = ldg_stream(reinterpret_cast<const float4*>(prow[r] + c));
    } else {
#pragma unroll
      for (int r = 0; r < MV_ROWS; ++r) {
        p[r].x = (c + 0 < c_end) ? prow[r][c + 0] : 0.f;
        p[r].y = (c + 1 < c_end) ? prow[r][c + 1] : 0.f;
        p[r].z = (c + 2 < c_end) ? prow[r][c + 2] : 0.f;
        p[r].w = (c + 3 < c_end) ? prow[r][c + 3] : 0.f;
      }
    }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
      const float4 x0 = xs[buf][cc * 32 + lane], x1 = xs[buf][(4 + cc) * 32 + lane];
#pragma unroll
      for (int r = 0; r < MV_ROWS; ++r) {
        const float pv = cc == 0 ? p[r].x : cc == 1 ? p[r].y : cc == 2 ? p[r].z : p[r].w;
        acc[r][0] = fmaf(pv, x0.x, acc[r][0]); acc[r][1] = fmaf(pv, x0.y, acc[r][1]);
        acc[r][2] = fmaf(pv, x0.z, acc[r][2]); acc[r][3] = fmaf(pv, x0.w, acc[r][3]);
        acc[r][4] = fmaf(pv, x1.x, acc[r][4]); acc[r][5] = fmaf(pv, x1.y, acc[r][5]);
        acc[r][6] = fmaf(pv, x1.z, acc[r][6]); acc[r][7] = fmaf(pv, x1.w, acc[r][7]);
      }
    }
    __syncthreads();
  }
  // warp-shuffle reduction in FP64; lane (r*8+j) % 32 keeps value (r, j)
  double* yp = Ypart + (long long)blockIdx.y * n_rows * MV_B;
#pragma unroll
  for (int r = 0; r < MV_ROWS; ++r) {
#pragma unroll
    for (int j = 0; j < MV_B; ++j) {
      const double v = warp_sum((double)acc[r][j]);
      if (lane == ((r * MV_B + j) & 31) && rbase + r < n_rows) yp[(rbase + r) * MV_B + j] = v;
    }
  }
}

__global__ void reduce_parts_kernel(const double* __restrict__ Ypart, long long n_rows, int nchunks, int b,
                                    double* __restrict__ Y) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_rows * MV_B; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / MV_B;
    const int j = (int)(e - i * MV_B);
    if (j >= b) continue;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += Ypart[(long long)c * n_rows * MV_B + e];
    Y[i * b + j] = s;
  }
}

static inline void mv_plan(long long n_rows, long long N, int* nchunks, long long* chunk_cols) {
  const long long row_groups = ceil_div64(n_rows, MV_WARPS * MV_ROWS);
  long long want = ceil_div64(6LL * kNumSMs, row_groups);     // aim for >= ~6 CTAs per SM
  long long max_chunks = N / 1024;                            // keep >= 1024 columns per chunk
  if (max_chunks < 1) max_chunks = 1;
  if (want > max_chunks) want = max_chunks;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  long long cc = ceil_div64(ceil_div64(N, want), MV_STRIP) * MV_STRIP;
  *chunk_cols = cc;
  *nchunks = (int)ceil_div64(N, cc);
}

}  // namespace v3s

using namespace v3s;

extern "C" long long v3s_block_matvec_workspace_bytes(long long n_rows, long long N) {
  int nchunks; long long cc;
  mv_plan(n_rows, N, &nchunks, &cc);
  return (long long)sizeof(float) * N * MV_B + (long long)sizeof(double) * nchunks * n_rows * MV_B + 256;
}

static int mv_run(const float* P_rows, long long ldp, long long n_rows, long long N, const float* Xf, int b,
                  double* Y_rows, double* Ypart, cudaStream_t st);

extern "C" int v3s_block_matvec(const float* P_rows, long long ldp, long long n_rows, long long N, const double* X, int b,
                                double* Y_rows, void* workspace, void* stream) {
  V3S_REQUIRE(b >= 1 && b <= MV_B, "v3s_block_matvec: block size b=%d outside [1,%d]", b, MV_B);
  V3S_REQUIRE(n_rows >= 0 && N >= 1 && ldp >= N && ldp % 4 == 0, "v3s_block_matvec: bad shape (ldp must be a multiple of 4)");
  V3S_REQUIRE(((uintptr_t)P_rows & 15) == 0, "v3s_block_matvec: P_rows must be 16-byte aligned");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  float* Xf = reinterpret_cast<float*>(workspace);
  double* Ypart = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + (((size_t)sizeof(float) * N * MV_B + 255) & ~(size_t)255));
  long long g = ceil_div64(N * MV_B, 256);
  if (g > 8LL * kNumSMs) g = 8LL * kNumSMs;
  x_to_f32_kernel<<<(int)g, 256, 0, st>>>(X, N, b, Xf);
  V3S_CHECK_LAUNCH();
  return mv_run(P_rows, ldp, n_rows, N, Xf, b, Y_rows, Ypart, st);
}

// Same product with the block already in FP32 [N,8] (the form the Lanczos step emits).
extern "C" int v3s_block_matvec_f32(const float* P_rows, long long ldp, long long n_rows, long long N, const float* Xf,
                                    double* Y_rows, void* workspace, void* stream) {
  V3S_REQUIRE(n_rows >= 0 && N >= 1 && ldp >= N && ldp % 4 == 0, "v3s_block_matvec_f32: bad shape (ldp must be a multiple of 4)");
  V3S_REQUIRE(((uintptr_t)P_rows & 15) == 0 && ((uintptr_t)Xf & 15) == 0, "v3s_block_matvec_f32: pointers must be 16-byte aligned");
  if (n_rows == 0) return 0;
  double* Ypart = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + (((size_t)sizeof(float) * N * MV_B + 255) & ~(size_t)255));
  return mv_run(P_rows, ldp, n_rows, N, Xf, MV_B, Y_rows, Ypart, (cudaStream_t)stream);
}

static int mv_run(const float* P_rows, long long ldp, long long n_rows, long long N, const float* Xf, int b,
                  double* Y_rows, double* Ypart, cudaStream_t st) {
  int nchunks; long long cc;
  mv_plan(n_rows, N, &nchunks, &cc);
  dim3 grid((unsigned)ceil_div64(n_rows, MV_WARPS * MV_ROWS), (unsigned)nchunks);
  block_matvec_kernel<<<grid, MV_THREADS, 0, st>>>(P_rows, ldp, n_rows, N, Xf, cc, Ypart);
  V3S_CHECK_LAUNCH();
  long long g2 = ceil_div64(n_rows * MV_B, 256);
  if (g2 > 8LL * kNumSMs) g2 = 8LL * kNumSMs;
  reduce_parts_kernel<<<(int)g2, 256, 0, st>>>(Ypart, n_rows, nchunks, b, Y_rows);
  V3S_CHECK_LAUNCH();
  return 0;
}
