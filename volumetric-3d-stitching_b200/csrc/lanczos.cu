// K3 glue -- one block-Lanczos step after the mat-vec, entirely on the device (FP64):
// two-pass block Gram-Schmidt against the whole basis, Cholesky-QR (twice) of the new block,
// append to the basis, emit the FP32 copy of the next mat-vec operand and the (A_j, B_j) blocks of
// the projected block-tridiagonal matrix.  One C call enqueues ~14 small kernels and never
// synchronises the host -- the replacement for the per-iteration work ARPACK does on the host
// behind scipy.sparse.linalg.eigs (diffusion.py:61).
//
// Layouts: basis V [max_dim, N] (each Lanczos vector contiguous), block W [N, 8] row-major,
// projections H [max_dim, 8].  Block size is fixed at 8 (= the mat-vec kernel's block).
#include "common.cuh"
#include "v3s.h"

namespace v3s {

constexpr int LB = 8;          // block size
constexpr int PJ_VECS = 4;     // basis vectors per CTA in the projection kernel
constexpr int PJ_T = 256;

// Hpart[split][i][0..8) = sum over the split's rows n of V[i][n] * W[n][:]
__global__ void __launch_bounds__(PJ_T)
lz_proj_kernel(const double* __restrict__ V, long long ldv, int m, const double* __restrict__ W, long long N,
               long long rows_per_split, double* __restrict__ Hpart) {
  __shared__ double red[PJ_T / 32][PJ_VECS * LB];
  const int i0 = blockIdx.x * PJ_VECS;
  const long long n0 = (long long)blockIdx.y * rows_per_split;
  long long n1 = n0 + rows_per_split;
  if (n1 > N) n1 = N;
  double acc[PJ_VECS][LB];
#pragma unroll
  for (int a = 0; a < PJ_VECS; ++a)
#pragma unroll
    for (int j = 0; j < LB; ++j) acc[a][j] = 0.0;
  const double* vrow[PJ_VECS];
#pragma unroll
  for (int a = 0; a < PJ_VECS; ++a) vrow[a] = V + (long long)(i0 + a < m ? i0 + a : m - 1) * ldv;
  for (long long n = n0 + threadIdx.x; n < n1; n += PJ_T) {
    const double2* w2 = reinterpret_cast<const double2*>(W + n * LB);
    const double2 w01 = w2[0], w23 = w2[1], w45 = w2[2], w67 = w2[3];
    const double w[LB] = {w01.x, w01.y, w23.x, w23.y, w45.x, w45.y, w67.x, w67.y};
#pragma unroll
    for (int a = 0; a < PJ_VECS; ++a) {
      const double v = vrow[a][n];
#pragma unroll
      for (int j = 0; j < LB; ++j) acc[a][j] = fma(v, w[j], acc[a][j]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < PJ_VECS; ++a)
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      const double s = warp_sum(acc[a][j]);
      if (lane == 0) red[warp][a * LB + j] = s;
    }
  __syncthreads();
  if (threadIdx.x < PJ_VECS * LB) {
    double s = 0.0;
    for (int w_ = 0; w_ < PJ_T / 32; ++w_) s += red[w_][threadIdx.x];
    const int a = threadIdx.x / LB, j = threadIdx.x % LB;
    if (i0 + a < m) Hpart[((long long)blockIdx.y * m + i0 + a) * LB + j] = s;
  }
}

// H[i][:] (+)= sum_splits Hpart ; accumulate = 0 overwrites (first pass), 1 adds (second pass)
__global__ void lz_reduce_h_kernel(const double* __restrict__ Hpart, int m, int nsplit, int accumulate,
                                   double* __restrict__ Hcur, double* __restrict__ Hsum) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= m * LB) return;
  double s = 0.0;
  for (int p = 0; p < nsplit; ++p) s += Hpart[(long long)p * m * LB + e];
  Hcur[e] = s;
  Hsum[e] = accumulate ? Hsum[e] + s : s;
}

// W[n][:] -= sum_i V[i][n] * H[i][:] ; CTA = 32 rows n x 4 warps, each warp a quarter of the basis.
// nsplit > 0: H is still in `nsplit` partial sums (Hpart, from lz_proj_kernel) and every CTA adds them up
// itself while staging H in shared memory (block 0 also stores H into Hsum: overwrite on the first
// Gram-Schmidt pass, accumulate on the second) -- one launch less per pass when the partials are small.
constexpr int UP_WARPS = 4;
__global__ void __launch_bounds__(UP_WARPS * 32)
lz_update_kernel(const double* __restrict__ V, long long ldv, int m, const double* __restrict__ H, int nsplit,
                 int accumulate, double* __restrict__ Hsum, double* __restrict__ W, long long N) {
  extern __shared__ double hs[];                         // m * 8
  __shared__ double part[UP_WARPS][32][LB + 1];
  if (nsplit > 0) {
    for (int e = threadIdx.x; e < m * LB; e += UP_WARPS * 32) {
      double sum = 0.0;
      for (int p = 0; p < nsplit; ++p) sum += H[(long long)p * m * LB + e];
      hs[e] = sum;
      if (blockIdx.x == 0) Hsum[e] = accumulate ? Hsum[e] + sum : sum;
    }
  } else {
    for (int e = threadIdx.x; e < m * LB; e += UP_WARPS * 32) hs[e] = H[e];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long n = (long long)blockIdx.x * 32 + lane;
  const bool ok = n < N;
  const int per = (m + UP_WARPS - 1) / UP_WARPS;
  const int ia = warp * per, ib = min(m, ia + per);
  double acc[LB];
#pragma unroll
  for (int j = 0; j < LB; ++j) acc[j] = 0.0;
  const double* vp = V + (ok ? n : 0);
  int i = ia;
  for (; i + 4 <= ib; i += 4) {
    const double v0 = vp[(long long)i * ldv], v1 = vp[(long long)(i + 1) * ldv], v2 = vp[(long long)(i + 2) * ldv],
                 v3 = vp[(long long)(i + 3) * ldv];
#pragma unroll
    for (int j = 0; j < LB; ++j) {
      acc[j] = fma(v0, hs[i * LB + j], acc[j]);
      acc[j] = fma(v1, hs[(i + 1) * LB + j], acc[j]);
      acc[j] = fma(v2, hs[(i + 2) * LB + j], acc[j]);
      acc[j] = fma(v3, hs[(i + 3) * LB + j], acc[j]);
    }
  }
  for (; i < ib; ++i) {
    const double v0 = vp[(long long)i * ldv];
#pragma unroll
    for (int j = 0; j < LB; ++j) acc[j] = fma(v0, hs[i * LB + j], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < LB; ++j) part[warp][lane][j] = acc[j];
  __syncthreads();
  // 128 threads finish 32 rows x 8 columns = 256 outputs
  for (int e = threadIdx.x; e < 32 * LB; e += UP_WARPS * 32) {
    const int r = e / LB, j = e % LB;
    const long long nn = (long long)blockIdx.x * 32 + r;
    if (nn < N) {
      double s = 0.0;
#pragma unroll
      for (int w_ = 0; w_ < UP_WARPS; ++w_) s += part[w_][r][j];
      W[nn * LB + j] -= s;
    }
  }
}

// 8 x 8 Cholesky-QR factors from the partial Gram sums, executed by threads t < 64 of ONE block (all
// threads of the block must call it: it synchronises).  G = sum of partials; G = L L^T (column by
// column, thread i owns row i); Linv (thread j owns column j); R = L^T.  small[0..63] = Linv
// (row-major), small[64..127] = R of this round, small[128..191] = R_total (R_round * R_total_prev when
// round > 0; thread (i,j) owns one entry).  Tblk != NULL (final round of a step): also
// Tblk[0..63] = sym(Hsum[m-8 .. m)), Tblk[64..127] = R_total, and the breakdown flag is published.
__device__ void chol8_block(const double* Gpart, int nparts, int round, double* __restrict__ small, int* __restrict__ status,
                            const double* __restrict__ Hsum, int m, double* __restrict__ Tblk, int* __restrict__ status_out) {
  __shared__ double G[LB][LB], L[LB][LB], Li[LB][LB], Rprev[LB][LB];
  __shared__ int bad;
  const int t = threadIdx.x;
  if (t == 0) bad = 0;
  if (t < 36) {
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += __ldcg(Gpart + (long long)p * 36 + t);
    int e = 0, ii = 0, jj = 0;
    for (int i = 0; i < LB; ++i) for (int j = i; j < LB; ++j) { if (e == t) { ii = i; jj = j; } ++e; }
    G[ii][jj] = s; G[jj][ii] = s;
  }
  if (t < 64) { const int i = t / LB, j = t % LB; L[i][j] = 0.0; Li[i][j] = 0.0; if (round > 0) Rprev[i][j] = small[128 + t]; }
  __syncthreads();
  // right-looking Cholesky: after step j, column j of L is final
  for (int j = 0; j < LB; ++j) {
    if (t == j) {
      double d = G[j][j];
      for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
      if (!(d > 0.0)) { bad = 1; d = 1.0; }
      L[j][j] = sqrt(d);
    }
    __syncthreads();
    if (t > j && t < LB) {
      double s = G[t][j];
      for (int k = 0; k < j; ++k) s -= L[t][k] * L[j][k];
      L[t][j] = s / L[j][j];
    }
    __syncthreads();
  }
  if (t < LB) {                                // inverse of lower-triangular L, column t
    const int j = t;
    Li[j][j] = 1.0 / L[j][j];
    for (int i = j + 1; i < LB; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s -= L[i][k] * Li[k][j];
      Li[i][j] = s / L[i][i];
    }
  }
  __syncthreads();
  if (t < 64) {
    const int i = t / LB, j = t % LB;
    small[t] = Li[i][j];
    small[64 + t] = L[j][i];                   // R = L^T
    double rt = L[j][i];
    if (round > 0) {                           // R_total = R_round * R_total_prev
      rt = 0.0;
      for (int k = 0; k < LB; ++k) rt += L[k][i] * Rprev[k][j];
    }
    small[128 + t] = rt;
    if (Tblk) {
      Tblk[t] = 0.5 * (Hsum[(m - LB + i) * LB + j] + Hsum[(m - LB + j) * LB + i]);
      Tblk[64 + t] = rt;
    }
  }
  if (t == 0) {
    if (bad) atomicExch(status, 1);
    if (status_out) *status_out = bad ? 1 : *status;
  }
}

// Gpart[cta][36] = upper triangle of W^T W over the CTA's rows; the block that finishes last (ticket
// counter) goes on to the Cholesky factors, so the Gram reduction and the 8 x 8 factorisation are one launch.
constexpr int GR_T = 256;
__global__ void __launch_bounds__(GR_T)
lz_gram8_chol_kernel(const double* __restrict__ W, long long N, double* __restrict__ Gpart, unsigned* __restrict__ ticket,
                     int round, double* __restrict__ small, int* __restrict__ status, const double* __restrict__ Hsum, int m,
                     double* __restrict__ Tblk, int* __restrict__ status_out) {
  __shared__ double red[GR_T / 32][36];
  __shared__ int is_last;
  double acc[36];
#pragma unroll
  for (int e = 0; e < 36; ++e) acc[e] = 0.0;
  for (long long n = (long long)blockIdx.x * GR_T + threadIdx.x; n < N; n += (long long)gridDim.x * GR_T) {
    const double2* w2 = reinterpret_cast<const double2*>(W + n * LB);
    const double2 a = w2[0], b = w2[1], c = w2[2], d = w2[3];
    const double w[LB] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
    int e = 0;
#pragma unroll
    for (int i = 0; i < LB; ++i)
#pragma unroll
      for (int j = i; j < LB; ++j) { acc[e] = fma(w[i], w[j], acc[e]); ++e; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < 36; ++e) {
    const double s = warp_sum(acc[e]);
    if (lane == 0) red[warp][e] = s;
  }
  __syncthreads();
  if (threadIdx.x < 36) {
    double s = 0.0;
    for (int w_ = 0; w_ < GR_T / 32; ++w_) s += red[w_][threadIdx.x];
    Gpart[(long long)blockIdx.x * 36 + threadIdx.x] = s;
    __threadfence();                            // the partial is visible before the ticket is taken
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned tk = atomicAdd(ticket, 1u);
    is_last = (tk == gridDim.x - 1);
    if (is_last) *ticket = 0;                   // ready for the next launch
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  chol8_block(Gpart, (int)gridDim.x, round, small, status, Hsum, m, Tblk, status_out);
}

// W <- W L^-T  (row n: w_j = sum_{k<=j} w_k Linv[j][k]); on the final round also append to the
// basis (V[m + j][n]) and emit the FP32 operand of the next mat-vec
__global__ void lz_scale_kernel(double* __restrict__ W, long long N, const double* __restrict__ small, int final_round,
                                double* __restrict__ Vnext, long long ldv, float* __restrict__ Xf) {
  __shared__ double Li[64];
  if (threadIdx.x < 64) Li[threadIdx.x] = small[threadIdx.x];
  __syncthreads();
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n >= N) return;
  double2* w2 = reinterpret_cast<double2*>(W + n * LB);
  const double2 a = w2[0], b = w2[1], c = w2[2], d = w2[3];
  const double w[LB] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y};
  double q[LB];
#pragma unroll
  for (int j = 0; j < LB; ++j) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k <= j; ++k) s = fma(w[k], Li[j * LB + k], s);
    q[j] = s;
  }
  w2[0] = make_double2(q[0], q[1]); w2[1] = make_double2(q[2], q[3]);
  w2[2] = make_double2(q[4], q[5]); w2[3] = make_double2(q[6], q[7]);
  if (final_round) {
    if (Vnext) {
#pragma unroll
      for (int j = 0; j < LB; ++j) Vnext[(long long)j * ldv + n] = q[j];
    }
    *reinterpret_cast<float4*>(Xf + xf_strip_index(n, 0)) = make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]);
    *reinterpret_cast<float4*>(Xf + xf_strip_index(n, 4)) = make_float4((float)q[4], (float)q[5], (float)q[6], (float)q[7]);
  }
}

// Chebyshev three-term recurrence on a block [N,8]: Y_out = alpha * PY + beta * Yk + gamma * Ykm1
// (FP64), plus the FP32 strip-layout copy the next mat-vec reads.  Y_out may alias Ykm1.
__global__ void lz_cheb_combine_kernel(const double* __restrict__ PY, const double* __restrict__ Yk,
                                       const double* Ykm1, double alpha, double beta, double gamma, double* Y_out,
                                       float* __restrict__ Xf_out, long long count) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < count; e += (long long)gridDim.x * blockDim.x) {
    double v = fma(alpha, PY[e], beta * Yk[e]);
    if (Ykm1) v = fma(gamma, Ykm1[e], v);
    Y_out[e] = v;
    if (Xf_out) Xf_out[xf_strip_index(e >> 3, (int)(e & 7))] = (float)v;
  }
}

struct LzWs {
  double *Hpart, *Hcur, *Hsum, *Gpart, *small;
  int* status;
  unsigned* ticket;
};
static inline LzWs lz_carve(double* ws, int max_dim) {
  LzWs w;
  w.Hpart = ws;
  w.Hcur = w.Hpart + 32LL * max_dim * LB;
  w.Hsum = w.Hcur + (long long)max_dim * LB;
  w.Gpart = w.Hsum + (long long)max_dim * LB;
  w.small = w.Gpart + 1024LL * 36;
  w.status = reinterpret_cast<int*>(w.small + 192);
  w.ticket = reinterpret_cast<unsigned*>(w.status + 1);
  return w;
}
}  // namespace v3s

using namespace v3s;

static inline int lz_nsplit(int m, long long N) {
  const int gx = (m + PJ_VECS - 1) / PJ_VECS;
  long long want = (2LL * kNumSMs + gx - 1) / gx;
  long long maxs = N / 2048;
  if (maxs < 1) maxs = 1;
  if (want > maxs) want = maxs;
  if (want > 32) want = 32;
  if (want < 1) want = 1;
  return (int)want;
}

extern "C" long long v3s_lanczos_workspace_doubles(long long N, int max_dim) {
  // Hpart [32][max_dim][8] + Hcur [max_dim][8] + Hsum [max_dim][8] + Gpart [1024][36] + small [192] + status
  return 32LL * max_dim * LB + 2LL * max_dim * LB + 1024LL * 36 + 192 + 8;
}

static int lz_cholqr(double* W, long long N, const LzWs& w, double* Vnext, long long ldv, float* Xf, const double* Hsum, int m,
                     double* Tblk, int* status_out, cudaStream_t st) {
  long long g = ceil_div64(N, GR_T * 4);
  if (g > 1024) g = 1024;
  if (g < 1) g = 1;
  for (int round = 0; round < 2; ++round) {
    const bool fin = round == 1;
    lz_gram8_chol_kernel<<<(int)g, GR_T, 0, st>>>(W, N, w.Gpart, w.ticket, round, w.small, w.status, Hsum, m,
                                                 fin ? Tblk : nullptr, fin ? status_out : nullptr);
    V3S_CHECK_LAUNCH();
    lz_scale_kernel<<<(int)ceil_div64(N, 256), 256, 0, st>>>(W, N, w.small, fin, Vnext, ldv, Xf);
    V3S_CHECK_LAUNCH();
  }
  return 0;
}

// Orthonormalise the start block W [N,8] in place; V[0..8) <- Q^T, Xf <- float(Q).
extern "C" int v3s_lanczos_start(double* V, long long ldv, long long N, int max_dim, double* W, float* Xf, double* ws,
                                 void* stream) {
  V3S_REQUIRE(N >= LB && ldv >= N && max_dim >= LB, "v3s_lanczos_start: bad shape");
  const LzWs w = lz_carve(ws, max_dim);
  cudaStream_t st = (cudaStream_t)stream;
  V3S_CUDA(cudaMemsetAsync(w.status, 0, 2 * sizeof(int), st));      // breakdown flag + the Gram kernel's ticket counter
  return lz_cholqr(W, N, w, V, ldv, Xf, nullptr, 0, nullptr, nullptr, st);
}

// One Lanczos step after W = P @ Q_j has been formed (and all-gathered): m = current basis size
// (multiple of 8, the last 8 rows of V hold Q_j).  On return W holds Q_{j+1}, V[m..m+8) too (when
// m + 8 <= max_dim), Xf its FP32 copy, Tblk [128] = {A_j (8x8), B_j (8x8)}; status_out != 0 after a
// Cholesky breakdown.
extern "C" int v3s_lanczos_step(double* V, long long ldv, long long N, int m, int max_dim, double* W, float* Xf,
                                double* Tblk, double* ws, int* status_out_dev, void* stream) {
  V3S_REQUIRE(m >= LB && m % LB == 0 && m <= max_dim && N >= LB && ldv >= N, "v3s_lanczos_step: bad shape (m=%d)", m);
  cudaStream_t st = (cudaStream_t)stream;
  const LzWs w = lz_carve(ws, max_dim);
  double *Hpart = w.Hpart, *Hcur = w.Hcur, *Hsum = w.Hsum;
  const int nsplit = lz_nsplit(m, N);
  const long long rows_per = ceil_div64(N, nsplit);
  const size_t up_smem = (size_t)m * LB * sizeof(double);
  // static (9 kB) + dynamic shared memory exceeds the 48 kB default from m ~ 620 on: opt in once
  static size_t up_set = 0;
  if (up_smem > up_set) {
    size_t want = (size_t)max_dim * LB * sizeof(double);
    if (want < 64 * 1024) want = 64 * 1024;
    V3S_REQUIRE(want <= 200 * 1024, "v3s_lanczos_step: max_dim=%d needs %zu B of shared memory", max_dim, want);
    V3S_CUDA(cudaFuncSetAttribute(lz_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want));
    up_set = want;
  }
  // every update CTA can add up the projection partials itself while that re-read stays small
  const bool fuse_reduce = (double)nsplit * m * LB * sizeof(double) * (double)ceil_div64(N, 32) <= 48e6;
  for (int pass = 0; pass < 2; ++pass) {
    lz_proj_kernel<<<dim3((m + PJ_VECS - 1) / PJ_VECS, nsplit), PJ_T, 0, st>>>(V, ldv, m, W, N, rows_per, Hpart);
    V3S_CHECK_LAUNCH();
    if (fuse_reduce) {
      lz_update_kernel<<<(int)ceil_div64(N, 32), UP_WARPS * 32, up_smem, st>>>(V, ldv, m, Hpart, nsplit, pass, Hsum, W, N);
    } else {
      lz_reduce_h_kernel<<<(m * LB + 255) / 256, 256, 0, st>>>(Hpart, m, nsplit, pass, Hcur, Hsum);
      V3S_CHECK_LAUNCH();
      lz_update_kernel<<<(int)ceil_div64(N, 32), UP_WARPS * 32, up_smem, st>>>(V, ldv, m, Hcur, 0, pass, Hsum, W, N);
    }
    V3S_CHECK_LAUNCH();
  }
  double* Vnext = (m + LB <= max_dim) ? V + (long long)m * ldv : nullptr;
  return lz_cholqr(W, N, w, Vnext, ldv, Xf, Hsum, m, Tblk, status_out_dev, st);
}

extern "C" int v3s_cheb_combine(const double* PY, const double* Yk, const double* Ykm1, double alpha, double beta,
                                double gamma, double* Y_out, float* Xf_out, long long count, void* stream) {
  V3S_REQUIRE(count >= 0, "v3s_cheb_combine: bad count");
  if (count == 0) return 0;
  long long g = ceil_div64(count, 256);
  if (g > 8LL * kNumSMs) g = 8LL * kNumSMs;
  lz_cheb_combine_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(PY, Yk, Ykm1, alpha, beta, gamma, Y_out, Xf_out, count);
  V3S_CHECK_LAUNCH();
  return 0;
}
