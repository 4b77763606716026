// K2c -- kNN lists -> normalised Markov matrix P_ep (row block, dense FP32) and D^-1/2.
//
// Restates diffusion.py:291-325 (s2_to_w_matrix), :89-105 (anisotropic_norm), :244-260 (dm_cov)
// without ever forming the intermediate dense W matrices or the reference's four dense-diagonal
// N^3 matmuls:
//   w_ij  = exp(-S2_ij / (eps * S2_Max))            on the thresholded kNN graph
//   q_i   = max(sum_j w_ij, 1/N)                      W1 = w_ij / (q_i q_j),  W2 = (W1 + W1^T)/2
//   d_i   = max(sum_j W2_ij, 1/N)                     P  = W2_ij / sqrt(d_i d_j)   (already symmetric)
// Bug-compatible scalars (the Python port indexes ROWS of S2 (N,k)):
//   S2_Max = median(S2[min(med_row, k-1), :]), Dist_Thr = max(S2[Index - idx_off, :]) with Python
//   negative-index wrap, and S2 is mutated in place (entries > Dist_Thr -> inf, diffusion.py:311).
// All reductions are FP64; P is rounded to FP32 once, on store.
#include "common.cuh"
#include "v3s.h"
#include <math_constants.h>

namespace v3s {

// scalars[0]=S2_Max [1]=Dist_Thr [2]=eps_eff [3]=status (0 ok, 1: S2_Max==0, 2: index ran off) [4]=Index
__global__ void __launch_bounds__(256)
s2_scalars_kernel(const double* __restrict__ S2, long long N, int k, double eps, int med_row, int idx_off,
                  double* __restrict__ scalars) {
  extern __shared__ double srt[];   // next pow2 >= k
  __shared__ double red[32];
  __shared__ double s_val;
  const int tid = threadIdx.x;
  int kp = 1;
  while (kp < k) kp <<= 1;
  // --- median of row min(med_row, k-1) ---
  const long long mrow = med_row < k - 1 ? med_row : k - 1;
  for (int i = tid; i < kp; i += 256) srt[i] = (i < k && mrow < N) ? S2[mrow * k + i] : CUDART_INF;
  __syncthreads();
  for (int size = 2; size <= kp; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < kp / 2; t += 256) {
        const int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const double a = srt[lo], b = srt[hi];
        if ((a > b) == up) { srt[lo] = b; srt[hi] = a; }
      }
      __syncthreads();
    }
  double s2max = (k & 1) ? srt[k / 2] : 0.5 * (srt[k / 2 - 1] + srt[k / 2]);
  __syncthreads();
  if (s2max == 0.0) {   // diffusion.py:300-301: fall back to the global max
    double m = -CUDART_INF;
    for (long long i = tid; i < N * k; i += 256) m = fmax(m, S2[i]);
    m = warp_max(m);
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) { double mm = red[0]; for (int w = 1; w < 8; ++w) mm = fmax(mm, red[w]); s_val = mm; }
    __syncthreads();
    s2max = s_val;
  }
  auto row_reduce = [&](long long row, bool is_max) -> double {
    double m = is_max ? -CUDART_INF : CUDART_INF;
    for (int i = tid; i < k; i += 256) { const double v = S2[row * k + i]; m = is_max ? fmax(m, v) : fmin(m, v); }
    m = is_max ? warp_max(m) : warp_min(m);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = m;
    __syncthreads();
    double mm = red[0];
    for (int w = 1; w < 8; ++w) mm = is_max ? fmax(mm, red[w]) : fmin(mm, red[w]);
    return mm;
  };
  int status = (s2max == 0.0 || mrow >= N) ? 1 : 0;
  long long Index = 0;
  double thr = CUDART_INF;
  if (k - 1 < N) {
    const double dist_max = row_reduce(k - 1, false);    // np.min(S2[d - 1, :])
    while (Index < N && row_reduce(Index, true) < dist_max) ++Index;
    if (Index >= N) status = status ? status : 2;
    else {
      long long row = Index - idx_off;
      if (row < 0) row += N;                             // Python negative index
      if (row < 0 || row >= N) status = status ? status : 2;
      else thr = row_reduce(row, true);
    }
  } else {
    status = status ? status : 2;
  }
  if (tid == 0) {
    scalars[0] = s2max; scalars[1] = thr; scalars[2] = eps * s2max; scalars[3] = (double)status; scalars[4] = (double)Index;
  }
}

__global__ void threshold_kernel(double* __restrict__ S2, long long n, const double* __restrict__ scalars) {
  const double thr = scalars[1];
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    if (S2[i] > thr) S2[i] = CUDART_INF;
}

// q_i = max(sum_j exp(-S2_ij/eps_eff), 1/N): one warp per row
__global__ void rowsum_q_kernel(const double* __restrict__ S2, long long N, int k, const double* __restrict__ scalars,
                                double* __restrict__ q) {
  const double inv_eps = 1.0 / scalars[2];
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp; i < N; i += nwarps) {
    double s = 0.0;
    for (int j = lane; j < k; j += 32) s += exp(-S2[i * k + j] * inv_eps);
    s = warp_sum(s);
    if (lane == 0) q[i] = fmax(s, 1.0 / (double)N);
  }
}

// d_i accumulates the row sums of W2 = (W1 + W1^T)/2
__global__ void colsum_d_kernel(const double* __restrict__ S2, const int* __restrict__ Nbr, long long N, int k,
                                const double* __restrict__ scalars, const double* __restrict__ q, double* __restrict__ d) {
  const double inv_eps = 1.0 / scalars[2];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < N * k; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / k;
    const int j = Nbr[e];
    const double h = 0.5 * exp(-S2[e] * inv_eps) / (q[i] * q[j]);
    if (h != 0.0) { atomicAdd(&d[i], h); atomicAdd(&d[j], h); }
  }
}

__global__ void finish_d_kernel(double* __restrict__ d, long long N, double* __restrict__ dinv_sqrt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const double v = fmax(d[i], 1.0 / (double)N);
    d[i] = v;
    dinv_sqrt[i] = 1.0 / sqrt(v);
  }
}

// P[i - row0][j] += h_ij * dis_i * dis_j for (i,j) and its transpose, restricted to the row block
__global__ void scatter_p_kernel(const double* __restrict__ S2, const int* __restrict__ Nbr, long long N, int k,
                                 const double* __restrict__ scalars, const double* __restrict__ q,
                                 const double* __restrict__ dis, long long row0, long long n_rows, float* __restrict__ P,
                                 long long ldp) {
  const double inv_eps = 1.0 / scalars[2];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < N * k; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / k;
    const long long j = Nbr[e];
    const bool own_i = (i >= row0 && i < row0 + n_rows), own_j = (j >= row0 && j < row0 + n_rows);
    if (!own_i && !own_j) continue;
    const double h = 0.5 * exp(-S2[e] * inv_eps) / (q[i] * q[j]) * dis[i] * dis[j];
    if (h == 0.0) continue;
    const float hf = (float)h;
    if (own_i) atomicAdd(&P[(i - row0) * ldp + j], hf);
    if (own_j) atomicAdd(&P[(j - row0) * ldp + i], hf);
  }
}

// ---- sparse form of P_ep: the kNN lists ARE the matrix (P = A + A^T, A_ij = h_ij on i's list) ----
// a_val[e] = g_e = exp(-S2_e/eps') / (2 q_i q_j), the entry before the D^-1/2 scaling; count[j] =
// in-degree of j over the non-zero entries (thresholded entries are exact zeros and carry no
// transposed copy).
__global__ void sp_values_kernel(const double* __restrict__ S2, const int* __restrict__ Nbr, long long N, int k,
                                 const double* __restrict__ scalars, const double* __restrict__ q,
                                 double* __restrict__ a_val, int* __restrict__ count) {
  const double inv_eps = 1.0 / scalars[2];
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < N * k; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / k;
    const long long j = Nbr[e];
    const double g = 0.5 * exp(-S2[e] * inv_eps) / (q[i] * q[j]);
    a_val[e] = g;
    if (g != 0.0) atomicAdd(&count[j], 1);
  }
}

// d_i = max(sum of row i of W2 = list part + transposed part, 1/N), dinv_sqrt_i = d_i^-1/2: one warp
// per row, fixed order (unlike the dense build's atomics this is reproducible to the last bit)
__global__ void sp_rowsum_d_kernel(const double* __restrict__ a_val, int k, const int* __restrict__ t_ptr,
                                   const double* __restrict__ t_val, long long N, double* __restrict__ d,
                                   double* __restrict__ dinv_sqrt) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp; i < N; i += nwarps) {
    double s = 0.0;
    for (int t = lane; t < k; t += 32) s += a_val[i * k + t];
    const int p0 = t_ptr[i], p1 = t_ptr[i + 1];
    for (int e = p0 + lane; e < p1; e += 32) s += t_val[e];
    s = warp_sum(s);
    if (lane == 0) {
      const double v = fmax(s, 1.0 / (double)N);
      d[i] = v;
      dinv_sqrt[i] = 1.0 / sqrt(v);
    }
  }
}

// h = (g dis_i) dis_j on both copies of every entry (the expression and order of scatter_p_kernel)
__global__ void sp_scale_kernel(double* __restrict__ a_val, const int* __restrict__ Nbr, int k, long long N,
                                const int* __restrict__ t_ptr, const int* __restrict__ t_col, double* __restrict__ t_val,
                                const double* __restrict__ dis) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp; i < N; i += nwarps) {
    const double di = dis[i];
    for (int t = lane; t < k; t += 32) a_val[i * k + t] = a_val[i * k + t] * di * dis[Nbr[i * k + t]];
    const int p0 = t_ptr[i], p1 = t_ptr[i + 1];
    for (int e = p0 + lane; e < p1; e += 32) t_val[e] = t_val[e] * dis[t_col[e]] * di;
  }
}

// exclusive scan of count[N] -> ptr[N+1]; one block, each thread owns a contiguous chunk
__global__ void __launch_bounds__(1024) sp_scan_kernel(const int* __restrict__ count, long long N, int* __restrict__ ptr) {
  __shared__ int wsum[32];
  __shared__ int total_before;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long chunk = (N + 1023) / 1024;
  const long long b0 = tid * chunk, b1 = b0 + chunk < N ? b0 + chunk : N;
  int s = 0;
  for (long long i = b0; i < b1; ++i) s += count[i];
  int inc = s;                                           // inclusive scan of the per-thread sums
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = wsum[lane], iv = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, iv, o); if (lane >= o) iv += t; }
    wsum[lane] = iv - v;                                 // exclusive offsets of the warps
    if (lane == 31) total_before = iv;
  }
  __syncthreads();
  int run = wsum[w] + inc - s;
  for (long long i = b0; i < b1; ++i) { ptr[i] = run; run += count[i]; }
  if (tid == 0) ptr[N] = total_before;
}

// entry e = (i -> j) lands somewhere in row j of the transposed list (order fixed by the sort below)
__global__ void sp_fill_kernel(const int* __restrict__ Nbr, const double* __restrict__ a_val, long long nk,
                               const int* __restrict__ ptr, int* __restrict__ cursor, int* __restrict__ tmp) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < nk; e += (long long)gridDim.x * blockDim.x) {
    if (a_val[e] == 0.0) continue;
    const int j = Nbr[e];
    tmp[ptr[j] + atomicAdd(&cursor[j], 1)] = (int)e;
  }
}

// Row j of the transposed list in ascending source row (entry indices e = i k + t are unique per
// source row i, so ordering by e orders by i): rank by counting, one block per row.  Makes the
// summation order of the mat-vec -- and with it every result -- independent of the atomics above.
constexpr int SP_SORT_THREADS = 128;
constexpr int SP_SORT_CHUNK = 2048;
__global__ void __launch_bounds__(SP_SORT_THREADS)
sp_sort_rows_kernel(const int* __restrict__ ptr, const int* __restrict__ tmp, const double* __restrict__ a_val, int k,
                    long long N, int* __restrict__ t_col, double* __restrict__ t_val) {
  __shared__ int sh[SP_SORT_CHUNK];
  for (long long j = blockIdx.x; j < N; j += gridDim.x) {
    const int p0 = ptr[j], d = ptr[j + 1] - p0;
    for (int m0 = 0; m0 < d; m0 += SP_SORT_THREADS) {
      const int mi = m0 + threadIdx.x;
      const int x = mi < d ? tmp[p0 + mi] : 0x7fffffff;
      int rank = 0;
      for (int c0 = 0; c0 < d; c0 += SP_SORT_CHUNK) {
        const int cn = d - c0 < SP_SORT_CHUNK ? d - c0 : SP_SORT_CHUNK;
        __syncthreads();
        for (int t = threadIdx.x; t < cn; t += SP_SORT_THREADS) sh[t] = tmp[p0 + c0 + t];
        __syncthreads();
        if (mi < d)
          for (int t = 0; t < cn; ++t) rank += (sh[t] < x);
      }
      if (mi < d) { t_col[p0 + rank] = x / k; t_val[p0 + rank] = a_val[x]; }
    }
  }
}

// ---- dense FP64 forms of the three reference functions (API fidelity at small N) ----------
__global__ void w_dense_scatter_kernel(const double* __restrict__ S2, const int* __restrict__ Nbr, long long N, int k,
                                       double eps_eff, double* __restrict__ W) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < N * k; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / k;
    atomicAdd(&W[i * N + Nbr[e]], exp(-S2[e] / eps_eff));   // duplicates sum, like csr_array(...).toarray()
  }
}
// sums over axis: axis=1 -> row sums, axis=0 -> column sums; clamp to 1/N
__global__ void dense_sums_kernel(const double* __restrict__ W, long long N, int axis, double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp; i < N; i += nwarps) {
    double s = 0.0;
    for (long long j = lane; j < N; j += 32) s += axis ? W[i * N + j] : W[j * N + i];
    s = warp_sum(s);
    if (lane == 0) out[i] = fmax(s, 1.0 / (double)N);
  }
}
// out_ij = ( s_i W_ij s_j + s_j W_ji s_i ) / 2   with s = 1/v (mode 0) or 1/sqrt(v) (mode 1)
__global__ void dense_scale_sym_kernel(const double* __restrict__ W, long long N, const double* __restrict__ v, int mode,
                                       double* __restrict__ out) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < N * N; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / N, j = e - i * N;
    const double si = mode ? 1.0 / sqrt(v[i]) : 1.0 / v[i], sj = mode ? 1.0 / sqrt(v[j]) : 1.0 / v[j];
    const double a = (si * W[i * N + j]) * sj, b = (sj * W[j * N + i]) * si;
    out[e] = (a + b) / 2;
  }
}
__global__ void inv_sqrt_kernel(const double* __restrict__ v, long long N, double* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
    out[i] = 1.0 / sqrt(v[i]);
}

static inline int grid_for(long long n, int threads) {
  long long g = (n + threads - 1) / threads;
  const long long cap = 32LL * kNumSMs;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace v3s

using namespace v3s;

extern "C" int v3s_s2_scalars(const double* S2, long long N, int k, double eps, int med_row, int idx_off,
                              double* scalars5, void* stream) {
  V3S_REQUIRE(N >= 1 && k >= 1 && k <= 4096, "v3s_s2_scalars: need N>=1 and 1<=k<=4096 (got N=%lld k=%d)", N, k);
  int kp = 1;
  while (kp < k) kp <<= 1;
  s2_scalars_kernel<<<1, 256, kp * sizeof(double), (cudaStream_t)stream>>>(S2, N, k, eps, med_row, idx_off, scalars5);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_s2_threshold(double* S2, long long N, int k, const double* scalars5, void* stream) {
  threshold_kernel<<<grid_for(N * k, 256), 256, 0, (cudaStream_t)stream>>>(S2, N * k, scalars5);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_build_markov(const double* S2_thresholded, const int* Nbr, long long N, int k, const double* scalars5,
                                long long row0, long long n_rows, float* P_rows, long long ldp, double* dinv_sqrt,
                                double* ws_2N, void* stream) {
  V3S_REQUIRE(N >= 1 && k >= 1 && row0 >= 0 && n_rows >= 0 && row0 + n_rows <= N && ldp >= N, "v3s_build_markov: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  double* q = ws_2N;
  double* d = ws_2N + N;
  V3S_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * N, st));
  rowsum_q_kernel<<<grid_for(N * 32, 256), 256, 0, st>>>(S2_thresholded, N, k, scalars5, q);
  V3S_CHECK_LAUNCH();
  colsum_d_kernel<<<grid_for(N * k, 256), 256, 0, st>>>(S2_thresholded, Nbr, N, k, scalars5, q, d);
  V3S_CHECK_LAUNCH();
  finish_d_kernel<<<grid_for(N, 256), 256, 0, st>>>(d, N, dinv_sqrt);
  V3S_CHECK_LAUNCH();
  if (n_rows > 0 && P_rows) {
    V3S_CUDA(cudaMemsetAsync(P_rows, 0, sizeof(float) * (size_t)n_rows * (size_t)ldp, st));
    scatter_p_kernel<<<grid_for(N * k, 256), 256, 0, st>>>(S2_thresholded, Nbr, N, k, scalars5, q, dinv_sqrt, row0, n_rows,
                                                          P_rows, ldp);
    V3S_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int v3s_build_markov_sparse(const double* S2_thresholded, const int* Nbr, long long N, int k,
                                       const double* scalars5, double* a_val, int* t_ptr, int* t_col, double* t_val,
                                       double* dinv_sqrt, double* ws_2N, int* iws, void* stream) {
  V3S_REQUIRE(N >= 1 && k >= 1 && N * (long long)k < (1LL << 31), "v3s_build_markov_sparse: need N>=1, k>=1 and N*k < 2^31");
  V3S_REQUIRE(a_val && t_ptr && t_col && t_val && dinv_sqrt && ws_2N && iws, "v3s_build_markov_sparse: null buffer");
  cudaStream_t st = (cudaStream_t)stream;
  double* q = ws_2N;
  double* d = ws_2N + N;
  int* count = iws;            // [N]   in-degrees, then the fill cursors
  int* tmp = iws + N;          // [N k] entry indices in arrival order
  V3S_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * N, st));
  rowsum_q_kernel<<<grid_for(N * 32, 256), 256, 0, st>>>(S2_thresholded, N, k, scalars5, q);
  V3S_CHECK_LAUNCH();
  sp_values_kernel<<<grid_for(N * k, 256), 256, 0, st>>>(S2_thresholded, Nbr, N, k, scalars5, q, a_val, count);
  V3S_CHECK_LAUNCH();
  sp_scan_kernel<<<1, 1024, 0, st>>>(count, N, t_ptr);
  V3S_CHECK_LAUNCH();
  V3S_CUDA(cudaMemsetAsync(count, 0, sizeof(int) * N, st));
  sp_fill_kernel<<<grid_for(N * k, 256), 256, 0, st>>>(Nbr, a_val, N * k, t_ptr, count, tmp);
  V3S_CHECK_LAUNCH();
  sp_sort_rows_kernel<<<grid_for(N * SP_SORT_THREADS, SP_SORT_THREADS), SP_SORT_THREADS, 0, st>>>(t_ptr, tmp, a_val, k, N,
                                                                                                t_col, t_val);
  V3S_CHECK_LAUNCH();
  sp_rowsum_d_kernel<<<grid_for(N * 32, 256), 256, 0, st>>>(a_val, k, t_ptr, t_val, N, d, dinv_sqrt);
  V3S_CHECK_LAUNCH();
  sp_scale_kernel<<<grid_for(N * 32, 256), 256, 0, st>>>(a_val, Nbr, k, N, t_ptr, t_col, t_val, dinv_sqrt);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_w_dense(const double* S2_thresholded, const int* Nbr, long long N, int k, double eps_eff, double* W,
                           void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  V3S_CUDA(cudaMemsetAsync(W, 0, sizeof(double) * (size_t)N * (size_t)N, st));
  w_dense_scatter_kernel<<<grid_for(N * k, 256), 256, 0, st>>>(S2_thresholded, Nbr, N, k, eps_eff, W);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_anisotropic_norm_dense(const double* W, long long N, double* out, double* ws_N, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  dense_sums_kernel<<<grid_for(N * 32, 256), 256, 0, st>>>(W, N, 1, ws_N);
  V3S_CHECK_LAUNCH();
  dense_scale_sym_kernel<<<grid_for(N * N, 256), 256, 0, st>>>(W, N, ws_N, 0, out);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_dm_cov_dense(const double* W, long long N, double* P_out, double* dinv_sqrt, double* ws_N, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  dense_sums_kernel<<<grid_for(N * 32, 256), 256, 0, st>>>(W, N, 0, ws_N);
  V3S_CHECK_LAUNCH();
  dense_scale_sym_kernel<<<grid_for(N * N, 256), 256, 0, st>>>(W, N, ws_N, 1, P_out);
  V3S_CHECK_LAUNCH();
  inv_sqrt_kernel<<<grid_for(N, 256), 256, 0, st>>>(ws_N, N, dinv_sqrt);
  V3S_CHECK_LAUNCH();
  return 0;
}
