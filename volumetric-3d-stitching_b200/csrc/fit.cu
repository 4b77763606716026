// K4 -- orientation retrieval: linear fit of the 9 non-trivial diffusion coordinates to the 9
// rotation-matrix elements, per-pattern 3x3 projection, quaternions, all-pairs geodesic error.
//
// Restates compute.py:103-112 (lsq_linear: M = ((b^T A) pinv(A^T A))^T), diffusion.py:144-174
// (Recon = X c; per-row polar_decompose; rot_mat_to_quat for estimate and truth),
// compute.py:51-55 (polar_decompose), rotation.py:165-210 (R -> axis -> quaternion) and
// rotation.py:214-244 (Sigma_T = 2 sum |acos(clip(q1_i.q1_j,0,1)) - acos(clip(q2_i.q2_j,0,1))| / (N(N-1))).
// Everything is FP64.  polar_mode: 0 = reference-compat U.Vh^T (compute.py:53, depends on the
// SVD sign convention -- ours is Jacobi's, not LAPACK's), 1 = polar factor
// U.Vh (src/octave/PolarDecompose.m), 2 = nearest rotation (det fixed to +1).
#include "common.cuh"
#include <type_traits>
#include "v3s.h"

namespace v3s {

constexpr int FIT_ROWS = 128;

// partial sums of X^T X (81) and X^T Y (81); X = Ps[:, 1:10] (ld = ldx), Y = R^T rows [N, 9]
__global__ void __launch_bounds__(192)
fit_gram_kernel(const double* __restrict__ Ps, int ldx, const double* __restrict__ Y, long long n_rows,
                double* __restrict__ out162) {
  __shared__ double sx[FIT_ROWS][9], sy[FIT_ROWS][9];
  const int tid = threadIdx.x;
  double acc = 0.0;
  const int which = tid / 81, e = tid % 81, a = e / 9, b = e % 9;   // tid < 162 active
  for (long long r0 = (long long)blockIdx.x * FIT_ROWS; r0 < n_rows; r0 += (long long)gridDim.x * FIT_ROWS) {
    __syncthreads();
    for (int i = tid; i < FIT_ROWS * 9; i += blockDim.x) {
      const int rr = i / 9, cc = i % 9;
      const long long r = r0 + rr;
      sx[rr][cc] = r < n_rows ? Ps[r * ldx + 1 + cc] : 0.0;
      sy[rr][cc] = r < n_rows ? Y[r * 9 + cc] : 0.0;
    }
    __syncthreads();
    if (tid < 162) {
      if (which == 0) for (int rr = 0; rr < FIT_ROWS; ++rr) acc += sx[rr][a] * sx[rr][b];
      else            for (int rr = 0; rr < FIT_ROWS; ++rr) acc += sx[rr][a] * sy[rr][b];
    }
  }
  if (tid < 162) atomicAdd(&out162[tid], acc);
}

// Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 9), single thread.
__device__ void jacobi_sym(double* A, double* V, int n) {
  for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) V[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; ++i) { diag += A[i * n + i] * A[i * n + i]; for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j]; }
    if (off <= 1e-60 || off <= 1e-34 * diag) break;
    for (int p = 0; p < n - 1; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A[p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq; A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk; A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq; V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
}

// The same cyclic Jacobi iteration on a 9 x 9 matrix in shared memory, executed by one warp: the rotation
// parameters are computed redundantly by every lane, lane k < 9 applies the rotation to row / column k.
// Same rotations in the same order and the same arithmetic per element as jacobi_sym (bit-identical).
__device__ void jacobi_sym9_warp(double (*A)[9], double (*V)[9]) {
  const int k = threadIdx.x;
  if (k < 9) for (int j = 0; j < 9; ++j) V[k][j] = (k == j) ? 1.0 : 0.0;
  __syncwarp();
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < 9; ++i) { diag += A[i][i] * A[i][i]; for (int j = i + 1; j < 9; ++j) off += A[i][j] * A[i][j]; }
    if (off <= 1e-60 || off <= 1e-34 * diag) break;
    for (int p = 0; p < 8; ++p)
      for (int q = p + 1; q < 9; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;                      // uniform across the warp
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
        __syncwarp();
        if (k < 9) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - sn * akq; A[k][q] = sn * akp + c * akq;
        }
        __syncwarp();
        if (k < 9) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - sn * aqk; A[q][k] = sn * apk + c * aqk;
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - sn * vkq; V[k][q] = sn * vkp + c * vkq;
        }
        __syncwarp();
      }
  }
  __syncwarp();
}

// c = pinv(X^T X) (X^T Y); numpy's pinv cut-off: singular values <= 1e-15 * max are dropped.  One warp.
__global__ void fit_solve_kernel(const double* __restrict__ in162, double* __restrict__ c81) {
  __shared__ double A[9][9], V[9][9], pinv[9][9];
  const int tid = threadIdx.x;
  if (blockIdx.x != 0) return;
  for (int e = tid; e < 81; e += 32) A[e / 9][e % 9] = in162[e];
  __syncwarp();
  if (tid < 9)
    for (int j = tid + 1; j < 9; ++j) { const double m = 0.5 * (A[tid][j] + A[j][tid]); A[tid][j] = m; A[j][tid] = m; }
  __syncwarp();
  jacobi_sym9_warp(A, V);
  double wmax = 0.0;
  for (int i = 0; i < 9; ++i) wmax = fmax(wmax, fabs(A[i][i]));
  for (int e = tid; e < 81; e += 32) {
    const int i = e / 9, j = e % 9;
    double s = 0.0;
    for (int k = 0; k < 9; ++k) {
      const double w = A[k][k];
      if (fabs(w) > 1e-15 * wmax) s += V[i][k] * V[j][k] / w;
    }
    pinv[i][j] = s;
  }
  __syncwarp();
  const double* xty = in162 + 81;
  for (int e = tid; e < 81; e += 32) {
    const int i = e / 9, j = e % 9;
    double s = 0.0;
    for (int k = 0; k < 9; ++k) s += pinv[i][k] * xty[k * 9 + j];
    c81[e] = s;
  }
}

__device__ __forceinline__ void quat_from_rot(const double* R, double* q) {
  // rotation.py:165-202
  const double x = R[7] - R[5], y = R[2] - R[6], z = R[3] - R[1];
  const double s = sqrt(x * x + y * y + z * z);
  if (s != 0.0) {
    const double theta = atan2(s, R[0] + R[4] + R[8] - 1.0);
    const double f = theta / s;
    const double ax = f * x, ay = f * y, az = f * z;
    const double th = sqrt(ax * ax + ay * ay + az * az);
    if (th == 0.0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
    const double g = sin(th / 2) / th;
    q[0] = cos(th / 2); q[1] = g * ax; q[2] = g * ay; q[3] = g * az;
  } else {
    q[0] = 1; q[1] = q[2] = q[3] = 0;
  }
}

__global__ void __launch_bounds__(128)
fit_project_kernel(const double* __restrict__ Ps, int ldx, const double* __restrict__ Y, long long n_rows,
                   const double* __restrict__ c81, int polar_mode, double* __restrict__ recon_pre,
                   double* __restrict__ rproj, double* __restrict__ quat_est, double* __restrict__ quat_true) {
  __shared__ double sc[81];
  if (threadIdx.x < 81) sc[threadIdx.x] = c81[threadIdx.x];
  __syncthreads();
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
    double x[9], M[9];
    for (int a = 0; a < 9; ++a) x[a] = Ps[r * ldx + 1 + a];
    for (int b = 0; b < 9; ++b) {
      double s = 0.0;
      for (int a = 0; a < 9; ++a) s += x[a] * sc[a * 9 + b];
      M[b] = s;
      if (recon_pre) recon_pre[r * 9 + b] = s;
    }
    // SVD of M (row-major 3x3) through the eigen-decomposition of M^T M
    double S[9], V[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) S[i * 3 + j] = M[0 + i] * M[0 + j] + M[3 + i] * M[3 + j] + M[6 + i] * M[6 + j];
    jacobi_sym(S, V, 3);
    // order singular values descending
    int ord[3] = {0, 1, 2};
    double w[3] = {S[0], S[4], S[8]};
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2 - i; ++j) if (w[ord[j]] < w[ord[j + 1]]) { const int t = ord[j]; ord[j] = ord[j + 1]; ord[j + 1] = t; }
    double Vs[9], U[9];
    for (int k = 0; k < 3; ++k) for (int i = 0; i < 3; ++i) Vs[i * 3 + k] = V[i * 3 + ord[k]];
    // U columns: u_k = M v_k / sigma_k; complete degenerate directions by orthogonalisation
    double sig[3];
    for (int k = 0; k < 3; ++k) {
      double u[3];
      for (int i = 0; i < 3; ++i) u[i] = M[i * 3 + 0] * Vs[0 * 3 + k] + M[i * 3 + 1] * Vs[1 * 3 + k] + M[i * 3 + 2] * Vs[2 * 3 + k];
      sig[k] = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
      for (int i = 0; i < 3; ++i) U[i * 3 + k] = u[i];
    }
    const double tiny = 1e-13 * fmax(sig[0], 1e-300);
    if (sig[0] > tiny) for (int i = 0; i < 3; ++i) U[i * 3 + 0] /= sig[0]; else { U[0] = 1; U[3] = 0; U[6] = 0; }
    if (sig[1] > tiny) {
      for (int i = 0; i < 3; ++i) U[i * 3 + 1] /= sig[1];
    } else {   // any unit vector orthogonal to u0
      const double a0 = U[0], a1 = U[3], a2 = U[6];
      double t[3] = {0, 0, 0};
      if (fabs(a0) <= fabs(a1) && fabs(a0) <= fabs(a2)) t[0] = 1; else if (fabs(a1) <= fabs(a2)) t[1] = 1; else t[2] = 1;
      const double dt = t[0] * a0 + t[1] * a1 + t[2] * a2;
      double v0 = t[0] - dt * a0, v1 = t[1] - dt * a1, v2 = t[2] - dt * a2;
      const double nv = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
      U[1] = v0 / nv; U[4] = v1 / nv; U[7] = v2 / nv;
    }
    if (sig[2] > tiny) {
      for (int i = 0; i < 3; ++i) U[i * 3 + 2] /= sig[2];
    } else {   // u2 = u0 x u1, sign chosen so that det(U) = det(V)
      double c0 = U[3] * U[7] - U[6] * U[4], c1 = U[6] * U[1] - U[0] * U[7], c2 = U[0] * U[4] - U[3] * U[1];
      const double detV = Vs[0] * (Vs[4] * Vs[8] - Vs[5] * Vs[7]) - Vs[1] * (Vs[3] * Vs[8] - Vs[5] * Vs[6]) + Vs[2] * (Vs[3] * Vs[7] - Vs[4] * Vs[6]);
      const double sgn = detV < 0 ? -1.0 : 1.0;
      U[2] = sgn * c0; U[5] = sgn * c1; U[8] = sgn * c2;
    }
    double Rp[9];
    if (polar_mode == 0) {            // U @ Vh.T == U @ V   (compute.py:53)
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rp[i * 3 + j] = U[i * 3 + 0] * Vs[0 * 3 + j] + U[i * 3 + 1] * Vs[1 * 3 + j] + U[i * 3 + 2] * Vs[2 * 3 + j];
    } else {                          // U @ Vh == U @ V^T
      double d3 = 1.0;
      if (polar_mode == 2) {
        const double detU = U[0] * (U[4] * U[8] - U[5] * U[7]) - U[1] * (U[3] * U[8] - U[5] * U[6]) + U[2] * (U[3] * U[7] - U[4] * U[6]);
        const double detV = Vs[0] * (Vs[4] * Vs[8] - Vs[5] * Vs[7]) - Vs[1] * (Vs[3] * Vs[8] - Vs[5] * Vs[6]) + Vs[2] * (Vs[3] * Vs[7] - Vs[4] * Vs[6]);
        d3 = (detU * detV < 0) ? -1.0 : 1.0;
      }
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Rp[i * 3 + j] = U[i * 3 + 0] * Vs[j * 3 + 0] + U[i * 3 + 1] * Vs[j * 3 + 1] + d3 * U[i * 3 + 2] * Vs[j * 3 + 2];
    }
    if (rproj) for (int i = 0; i < 9; ++i) rproj[r * 9 + i] = Rp[i];
    double q[4];
    quat_from_rot(Rp, q);
    for (int i = 0; i < 4; ++i) quat_est[r * 4 + i] = q[i];
    double Rt[9];
    for (int i = 0; i < 9; ++i) Rt[i] = Y[r * 9 + i];
    quat_from_rot(Rt, q);
    for (int i = 0; i < 4; ++i) quat_true[r * 4 + i] = q[i];
  }
}

// sum over i in [row0,row0+n_rows), j in [0,N) of |acos(clip(q1_i.q1_j)) - acos(clip(q2_i.q2_j))|.
// The summand is symmetric in (i,j) and zero on the diagonal, so each unordered pair is evaluated
// once and counted twice.  Who evaluates it is decided on 32-index blocks (bi = i/32, bj = j/32) so
// the choice is uniform across a warp: bi != bj -> the row whose block satisfies (bi < bj and bi+bj
// even) or (bi > bj and bi+bj odd) (balanced over row blocks); bi == bj -> the row with i < j.
constexpr int SG_T = 256;
// FP64MATH = true: the reference's arithmetic (FP64 dot products and acos).  false (N > kSigmaFp64MaxN): the N^2
// pair terms in FP32 (dot, clip, acosf) accumulated in FP64 -- 1e10 double-precision acos at N = 10^5 were
// 10 % of the whole pass for a statistic the reference itself prints with 2 digits; per-term error <= 2e-7 rad
// except for the (measure ~theta^3) pairs of nearly identical orientations, mean error of the sum ~1e-7 relative.
constexpr long long kSigmaFp64MaxN = 20000;
template <bool FP64MATH>
__global__ void __launch_bounds__(SG_T)
sigma_pairs_kernel(const double* __restrict__ q1, const double* __restrict__ q2, long long N, long long row0,
                   long long n_rows, double* __restrict__ partials) {
  using T = typename std::conditional<FP64MATH, double, float>::type;
  __shared__ T sj1[SG_T][4], sj2[SG_T][4];
  __shared__ double red[32];
  const int tid = threadIdx.x;
  const long long i = row0 + (long long)blockIdx.x * SG_T + tid;
  const bool i_ok = i < row0 + n_rows;
  T a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  if (i_ok) for (int t = 0; t < 4; ++t) { a1[t] = (T)q1[i * 4 + t]; a2[t] = (T)q2[i * 4 + t]; }
  double acc = 0.0;
  const long long j_per = ceil_div64(N, gridDim.y);
  const long long j0 = (long long)blockIdx.y * j_per;
  long long j1 = j0 + j_per;
  if (j1 > N) j1 = N;
  for (long long jb = j0; jb < j1; jb += SG_T) {
    __syncthreads();
    const long long j = jb + tid;
    for (int t = 0; t < 4; ++t) { sj1[tid][t] = j < j1 ? (T)q1[j * 4 + t] : T(0); sj2[tid][t] = j < j1 ? (T)q2[j * 4 + t] : T(0); }
    __syncthreads();
    const int lim = (int)((j1 - jb) < SG_T ? (j1 - jb) : SG_T);
    if (i_ok) {
      T part = 0;                                  // FP32 mode: 256 terms per partial, then into the FP64 accumulator
      for (int jj = 0; jj < lim; ++jj) {
        const long long jg = jb + jj;
        const long long bi = i >> 5, bj = jg >> 5;
        const bool mine = (bi == bj) ? (i < jg) : ((bi < bj) ? (((bi + bj) & 1) == 0) : (((bi + bj) & 1) == 1));
        if (!mine) continue;
        T d1 = a1[0] * sj1[jj][0] + a1[1] * sj1[jj][1] + a1[2] * sj1[jj][2] + a1[3] * sj1[jj][3];
        T d2 = a2[0] * sj2[jj][0] + a2[1] * sj2[jj][1] + a2[2] * sj2[jj][2] + a2[3] * sj2[jj][3];
        if constexpr (FP64MATH) {
          d1 = fmin(fmax(d1, 0.0), 1.0);
          d2 = fmin(fmax(d2, 0.0), 1.0);
          acc += 2.0 * fabs(acos(d1) - acos(d2));
        } else {
          d1 = fminf(fmaxf(d1, 0.f), 1.f);
          d2 = fminf(fmaxf(d2, 0.f), 1.f);
          part += fabsf(acosf(d1) - acosf(d2));
        }
      }
      if constexpr (!FP64MATH) acc += 2.0 * (double)part;
    }
  }
  acc = block_sum(acc, red);
  if (tid == 0) partials[(long long)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

// The FP32 form for large N, restructured: (1) ownership of a pair is decided once per 32 x 32 block of (i, j) -- rows
// are mapped to threads from a 32-aligned base, so the decision is uniform across a warp and a block that belongs to
// the transposed side is skipped without touching its elements; (2) two columns per instruction: the j-side
// quaternions sit in shared memory as float2 pairs (j, j+1) and dots, clip-free polynomial and products run as packed
// FP32 (FFMA2 / FADD2 / FMUL2); (3) acos(x) on [0,1] as sqrt(1-x) * p7(x) (Abramowitz & Stegun 4.4.46, |error| <= 2e-8)
// with the approximate square root (1 ulp) instead of the library acosf.  Per-term error ~1e-7 rad like the form above.
__device__ __forceinline__ float2 sg_acos01_2(float2 x) {
  const float2 one = make_float2(1.f, 1.f);
  float2 p = make_float2(-0.0012624911f, -0.0012624911f);
  p = __ffma2_rn(p, x, make_float2(0.0066700901f, 0.0066700901f));
  p = __ffma2_rn(p, x, make_float2(-0.0170881256f, -0.0170881256f));
  p = __ffma2_rn(p, x, make_float2(0.0308918810f, 0.0308918810f));
  p = __ffma2_rn(p, x, make_float2(-0.0501743046f, -0.0501743046f));
  p = __ffma2_rn(p, x, make_float2(0.0889789874f, 0.0889789874f));
  p = __ffma2_rn(p, x, make_float2(-0.2145988016f, -0.2145988016f));
  p = __ffma2_rn(p, x, make_float2(1.5707963050f, 1.5707963050f));
  const float2 y = __ffma2_rn(x, make_float2(-1.f, -1.f), one);      // 1 - x >= 0 after the clip
  float sx, sy;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sx) : "f"(y.x));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sy) : "f"(y.y));
  return __fmul2_rn(make_float2(sx, sy), p);
}
__global__ void __launch_bounds__(SG_T)
sigma_pairs_f32_kernel(const double* __restrict__ q1, const double* __restrict__ q2, long long N, long long row0,
                       long long n_rows, long long i_base, long long j_per, double* __restrict__ partials) {
  // sj[q][pair][t] = (component t of quaternion j = 2 pair, of j = 2 pair + 1); q = 0: estimated, 1: true
  __shared__ __align__(16) float2 sj[2][SG_T / 2][4];
  __shared__ double red[32];
  const int tid = threadIdx.x, lane = tid & 31;
  const long long i = i_base + (long long)blockIdx.x * SG_T + tid;       // i_base is a multiple of 32
  const bool i_ok = i >= row0 && i < row0 + n_rows;
  float a1[4] = {0, 0, 0, 0}, a2[4] = {0, 0, 0, 0};
  if (i_ok) for (int t = 0; t < 4; ++t) { a1[t] = (float)q1[i * 4 + t]; a2[t] = (float)q2[i * 4 + t]; }
  const long long bi = i >> 5;                                           // uniform across the warp
  double acc = 0.0;
  const long long j0 = (long long)blockIdx.y * j_per;                    // j_per is a multiple of SG_T
  long long j1 = j0 + j_per;
  if (j1 > N) j1 = N;
  for (long long jb = j0; jb < j1; jb += SG_T) {
    __syncthreads();
    {
      const long long j = jb + tid;
      float* dst1 = reinterpret_cast<float*>(&sj[0][tid >> 1][0]) + (tid & 1);
      float* dst2 = reinterpret_cast<float*>(&sj[1][tid >> 1][0]) + (tid & 1);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        dst1[2 * t] = j < j1 ? (float)q1[j * 4 + t] : 0.f;
        dst2[2 * t] = j < j1 ? (float)q2[j * 4 + t] : 0.f;
      }
    }
    __syncthreads();
    float2 part = make_float2(0.f, 0.f);           // <= 256 terms per partial, then into the FP64 accumulator
    for (int blk = 0; blk < SG_T / 32; ++blk) {
      const long long jblk = jb + 32 * blk;
      if (jblk >= j1) break;
      const long long bj = jblk >> 5;
      const bool diag = bi == bj;
      if (!diag && !((bi < bj) ? (((bi + bj) & 1) == 0) : (((bi + bj) & 1) == 1))) continue;   // the other side's block
      const int n_in = (int)((j1 - jblk) < 32 ? (j1 - jblk) : 32);
#pragma unroll 4
      for (int pp = 0; pp < 16; ++pp) {
        const int pr = blk * 16 + pp;
        const float4 u01 = *reinterpret_cast<const float4*>(&sj[0][pr][0]);   // (t0.j, t0.j+1, t1.j, t1.j+1)
        const float4 u23 = *reinterpret_cast<const float4*>(&sj[0][pr][2]);
        const float4 v01 = *reinterpret_cast<const float4*>(&sj[1][pr][0]);
        const float4 v23 = *reinterpret_cast<const float4*>(&sj[1][pr][2]);
        float2 d1 = __fmul2_rn(make_float2(a1[0], a1[0]), make_float2(u01.x, u01.y));
        d1 = __ffma2_rn(make_float2(a1[1], a1[1]), make_float2(u01.z, u01.w), d1);
        d1 = __ffma2_rn(make_float2(a1[2], a1[2]), make_float2(u23.x, u23.y), d1);
        d1 = __ffma2_rn(make_float2(a1[3], a1[3]), make_float2(u23.z, u23.w), d1);
        float2 d2 = __fmul2_rn(make_float2(a2[0], a2[0]), make_float2(v01.x, v01.y));
        d2 = __ffma2_rn(make_float2(a2[1], a2[1]), make_float2(v01.z, v01.w), d2);
        d2 = __ffma2_rn(make_float2(a2[2], a2[2]), make_float2(v23.x, v23.y), d2);
        d2 = __ffma2_rn(make_float2(a2[3], a2[3]), make_float2(v23.z, v23.w), d2);
        d1.x = fminf(fmaxf(d1.x, 0.f), 1.f); d1.y = fminf(fmaxf(d1.y, 0.f), 1.f);      // rotation.py:220 (clip to [0,1])
        d2.x = fminf(fmaxf(d2.x, 0.f), 1.f); d2.y = fminf(fmaxf(d2.y, 0.f), 1.f);
        const float2 g1 = sg_acos01_2(d1), g2 = sg_acos01_2(d2);
        float tx = fabsf(g1.x - g2.x), ty = fabsf(g1.y - g2.y);
        // columns past the end, and on the diagonal block the pairs with j <= i, do not count
        const int jx = 2 * pp, jy = 2 * pp + 1;
        if (jx >= n_in || (diag && jblk + jx <= i)) tx = 0.f;
        if (jy >= n_in || (diag && jblk + jy <= i)) ty = 0.f;
        part.x += tx;
        part.y += ty;
      }
    }
    if (i_ok) acc += 2.0 * ((double)part.x + (double)part.y);
  }
  (void)lane;
  acc = block_sum(acc, red);
  if (tid == 0) partials[(long long)blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

__global__ void sum_partials_kernel(const double* __restrict__ partials, int n, double scale, double* __restrict__ out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += partials[i];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = acc * scale;
}

// ---- rotation-matrix constraint functional of the fit WITHOUT known orientations -----------------
// DiffusionMapSolver.or_func (diffusion.py:265-287; src/octave/CalculateDiffusion.m:136-157):
//   R_l = reshape(C psi_l, 3, 3),  T = R_l^T R_l - I,
//   form 0 (the Python expression):  g_l = (sum T^2 + |det R_l - 1|)^(1/4) / sqrt(r)
//   form 1 (the Octave expression):  g_l = (sqrt(sum T^2) + |det R_l - 1|)^(1/2) / sqrt(r)
// C = c.reshape(9,9) row-major (numpy), psi_l = Ps[l, 1:10].  ||R^T R - I||_F and det R do not depend on
// whether the 9-vector is folded row- or column-major, so Octave's column-major reshape gives the same g.
__device__ __forceinline__ double or_value(const double* R, int form, double inv_sqrt_r) {
  double s2 = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const double t = R[a] * R[b] + R[3 + a] * R[3 + b] + R[6 + a] * R[6 + b] - (a == b ? 1.0 : 0.0);   // (R^T R - I)[a][b]
      s2 = fma(t, t, s2);
    }
  const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
  const double d = fabs(det - 1.0);
  const double inner = form == 0 ? sqrt(s2 + d) : sqrt(s2) + d;
  return sqrt(inner) * inv_sqrt_r;
}

// G [n] and (optional) the forward-difference Jacobian J [n,81] with step h_p = rel_step * max(1,|c_p|)
// (scipy.optimize.least_squares '2-point'); one thread per pattern, c in shared memory.
__global__ void __launch_bounds__(128)
or_func_kernel(const double* __restrict__ c81, const double* __restrict__ Ps, int ldx, long long n, long long r_total,
               int form, double rel_step, double* __restrict__ G, double* __restrict__ J) {
  __shared__ double cs[81];
  if (threadIdx.x < 81) cs[threadIdx.x] = c81[threadIdx.x];
  __syncthreads();
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (l >= n) return;
  const double inv_sqrt_r = 1.0 / sqrt((double)r_total);
  double psi[9], R[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) psi[q] = Ps[l * ldx + 1 + q];
#pragma unroll
  for (int a = 0; a < 9; ++a) {
    double v = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) v = fma(cs[a * 9 + q], psi[q], v);
    R[a] = v;
  }
  const double g0 = or_value(R, form, inv_sqrt_r);
  G[l] = g0;
  if (!J) return;
  for (int a = 0; a < 9; ++a) {
    const double keep = R[a];
    for (int q = 0; q < 9; ++q) {
      const double cp = cs[a * 9 + q];
      const double h = rel_step * fmax(1.0, fabs(cp));
      const double cph = cp + h;                       // the step actually taken in floating point
      R[a] = keep + (cph - cp) * psi[q];
      J[l * 81 + a * 9 + q] = (or_value(R, form, inv_sqrt_r) - g0) / (cph - cp);
    }
    R[a] = keep;
  }
}

}  // namespace v3s

using namespace v3s;

extern "C" int v3s_fit_gram(const double* Ps, int ldx, const double* Y, long long n_rows, double* out162, void* stream) {
  V3S_REQUIRE(ldx >= 10 && n_rows >= 0, "v3s_fit_gram: Ps needs >= 10 columns");
  cudaStream_t st = (cudaStream_t)stream;
  V3S_CUDA(cudaMemsetAsync(out162, 0, sizeof(double) * 162, st));
  if (n_rows == 0) return 0;
  long long g = ceil_div64(n_rows, FIT_ROWS);
  if (g > 2LL * kNumSMs) g = 2LL * kNumSMs;
  fit_gram_kernel<<<(int)g, 192, 0, st>>>(Ps, ldx, Y, n_rows, out162);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_fit_solve(const double* in162, double* c81, void* stream) {
  fit_solve_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(in162, c81);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_fit_project(const double* Ps, int ldx, const double* Y, long long n_rows, const double* c81,
                               int polar_mode, double* recon_pre, double* rproj, double* quat_est, double* quat_true,
                               void* stream) {
  V3S_REQUIRE(ldx >= 10 && n_rows >= 0 && polar_mode >= 0 && polar_mode <= 2, "v3s_fit_project: bad arguments");
  if (n_rows == 0) return 0;
  long long g = ceil_div64(n_rows, 128);
  if (g > 16LL * kNumSMs) g = 16LL * kNumSMs;
  fit_project_kernel<<<(int)g, 128, 0, (cudaStream_t)stream>>>(Ps, ldx, Y, n_rows, c81, polar_mode, recon_pre, rproj,
                                                             quat_est, quat_true);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_sigma_pairs(const double* quat_est, const double* quat_true, long long N, long long row0,
                               long long n_rows, double* partials_ws, double* out_sum, void* stream) {
  V3S_REQUIRE(N >= 1 && row0 >= 0 && n_rows >= 0 && row0 + n_rows <= N, "v3s_sigma_pairs: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rows == 0) { V3S_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(double), st)); return 0; }
  const long long maxy = ceil_div64(N, SG_T);
  if (N <= kSigmaFp64MaxN) {
    const int gx = (int)ceil_div64(n_rows, SG_T);
    int gy = (int)ceil_div64(8LL * kNumSMs, gx);
    if (gy > maxy) gy = (int)maxy;
    if (gy < 1) gy = 1;
    V3S_REQUIRE((long long)gx * gy <= 65536, "v3s_sigma_pairs: partials workspace (65536 doubles) too small");
    sigma_pairs_kernel<true><<<dim3(gx, gy), SG_T, 0, st>>>(quat_est, quat_true, N, row0, n_rows, partials_ws);
    V3S_CHECK_LAUNCH();
    sum_partials_kernel<<<1, 256, 0, st>>>(partials_ws, gx * gy, 1.0, out_sum);
    V3S_CHECK_LAUNCH();
    return 0;
  }
  // FP32 form: threads <-> rows from a 32-aligned base, column ranges in multiples of the 256-column chunk
  const long long i_base = row0 & ~31LL;
  const int gx = (int)ceil_div64(row0 + n_rows - i_base, SG_T);
  int gy = (int)ceil_div64(16LL * kNumSMs, gx);
  if (gy > maxy) gy = (int)maxy;
  if (gy < 1) gy = 1;
  const long long j_per = ceil_div64(ceil_div64(N, gy), SG_T) * SG_T;
  gy = (int)ceil_div64(N, j_per);
  V3S_REQUIRE((long long)gx * gy <= 65536, "v3s_sigma_pairs: partials workspace (65536 doubles) too small");
  sigma_pairs_f32_kernel<<<dim3(gx, gy), SG_T, 0, st>>>(quat_est, quat_true, N, row0, n_rows, i_base, j_per, partials_ws);
  V3S_CHECK_LAUNCH();
  sum_partials_kernel<<<1, 256, 0, st>>>(partials_ws, gx * gy, 1.0, out_sum);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_or_func(const double* c81, const double* Ps, int ldx, long long n_rows, long long r_total, int form,
                           double rel_step, double* G, double* J, void* stream) {
  V3S_REQUIRE(c81 && Ps && G && ldx >= 10 && n_rows >= 0 && r_total >= 1 && (form == 0 || form == 1),
              "v3s_or_func: bad arguments (Ps needs the trivial + 9 non-trivial columns)");
  if (n_rows == 0) return 0;
  or_func_kernel<<<(int)ceil_div64(n_rows, 128), 128, 0, (cudaStream_t)stream>>>(c81, Ps, ldx, n_rows, r_total, form, rel_step, G, J);
  V3S_CHECK_LAUNCH();
  return 0;
}
