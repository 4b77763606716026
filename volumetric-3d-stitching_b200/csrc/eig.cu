// D8 through the C-ABI alone -- v3s_eig_topk: the largest-magnitude eigenpairs of the symmetric Markov matrix
// P_ep, the replacement for `scipy.sparse.linalg.eigs(P_ep, k)` (ARPACK) at diffusion.py:61.
//
// A host-side driver in the library (no Python, no LAPACK): block Lanczos (b = 8, full two-pass
// re-orthogonalisation, lanczos.cu) on the fused mat-vec kernels (matvec.cu), Chebyshev-filtered for N >= 2048
// exactly like v3s/eig.py, with the Ritz pairs of the small projected matrix from a Householder tridiagonalisation
// + implicit QL iteration written out below.  It synchronises the stream at every Ritz check (a few small
// device-to-host copies) -- the one entry point of this header that does.
//   probe   : `probe_steps` plain steps; Ritz values are lower bounds of the eigenvalues (Cauchy interlacing), so
//             cut = theta_nev(probe) <= lambda_nev never damps a wanted eigenvalue
//   filter  : Lanczos on T_degree((P - c)/e), [c-e, c+e] = [lo, cut], started from the probe's best Ritz block
//   finish  : Rayleigh-Ritz of P itself on the converged block (eigenvalues of P, true residuals)
// Degenerate clusters of multiplicity >= 8 (disconnected graphs) saturate a block Krylov space: reported through
// info[5]; v3s/eig.py's locking loop (topk_with_locking) is the remedy and stays on the Python side.
#include "common.cuh"
#include "v3s.h"
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace v3s {

// ---- small dense symmetric eigen-solver (host): Householder tridiagonalisation + implicit QL ----------------
// One plane rotation of the QL iteration applied to two contiguous vectors (6 n^3 flops of the whole solve are here):
// an AVX2 + FMA build of the loop is used where the host CPU has it (checked once at run time).
__attribute__((target("avx2,fma"))) static void ql_rotate_avx2(double* __restrict__ zi, double* __restrict__ zi1, int len,
                                                                double c, double s) {
  for (int k = 0; k < len; ++k) {
    const double f = zi1[k], g = zi[k];
    zi1[k] = s * g + c * f;
    zi[k] = c * g - s * f;
  }
}
static void ql_rotate_generic(double* __restrict__ zi, double* __restrict__ zi1, int len, double c, double s) {
  for (int k = 0; k < len; ++k) {
    const double f = zi1[k], g = zi[k];
    zi1[k] = s * g + c * f;
    zi[k] = c * g - s * f;
  }
}
static inline void ql_rotate(double* zi, double* zi1, int len, double c, double s) {
  static const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  if (avx2) ql_rotate_avx2(zi, zi1, len, c, s);
  else ql_rotate_generic(zi, zi1, len, c, s);
}
// z is stored TRANSPOSED here (row i = vector i): the plane rotations then update two contiguous rows
// k0 > 0: only components k0 .. n-1 of the vectors are carried through the rotations (the convergence check of the
// block Lanczos run needs the last block of every Ritz vector, not the vectors)
static int tridiagonal_ql(std::vector<double>& d, std::vector<double>& e, int n, int ld, std::vector<double>& z, int k0 = 0) {
  auto Z = [&](int i, int j) -> double& { return z[(size_t)j * ld + i]; };
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  // deflation threshold relative to the matrix norm as well as to the neighbouring diagonal entries: a cluster of
  // (numerically) zero eigenvalues -- rank-deficient blocks of a disconnected graph's operator -- leaves off-diagonal
  // noise of the size of its diagonal noise, which a purely relative test never accepts
  double anorm = 0.0;
  for (int i = 0; i < n; ++i) anorm = fmax(anorm, fabs(d[i]) + fabs(e[i]));
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = fabs(d[m]) + fabs(d[m + 1]);
        if (fabs(e[m]) <= 2.3e-16 * (dd + anorm)) break;
      }
      if (m != l) {
        if (iter++ == 60) return 1;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = sqrt(g * g + 1.0);
        if (!(r < 1e150)) r = hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          r = sqrt(f * f + g * g);                         // (the scalar recurrence is the critical path of the iteration:
          if (!(r > 1e-150 && r < 1e150)) r = hypot(f, g); //  plain sqrt unless the squares could over- / underflow)
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - p;
          r = (d[i] - g) * s + 2.0 * c * b;
          d[i + 1] = g + (p = s * r);
          g = c * r - b;
          ql_rotate(&Z(k0, i), &Z(k0, i + 1), n - k0, c, s);
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return 0;
}
// Householder tridiagonalisation on the FULL symmetric matrix with row-contiguous inner loops (symmetric mat-vec +
// rank-2 update per step, backward accumulation of Q), so that the compiler vectorises them; the body is compiled
// twice, for AVX2 + FMA and for the baseline ISA, and picked at run time.
//   a [n, ld] row-major symmetric (destroyed) -> d [n] diagonal, e [n] with e[i] = T[i][i-1] (e[0] = 0),
//   qt [n, ld]: Q TRANSPOSED (row i = column i of Q, the layout tridiagonal_ql rotates), A = Q T Q^T.
#define V3S_TRIDIAG_BODY                                                                                              \
  std::vector<double> beta((size_t)n, 0.0), p((size_t)n), t((size_t)n);                                               \
  for (int k = 0; k + 2 < n; ++k) {                                                                                    \
    double* x = &a[(size_t)k * ld + k + 1];                                                                            \
    const int len = n - k - 1;                                                                                         \
    double sigma = 0.0;                                                                                                \
    for (int i = 1; i < len; ++i) sigma += x[i] * x[i];                                                                \
    const double x0 = x[0];                                                                                            \
    d[k] = a[(size_t)k * ld + k];                                                                                      \
    if (sigma == 0.0) { e[k + 1] = x0; beta[k] = 0.0; continue; }                                                      \
    const double mu = sqrt(x0 * x0 + sigma);                                                                           \
    const double v0 = x0 <= 0.0 ? x0 - mu : -sigma / (x0 + mu);                                                        \
    const double b = 2.0 * v0 * v0 / (sigma + v0 * v0);                                                                \
    const double inv = 1.0 / v0;                                                                                       \
    x[0] = 1.0;                                                                                                        \
    for (int i = 1; i < len; ++i) x[i] *= inv;               /* x is now v (v[0] = 1), kept in row k for Q */          \
    beta[k] = b;                                                                                                       \
    e[k + 1] = mu;                                                                                                     \
    /* p = b A_sub v */                                                                                                \
    double pv = 0.0;                                                                                                   \
    for (int i = 0; i < len; ++i) {                                                                                    \
      const double* row = &a[(size_t)(k + 1 + i) * ld + k + 1];                                                        \
      double acc = 0.0;                                                                                                \
      for (int j = 0; j < len; ++j) acc += row[j] * x[j];                                                              \
      p[i] = b * acc;                                                                                                  \
      pv += p[i] * x[i];                                                                                               \
    }                                                                                                                  \
    const double K = 0.5 * b * pv;                                                                                     \
    for (int i = 0; i < len; ++i) p[i] -= K * x[i];          /* w */                                                   \
    for (int i = 0; i < len; ++i) {                          /* A_sub -= v w^T + w v^T */                              \
      double* row = &a[(size_t)(k + 1 + i) * ld + k + 1];                                                              \
      const double vi = x[i], wi = p[i];                                                                               \
      for (int j = 0; j < len; ++j) row[j] -= vi * p[j] + wi * x[j];                                                   \
    }                                                                                                                  \
  }                                                                                                                    \
  e[0] = 0.0;                                                                                                          \
  if (n >= 2) {                                                                                                        \
    d[n - 2] = a[(size_t)(n - 2) * ld + n - 2];                                                                        \
    e[n - 1] = a[(size_t)(n - 2) * ld + n - 1];                                                                        \
  }                                                                                                                    \
  d[n - 1] = a[(size_t)(n - 1) * ld + n - 1];                                                                          \
  /* Q = H_0 H_1 ... by backward accumulation, row-major in q; then transposed into qt */                             \
  std::vector<double> q((size_t)n * ld, 0.0);                                                                          \
  for (int i = 0; i < n; ++i) q[(size_t)i * ld + i] = 1.0;                                                             \
  for (int k = n - 3; k >= 0; --k) {                                                                                   \
    const double b = beta[k];                                                                                          \
    if (b == 0.0) continue;                                                                                            \
    const double* v = &a[(size_t)k * ld + k + 1];                                                                      \
    const int len = n - k - 1;                                                                                         \
    for (int j = 0; j < len; ++j) t[j] = 0.0;                                                                          \
    for (int i = 0; i < len; ++i) {                                                                                    \
      const double* row = &q[(size_t)(k + 1 + i) * ld + k + 1];                                                        \
      const double vi = v[i];                                                                                          \
      for (int j = 0; j < len; ++j) t[j] += vi * row[j];                                                               \
    }                                                                                                                  \
    for (int i = 0; i < len; ++i) {                                                                                    \
      double* row = &q[(size_t)(k + 1 + i) * ld + k + 1];                                                              \
      const double bv = b * v[i];                                                                                      \
      for (int j = 0; j < len; ++j) row[j] -= bv * t[j];                                                               \
    }                                                                                                                  \
  }                                                                                                                    \
  for (int i = 0; i < n; ++i)                                                                                          \
    for (int j = 0; j < n; ++j) qt[(size_t)j * ld + i] = q[(size_t)i * ld + j];

__attribute__((target("avx2,fma"))) static void tridiag_full_avx2(std::vector<double>& a, int n, int ld, std::vector<double>& d,
                                                                   std::vector<double>& e, std::vector<double>& qt) {
  V3S_TRIDIAG_BODY
}
static void tridiag_full_generic(std::vector<double>& a, int n, int ld, std::vector<double>& d, std::vector<double>& e,
                                 std::vector<double>& qt) {
  V3S_TRIDIAG_BODY
}
#undef V3S_TRIDIAG_BODY

// eigenpairs of the symmetric n x n matrix a (row-major; destroyed): w ascending, vectors = columns of a
// rows_from > 0: only rows rows_from .. n-1 of the eigenvector matrix are valid on return (the QL rotations carry
// those components only)
static int sym_eig(std::vector<double>& a, int n, std::vector<double>& w, int rows_from = 0) {
  std::vector<double> e((size_t)n, 0.0);
  w.assign((size_t)n, 0.0);
  if (n == 1) { w[0] = a[0]; a[0] = 1.0; return 0; }
  // row stride padded away from multiples of 256 B (cache sets)
  const int ld = (n >= 64 && n % 32 == 0) ? n + 8 : n;
  {
    std::vector<double> ap((size_t)n * ld, 0.0), zt((size_t)n * ld);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) ap[(size_t)i * ld + j] = a[(size_t)i * n + j];
    static const bool avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
    if (avx2) tridiag_full_avx2(ap, n, ld, w, e, zt);
    else tridiag_full_generic(ap, n, ld, w, e, zt);
    if (tridiagonal_ql(w, e, n, ld, zt, rows_from)) return 1;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) a[(size_t)i * n + j] = zt[(size_t)j * ld + i];
  }
  std::vector<int> idx((size_t)n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::sort(idx.begin(), idx.end(), [&](int x, int y) { return w[x] < w[y]; });
  std::vector<double> w2((size_t)n), a2((size_t)n * n);
  for (int j = 0; j < n; ++j) {
    w2[j] = w[idx[j]];
    for (int i = 0; i < n; ++i) a2[(size_t)i * n + j] = a[(size_t)i * n + idx[j]];
  }
  w.swap(w2);
  a.swap(a2);
  return 0;
}

// ---- small device helpers ----------------------------------------------------------------------------------
// out[n, j] = sum_i V[i, n] S[i, j]   (V: m basis vectors as rows, ld ldv; S: m x c row-major, c <= 16)
__global__ void basis_combine_kernel(const double* __restrict__ V, long long ldv, int m, const double* __restrict__ S, int c,
                                     double* __restrict__ out, int ldo, long long N) {
  extern __shared__ double sS[];
  for (int i = threadIdx.x; i < m * c; i += blockDim.x) sS[i] = S[i];
  __syncthreads();
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n >= N) return;
  double acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.0;
  for (int i = 0; i < m; ++i) {
    const double v = V[(long long)i * ldv + n];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < c) acc[j] = fma(v, sS[i * c + j], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j)
    if (j < c) out[n * ldo + j] = acc[j];
}
// out[n, j] = sum_i A[n, i] Z[i, j]    (A: N x ca row-major ld lda; Z ca x c)
__global__ void rows_times_small_kernel(const double* __restrict__ A, int lda, int ca, const double* __restrict__ Z, int c,
                                        double* __restrict__ out, int ldo, long long N) {
  __shared__ double sZ[256];
  for (int i = threadIdx.x; i < ca * c; i += blockDim.x) sZ[i] = Z[i];
  __syncthreads();
  const long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n >= N) return;
  double a[16];
  for (int i = 0; i < ca; ++i) a[i] = A[n * lda + i];
  for (int j = 0; j < c; ++j) {
    double acc = 0.0;
    for (int i = 0; i < ca; ++i) acc = fma(a[i], sZ[i * c + j], acc);
    out[n * ldo + j] = acc;
  }
}
// H[i, j] += sum_n A[n, i] B[n, j]   (both N x c, c <= 16; H zero-initialised, 16 x 16 row-major)
__global__ void small_gram_kernel(const double* __restrict__ A, int lda, const double* __restrict__ B, int ldb, int c,
                                  long long N, double* __restrict__ H) {
  __shared__ double red[32];
  for (int i = 0; i < c; ++i)
    for (int j = 0; j < c; ++j) {
      double acc = 0.0;
      for (long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x)
        acc = fma(A[n * lda + i], B[n * ldb + j], acc);
      acc = block_sum(acc, red);
      if (threadIdx.x == 0) atomicAdd(&H[i * 16 + j], acc);
    }
}
// block [N,8] <- columns [j0, j0+8) of Y [N, ld] (zero beyond c)
__global__ void take_cols_kernel(const double* __restrict__ Y, int ld, int c, int j0, double* __restrict__ blk, long long N) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= N * 8) return;
  const long long n = e >> 3;
  const int j = (int)(e & 7);
  blk[e] = (j0 + j < c) ? Y[n * ld + j0 + j] : 0.0;
}
// dst [N, ldd] columns [j0, j0+8) <- blk [N,8] (only columns < c)
__global__ void put_cols_kernel(const double* __restrict__ blk, double* __restrict__ dst, int ldd, int c, int j0, long long N) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= N * 8) return;
  const long long n = e >> 3;
  const int j = (int)(e & 7);
  if (j0 + j < c) dst[n * ldd + j0 + j] = blk[e];
}

struct EigCtx {
  v3s_eig_problem* pb;
  cudaStream_t st;
  long long N;
  int max_dim;
  double *V, *lzws, *T, *Yprobe, *Y, *PY, *YZ, *blk, *small;   // device
  float* Xf;
  int* status;                                            // device (breakdown flag)
  int matvecs, checks;
  bool sparse;
};

static double* ybuf(const EigCtx& c, int b) { return reinterpret_cast<double*>(c.pb->ybuf_ptrs[(size_t)b * c.pb->world + c.pb->rank]); }
static int ybuf_index(const EigCtx& c, const double* q) {
  for (int b = 0; b < 3; ++b)
    if (ybuf(c, b) == q) return b;
  return -1;
}
static int eig_sync(EigCtx& c) {
  v3s_eig_problem* p = c.pb;
  if (p->world > 1) {
    p->epoch += 1;
    return v3s_peer_barrier(p->flag_ptrs, p->rank, p->world, p->epoch, p->status_dev, c.st);
  }
  return 0;
}
// W = P Q (complete on every rank) in a rotating buffer other than Q; Xf = float(Q) in strip layout (dense operator)
static int apply_plain(EigCtx& c, const double* Q, double** W) {
  v3s_eig_problem* p = c.pb;
  const int qi = ybuf_index(c, Q);
  int out = 0;
  while (out == qi) ++out;
  if (p->world > 1) p->epoch += 1;
  int rc;
  if (c.sparse)
    rc = v3s_sparse_matvec_combine(p->sparse, p->n_rows, p->N, p->row0, Q, nullptr, nullptr, 1.0, 0.0, 0.0,
                                   p->ybuf_ptrs + (size_t)out * p->world, nullptr, p->flag_ptrs, p->rank, p->world, p->epoch,
                                   p->ticket_dev, p->status_dev, p->mv_workspace, c.st);
  else
    rc = v3s_matvec_combine(p->P_rows, p->ldp, p->n_rows, p->N, p->row0, c.Xf, nullptr, nullptr, 1.0, 0.0, 0.0,
                            p->ybuf_ptrs + (size_t)out * p->world, nullptr, p->flag_ptrs, p->rank, p->world, p->epoch,
                            p->ticket_dev, p->status_dev, p->mv_workspace, c.st);
  c.matvecs += 1;
  *W = ybuf(c, out);
  return rc;
}
static int apply_cheb(EigCtx& c, const double* Q, int degree, double cc, double ee, double** W) {
  v3s_eig_problem* p = c.pb;
  int res = 0, rc;
  if (c.sparse)
    rc = v3s_sparse_cheb_filter(p->sparse, p->n_rows, p->N, p->row0, Q, ybuf_index(c, Q), degree, cc, ee, p->ybuf_ptrs, p->flag_ptrs,
                                p->rank, p->world, p->epoch, p->ticket_dev, p->status_dev, p->mv_workspace, &res, c.st);
  else
    rc = v3s_cheb_filter(p->P_rows, p->ldp, p->n_rows, p->N, p->row0, c.Xf, Q, ybuf_index(c, Q), degree, cc, ee, p->ybuf_ptrs,
                         p->xf_ptrs, p->flag_ptrs, p->rank, p->world, p->epoch, p->ticket_dev, p->status_dev, p->mv_workspace,
                         &res, c.st);
  if (p->world > 1) p->epoch += degree;
  c.matvecs += degree;
  *W = ybuf(c, res);
  return rc;
}

struct RitzOut {
  std::vector<double> w, S, res;     // nev values, m x nev vectors (row-major), residual estimates
  int m = 0;
  bool converged = false;
  double theta_min = 0.0, theta_max = 0.0;
};

// Block Lanczos on `plain` (degree 0) or the Chebyshev-filtered operator, from the start block W0 [N,8]
// (device, overwritten).  both_ends: pick the nev largest |theta| (ARPACK 'LM'), else the algebraically largest.
static int lanczos_run(EigCtx& c, double* W0, int run_dim, int first_check, int gap, bool both_ends, int nev, double tol,
                       int degree, double cc, double ee, RitzOut& out) {
  const long long N = c.N;
  const int b = 8;
  run_dim = std::max<int>(b, (int)((std::min<long long>(run_dim, N) / b) * b));
  if (int rc = eig_sync(c)) return rc;
  if (int rc = v3s_lanczos_start(c.V, N, N, run_dim, W0, c.Xf, c.lzws, c.st)) return rc;
  V3S_CUDA(cudaMemsetAsync(c.status, 0, sizeof(int), c.st));
  const double* Q = W0;
  int steps = 0, m = b, next_check = std::min(first_check, run_dim / b);
  std::vector<double> Th((size_t)(run_dim / b) * 128);
  int copied = 0;
  for (;;) {
    double* W = nullptr;
    if (int rc = degree > 0 ? apply_cheb(c, Q, degree, cc, ee, &W) : apply_plain(c, Q, &W)) return rc;
    if (int rc = v3s_lanczos_step(c.V, N, N, m, run_dim, W, c.Xf, c.T + (size_t)steps * 128, c.lzws, c.status, c.st)) return rc;
    Q = W;
    ++steps;
    const bool last = (m + b > run_dim);
    if (steps >= next_check || last) {
      int hstatus = 0;
      // (through the pinned mailbox, not the copy engine: a D2H memcpy would queue behind a bulk transfer of the caller)
      if (int rc = read_back(&hstatus, c.status, sizeof(int), c.st)) return rc;
      if (int rc = read_back(Th.data() + (size_t)copied * 128, c.T + (size_t)copied * 128, sizeof(double) * 128 * (steps - copied), c.st)) return rc;
      copied = steps;
      V3S_REQUIRE(hstatus == 0, "block Lanczos broke down (Cholesky-QR of a rank-deficient block)");
      // The host now spends ~1 ms on the Ritz values: keep the GPU busy with the NEXT step meanwhile.  It is
      // speculative -- discarded when this check finds convergence (its basis rows / T block lie beyond m / steps,
      // nothing reads them) and adopted otherwise; every rank takes the same decisions from the same numbers.
      double* Wspec = nullptr;
      if (!last && m + 2 * b <= run_dim) {                      // (not when it would be the run's last step: that one ends with a check)
        if (int rc = degree > 0 ? apply_cheb(c, Q, degree, cc, ee, &Wspec) : apply_plain(c, Q, &Wspec)) return rc;
        if (int rc = v3s_lanczos_step(c.V, N, N, m + b, run_dim, Wspec, c.Xf, c.T + (size_t)steps * 128, c.lzws, c.status, c.st)) return rc;
      }
      // projected block-tridiagonal matrix: A_j on the diagonal, B_j (upper triangular) below it
      std::vector<double> Tm((size_t)m * m, 0.0), w;
      for (int j = 0; j < steps; ++j) {
        const double* A = Th.data() + (size_t)j * 128;
        const double* B = A + 64;
        for (int r = 0; r < b; ++r)
          for (int q = 0; q <= r; ++q) {
            Tm[(size_t)(j * b + r) * m + j * b + q] = A[r * b + q];
            Tm[(size_t)(j * b + q) * m + j * b + r] = A[r * b + q];
          }
        if (j + 1 < steps)
          for (int r = 0; r < b; ++r)
            for (int q = r; q < b; ++q) {
              Tm[(size_t)((j + 1) * b + r) * m + j * b + q] = B[r * b + q];
              Tm[(size_t)(j * b + q) * m + (j + 1) * b + r] = B[r * b + q];
            }
      }
      // first the eigenvalues and the LAST block of the eigenvectors only (all the residual estimates need); the
      // vectors themselves are extracted once, when the run ends
      std::vector<double> Tfull;
      const bool last_possible = last;
      if (!last_possible) Tfull = Tm;
      V3S_REQUIRE(sym_eig(Tm, m, w, last_possible ? 0 : m - b) == 0,
                  "v3s_eig_topk: the QL iteration of the projected matrix did not converge");
      c.checks += 1;
      const int want = std::min(nev, m);
      std::vector<int> idx((size_t)m);
      for (int i = 0; i < m; ++i) idx[i] = i;
      if (both_ends) std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return fabs(w[x]) > fabs(w[y]); });
      else std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return w[x] > w[y]; });
      out.m = m;
      out.theta_min = w[0];
      out.theta_max = w[m - 1];
      out.w.assign((size_t)want, 0.0);
      out.S.assign((size_t)m * want, 0.0);
      out.res.assign((size_t)want, 0.0);
      const double* Bl = Th.data() + (size_t)(steps - 1) * 128 + 64;
      double wmax = 0.0, rmax = 0.0;
      for (int jv = 0; jv < want; ++jv) {
        const int col = idx[jv];
        out.w[jv] = w[col];
        wmax = std::max(wmax, fabs(w[col]));
        for (int i = 0; i < m; ++i) out.S[(size_t)i * want + jv] = Tm[(size_t)i * m + col];
        double r2 = 0.0;
        for (int r = 0; r < b; ++r) {
          double acc = 0.0;
          for (int q = r; q < b; ++q) acc += Bl[r * b + q] * Tm[(size_t)(m - b + q) * m + col];
          r2 += acc * acc;
        }
        out.res[jv] = sqrt(r2);
        V3S_REQUIRE(isfinite(out.res[jv]), "block Lanczos broke down (non-finite residuals)");
        rmax = std::max(rmax, out.res[jv]);
      }
      out.converged = rmax <= tol * std::max(wmax, 1e-300);
      if (out.converged && !last_possible) {
        // converged at an intermediate check: now the Ritz vectors (same eigenvalue order: same matrix, same method)
        std::vector<double> w2;
        V3S_REQUIRE(sym_eig(Tfull, m, w2) == 0, "v3s_eig_topk: the QL iteration of the projected matrix did not converge");
        for (int jv = 0; jv < want; ++jv) {
          // match by value: the partial and the full run order equal eigenvalues identically, but stay safe
          int col = idx[jv];
          if (w2[col] != w[col]) {
            double best = 1e300;
            for (int i = 0; i < m; ++i)
              if (fabs(w2[i] - w[idx[jv]]) < best) { best = fabs(w2[i] - w[idx[jv]]); col = i; }
          }
          for (int i = 0; i < m; ++i) out.S[(size_t)i * want + jv] = Tfull[(size_t)i * m + col];
        }
      }
      if (out.converged || last) {
        if (Wspec) c.matvecs -= degree > 0 ? degree : 1;      // (the discarded step is not part of the solve's count)
        break;
      }
      next_check = steps + gap;
      if (Wspec) {                                            // the speculative step becomes a regular one
        m += b;
        Q = Wspec;
        ++steps;
      }
    }
    m += b;
  }
  return 0;
}

}  // namespace v3s

using namespace v3s;

// The host eigen-solver of the projected matrices, exposed for testing: a [n,n] symmetric row-major (overwritten
// with the eigenvectors in its columns), w [n] ascending eigenvalues.  Host pointers, no GPU involved.
extern "C" int v3s_sym_eig_host(double* a, int n, double* w) {
  V3S_REQUIRE(a && w && n >= 1 && n <= 4096, "v3s_sym_eig_host: bad arguments");
  std::vector<double> A(a, a + (size_t)n * n), W;
  V3S_REQUIRE(sym_eig(A, n, W) == 0, "v3s_sym_eig_host: QL iteration did not converge");
  std::copy(A.begin(), A.end(), a);
  std::copy(W.begin(), W.end(), w);
  return 0;
}

extern "C" long long v3s_eig_topk_workspace_doubles(long long N, int max_dim) {
  if (N < 8 || max_dim < 8) return 0;
  const long long md = (std::min<long long>(max_dim, N) / 8) * 8;
  return md * N + v3s_lanczos_workspace_doubles(N, (int)md) + (md / 8) * 128 + (v3s_xf_floats(N) + 1) / 2 +
         4 * N * 16 + N * 8 + 1024 + 64;
}

extern "C" int v3s_eig_topk(v3s_eig_problem* pb, int nev, double tol, int max_dim, int degree, int probe_steps,
                            unsigned long long seed, double* ws, double* lambda_host, double* Y_dev, double* resid_host,
                            int* info_host, void* stream) {
  V3S_REQUIRE(pb && ws && lambda_host && Y_dev && info_host, "v3s_eig_topk: null argument");
  V3S_REQUIRE((pb->P_rows != nullptr) != (pb->sparse != nullptr), "v3s_eig_topk: give the dense row block OR the sparse operator");
  V3S_REQUIRE(pb->N >= 64 && nev >= 1 && nev <= 16 && nev < pb->N - 1, "No spare-matrix for eigs-calculation");   // diffusion.py:59-60
  V3S_REQUIRE(pb->world >= 1 && pb->world <= 8 && pb->ybuf_ptrs && (pb->sparse || pb->xf_ptrs) && pb->mv_workspace,
              "v3s_eig_topk: the rotating result buffers / operand buffers / mat-vec workspace are required");
  V3S_REQUIRE(max_dim >= 16 * ((nev + 7) / 8) && degree >= 0 && probe_steps >= 2 && tol > 0, "v3s_eig_topk: bad solver parameters");
  cudaStream_t st = (cudaStream_t)stream;
  const long long N = pb->N;
  const int md = (int)((std::min<long long>(max_dim, N) / 8) * 8);
  EigCtx c{};
  c.pb = pb; c.st = st; c.N = N; c.max_dim = md; c.sparse = pb->sparse != nullptr;
  double* p = ws;
  c.V = p; p += (long long)md * N;
  c.lzws = p; p += v3s_lanczos_workspace_doubles(N, md);
  c.T = p; p += (md / 8) * 128;
  c.Xf = reinterpret_cast<float*>(p); p += (v3s_xf_floats(N) + 1) / 2;
  c.Yprobe = p; p += N * 16;
  c.Y = p; p += N * 16;
  c.PY = p; p += N * 16;
  c.YZ = p; p += N * 16;
  c.blk = p; p += N * 8;
  c.small = p; p += 1024;
  c.status = reinterpret_cast<int*>(p);
  V3S_CUDA(cudaMemsetAsync(c.Xf, 0, sizeof(float) * (size_t)v3s_xf_floats(N), st));      // strip padding columns stay zero
  // seeded start block (host xorshift + Box-Muller; the reference starts ARPACK from an unseeded random vector)
  {
    // (kept between calls: 8 N normal deviates are ~16 ms of host time at N = 10^5, the same block every pass)
    static std::vector<double> h;
    static long long h_n = -1;
    static unsigned long long h_seed = 0;
    if (h_n != N || h_seed != seed) {
      h.assign((size_t)N * 8, 0.0);
      unsigned long long s = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
      auto next = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (double)((s >> 11) + 1) / 9007199254740993.0; };
      for (size_t i = 0; i + 1 < h.size(); i += 2) {
        const double r = sqrt(-2.0 * log(next())), a = 6.283185307179586 * next();
        h[i] = r * cos(a);
        h[i + 1] = r * sin(a);
      }
      h_n = N;
      h_seed = seed;
    }
    V3S_CUDA(cudaMemcpyAsync(c.blk, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice, st));
    V3S_CUDA(cudaStreamSynchronize(st));
  }
  const int grid_n = (int)ceil_div64(N, 256), grid_n8 = (int)ceil_div64(N * 8, 256);
  auto ritz_vectors = [&](const RitzOut& r, double* dst) -> int {        // dst [N,16] <- V^T S
    const int want = (int)r.w.size();
    V3S_CUDA(cudaMemcpyAsync(c.small, r.S.data(), sizeof(double) * r.S.size(), cudaMemcpyHostToDevice, st));
    V3S_REQUIRE((size_t)r.m * want * 8 <= 200 * 1024, "v3s_eig_topk: Krylov dimension %d too large for the Ritz-vector kernel", r.m);
    static size_t set = 0;
    const size_t sm = (size_t)r.m * want * 8;
    if (sm > set) { V3S_CUDA(cudaFuncSetAttribute(basis_combine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(sm, 64 * 1024))); set = std::max<size_t>(sm, 64 * 1024); }
    basis_combine_kernel<<<grid_n, 256, sm, st>>>(c.V, N, r.m, c.small, want, dst, 16, N);
    V3S_CHECK_LAUNCH();
    V3S_CUDA(cudaStreamSynchronize(st));                                   // r.S (host) was the source of an async copy
    return 0;
  };
  RitzOut probe;
  const bool use_cheb = degree > 0 && N >= 2048;
  const int probe_dim = use_cheb ? std::min(md, probe_steps * 8) : md;
  if (int rc = lanczos_run(c, c.blk, probe_dim, use_cheb ? probe_steps : 6, 4, true, nev, tol, 0, 0.0, 0.0, probe)) return rc;
  std::vector<double> lam = probe.w, res = probe.res;
  int driver = 0, krylov = probe.m;
  bool converged = probe.converged;
  if (int rc = ritz_vectors(probe, c.Yprobe)) return rc;
  double* Yfinal = c.Yprobe;
  bool negative = false;
  for (double v : probe.w) negative |= v < 0.0;
  if (use_cheb && !probe.converged && !negative && (int)probe.w.size() == nev) {
    driver = 1;
    for (int attempt = 0; attempt < 2; ++attempt) {
      const double cut = *std::min_element(probe.w.begin(), probe.w.end());
      double lo = -1.0;
      if (attempt == 0) lo = std::max(-1.0, probe.theta_min - 0.15 * (probe.theta_max - probe.theta_min));
      const double cc = 0.5 * (cut + lo), ee = 0.5 * (cut - lo);
      take_cols_kernel<<<grid_n8, 256, 0, st>>>(c.Yprobe, 16, nev, 0, c.blk, N);     // the probe's best Ritz block
      V3S_CHECK_LAUNCH();
      RitzOut fr;
      if (int rc = lanczos_run(c, c.blk, md, 8, 4, false, nev, tol, degree, cc, ee, fr)) return rc;
      if (int rc = ritz_vectors(fr, c.Y)) return rc;
      krylov = fr.m;
      converged = fr.converged;
      // Rayleigh-Ritz with P itself on span(Y)
      for (int j0 = 0; j0 < nev; j0 += 8) {
        take_cols_kernel<<<grid_n8, 256, 0, st>>>(c.Y, 16, nev, j0, c.blk, N);
        V3S_CHECK_LAUNCH();
        if (!c.sparse)     // Xf <- float(block) in strip layout (the basis V is free scratch by now)
          if (int rc = v3s_cheb_combine(c.blk, c.blk, nullptr, 1.0, 0.0, 0.0, c.V, c.Xf, N * 8, st)) return rc;
        double* W = nullptr;
        if (int rc = apply_plain(c, c.blk, &W)) return rc;
        put_cols_kernel<<<grid_n8, 256, 0, st>>>(W, c.PY, 16, nev, j0, N);
        V3S_CHECK_LAUNCH();
      }
      V3S_CUDA(cudaMemsetAsync(c.small, 0, sizeof(double) * 256, st));
      small_gram_kernel<<<64, 256, 0, st>>>(c.Y, 16, c.PY, 16, nev, N, c.small);
      V3S_CHECK_LAUNCH();
      std::vector<double> H(256), Hn((size_t)nev * nev), wz;
      if (int rc = read_back(H.data(), c.small, sizeof(double) * 256, st)) return rc;
      for (int i = 0; i < nev; ++i)
        for (int j = 0; j < nev; ++j) Hn[(size_t)i * nev + j] = 0.5 * (H[i * 16 + j] + H[j * 16 + i]);
      V3S_REQUIRE(sym_eig(Hn, nev, wz) == 0, "v3s_eig_topk: Rayleigh-Ritz eigenproblem did not converge");
      // Y <- Y Z, PY <- PY Z (into Yprobe / blk-sized scratch), residual norms
      V3S_CUDA(cudaMemcpyAsync(c.small, Hn.data(), sizeof(double) * nev * nev, cudaMemcpyHostToDevice, st));
      rows_times_small_kernel<<<grid_n, 256, 0, st>>>(c.Y, 16, nev, c.small, nev, c.YZ, 16, N);
      V3S_CHECK_LAUNCH();
      rows_times_small_kernel<<<grid_n, 256, 0, st>>>(c.PY, 16, nev, c.small, nev, c.Y, 16, N);       // c.Y now holds (P Y) Z
      V3S_CHECK_LAUNCH();
      // R = (P Y) Z - (Y Z) diag(w):  ||R_j||^2 = ||PYZ_j||^2 - 2 w_j <PYZ_j, YZ_j> + w_j^2 ||YZ_j||^2
      V3S_CUDA(cudaMemsetAsync(c.small + 256, 0, sizeof(double) * 768, st));
      small_gram_kernel<<<64, 256, 0, st>>>(c.Y, 16, c.Y, 16, nev, N, c.small + 256);
      V3S_CHECK_LAUNCH();
      small_gram_kernel<<<64, 256, 0, st>>>(c.Y, 16, c.YZ, 16, nev, N, c.small + 512);
      V3S_CHECK_LAUNCH();
      small_gram_kernel<<<64, 256, 0, st>>>(c.YZ, 16, c.YZ, 16, nev, N, c.small + 768);
      V3S_CHECK_LAUNCH();
      std::vector<double> G(768);
      if (int rc = read_back(G.data(), c.small + 256, sizeof(double) * 768, st)) return rc;
      lam.assign((size_t)nev, 0.0);
      res.assign((size_t)nev, 0.0);
      std::vector<int> idx((size_t)nev);
      for (int j = 0; j < nev; ++j) idx[j] = j;
      std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return fabs(wz[x]) > fabs(wz[y]); });
      std::vector<double> perm((size_t)nev * nev, 0.0);
      for (int j = 0; j < nev; ++j) {
        const int cj = idx[j];
        lam[j] = wz[cj];
        const double r2 = G[cj * 16 + cj] - 2.0 * wz[cj] * G[256 + cj * 16 + cj] + wz[cj] * wz[cj] * G[512 + cj * 16 + cj];
        res[j] = sqrt(std::max(r2, 0.0));
        perm[(size_t)cj * nev + j] = 1.0;
      }
      V3S_CUDA(cudaMemcpyAsync(c.small, perm.data(), sizeof(double) * nev * nev, cudaMemcpyHostToDevice, st));
      rows_times_small_kernel<<<grid_n, 256, 0, st>>>(c.YZ, 16, nev, c.small, nev, c.PY, 16, N);   // sorted eigenvectors
      V3S_CHECK_LAUNCH();
      V3S_CUDA(cudaStreamSynchronize(st));
      Yfinal = c.PY;
      const double lmin = *std::min_element(lam.begin(), lam.end());
      if (lo > -1.0 && lmin < cut - 1e-6 * std::max(1.0, fabs(cut))) continue;   // something below `lo` was amplified: rerun with lo = -1
      break;
    }
  }
  // result: Y_dev [N, nev] row-major
  for (int j0 = 0; j0 < nev; j0 += 8) {
    take_cols_kernel<<<grid_n8, 256, 0, st>>>(Yfinal, 16, nev, j0, c.blk, N);
    V3S_CHECK_LAUNCH();
    put_cols_kernel<<<grid_n8, 256, 0, st>>>(c.blk, Y_dev, nev, nev, j0, N);
    V3S_CHECK_LAUNCH();
  }
  if (int rc = eig_sync(c)) return rc;
  V3S_CUDA(cudaStreamSynchronize(st));
  int saturated = 0;
  for (int i = 0; i < nev; ++i) {
    int same = 0;
    for (int j = 0; j < nev; ++j) same += fabs(fabs(lam[i]) - fabs(lam[j])) <= 1e-6 * std::max(fabs(lam[i]), 1e-300);
    saturated |= same >= 8;
  }
  for (int i = 0; i < nev; ++i) {
    lambda_host[i] = lam[i];
    if (resid_host) resid_host[i] = res[i];
  }
  info_host[0] = converged ? 1 : 0;
  info_host[1] = c.matvecs;
  info_host[2] = krylov;
  info_host[3] = driver;
  info_host[4] = c.checks;
  info_host[5] = saturated;
  return 0;
}
