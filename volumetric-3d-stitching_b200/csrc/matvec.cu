// K3 -- block mat-vec  Y = P_rows . X  for the block-Lanczos eigen-solver (HBM-bound).
//
// Replaces the dense `P_ep @ v` that ARPACK drives inside scipy.sparse.linalg.eigs
// (diffusion.py:61).  P_rows is the FP32 row block [n_rows, N] of the Markov matrix, X the Lanczos
// block [N, 8].  Algorithmic traffic per call: 4 n_rows N bytes of P, read exactly once.
//
// Kernel: 2 persistent CTAs per SM, each owning an equal contiguous range of (32-row group,
// 128-column strip) units (no tail wave).  A producer thread streams the range's
// 32 x 128 FP32 tiles of P (TMA 2-D tensor copy, 16 KB) and the matching 4 KB strip of X (TMA bulk
// copy, strip layout, common.cuh) through a 4-stage mbarrier ring -- up to 8 tiles in flight per SM
// with two resident CTAs, independent of register pressure.  Four consumer warps own 8 rows each:
// lane l reads columns 4l..4l+3 of its rows and the 8 X values of those columns with 128-bit
// shared loads, accumulates 8 rows x 8 vectors in FP32 registers, and the per-lane partial sums
// are combined with warp shuffles in FP64 ("warp-shuffle-reduced") at the end of each row-group
// segment.  Segment partials are added in a fixed order by a second tiny kernel, so results are
// run-to-run deterministic.
#include "common.cuh"
#include "v3s.h"
#include "tma.cuh"
#include <vector>

namespace v3s {

constexpr int MV_B = 8;            // vectors per block
constexpr int MV_ROWS = 8;         // rows per warp
constexpr int MV_WARPS = 4;        // consumer warps per CTA
constexpr int MV_TILE_ROWS = MV_WARPS * MV_ROWS;
constexpr int MV_THREADS = (MV_WARPS + 1) * 32;
constexpr int MV_STRIP = 128;      // columns per strip (32 lanes x float4)
constexpr int MV_STAGES = 4;
constexpr int MV_P_BYTES = MV_TILE_ROWS * MV_STRIP * 4;   // 16 KB
constexpr int MV_X_BYTES = MV_STRIP * MV_B * 4;           // 4 KB
constexpr int MV_STAGE_BYTES = MV_P_BYTES + MV_X_BYTES;
constexpr int MV_SMEM = MV_STAGES * MV_STAGE_BYTES + 1024 + 128;

__global__ void x_to_f32_kernel(const double* __restrict__ X, long long N, int b, float* __restrict__ Xf) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < N * MV_B; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / MV_B;
    const int j = (int)(e - i * MV_B);
    Xf[xf_strip_index(i, j)] = j < b ? (float)X[i * b + j] : 0.f;
  }
}

// Work unit = (32-row group, 128-column strip); units are numbered row-group-major and split into
// equal contiguous ranges, one per CTA (grid = 2 x SM count, all CTAs co-resident: no tail wave).
// A CTA's range may cross row-group boundaries; every (CTA, row-group) segment writes its partial
// sums to its own slot and `reduce_segments_kernel` adds the slots of a row group in a fixed order.

__global__ void __launch_bounds__(MV_THREADS, 2)
block_matvec_kernel(const __grid_constant__ CUtensorMap tmP, const float* __restrict__ XfS, long long n_rows, long long N,
                    int nstrips_row, long long units_per_cta, long long total_units, int maxseg,
                    double* __restrict__ Ypart, int reverse) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + MV_STAGES * MV_STAGE_BYTES);
  uint64_t* empty = full + MV_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // serpentine order: consecutive launches walk P in opposite directions, so the tail of the previous
  // pass (still in the 126 MB L2) is what the next pass reads first
  const long long cta = reverse ? (long long)gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const long long u0 = cta * units_per_cta;
  long long u1 = u0 + units_per_cta;
  if (u1 > total_units) u1 = total_units;

  if (tid == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmP) : "memory");
    for (int i = 0; i < MV_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], MV_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (u0 >= u1) return;

  if (warp == MV_WARPS) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long rg = u0 / nstrips_row;
      int s = (int)(u0 - rg * nstrips_row);
      for (long long u = u0; u < u1; ++u) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], MV_STAGE_BYTES);
        const uint32_t dst = smem_u32(smem + stage * MV_STAGE_BYTES);
        tma_load_2d(dst, &tmP, s * MV_STRIP, (int)(rg * MV_TILE_ROWS), &full[stage]);
        tma_load_1d(dst + MV_P_BYTES, XfS + (long long)s * 1024, MV_X_BYTES, &full[stage]);
        if (++stage == MV_STAGES) { stage = 0; phase ^= 1; }
        if (++s == nstrips_row) { s = 0; ++rg; }
      }
    }
    return;
  }

  float acc[MV_ROWS][MV_B];
#pragma unroll
  for (int r = 0; r < MV_ROWS; ++r)
#pragma unroll
    for (int j = 0; j < MV_B; ++j) acc[r][j] = 0.f;

  int stage = 0;
  uint32_t phase = 0;
  const long long rg_first = u0 / nstrips_row;
  long long rg = rg_first;
  int s = (int)(u0 - rg * nstrips_row);
  for (long long u = u0; u < u1; ++u) {
    mbar_wait(&full[stage], phase);
    const float4* pt = reinterpret_cast<const float4*>(smem + stage * MV_STAGE_BYTES) + (warp * MV_ROWS) * (MV_STRIP / 4) + lane;
    const float4* xt = reinterpret_cast<const float4*>(smem + stage * MV_STAGE_BYTES + MV_P_BYTES) + lane;
    float4 p[MV_ROWS], x0[4], x1[4];
#pragma unroll
    for (int r = 0; r < MV_ROWS; ++r) p[r] = pt[r * (MV_STRIP / 4)];
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) { x0[cc] = xt[cc * 32]; x1[cc] = xt[(4 + cc) * 32]; }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);            // tile is in registers: free the stage
    if (++stage == MV_STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
      for (int r = 0; r < MV_ROWS; ++r) {
        const float pv = cc == 0 ? p[r].x : cc == 1 ? p[r].y : cc == 2 ? p[r].z : p[r].w;
        acc[r][0] = fmaf(pv, x0[cc].x, acc[r][0]); acc[r][1] = fmaf(pv, x0[cc].y, acc[r][1]);
        acc[r][2] = fmaf(pv, x0[cc].z, acc[r][2]); acc[r][3] = fmaf(pv, x0[cc].w, acc[r][3]);
        acc[r][4] = fmaf(pv, x1[cc].x, acc[r][4]); acc[r][5] = fmaf(pv, x1[cc].y, acc[r][5]);
        acc[r][6] = fmaf(pv, x1[cc].z, acc[r][6]); acc[r][7] = fmaf(pv, x1[cc].w, acc[r][7]);
      }
    }
    ++s;
    if (s == nstrips_row || u + 1 == u1) {
      // end of this row group's segment: warp-shuffle reduction in FP64 into the segment's slot;
      // lane (r*8+j) % 32 keeps value (r, j)
      double* slot = Ypart + ((cta * maxseg + (rg - rg_first)) * MV_TILE_ROWS + warp * MV_ROWS) * MV_B;
#pragma unroll
      for (int r = 0; r < MV_ROWS; ++r) {
#pragma unroll
        for (int j = 0; j < MV_B; ++j) {
          const double v = warp_sum((double)acc[r][j]);
          if (lane == ((r * MV_B + j) & 31)) slot[r * MV_B + j] = v;
          acc[r][j] = 0.f;
        }
      }
      s = 0;
      ++rg;
    }
  }
}

// Y[row][j] = sum over the CTAs whose unit range intersects the row's group, in CTA order
__global__ void reduce_segments_kernel(const double* __restrict__ Ypart, long long n_rows, int nstrips_row,
                                       long long units_per_cta, int maxseg, int b, double* __restrict__ Y) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_rows * MV_B; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / MV_B;
    const int j = (int)(e - i * MV_B);
    if (j >= b) continue;
    const long long rg = i / MV_TILE_ROWS;
    const int rr = (int)(i - rg * MV_TILE_ROWS);
    const long long c_first = (rg * nstrips_row) / units_per_cta;
    const long long c_last = ((rg + 1) * nstrips_row - 1) / units_per_cta;
    double sum = 0.0;
    for (long long c = c_first; c <= c_last; ++c) {
      const long long first_rg = (c * units_per_cta) / nstrips_row;
      sum += Ypart[((c * maxseg + (rg - first_rg)) * MV_TILE_ROWS + rr) * MV_B + j];
    }
    Y[i * b + j] = sum;
  }
}

// ---- sparse operator ------------------------------------------------------------------------
// P_ep has at most k + in-degree(i) non-zeros in row i: P = A + A^T with A the kNN lists themselves
// (v3s_build_markov_sparse).  Y[i,:] = sum_t a[i,t] X[nbr[i,t],:] + sum_{e in T(i)} t_val[e] X[t_col[e],:],
// FP64 throughout.  One warp per row: a warp-wide coalesced load fetches 32 (column, value) pairs,
// which are then broadcast by shuffles to 4 groups of 8 lanes -- lane j of a group gathers X[col, j],
// so a group reads one 64-byte row of X per entry and 8 independent gathers per lane are in flight.
// Fixed entry order per lane group + fixed shuffle tree => run-to-run deterministic and independent
// of the sharding.  Traffic: 64 B of L2 gathers per non-zero (X itself is N x 64 B and stays in L2)
// plus 12 B of list data.  Measured at cfg 2 (3.0 M entries): 36 us per launch, 30 MB from DRAM, L2
// throughput 12 %, warps active 34 % (profiles/r01_ncu_full_summary.txt) -- latency-bound; fetching
// the next 32 entries ahead of the gathers and one row per warp without a grid-stride tail were tried
// and changed nothing (36 -> 38 us), so the simple form stays.
constexpr int SP_THREADS = 1024;   // 32 warps work on 32 consecutive rows of the processing order: with the neighbourhood order
                                   // of the kNN stage these share most of their gathers (L1 hits); two CTAs per SM

__device__ __forceinline__ double sp_segment(const int* __restrict__ cols, const double* __restrict__ vals, int len,
                                             const double* __restrict__ X, int ldx, int j, int g, int lane, double acc) {
  for (int base = 0; base < len; base += 32) {
    const int e = base + lane;
    const int c = e < len ? cols[e] : 0;
    const double v = e < len ? vals[e] : 0.0;
    double xv[8], vv[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int src = it * 4 + g;
      const int cc = __shfl_sync(0xffffffffu, c, src);
      vv[it] = __shfl_sync(0xffffffffu, v, src);
      xv[it] = (vv[it] != 0.0 && j < ldx) ? X[(long long)cc * ldx + j] : 0.0;   // padding / thresholded entries: no load
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) acc = fma(vv[it], xv[it], acc);
  }
  return acc;
}

// entries [e0, e0+32) of a row's unified range (list part [0,k), transposed part [k,L)): lane -> (column, value)
__device__ __forceinline__ void sp_fetch(const int* __restrict__ colA, const double* __restrict__ valA, int k,
                                         const int* __restrict__ colT, const double* __restrict__ valT, int L, int e,
                                         int& c, double& v) {
  if (e < k) { c = colA[e]; v = valA[e]; }
  else if (e < L) { c = colT[e - k]; v = valT[e - k]; }
  else { c = 0; v = 0.0; }
}

// 32 entries held one per lane -> 8 x (4 entries, 8 vectors): shuffle-broadcast, gather, accumulate
__device__ __forceinline__ double sp_chunk(int c, double v, const double* __restrict__ X, int ldx, int j, int g, double acc) {
  double xv[8], vv[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int src = it * 4 + g;
    const int cc = __shfl_sync(0xffffffffu, c, src);
    vv[it] = __shfl_sync(0xffffffffu, v, src);
    xv[it] = (vv[it] != 0.0 && j < ldx) ? X[(long long)cc * ldx + j] : 0.0;   // padding / thresholded entries: no load
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) acc = fma(vv[it], xv[it], acc);
  return acc;
}

constexpr int SP_LONG = 2048;      // rows with more stored entries are shared by the 8 warps of a block (phase 2)

__global__ void __launch_bounds__(SP_THREADS)
sparse_block_matvec_kernel(const int* __restrict__ nbr, const double* __restrict__ a_val, int k,
                           const int* __restrict__ t_ptr, const int* __restrict__ t_col, const double* __restrict__ t_val,
                           long long row0, long long n_rows, const double* __restrict__ X, int ldx,
                           double* __restrict__ Y, int ldy, const int* __restrict__ order) {
  const int lane = threadIdx.x & 31, j = lane & 7, g = lane >> 3, w = threadIdx.x >> 5;
  // phase 1: one warp per row; every CTA walks one contiguous chunk of the processing order
  const long long per_cta = ceil_div64(n_rows, gridDim.x);
  const long long c0 = blockIdx.x * per_cta;
  const long long c1 = c0 + per_cta < n_rows ? c0 + per_cta : n_rows;
  for (long long idx = c0 + w; idx < c1; idx += SP_THREADS / 32) {
    const long long r = order ? order[idx] : idx;
    const long long i = row0 + r;
    const int p0 = t_ptr[i];
    const int d = t_ptr[i + 1] - p0;
    if (k + d > SP_LONG) continue;                     // a hub of the kNN graph: phase 2
    double acc = sp_segment(nbr + i * k, a_val + i * k, k, X, ldx, j, g, lane, 0.0);
    acc = sp_segment(t_col + p0, t_val + p0, d, X, ldx, j, g, lane, acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    if (g == 0 && j < ldy) Y[r * ldy + j] = acc;
  }
  // phase 2: long rows (in-degree in the thousands: hubs of noisy high-dimensional data), one block per row --
  // warp w takes the 32-entry chunks w, w+8, ... of the row's unified entry range, the 8 partial sums are added in
  // warp order.  The order depends on the row alone, so results stay independent of the sharding.
  __shared__ double part[SP_THREADS / 32][8];
  const long long per_block = ceil_div64(n_rows, gridDim.x);       // the block looks for long rows in its own contiguous range
  const long long rb0 = blockIdx.x * per_block;
  const long long rb1 = rb0 + per_block < n_rows ? rb0 + per_block : n_rows;
  for (long long rb = rb0; rb < rb1; rb += SP_THREADS) {
    const long long rt = rb + threadIdx.x;
    int is_long = 0;
    if (rt < rb1) is_long = (k + (t_ptr[row0 + rt + 1] - t_ptr[row0 + rt])) > SP_LONG;
    if (!__syncthreads_or(is_long)) continue;          // the usual case: one test per 256 rows
    const long long re = rb + SP_THREADS < rb1 ? rb + SP_THREADS : rb1;
    for (long long r = rb; r < re; ++r) {
      const long long i = row0 + r;
      const int p0 = t_ptr[i];
      const int L = k + (t_ptr[i + 1] - p0);
      if (L <= SP_LONG) continue;                      // block-uniform
      double acc = 0.0;
      for (int base = w * 32; base < L; base += SP_THREADS) {
        int c;
        double v;
        sp_fetch(nbr + i * k, a_val + i * k, k, t_col + p0, t_val + p0, L, base + lane, c, v);
        acc = sp_chunk(c, v, X, ldx, j, g, acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 8);
      acc += __shfl_xor_sync(0xffffffffu, acc, 16);
      if (g == 0) part[w][j] = acc;
      __syncthreads();
      if (w == 0 && g == 0 && j < ldy) {
        double s_ = part[0][j];
#pragma unroll
        for (int q = 1; q < SP_THREADS / 32; ++q) s_ += part[q][j];
        Y[r * ldy + j] = s_;
      }
      __syncthreads();
    }
  }
}

struct MvPlan { int grid; int nstrips_row; int maxseg; long long units_per_cta, total_units; };
static inline MvPlan mv_plan(long long n_rows, long long N) {
  MvPlan p;
  p.nstrips_row = (int)ceil_div64(N, MV_STRIP);
  const long long row_groups = ceil_div64(n_rows, MV_TILE_ROWS);
  p.total_units = row_groups * p.nstrips_row;
  long long ctas = 2LL * kNumSMs;
  long long upc = ceil_div64(p.total_units, ctas);
  // keep segments long enough to amortise the reduction (1/8 of a row group: measured 23.5 -> 20.2 us on a
  // 50 MB row block against 1/4, no difference on larger blocks)
  const long long min_upc = ceil_div64(p.nstrips_row, 8);
  if (upc < min_upc) upc = min_upc;
  if (upc < 1) upc = 1;
  p.units_per_cta = upc;
  p.grid = (int)ceil_div64(p.total_units, upc);
  p.maxseg = (int)ceil_div64(upc, p.nstrips_row) + 1;          // row groups a contiguous range can touch
  return p;
}

static inline size_t mv_xf_bytes(long long N) { return ((size_t)xf_strip_floats(N) * sizeof(float) + 255) & ~(size_t)255; }

// ---- per-launch device timing of the mat-vec kernel (bench.py's roofline): CUDA events recorded on
// the launching stream around block_matvec_kernel alone, owned by the library ------------------
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_ev;      // pairs (start, stop)
static size_t g_prof_used = 0;

static int prof_begin(cudaStream_t st, cudaEvent_t* e1) {
  *e1 = nullptr;
  if (!g_prof_on) return 0;
  if (g_prof_used + 2 > g_prof_ev.size()) {
    for (int i = 0; i < 2; ++i) { cudaEvent_t e; V3S_CUDA(cudaEventCreate(&e)); g_prof_ev.push_back(e); }
  }
  cudaEvent_t e0 = g_prof_ev[g_prof_used];
  *e1 = g_prof_ev[g_prof_used + 1];
  g_prof_used += 2;
  V3S_CUDA(cudaEventRecord(e0, st));
  return 0;
}

static int mv_launch(const float* P_rows, long long ldp, long long n_rows, long long N, const float* XfS, double* Ypart,
                     const MvPlan& pl, cudaStream_t st) {
  CUtensorMap tm;
  if (make_tmap_2d(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, P_rows, n_rows, N, ldp * 4, MV_STRIP, MV_TILE_ROWS,
                   CU_TENSOR_MAP_SWIZZLE_NONE)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    V3S_CUDA(cudaFuncSetAttribute(block_matvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MV_SMEM));
    attr_set = true;
  }
  static unsigned launch_parity = 0;
  cudaEvent_t e1 = nullptr;
  if (int rc = prof_begin(st, &e1)) return rc;
  block_matvec_kernel<<<pl.grid, MV_THREADS, MV_SMEM, st>>>(tm, XfS, n_rows, N, pl.nstrips_row, pl.units_per_cta,
                                                           pl.total_units, pl.maxseg, Ypart, (int)(launch_parity++ & 1));
  V3S_CHECK_LAUNCH();
  if (e1) V3S_CUDA(cudaEventRecord(e1, st));
  return 0;
}

// sparse operator: Y [n_rows, ldy] = P[row0 .. row0+n_rows, :] X,  X [N, ldx] FP64 (ldx = ldy = b <= 8)
static int sp_launch(const v3s_sparse_markov& op, long long row0, long long n_rows, const double* X, int ldx, double* Y,
                     int ldy, cudaStream_t st) {
  long long g = ceil_div64(n_rows, SP_THREADS / 32);
  if (g > 2LL * kNumSMs) g = 2LL * kNumSMs;
  const int* order = (op.row_order && op.order_row0 == row0 && op.order_rows == n_rows) ? op.row_order : nullptr;
  cudaEvent_t e1 = nullptr;
  if (int rc = prof_begin(st, &e1)) return rc;
  sparse_block_matvec_kernel<<<(int)g, SP_THREADS, 0, st>>>(op.nbr, op.a_val, op.k, op.t_ptr, op.t_col, op.t_val, row0, n_rows,
                                                          X, ldx, Y, ldy, order);
  V3S_CHECK_LAUNCH();
  if (e1) V3S_CUDA(cudaEventRecord(e1, st));
  return 0;
}

static int mv_run(const float* P_rows, long long ldp, long long n_rows, long long N, const float* XfS, int b,
                  double* Y_rows, double* Ypart, cudaStream_t st) {
  const MvPlan pl = mv_plan(n_rows, N);
  if (int rc = mv_launch(P_rows, ldp, n_rows, N, XfS, Ypart, pl, st)) return rc;
  long long g2 = ceil_div64(n_rows * MV_B, 256);
  if (g2 > 8LL * kNumSMs) g2 = 8LL * kNumSMs;
  reduce_segments_kernel<<<(int)g2, 256, 0, st>>>(Ypart, n_rows, pl.nstrips_row, pl.units_per_cta, pl.maxseg, b, Y_rows);
  V3S_CHECK_LAUNCH();
  return 0;
}

// Fused tail of a sharded mat-vec inside the Lanczos / Chebyshev recurrence: segment reduction of
// the rank's rows, the three-term combination  v = alpha * (P X) + beta * Yk + gamma * Ykm1  (FP64),
// and the "all-gather": every value is stored straight into the [N,8] result block -- and, as FP32,
// into the strip-layout operand of the next mat-vec -- of EVERY rank (local HBM or NVLink peer
// mappings), so no collective follows; `v3s_peer_barrier` orders the steps.
struct PeerPtrs { unsigned long long p[8]; };

// thread r of the caller: publish `epoch` in rank r's flag word for this rank, then wait until rank r has
// published at least `epoch` in ours.  flags.p[r] = rank r's int[world] flag array (zero-initialised once);
// epochs grow monotonically.  A peer that never arrives (a crashed rank) ends the wait after ~2 s and
// raises *status.
__device__ __forceinline__ void peer_signal_wait(const PeerPtrs& flags, int rank, int r, int epoch, int* status) {
  __threadfence_system();
  int* dst = reinterpret_cast<int*>(flags.p[r]) + rank;
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
  const int* src = reinterpret_cast<const int*>(flags.p[rank]) + r;
  int v = 0;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
    if (v - epoch >= 0) break;
    if (clock64() - t0 > 4000000000LL) { if (status) atomicExch(status, 3); break; }
  }
}

__global__ void reduce_combine_scatter_kernel(const double* __restrict__ Ypart, long long n_rows, long long row0,
                                              int nstrips_row, long long units_per_cta, int maxseg,
                                              const double* __restrict__ Yk, const double* __restrict__ Ykm1, double alpha,
                                              double beta, double gamma, PeerPtrs yout, PeerPtrs xfout, int world,
                                              int write_xf, PeerPtrs flags, int rank, int epoch, unsigned* ticket,
                                              int* status) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n_rows * MV_B; e += (long long)gridDim.x * blockDim.x) {
    const long long i = e / MV_B;
    const int j = (int)(e - i * MV_B);
    const long long rg = i / MV_TILE_ROWS;
    const int rr = (int)(i - rg * MV_TILE_ROWS);
    const long long c_first = (rg * nstrips_row) / units_per_cta;
    const long long c_last = ((rg + 1) * nstrips_row - 1) / units_per_cta;
    double py = 0.0;
    for (long long c = c_first; c <= c_last; ++c) {
      const long long first_rg = (c * units_per_cta) / nstrips_row;
      py += Ypart[((c * maxseg + (rg - first_rg)) * MV_TILE_ROWS + rr) * MV_B + j];
    }
    const long long ge = (row0 + i) * MV_B + j;
    double v = Yk ? fma(alpha, py, beta * Yk[ge]) : alpha * py;
    if (Ykm1) v = fma(gamma, Ykm1[ge], v);
#pragma unroll 1
    for (int r = 0; r < world; ++r) reinterpret_cast<double*>(yout.p[r])[ge] = v;
    if (write_xf) {
      const long long xi = xf_strip_index(row0 + i, j);
      const float f = (float)v;
#pragma unroll 1
      for (int r = 0; r < world; ++r) reinterpret_cast<float*>(xfout.p[r])[xi] = f;
    }
  }
  if (!ticket) return;
  // Sharded job: the block that finishes last signals every peer (all stores of this grid are ordered
  // before it by the fences around the ticket) and waits for every peer's signal, so the kernel -- and
  // with it the stream -- only moves on when all rows of the block have arrived from all ranks.
  __shared__ int is_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned tk = atomicAdd(ticket, 1u);
    is_last = (tk == gridDim.x - 1);
    if (is_last) *ticket = 0;
  }
  __syncthreads();
  if (is_last && threadIdx.x < world) peer_signal_wait(flags, rank, threadIdx.x, epoch, status);
}

static int mv_combine(const float* P_rows, long long ldp, long long n_rows, long long N, long long row0, const float* XfS,
                      const double* Yk, const double* Ykm1, double alpha, double beta, double gamma, const PeerPtrs& yout,
                      const PeerPtrs& xfout, int world, int write_xf, double* Ypart, const PeerPtrs& flags, int rank, int epoch,
                      unsigned* ticket, int* status, cudaStream_t st) {
  const MvPlan pl = mv_plan(n_rows, N);
  if (int rc = mv_launch(P_rows, ldp, n_rows, N, XfS, Ypart, pl, st)) return rc;
  long long g2 = ceil_div64(n_rows * MV_B, 256);
  if (g2 > 8LL * kNumSMs) g2 = 8LL * kNumSMs;
  reduce_combine_scatter_kernel<<<(int)g2, 256, 0, st>>>(Ypart, n_rows, row0, pl.nstrips_row, pl.units_per_cta, pl.maxseg, Yk,
                                                         Ykm1, alpha, beta, gamma, yout, xfout, world, write_xf, flags, rank,
                                                         epoch, world > 1 ? ticket : nullptr, status);
  V3S_CHECK_LAUNCH();
  return 0;
}

// The same fused tail behind the sparse mat-vec: its [n_rows,8] result is read as one segment per row
// group (nstrips_row = units_per_cta = maxseg = 1).
static int sp_combine(const v3s_sparse_markov& op, long long n_rows, long long row0, const double* X, const double* Yk,
                      const double* Ykm1, double alpha, double beta, double gamma, const PeerPtrs& yout, const PeerPtrs& xfout,
                      int world, int write_xf, double* PY, const PeerPtrs& flags, int rank, int epoch, unsigned* ticket,
                      int* status, cudaStream_t st) {
  if (int rc = sp_launch(op, row0, n_rows, X, MV_B, PY, MV_B, st)) return rc;
  long long g2 = ceil_div64(n_rows * MV_B, 256);
  if (g2 > 8LL * kNumSMs) g2 = 8LL * kNumSMs;
  reduce_combine_scatter_kernel<<<(int)g2, 256, 0, st>>>(PY, n_rows, row0, 1, 1, 1, Yk, Ykm1, alpha, beta, gamma, yout, xfout,
                                                         world, write_xf, flags, rank, epoch, world > 1 ? ticket : nullptr,
                                                         status);
  V3S_CHECK_LAUNCH();
  return 0;
}

// Stand-alone form of the same barrier (phase changes of the solver).
__global__ void peer_barrier_kernel(PeerPtrs flags, int rank, int world, int epoch, int* status) {
  if (threadIdx.x < world) peer_signal_wait(flags, rank, threadIdx.x, epoch, status);
}

}  // namespace v3s

using namespace v3s;

extern "C" long long v3s_xf_floats(long long N) { return xf_strip_floats(N); }

extern "C" long long v3s_block_matvec_workspace_bytes(long long n_rows, long long N) {
  const MvPlan pl = mv_plan(n_rows, N);
  return (long long)mv_xf_bytes(N) + (long long)sizeof(double) * pl.grid * pl.maxseg * MV_TILE_ROWS * MV_B + 256;
}

extern "C" int v3s_block_matvec(const float* P_rows, long long ldp, long long n_rows, long long N, const double* X, int b,
                                double* Y_rows, void* workspace, void* stream) {
  V3S_REQUIRE(b >= 1 && b <= MV_B, "v3s_block_matvec: block size b=%d outside [1,%d]", b, MV_B);
  V3S_REQUIRE(n_rows >= 0 && N >= 1 && ldp >= N && ldp % 4 == 0, "v3s_block_matvec: bad shape (ldp must be a multiple of 4)");
  V3S_REQUIRE(((uintptr_t)P_rows & 15) == 0 && ((uintptr_t)workspace & 15) == 0, "v3s_block_matvec: pointers must be 16-byte aligned");
  V3S_REQUIRE(N < (1LL << 31) && n_rows < (1LL << 31), "v3s_block_matvec: extents must fit int32");
  if (n_rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  float* Xf = reinterpret_cast<float*>(workspace);
  double* Ypart = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + mv_xf_bytes(N));
  V3S_CUDA(cudaMemsetAsync(Xf, 0, mv_xf_bytes(N), st));       // zero the padding columns of the last strip
  long long g = ceil_div64(N * MV_B, 256);
  if (g > 8LL * kNumSMs) g = 8LL * kNumSMs;
  x_to_f32_kernel<<<(int)g, 256, 0, st>>>(X, N, b, Xf);
  V3S_CHECK_LAUNCH();
  return mv_run(P_rows, ldp, n_rows, N, Xf, b, Y_rows, Ypart, st);
}

// Same product with the block already in FP32 strip layout (the form the Lanczos step emits;
// v3s_xf_floats(N) floats, padding columns zero).
extern "C" int v3s_block_matvec_f32(const float* P_rows, long long ldp, long long n_rows, long long N, const float* XfS,
                                    double* Y_rows, void* workspace, void* stream) {
  V3S_REQUIRE(n_rows >= 0 && N >= 1 && ldp >= N && ldp % 4 == 0, "v3s_block_matvec_f32: bad shape (ldp must be a multiple of 4)");
  V3S_REQUIRE(((uintptr_t)P_rows & 15) == 0 && ((uintptr_t)XfS & 15) == 0, "v3s_block_matvec_f32: pointers must be 16-byte aligned");
  V3S_REQUIRE(N < (1LL << 31) && n_rows < (1LL << 31), "v3s_block_matvec_f32: extents must fit int32");
  if (n_rows == 0) return 0;
  double* Ypart = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + mv_xf_bytes(N));
  return mv_run(P_rows, ldp, n_rows, N, XfS, MV_B, Y_rows, Ypart, (cudaStream_t)stream);
}

// ---- fused (sharded) forms used inside the eigen-solver ---------------------------------------

static inline int load_peers(PeerPtrs& dst, const unsigned long long* src, int world) {
  for (int r = 0; r < 8; ++r) dst.p[r] = (src && r < world) ? src[r] : 0ULL;
  return 0;
}

extern "C" int v3s_peer_barrier(const unsigned long long* flag_ptrs, int rank, int world, int epoch, int* status_dev,
                                void* stream) {
  V3S_REQUIRE(flag_ptrs && world >= 1 && world <= 8 && rank >= 0 && rank < world, "v3s_peer_barrier: bad arguments");
  if (world == 1) return 0;
  PeerPtrs f;
  load_peers(f, flag_ptrs, world);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(f, rank, world, epoch, status_dev);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_matvec_combine(const float* P_rows, long long ldp, long long n_rows, long long N, long long row0,
                                  const float* XfS, const double* Yk, const double* Ykm1, double alpha, double beta,
                                  double gamma, const unsigned long long* yout_ptrs, const unsigned long long* xfout_ptrs,
                                  const unsigned long long* flag_ptrs, int rank, int world, int epoch, unsigned* ticket_dev,
                                  int* status_dev, void* workspace, void* stream) {
  V3S_REQUIRE(n_rows >= 1 && N >= 1 && ldp >= N && ldp % 4 == 0 && row0 >= 0 && row0 + n_rows <= N,
              "v3s_matvec_combine: bad shape (ldp must be a multiple of 4, at least one row per rank)");
  V3S_REQUIRE(((uintptr_t)P_rows & 15) == 0 && ((uintptr_t)XfS & 15) == 0, "v3s_matvec_combine: pointers must be 16-byte aligned");
  V3S_REQUIRE(N < (1LL << 31) && n_rows < (1LL << 31), "v3s_matvec_combine: extents must fit int32");
  V3S_REQUIRE(yout_ptrs && world >= 1 && world <= 8 && rank >= 0 && rank < world && (world == 1 || (flag_ptrs && ticket_dev)),
              "v3s_matvec_combine: 1..8 ranks, result pointers (and flag words + ticket when sharded) required");
  PeerPtrs yo, xo, fl;
  load_peers(yo, yout_ptrs, world);
  load_peers(xo, xfout_ptrs, world);
  load_peers(fl, flag_ptrs, world);
  unsigned char* wsb = reinterpret_cast<unsigned char*>(workspace);
  double* Ypart = reinterpret_cast<double*>(wsb + mv_xf_bytes(N));
  return mv_combine(P_rows, ldp, n_rows, N, row0, XfS, Yk, Ykm1, alpha, beta, gamma, yo, xo, world, xfout_ptrs != nullptr,
                    Ypart, fl, rank, epoch, ticket_dev, status_dev, (cudaStream_t)stream);
}

// One application of the Chebyshev-filtered operator T_degree((P - c)/e) to the block Q, enqueued by
// ONE host call:  Y1 = (P Q - c Q)/e,  Y_{k+1} = 2 (P Y_k - c Y_k)/e - Y_{k-1}.  Every step is
// mat-vec kernel -> fused reduce/combine/scatter kernel -> (sharded) peer barrier; no collective.
// ybuf_ptrs[b * world + r] = address of rotating result buffer b (0..2) on rank r ([N,8] FP64),
// xf_ptrs[p * world + r]   = address of ping-pong FP32 strip operand p (0..1) on rank r.
// Q may be one of the rotating buffers (q_index 0..2) or a separate block (q_index = -1).
// *result_index = the rotating buffer that holds the result.
extern "C" int v3s_cheb_filter(const float* P_rows, long long ldp, long long n_rows, long long N, long long row0,
                               const float* Xf0, const double* Q, int q_index, int degree, double c, double e,
                               const unsigned long long* ybuf_ptrs, const unsigned long long* xf_ptrs,
                               const unsigned long long* flag_ptrs, int rank, int world, int epoch0, unsigned* ticket_dev,
                               int* status_dev, void* workspace, int* result_index, void* stream) {
  V3S_REQUIRE(degree >= 1 && q_index >= -1 && q_index <= 2 && e != 0.0, "v3s_cheb_filter: bad degree / buffer index / interval");
  V3S_REQUIRE(ybuf_ptrs && xf_ptrs && world >= 1 && world <= 8 && rank >= 0 && rank < world &&
                  (world == 1 || (flag_ptrs && ticket_dev)),
              "v3s_cheb_filter: bad rank layout");
  V3S_REQUIRE(n_rows >= 1 && N >= 1 && ldp >= N && ldp % 4 == 0 && row0 >= 0 && row0 + n_rows <= N && N < (1LL << 31),
              "v3s_cheb_filter: bad shape (at least one row per rank)");
  cudaStream_t st = (cudaStream_t)stream;
  double* Ypart = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(workspace) + mv_xf_bytes(N));
  auto local_y = [&](int b) { return reinterpret_cast<const double*>(ybuf_ptrs[(size_t)b * world + rank]); };
  auto local_x = [&](int p_) { return reinterpret_cast<const float*>(xf_ptrs[(size_t)p_ * world + rank]); };
  int prev = -2, cur = q_index;                    // -2: none, -1: the external Q
  const float* xin = Xf0;
  for (int s = 1; s <= degree; ++s) {
    int out = 0;
    while (out == cur || out == prev) ++out;       // lowest rotating buffer not read by this step
    PeerPtrs yo, xo;
    load_peers(yo, ybuf_ptrs + (size_t)out * world, world);
    const int write_xf = s < degree;
    load_peers(xo, xf_ptrs + (size_t)(s & 1) * world, world);
    const double* Yk = cur == -1 ? Q : local_y(cur);
    const double* Ykm1 = prev == -2 ? nullptr : (prev == -1 ? Q : local_y(prev));
    const double alpha = (s == 1 ? 1.0 : 2.0) / e, beta = (s == 1 ? -1.0 : -2.0) * c / e, gamma = s == 1 ? 0.0 : -1.0;
    PeerPtrs fl;
    load_peers(fl, flag_ptrs, world);
    if (int rc = mv_combine(P_rows, ldp, n_rows, N, row0, xin, Yk, Ykm1, alpha, beta, gamma, yo, xo, world, write_xf, Ypart, fl,
                            rank, epoch0 + s, ticket_dev, status_dev, st))
      return rc;
    xin = local_x(s & 1);
    prev = cur;
    cur = out;
  }
  if (result_index) *result_index = cur;
  return 0;
}

// ---- sparse-operator forms (same contracts; the operand is the FP64 block itself) -------------

static inline bool sp_valid(const v3s_sparse_markov* op) {
  return op && op->nbr && op->a_val && op->t_ptr && op->t_col && op->t_val && op->k >= 1;
}

extern "C" int v3s_sparse_block_matvec(const v3s_sparse_markov* op, long long row0, long long n_rows, long long N,
                                       const double* X, int b, double* Y_rows, void* stream) {
  V3S_REQUIRE(sp_valid(op), "v3s_sparse_block_matvec: incomplete operator descriptor");
  V3S_REQUIRE(b >= 1 && b <= MV_B, "v3s_sparse_block_matvec: block size b=%d outside [1,%d]", b, MV_B);
  V3S_REQUIRE(n_rows >= 0 && row0 >= 0 && row0 + n_rows <= N && N < (1LL << 31), "v3s_sparse_block_matvec: bad shape");
  if (n_rows == 0) return 0;
  return sp_launch(*op, row0, n_rows, X, b, Y_rows, b, (cudaStream_t)stream);
}

extern "C" int v3s_sparse_matvec_combine(const v3s_sparse_markov* op, long long n_rows, long long N, long long row0,
                                         const double* X, const double* Yk, const double* Ykm1, double alpha, double beta,
                                         double gamma, const unsigned long long* yout_ptrs,
                                         const unsigned long long* xfout_ptrs, const unsigned long long* flag_ptrs, int rank,
                                         int world, int epoch, unsigned* ticket_dev, int* status_dev, void* workspace,
                                         void* stream) {
  V3S_REQUIRE(sp_valid(op), "v3s_sparse_matvec_combine: incomplete operator descriptor");
  V3S_REQUIRE(n_rows >= 1 && N >= 1 && row0 >= 0 && row0 + n_rows <= N && N < (1LL << 31) && X && workspace,
              "v3s_sparse_matvec_combine: bad shape (at least one row per rank)");
  V3S_REQUIRE(yout_ptrs && world >= 1 && world <= 8 && rank >= 0 && rank < world && (world == 1 || (flag_ptrs && ticket_dev)),
              "v3s_sparse_matvec_combine: 1..8 ranks, result pointers (and flag words + ticket when sharded) required");
  PeerPtrs yo, xo, fl;
  load_peers(yo, yout_ptrs, world);
  load_peers(xo, xfout_ptrs, world);
  load_peers(fl, flag_ptrs, world);
  return sp_combine(*op, n_rows, row0, X, Yk, Ykm1, alpha, beta, gamma, yo, xo, world, xfout_ptrs != nullptr,
                    reinterpret_cast<double*>(workspace), fl, rank, epoch, ticket_dev, status_dev, (cudaStream_t)stream);
}

extern "C" int v3s_sparse_cheb_filter(const v3s_sparse_markov* op, long long n_rows, long long N, long long row0,
                                      const double* Q, int q_index, int degree, double c, double e,
                                      const unsigned long long* ybuf_ptrs, const unsigned long long* flag_ptrs, int rank,
                                      int world, int epoch0, unsigned* ticket_dev, int* status_dev, void* workspace,
                                      int* result_index, void* stream) {
  V3S_REQUIRE(sp_valid(op), "v3s_sparse_cheb_filter: incomplete operator descriptor");
  V3S_REQUIRE(degree >= 1 && q_index >= -1 && q_index <= 2 && e != 0.0 && Q, "v3s_sparse_cheb_filter: bad degree / buffer index / interval");
  V3S_REQUIRE(ybuf_ptrs && world >= 1 && world <= 8 && rank >= 0 && rank < world && (world == 1 || (flag_ptrs && ticket_dev)),
              "v3s_sparse_cheb_filter: bad rank layout");
  V3S_REQUIRE(n_rows >= 1 && N >= 1 && row0 >= 0 && row0 + n_rows <= N && N < (1LL << 31) && workspace,
              "v3s_sparse_cheb_filter: bad shape (at least one row per rank)");
  cudaStream_t st = (cudaStream_t)stream;
  double* PY = reinterpret_cast<double*>(workspace);
  auto local_y = [&](int b) { return reinterpret_cast<const double*>(ybuf_ptrs[(size_t)b * world + rank]); };
  int prev = -2, cur = q_index;                    // -2: none, -1: the external Q
  PeerPtrs fl, xo;
  load_peers(fl, flag_ptrs, world);
  load_peers(xo, nullptr, world);
  for (int s = 1; s <= degree; ++s) {
    int out = 0;
    while (out == cur || out == prev) ++out;       // lowest rotating buffer not read by this step
    PeerPtrs yo;
    load_peers(yo, ybuf_ptrs + (size_t)out * world, world);
    const double* Yk = cur == -1 ? Q : local_y(cur);
    const double* Ykm1 = prev == -2 ? nullptr : (prev == -1 ? Q : local_y(prev));
    const double alpha = (s == 1 ? 1.0 : 2.0) / e, beta = (s == 1 ? -1.0 : -2.0) * c / e, gamma = s == 1 ? 0.0 : -1.0;
    if (int rc = sp_combine(*op, n_rows, row0, Yk, Yk, Ykm1, alpha, beta, gamma, yo, xo, world, 0, PY, fl, rank, epoch0 + s,
                            ticket_dev, status_dev, st))
      return rc;
    prev = cur;
    cur = out;
  }
  if (result_index) *result_index = cur;
  return 0;
}

extern "C" int v3s_profile_matvec(int enable) {
  g_prof_on = enable != 0;
  g_prof_used = 0;
  return 0;
}

// Milliseconds between the END of recorded mat-vec launch i and the START of launch i+1 (what follows a mat-vec on
// the stream: the fused reduction / recurrence / exchange kernel with its cross-rank barrier, Lanczos steps, host
// gaps).  Call before v3s_profile_matvec_read (which resets the record); returns the number of gaps.
extern "C" int v3s_profile_matvec_gaps(float* ms_out, int cap) {
  const int n = (int)(g_prof_used / 2);
  if (n > 0) cudaEventSynchronize(g_prof_ev[g_prof_used - 1]);
  for (int i = 0; i + 1 < n && i < cap; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof_ev[2 * i + 1], g_prof_ev[2 * i + 2]) != cudaSuccess) ms = -1.f;
    ms_out[i] = ms;
  }
  return n > 0 ? n - 1 : 0;
}

// Elapsed milliseconds of the block_matvec_kernel launches recorded since v3s_profile_matvec(1)
// (synchronises on the last event); returns the number of launches, writes min(count, cap) values.
extern "C" int v3s_profile_matvec_read(float* ms_out, int cap) {
  const int n = (int)(g_prof_used / 2);
  if (n > 0) cudaEventSynchronize(g_prof_ev[g_prof_used - 1]);
  for (int i = 0; i < n && i < cap; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof_ev[2 * i], g_prof_ev[2 * i + 1]) != cudaSuccess) ms = -1.f;
    ms_out[i] = ms;
  }
  g_prof_used = 0;
  return n;
}
