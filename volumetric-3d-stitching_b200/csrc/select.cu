// K2b (second half) -- k nearest neighbours from the Gram row panel.
//
// Replaces distance.py:56-59 (arccos epilogue) and distance.py:26-48 (per-row full argsort).
// arccos is monotone, so the k nearest patterns of row r are its k largest Gram entries:
//   1. select_kernel : exact radix select (4 x 8-bit passes over the row, L2-resident) of the
//                      kk = k + margin largest g; ties broken towards the lower index
//   2. refine_sort_kernel : recompute the kk selected distances from the FP32 unit rows with the
//                      well-conditioned form d = 2 asin(|a_r - a_c| / 2) (== arccos(<a_r,a_c>) for
//                      unit vectors; arccos(g) itself amplifies FP32 error by 1/(d sin d)),
//                      bitonic-sort by (d, index) and keep the first k.
//                      With a distance workspace the refinement runs in two launches and every
//                      unordered pair (i, j) of rows of the same call that list each other is
//                      evaluated ONCE (by the lower row; the other row fetches the value) -- the
//                      refinement is bound by the reads of the candidate rows, and most kNN pairs
//                      are mutual.  Values are bit-identical to the one-launch form.
// Zero-norm rows (0/0 -> NaN in the reference, then NaN -> 0 at distance.py:58) get distance 0 to
// every pattern, as in the reference.
#include "common.cuh"
#include "v3s.h"

namespace v3s {

constexpr int kSelThreads = 256;

__device__ __forceinline__ float load_g(const float* __restrict__ grow, int c, bool row_zero,
                                        const int* __restrict__ zero_flag) {
  float g = grow[c];
  if (zero_flag && (row_zero || zero_flag[c])) g = 1.0f;
  return g;
}

__global__ void __launch_bounds__(kSelThreads)
select_kernel(const float* __restrict__ G, long long ldg, int n_rows, int n_cols, long long row0, int kk,
              float eps_g, int cap, const int* __restrict__ zero_flag, int* __restrict__ cand,
              int* __restrict__ cand_count, int* __restrict__ overflow) {
  __shared__ int hist[256];
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining, s_count_eq, s_slot, s_eq_taken;
  __shared__ int s_warp_cnt[kSelThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    const float* grow = G + (long long)r * ldg;
    const bool row_zero = zero_flag ? (zero_flag[row0 + r] != 0) : false;
    uint32_t prefix = 0, mask = 0;
    int remaining = kk;
    for (int shift = 24; shift >= 0; shift -= 8) {
      __syncthreads();
      hist[tid] = 0;
      __syncthreads();
      for (int c = tid; c < n_cols; c += kSelThreads) {
        const uint32_t key = float_to_key(load_g(grow, c, row_zero, zero_flag));
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, d = 255;
        for (; d > 0; --d) {
          if (cum + hist[d] >= remaining) break;
          cum += hist[d];
        }
        s_prefix = prefix | ((uint32_t)d << shift);
        s_remaining = remaining - cum;
        s_count_eq = hist[d];
      }
      __syncthreads();
      prefix = s_prefix;
      remaining = s_remaining;
      mask |= 255u << shift;
    }
    // prefix == key of the kk-th largest entry; `remaining` entries equal to it are still needed
    const uint32_t T = prefix;
    bool ties = (s_count_eq != remaining);
    __syncthreads();
    if (tid == 0) { s_slot = 0; s_eq_taken = 0; }
    __syncthreads();
    int* out = cand + (long long)r * cap;
    bool done = false;
    if (eps_g > 0.f && cap > kk) {
      // value margin: keep EVERY entry within eps_g of the kk-th largest.  With |G - G_exact| <= eps_g/2
      // this provably contains the exact top-kk set, whatever the tie / near-tie structure of the row.
      const float g_lo = key_to_float(T) - eps_g;
      for (int c = tid; c < n_cols; c += kSelThreads) {
        if (load_g(grow, c, row_zero, zero_flag) >= g_lo) {
          const int slot = atomicAdd(&s_slot, 1);
          if (slot < cap) out[slot] = c;
        }
      }
      __syncthreads();
      if (s_slot <= cap) {
        done = true;
        if (tid == 0) cand_count[r] = s_slot;
      } else {
        if (tid == 0 && overflow) atomicAdd(overflow, 1);   // band larger than the buffer: fall back to exactly kk
        ties = true;
      }
      __syncthreads();
      if (tid == 0) { s_slot = 0; s_eq_taken = 0; }
      __syncthreads();
    }
    if (done) continue;
    if (tid == 0) cand_count[r] = kk;
    if (!ties) {
      for (int c = tid; c < n_cols; c += kSelThreads) {
        const uint32_t key = float_to_key(load_g(grow, c, row_zero, zero_flag));
        if (key >= T) out[atomicAdd(&s_slot, 1)] = c;
      }
    } else {
      // ordered pass: entries > T unconditionally, entries == T lowest index first
      for (int c0 = 0; c0 < n_cols; c0 += kSelThreads) {
        const int c = c0 + tid;
        uint32_t key = 0;
        if (c < n_cols) key = float_to_key(load_g(grow, c, row_zero, zero_flag));
        const bool gt = (c < n_cols) && key > T;
        const bool eq = (c < n_cols) && key == T;
        if (gt) out[atomicAdd(&s_slot, 1)] = c;
        const unsigned bal = __ballot_sync(0xffffffffu, eq);
        if (lane == 0) s_warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int before = s_eq_taken;
        for (int w = 0; w < warp; ++w) before += s_warp_cnt[w];
        before += __popc(bal & ((1u << lane) - 1u));
        if (eq && before < remaining) out[atomicAdd(&s_slot, 1)] = c;
        __syncthreads();
        if (tid == 0) {
          int tot = 0;
          for (int w = 0; w < kSelThreads / 32; ++w) tot += s_warp_cnt[w];
          s_eq_taken += tot;
        }
        __syncthreads();
      }
    }
  }
}

constexpr int kRefThreads = 256;

// Sharded tiled refinement: the pair sums of rank o's rows live in dss[o] ([cnt_o, cap] FP64, NVLink peer mappings);
// rank o owns cnt_base + (o < rem) consecutive patterns.  world = 0: one rank, everything in the local workspace.
struct RefPeers {
  const double* dss[8];
  int world, rem;
  long long cnt_base;
};
__device__ __forceinline__ const double* ref_peer_row(const RefPeers& pr, long long x, int cap) {
  const long long big = pr.cnt_base + 1, split = big * pr.rem;
  int o;
  long long lo;
  if (x < split) { o = (int)(x / big); lo = o * big; }
  else { o = pr.rem + (int)((x - split) / pr.cnt_base); lo = split + (long long)(o - pr.rem) * pr.cnt_base; }
  return pr.dss[o] + (x - lo) * cap;
}

__global__ void __launch_bounds__(kRefThreads)
refine_sort_kernel(const float* __restrict__ a_rows, const float* __restrict__ a_all, int P, long long row0, int n_rows,
                   const int* __restrict__ cand, const int* __restrict__ cand_count, int cap, int kk_pad, int k,
                   const int* __restrict__ zero_flag,
                   int cache_row, const double* __restrict__ Bd, long long ldb, double* __restrict__ S2,
                   int* __restrict__ Nbr, double* Dws, int phase, const int* __restrict__ order,
                   const int* __restrict__ pref = nullptr, const RefPeers peers = RefPeers{}) {
  // phase 0: distances + sort in one launch.  phase 1: distances only, into Dws [n_rows, cap] in candidate
  // order; a pair whose partner row j < i of this call lists i too is skipped and stored as -(t + 1),
  // t = position of i in j's list.  phase 2: fetch the skipped values from the partner rows, sort, emit.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sd = reinterpret_cast<double*>(smem_raw);          // kk_pad
  int* si = reinterpret_cast<int*>(sd + kk_pad);              // kk_pad
  float* srow = reinterpret_cast<float*>(si + kk_pad);        // P (if cache_row)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool vec4 = (P % 4 == 0);

  // `order` (optional): processing order of the rows -- rows handled at the same time by the grid then
  // share candidates (neighbouring patterns), so their row reads hit L2; results do not depend on it
  for (int it = blockIdx.x; it < n_rows; it += gridDim.x) {
    const int r = order ? order[it] : it;
    __syncthreads();
    const int kk = cand_count[r];
    const float* arow = a_rows ? a_rows + (long long)r * P : nullptr;
    if (Bd) {   // dense mode: exact FP64 values gathered from the caller's matrix
      for (int s = tid; s < kk_pad; s += kRefThreads) {
        const bool ok = s < kk;
        const int j = ok ? cand[(long long)r * cap + s] : 0x7fffffff;
        sd[s] = ok ? Bd[(long long)r * ldb + j] : 1e300;
        si[s] = j;
      }
    }
    if (!Bd && cache_row && phase != 2) {
      for (int p = tid; p < P; p += kRefThreads) srow[p] = arow[p];
    }
    if (!Bd) {
      for (int s = tid; s < kk_pad; s += kRefThreads) {
        sd[s] = 1e300;   // padding sorts last
        si[s] = 0x7fffffff;
      }
    }
    __syncthreads();
    const float* xr = cache_row ? srow : arow;
    const bool rz = zero_flag ? (zero_flag[row0 + r] != 0) : false;
    if (phase == 2) {
      for (int s = tid; s < kk; s += kRefThreads) {
        const int j = cand[(long long)r * cap + s];
        double d = Dws[(long long)r * cap + s];
        if (d < 0.0) d = Dws[(long long)(j - row0) * cap + (long long)(-d) - 1];   // evaluated by the partner row
        sd[s] = d;
        si[s] = j;
      }
    }
    if (phase == 3) {
      // tiled refinement (refine_tiled_kernel): Dws holds |a_r - a_j|^2 of the pairs this row owns, pref the slot
      // of the partner row that owns the others
      for (int s = tid; s < kk; s += kRefThreads) {
        const int j = cand[(long long)r * cap + s];
        const int pr = pref[(long long)r * cap + s];
        // (a pair owned by the partner row: its sum is in that row's workspace -- on this rank, or read from the
        // owner rank's memory over NVLink in a sharded job)
        const double ss = pr < 0 ? Dws[(long long)r * cap + s]
                                 : (peers.world > 0 ? ref_peer_row(peers, j, cap)[pr] : Dws[(long long)(j - row0) * cap + pr]);
        double half = 0.5 * sqrt(ss);
        if (half > 1.0) half = 1.0;
        double d = 2.0 * asin(half);
        if (zero_flag && (rz || zero_flag[j])) d = 0.0;
        sd[s] = d;
        si[s] = j;
      }
    }
    for (int s = warp; s < ((Bd || phase >= 2) ? 0 : kk); s += kRefThreads / 32) {
      const int j = cand[(long long)r * cap + s];
      if (phase == 1) {
        const long long lj = (long long)j - row0;
        if (lj >= 0 && lj < r) {                              // partner row of this call with a lower index
          const int* cj = cand + lj * cap;
          const int cntj = cand_count[lj];
          const int me = (int)(row0 + r);
          int pos = -1;
          for (int t0 = 0; t0 < cntj && pos < 0; t0 += 32) {
            const int t = t0 + lane;
            const unsigned hit = __ballot_sync(0xffffffffu, t < cntj && cj[t] == me);
            if (hit) pos = t0 + __ffs(hit) - 1;
          }
          if (pos >= 0) {
            if (lane == 0) Dws[(long long)r * cap + s] = -(double)(pos + 1);
            continue;
          }
        }
      }
      const float* aj = a_all + (long long)j * P;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      if (vec4) {
        const float4* x4 = reinterpret_cast<const float4*>(xr);
        const float4* y4 = reinterpret_cast<const float4*>(aj);
        const int n4 = P / 4;
        int p = lane;
        // 4 independent 128-bit loads in flight per lane (2 kB per warp): the gather is HBM/L2-latency bound
        for (; p + 96 < n4; p += 128) {
          const float4 y0 = __ldg(&y4[p]), y1 = __ldg(&y4[p + 32]), y2 = __ldg(&y4[p + 64]), y3 = __ldg(&y4[p + 96]);
          const float4 x0 = x4[p], x1 = x4[p + 32], x2 = x4[p + 64], x3 = x4[p + 96];
          float d;
          d = x0.x - y0.x; a0 = fmaf(d, d, a0); d = x0.y - y0.y; a1 = fmaf(d, d, a1); d = x0.z - y0.z; a2 = fmaf(d, d, a2); d = x0.w - y0.w; a3 = fmaf(d, d, a3);
          d = x1.x - y1.x; a0 = fmaf(d, d, a0); d = x1.y - y1.y; a1 = fmaf(d, d, a1); d = x1.z - y1.z; a2 = fmaf(d, d, a2); d = x1.w - y1.w; a3 = fmaf(d, d, a3);
          d = x2.x - y2.x; a0 = fmaf(d, d, a0); d = x2.y - y2.y; a1 = fmaf(d, d, a1); d = x2.z - y2.z; a2 = fmaf(d, d, a2); d = x2.w - y2.w; a3 = fmaf(d, d, a3);
          d = x3.x - y3.x; a0 = fmaf(d, d, a0); d = x3.y - y3.y; a1 = fmaf(d, d, a1); d = x3.z - y3.z; a2 = fmaf(d, d, a2); d = x3.w - y3.w; a3 = fmaf(d, d, a3);
        }
        for (; p < n4; p += 32) {
          const float4 x = x4[p], y = __ldg(&y4[p]);
          const float d0 = x.x - y.x, d1 = x.y - y.y, d2 = x.z - y.z, d3 = x.w - y.w;
          a0 = fmaf(d0, d0, a0); a1 = fmaf(d1, d1, a1); a2 = fmaf(d2, d2, a2); a3 = fmaf(d3, d3, a3);
        }
      } else {
        for (int p = lane; p < P; p += 32) {
          const float d0 = xr[p] - __ldg(&aj[p]);
          a0 = fmaf(d0, d0, a0);
        }
      }
      double ss = ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
      ss = warp_sum(ss);
      if (lane == 0) {
        double half = 0.5 * sqrt(ss);
        if (half > 1.0) half = 1.0;
        double d = 2.0 * asin(half);
        if (zero_flag && (rz || zero_flag[j])) d = 0.0;
        if (phase == 1) {
          Dws[(long long)r * cap + s] = d;
        } else {
          sd[s] = d;
          si[s] = j;
        }
      }
    }
    if (phase == 1) continue;
    __syncthreads();
    // bitonic sort of (d, idx) pairs, ascending by d then idx, over the smallest power of two that holds this row's
    // kk candidates (the padding up to kk_pad sorts last and is never read: kk >= k)
    int kp = 4;
    while (kp < kk) kp <<= 1;
    if (kp > kk_pad) kp = kk_pad;
    for (int size = 2; size <= kp; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < kp / 2; t += kRefThreads) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const double dl = sd[lo], dh = sd[hi];
          const int il = si[lo], ih = si[hi];
          const bool gt = (dl > dh) || (dl == dh && il > ih);
          if (gt == up) { sd[lo] = dh; sd[hi] = dl; si[lo] = ih; si[hi] = il; }
        }
        __syncthreads();
      }
    }
    for (int s = tid; s < k; s += kRefThreads) {
      S2[(long long)r * k + s] = sd[s];
      Nbr[(long long)r * k + s] = si[s];
    }
  }
}

}  // namespace v3s

using namespace v3s;

static size_t g_refine_smem_set = 0;   // largest dynamic shared-memory size refine_sort_kernel was opted into

extern "C" int v3s_knn_from_gram(const float* G, long long ldg, long long n_rows, long long n_cols, long long row0,
                                 const float* a_rows, const float* a_all, int P, int k, int margin, float eps_g, int cap,
                                 const int* zero_flag, int* cand_ws, int* count_ws, int* overflow_flag, double* S2,
                                 int* Nbr, double* dist_ws, void* stream) {
  V3S_REQUIRE(k >= 1 && margin >= 0, "v3s_knn_from_gram: bad k/margin");
  V3S_REQUIRE(k <= n_cols, "Number of nearest neighbors to preserve is too high.");   // distance.py:33-34
  V3S_REQUIRE(n_rows >= 0 && n_cols < (1LL << 31) && n_rows < (1LL << 31) && ldg >= n_cols, "v3s_knn_from_gram: bad shape");
  if (n_rows == 0) return 0;
  long long kk = (long long)k + margin;
  if (kk > n_cols) kk = n_cols;
  if (cap < kk) cap = (int)kk;
  if (cap > n_cols) cap = (int)n_cols;
  int kk_pad = 4;
  while (kk_pad < cap) kk_pad <<= 1;
  V3S_REQUIRE(kk_pad <= 4096, "v3s_knn_from_gram: candidate capacity %d too large (max 4096)", cap);
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (int)(n_rows < 16LL * kNumSMs ? n_rows : 16LL * kNumSMs);
  select_kernel<<<grid, kSelThreads, 0, st>>>(G, ldg, (int)n_rows, (int)n_cols, row0, (int)kk, eps_g, cap, zero_flag,
                                              cand_ws, count_ws, overflow_flag);
  V3S_CHECK_LAUNCH();
  const size_t base = (size_t)kk_pad * (sizeof(double) + sizeof(int));
  const int cache_row = (base + (size_t)P * 4 <= 200 * 1024) ? 1 : 0;
  const size_t smem = base + (cache_row ? (size_t)P * 4 : 0);
  if (smem > g_refine_smem_set) {
    V3S_CUDA(cudaFuncSetAttribute(refine_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    g_refine_smem_set = smem;
  }
  const int grid2 = (int)(n_rows < 8LL * kNumSMs ? n_rows : 8LL * kNumSMs);
  if (dist_ws) {
    refine_sort_kernel<<<grid2, kRefThreads, smem, st>>>(a_rows, a_all, P, row0, (int)n_rows, cand_ws, count_ws, cap, kk_pad,
                                                        k, zero_flag, cache_row, nullptr, 0, S2, Nbr, dist_ws, 1, nullptr);
    V3S_CHECK_LAUNCH();
    refine_sort_kernel<<<grid2, kRefThreads, base, st>>>(a_rows, a_all, P, row0, (int)n_rows, cand_ws, count_ws, cap, kk_pad,
                                                        k, zero_flag, 0, nullptr, 0, S2, Nbr, dist_ws, 2, nullptr);
    V3S_CHECK_LAUNCH();
    return 0;
  }
  refine_sort_kernel<<<grid2, kRefThreads, smem, st>>>(a_rows, a_all, P, row0, (int)n_rows, cand_ws, count_ws, cap, kk_pad,
                                                      k, zero_flag, cache_row, nullptr, 0, S2, Nbr, nullptr, 0, nullptr);
  V3S_CHECK_LAUNCH();
  return 0;
}

// Exact distances + sort for candidate lists that are already selected (v3s_gram_topk +
// v3s_knn_merge_lists): the refinement half of v3s_knn_from_gram.  cand_ws int [n_rows,cap] /
// count_ws int [n_rows] are inputs; dist_ws FP64 [n_rows,cap] (NULL: one launch, no pair sharing).
// order (optional, int [n_rows]): processing order of the rows (a locality hint, not a result).
extern "C" int v3s_knn_refine(const float* a_rows, const float* a_all, int P, long long row0, long long n_rows,
                              long long n_cols, int k, int cap, const int* zero_flag, const int* cand_ws,
                              const int* count_ws, double* S2, int* Nbr, double* dist_ws, const int* order, void* stream) {
  V3S_REQUIRE(k >= 1 && cap >= k && n_rows >= 0 && n_cols < (1LL << 31) && n_rows < (1LL << 31), "v3s_knn_refine: bad arguments");
  V3S_REQUIRE(k <= n_cols, "Number of nearest neighbors to preserve is too high.");   // distance.py:33-34
  if (n_rows == 0) return 0;
  int kk_pad = 4;
  while (kk_pad < cap) kk_pad <<= 1;
  V3S_REQUIRE(kk_pad <= 4096, "v3s_knn_refine: candidate capacity %d too large (max 4096)", cap);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t base = (size_t)kk_pad * (sizeof(double) + sizeof(int));
  const int cache_row = (base + (size_t)P * 4 <= 200 * 1024) ? 1 : 0;
  const size_t smem = base + (cache_row ? (size_t)P * 4 : 0);
  if (smem > g_refine_smem_set) {
    V3S_CUDA(cudaFuncSetAttribute(refine_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    g_refine_smem_set = smem;
  }
  const int grid2 = (int)(n_rows < 8LL * kNumSMs ? n_rows : 8LL * kNumSMs);
  if (!dist_ws) {       // one launch, every pair evaluated by both of its rows
    refine_sort_kernel<<<grid2, kRefThreads, smem, st>>>(a_rows, a_all, P, row0, (int)n_rows, cand_ws, count_ws, cap, kk_pad, k,
                                                        zero_flag, cache_row, nullptr, 0, S2, Nbr, nullptr, 0, order);
    V3S_CHECK_LAUNCH();
    return 0;
  }
  refine_sort_kernel<<<grid2, kRefThreads, smem, st>>>(a_rows, a_all, P, row0, (int)n_rows, cand_ws, count_ws, cap, kk_pad, k,
                                                      zero_flag, cache_row, nullptr, 0, S2, Nbr, dist_ws, 1, order);
  V3S_CHECK_LAUNCH();
  refine_sort_kernel<<<grid2, kRefThreads, base, st>>>(a_rows, a_all, P, row0, (int)n_rows, cand_ws, count_ws, cap, kk_pad, k,
                                                      zero_flag, 0, nullptr, 0, S2, Nbr, dist_ws, 2, order);
  V3S_CHECK_LAUNCH();
  return 0;
}

namespace v3s {
// ---- tiled exact-distance refinement ---------------------------------------------------------------------
// The refinement reads one candidate row (P floats) per (row, candidate) pair: ~N k P 4 bytes, and consecutive
// rows have unrelated candidates when the patterns come in random order.  Here the rows are processed in TILES
// of neighbouring patterns (rows that share their nearest anchor pattern, v3s_knn_merge_lists' `order`): the
// tile's pairs are grouped by candidate, so a candidate row is read ONCE per tile and compared with every row
// of the tile that lists it, from a shared-memory copy of the tile's rows, chunk by chunk of the pixel axis.
//   refine_prepare_kernel : per tile, the pairs it owns (a pair listed from both sides by rows of this call is owned
//                           by the row with the lower index; the other row records the owner's slot in `pref`),
//                           sorted by candidate -> keys [tile][RT_KEYS] = candidate << 32 | row-in-tile << 16 | slot
//   refine_tiled_kernel   : per tile and chunk of RT_CH pixels: the tile's rows -> shared memory; each warp walks
//                           its share of the sorted pairs, keeps the current candidate's chunk in registers, adds
//                           the chunk's |a_r - a_c|^2 (FP32, warp-shuffle sum) into Dss [n_rows, cap] (FP64 RED)
//   refine_sort_kernel phase 3 : d = 2 asin(sqrt(ss)/2), sort by (d, index), emit the first k
constexpr int RT_CH = 512;
constexpr int RT_KEYS = 16384;
constexpr int RT_QMAX = 48;
constexpr int RT_THREADS = 256;      // prepare kernel
constexpr int RT_MAIN_THREADS = 512; // main kernel: 16 warps x 2 CTAs per SM hide the candidate-row load latency

__global__ void __launch_bounds__(RT_THREADS)
refine_prepare_kernel(const int* __restrict__ order, const int* __restrict__ tiles, long long row0,
                      const int* __restrict__ cand, const int* __restrict__ cand_count, int cap,
                      unsigned long long* __restrict__ keys, int* __restrict__ np_out, int* __restrict__ pref, int dedup,
                      const int* __restrict__ cand_all, const int* __restrict__ count_all) {
  // cand_all / count_all (sharded job, else NULL): the candidate lists of ALL patterns (all-gathered) -- a pair listed
  // from both sides is then evaluated once whichever ranks own its two rows: by the lower index when the indices
  // have equal parity, else by the higher (both directions carry about half of every row's mutual pairs)
  extern __shared__ unsigned long long sk[];       // RT_KEYS
  __shared__ int s_np;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  if (tile >= tiles[0]) return;
  const int q0 = tiles[1 + 2 * tile];
  const int nq = tiles[2 + 2 * tile];
  if (tid == 0) s_np = 0;
  __syncthreads();
  for (int qi = warp; qi < nq; qi += RT_THREADS / 32) {
    const int r = order[q0 + qi];
    const int cnt = cand_count[r];
    const int me = (int)(row0 + r);
    for (int s0 = 0; s0 < cnt; s0 += 32) {
      const int s = s0 + lane;
      const bool valid = s < cnt;
      const int j = valid ? cand[(long long)r * cap + s] : -1;
      bool owned = valid;
      if (dedup && valid) {
        const long long lj = (long long)j - row0;
        const bool partner_first = cand_all ? (j != me && ((((me ^ j) & 1) == 0) == (j < me)))   // the partner would own it
                                            : (lj >= 0 && lj < r);                                // partner row of this call with a lower index
        if (partner_first) {
          const int* cj = cand_all ? cand_all + (long long)j * cap : cand + lj * cap;
          const int cntj = cand_all ? count_all[j] : cand_count[lj];
          int lo = 0, hi = cntj;                              // lists are sorted by column index (knn_merge_kernel)
          while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (cj[mid] < me) lo = mid + 1; else hi = mid;
          }
          if (lo < cntj && cj[lo] == me) { owned = false; pref[(long long)r * cap + s] = lo; }
        }
      }
      const unsigned ob = __ballot_sync(0xffffffffu, owned);
      int base = 0;
      if (lane == 0 && ob) base = atomicAdd(&s_np, __popc(ob));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (owned) sk[base + __popc(ob & ((1u << lane) - 1u))] = ((unsigned long long)(unsigned)j << 32) | ((unsigned)qi << 16) | (unsigned)s;
    }
  }
  __syncthreads();
  const int np = s_np;
  int npad = 2;
  while (npad < np) npad <<= 1;
  for (int i = np + tid; i < npad; i += RT_THREADS) sk[i] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= npad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < npad / 2; t += RT_THREADS) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = sk[lo], b = sk[hi];
        if ((a > b) == up) { sk[lo] = b; sk[hi] = a; }
      }
      __syncthreads();
    }
  }
  unsigned long long* out = keys + (long long)tile * RT_KEYS;
  for (int i = tid; i < np; i += RT_THREADS) out[i] = sk[i];
  if (tid == 0) np_out[tile] = np;
}

__global__ void __launch_bounds__(RT_MAIN_THREADS, 2)
refine_tiled_kernel(const float* __restrict__ a_rows, const float* __restrict__ a_all, int P, const int* __restrict__ order,
                    const int* __restrict__ tiles, const unsigned long long* __restrict__ keys,
                    const int* __restrict__ np_arr, int cap, double* __restrict__ Dss) {
  extern __shared__ __align__(16) float sq[];       // qt * RT_CH
  __shared__ int s_rows[RT_QMAX];
  constexpr int RT_THREADS = RT_MAIN_THREADS;       // (shadows the prepare kernel's block size inside this kernel)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_tiles = tiles[0];
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int q0 = tiles[1 + 2 * tile];
    const int nq = tiles[2 + 2 * tile];
    const int np = np_arr[tile];
    const unsigned long long* kb = keys + (long long)tile * RT_KEYS;
    __syncthreads();
    if (tid < nq) s_rows[tid] = order[q0 + tid];
    const int i0 = (int)((long long)np * warp / (RT_THREADS / 32)), i1 = (int)((long long)np * (warp + 1) / (RT_THREADS / 32));
    for (int c0 = 0; c0 < P; c0 += RT_CH) {
      __syncthreads();
      // the NEXT chunk of the tile's rows -> L2 now (one 128-byte line per request): its load phase, which all warps
      // of the CTA wait for, then finds the lines on chip
      if (c0 + RT_CH < P) {
        for (int idx = tid; idx < nq * (RT_CH / 32); idx += RT_THREADS) {
          const int qi = idx / (RT_CH / 32), l = idx - qi * (RT_CH / 32);
          const int col = c0 + RT_CH + 32 * l;
          if (col < P) asm volatile("prefetch.global.L2 [%0];" ::"l"(a_rows + (long long)s_rows[qi] * P + col));
        }
      }
      for (int idx = tid; idx < nq * (RT_CH / 4); idx += RT_THREADS) {
        const int qi = idx / (RT_CH / 4), v = idx - qi * (RT_CH / 4);
        const int col = c0 + 4 * v;
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < P) val = *reinterpret_cast<const float4*>(a_rows + (long long)s_rows[qi] * P + col);
        *reinterpret_cast<float4*>(sq + qi * RT_CH + 4 * v) = val;
      }
      __syncthreads();
      // The warp walks its pairs in batches of 32 keys (one coalesced load, then shuffles): no memory latency per
      // pair.  The candidate whose chunk the NEXT segment needs is fetched into the second register buffer
      // while the current segment is being compared (the loads are what the kernel waits for: ~6 pairs of
      // arithmetic per 2 kB row chunk).
      int cur_u = -1;
      float4 ua[RT_CH / 128], ub[RT_CH / 128];
      const float* a_chunk = a_all + c0;
      auto load_u = [&](float4 (&dst)[RT_CH / 128], int u) {
        const float* up = a_chunk + (long long)u * P;
#pragma unroll
        for (int m = 0; m < RT_CH / 128; ++m) {
          const int col = 4 * (lane + 32 * m);
          dst[m] = (c0 + col < P) ? __ldg(reinterpret_cast<const float4*>(up + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      // |q - u|^2 of one pair over this lane's 16 pixels of the chunk: packed FP32 (FFMA2: two lanes per
      // instruction -- d = fma(u, -1, q) is the exactly rounded difference, then fma(d, d, acc)); four running sums
      // in the same element order as four scalar accumulators
      auto pair_partial = [&](const float4 (&uu)[RT_CH / 128], int qi) -> float {
        const float* q = sq + qi * RT_CH;
        float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
        const float2 neg1 = make_float2(-1.f, -1.f);
#pragma unroll
        for (int m = 0; m < RT_CH / 128; ++m) {
          const float4 x = *reinterpret_cast<const float4*>(q + 4 * (lane + 32 * m));
          const float2 d01 = __ffma2_rn(make_float2(uu[m].x, uu[m].y), neg1, make_float2(x.x, x.y));
          const float2 d23 = __ffma2_rn(make_float2(uu[m].z, uu[m].w), neg1, make_float2(x.z, x.w));
          a01 = __ffma2_rn(d01, d01, a01);
          a23 = __ffma2_rn(d23, d23, a23);
        }
        return (a01.x + a01.y) + (a23.x + a23.y);
      };
      // pairs j0 .. j1-1 of the current batch (all with the candidate chunk `uu`), four at a time: the four per-lane
      // partial sums are reduced across the warp together -- 6 shuffles for 4 pairs instead of 5 per pair, every
      // step adding the same two operands as the xor butterfly of warp_sum (bit-identical totals) -- and end up in
      // lanes 0 / 8 / 16 / 24, which add them to the pairs' FP64 sums in parallel
      auto seg_pairs = [&](const float4 (&uu)[RT_CH / 128], unsigned my_lo, int j0, int j1) {
        const bool b4 = lane & 16, b3 = lane & 8;
        for (int j = j0; j < j1; j += 4) {
          float p[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            p[t] = 0.f;
            if (j + t < j1) {                                    // warp-uniform
              const unsigned lo = __shfl_sync(0xffffffffu, my_lo, j + t);
              p[t] = pair_partial(uu, (int)(lo >> 16));
            }
          }
          const float r0 = __shfl_xor_sync(0xffffffffu, b4 ? p[0] : p[2], 16);
          const float r1 = __shfl_xor_sync(0xffffffffu, b4 ? p[1] : p[3], 16);
          const float k0 = (b4 ? p[2] : p[0]) + r0;              // pairs 0 / 2 (lanes 0-15 / 16-31)
          const float k1 = (b4 ? p[3] : p[1]) + r1;              // pairs 1 / 3
          float v = (b3 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, b3 ? k0 : k1, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          const int t = ((lane >> 4) << 1) | ((lane >> 3) & 1);  // the pair whose total this lane holds
          const unsigned lo = __shfl_sync(0xffffffffu, my_lo, min(j + t, 31));
          if ((lane & 7) == 0 && j + t < j1)
            atomicAdd(&Dss[(long long)s_rows[lo >> 16] * cap + (int)(lo & 0xffffu)], (double)v);
        }
      };
      for (int base = i0; base < i1; base += 32) {
        const int nb = min(32, i1 - base);
        const unsigned long long my_key = (lane < nb) ? __ldg(&kb[base + lane]) : ~0ull;
        const int my_u = (int)(my_key >> 32);
        const unsigned my_lo = (unsigned)my_key;                 // row in tile << 16 | slot
        const int prev_u = __shfl_up_sync(0xffffffffu, my_u, 1);
        // segment starts of this batch: a key whose candidate differs from its predecessor's (lane 0: from the
        // candidate still held in `ua` from the previous batch)
        unsigned starts = __ballot_sync(0xffffffffu, lane < nb && (lane == 0 ? (my_u != cur_u) : (my_u != prev_u)));
        // every candidate chunk this batch will need -> L2 now: the register double buffer below covers one segment
        // (a few pairs) of latency, not a DRAM access; the batch's later segments then hit in L2
        for (unsigned rest = starts; rest;) {
          const int js = __ffs(rest) - 1;
          rest &= rest - 1;
          const int u = __shfl_sync(0xffffffffu, my_u, js);
          if (lane < RT_CH / 32 && c0 + 32 * lane < P)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a_chunk + (long long)u * P + 32 * lane));
        }
        int j = 0;
        if (!(starts & 1u)) {
          // the batch begins inside the segment whose chunk is already in `ua`
          const int jn = starts ? (__ffs(starts) - 1) : nb;
          if (starts) load_u(ub, __shfl_sync(0xffffffffu, my_u, jn));
          seg_pairs(ua, my_lo, j, jn);
          j = jn;
          if (starts) {
#pragma unroll
            for (int m = 0; m < RT_CH / 128; ++m) ua[m] = ub[m];
            cur_u = __shfl_sync(0xffffffffu, my_u, jn);
            starts &= starts - 1;
          }
        } else {
          cur_u = __shfl_sync(0xffffffffu, my_u, 0);
          load_u(ua, cur_u);
          starts &= starts - 1;
        }
        while (j < nb) {
          // `ua` holds the chunk of the segment starting at j; fetch the next segment's while comparing this one
          const int jn = starts ? (__ffs(starts) - 1) : nb;
          if (starts) load_u(ub, __shfl_sync(0xffffffffu, my_u, jn));
          seg_pairs(ua, my_lo, j, jn);
          j = jn;
          if (starts) {
#pragma unroll
            for (int m = 0; m < RT_CH / 128; ++m) ua[m] = ub[m];
            cur_u = __shfl_sync(0xffffffffu, my_u, jn);
            starts &= starts - 1;
          }
        }
      }
    }
  }
}

}  // namespace v3s

// rows per tile of the tiled refinement for a candidate capacity (0: capacity too large for it)
extern "C" int v3s_knn_refine_tile_rows(int cap) {
  if (cap < 1 || cap > v3s::RT_KEYS) return 0;
  const int qt = v3s::RT_KEYS / cap;
  return qt < v3s::RT_QMAX ? qt : v3s::RT_QMAX;
}

// The tiled refinement in two halves (a sharded job synchronises its ranks in between):
//   v3s_knn_refine_tiled_pairs : |a_r - a_c|^2 of every pair this rank owns -> dist_ws (FP64 sums, [n_rows, cap]),
//                                pref_ws = slot of the partner row for pairs the partner owns.  cand_all / count_all
//                                (may be NULL): the candidate lists of ALL n_cols patterns, all-gathered -- pairs listed
//                                from both sides are then shared ACROSS ranks too (see refine_prepare_kernel).
//   v3s_knn_refine_tiled_emit  : d = 2 asin(sqrt(ss)/2), sort by (d, index), first k -> S2, Nbr; sums owned by a partner
//                                row are read from peer_dss[o] (host array [world] of the ranks' dist_ws as mapped in
//                                this process; NULL / world <= 1: the local dist_ws).
extern "C" int v3s_knn_refine_tiled_pairs(const float* a_rows, const float* a_all, int P, long long row0, long long n_rows,
                                          long long n_cols, int cap, const int* cand_ws, const int* count_ws,
                                          double* dist_ws, const int* order, const int* tiles, int max_tiles, int* pref_ws,
                                          unsigned long long* keys_ws, int* np_ws, int dedup, const int* cand_all,
                                          const int* count_all, void* stream) {
  V3S_REQUIRE(cap >= 1 && n_rows >= 0 && n_cols < (1LL << 31) && n_rows < (1LL << 31), "v3s_knn_refine_tiled: bad arguments");
  V3S_REQUIRE(P > 0 && P % 4 == 0, "v3s_knn_refine_tiled: P = %d must be a multiple of 4", P);
  V3S_REQUIRE(dist_ws && order && tiles && max_tiles >= 1 && pref_ws && keys_ws && np_ws,
              "v3s_knn_refine_tiled: workspaces, order and the tile table are required");
  V3S_REQUIRE((cand_all == nullptr) == (count_all == nullptr), "v3s_knn_refine_tiled: cand_all and count_all go together");
  const int qt = v3s_knn_refine_tile_rows(cap);
  V3S_REQUIRE(qt >= 1, "v3s_knn_refine_tiled: candidate capacity %d too large", cap);
  if (n_rows == 0) return 0;
  using namespace v3s;
  cudaStream_t st = (cudaStream_t)stream;
  V3S_CUDA(cudaMemsetAsync(dist_ws, 0, sizeof(double) * (size_t)n_rows * cap, st));
  V3S_CUDA(cudaMemsetAsync(pref_ws, 0xFF, sizeof(int) * (size_t)n_rows * cap, st));
  static bool attr_set = false;
  if (!attr_set) {
    V3S_CUDA(cudaFuncSetAttribute(refine_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_KEYS * 8));
    V3S_CUDA(cudaFuncSetAttribute(refine_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_QMAX * RT_CH * 4));
    attr_set = true;
  }
  refine_prepare_kernel<<<max_tiles, RT_THREADS, RT_KEYS * 8, st>>>(order, tiles, row0, cand_ws, count_ws, cap, keys_ws, np_ws,
                                                                   pref_ws, dedup, cand_all, count_all);
  V3S_CHECK_LAUNCH();
  const int grid = max_tiles < 2 * kNumSMs ? max_tiles : 2 * kNumSMs;
  refine_tiled_kernel<<<grid, RT_MAIN_THREADS, (size_t)qt * RT_CH * 4, st>>>(a_rows, a_all, P, order, tiles, keys_ws, np_ws, cap,
                                                                            dist_ws);
  V3S_CHECK_LAUNCH();
  return 0;
}

extern "C" int v3s_knn_refine_tiled_emit(long long row0, long long n_rows, long long n_cols, int k, int cap,
                                         const int* zero_flag, const int* cand_ws, const int* count_ws, double* S2, int* Nbr,
                                         double* dist_ws, const int* pref_ws, const unsigned long long* peer_dss /* host [world] */,
                                         int world, long long cnt_base, int rem, void* stream) {
  V3S_REQUIRE(k >= 1 && cap >= k && n_rows >= 0 && n_rows < (1LL << 31), "v3s_knn_refine_tiled: bad arguments");
  V3S_REQUIRE(k <= n_cols, "Number of nearest neighbors to preserve is too high.");   // distance.py:33-34
  V3S_REQUIRE(dist_ws && pref_ws, "v3s_knn_refine_tiled_emit: workspaces are required");
  if (n_rows == 0) return 0;
  using namespace v3s;
  cudaStream_t st = (cudaStream_t)stream;
  RefPeers peers{};
  if (peer_dss && world > 1) {
    V3S_REQUIRE(world <= 8 && cnt_base >= 1 && rem >= 0 && rem < world && cnt_base * world + rem == n_cols,
                "v3s_knn_refine_tiled_emit: bad sharding");
    for (int i = 0; i < world; ++i) peers.dss[i] = reinterpret_cast<const double*>(peer_dss[i]);
    peers.world = world; peers.rem = rem; peers.cnt_base = cnt_base;
  }
  int kk_pad = 4;
  while (kk_pad < cap) kk_pad <<= 1;
  V3S_REQUIRE(kk_pad <= 4096, "v3s_knn_refine_tiled: candidate capacity %d too large (max 4096)", cap);
  const size_t base = (size_t)kk_pad * (sizeof(double) + sizeof(int));
  if (base > g_refine_smem_set) {
    V3S_CUDA(cudaFuncSetAttribute(refine_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base));
    g_refine_smem_set = base;
  }
  const int grid2 = (int)(n_rows < 8LL * kNumSMs ? n_rows : 8LL * kNumSMs);
  refine_sort_kernel<<<grid2, kRefThreads, base, st>>>(nullptr, nullptr, 0, row0, (int)n_rows, cand_ws, count_ws, cap, kk_pad, k,
                                                      zero_flag, 0, nullptr, 0, S2, Nbr, dist_ws, 3, nullptr, pref_ws, peers);
  V3S_CHECK_LAUNCH();
  return 0;
}

// v3s_knn_refine with the pair evaluation tiled over neighbouring rows (see above): both halves in one call (one
// rank).  order int [n_rows] (rows of a neighbourhood contiguous, from v3s_knn_merge_lists) is required; P must be a
// multiple of 4.  Workspaces: dist_ws FP64 [n_rows,cap], pref_ws int [n_rows,cap], keys_ws 64-bit [max_tiles * 16384],
// np_ws int [max_tiles]; tiles (device, int [1 + 2 max_tiles]): the tile table written by v3s_knn_merge_lists.
// dedup != 0: pairs listed from both sides are evaluated once.
extern "C" int v3s_knn_refine_tiled(const float* a_rows, const float* a_all, int P, long long row0, long long n_rows,
                                    long long n_cols, int k, int cap, const int* zero_flag, const int* cand_ws,
                                    const int* count_ws, double* S2, int* Nbr, double* dist_ws, const int* order,
                                    const int* tiles, int max_tiles, int* pref_ws, unsigned long long* keys_ws, int* np_ws,
                                    int dedup, void* stream) {
  V3S_REQUIRE(k >= 1 && cap >= k, "v3s_knn_refine_tiled: bad arguments");
  V3S_REQUIRE(k <= n_cols, "Number of nearest neighbors to preserve is too high.");   // distance.py:33-34
  if (int rc = v3s_knn_refine_tiled_pairs(a_rows, a_all, P, row0, n_rows, n_cols, cap, cand_ws, count_ws, dist_ws, order, tiles,
                                          max_tiles, pref_ws, keys_ws, np_ws, dedup, nullptr, nullptr, stream)) return rc;
  return v3s_knn_refine_tiled_emit(row0, n_rows, n_cols, k, cap, zero_flag, cand_ws, count_ws, S2, Nbr, dist_ws, pref_ws,
                                   nullptr, 1, n_cols, 0, stream);
}

namespace v3s {
__global__ void neg_key_kernel(const double* __restrict__ B, long long n, float* __restrict__ key) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
    key[e] = -(float)B[e];
}
}  // namespace v3s

// DistanceMatrix.knn_from_round_distance (distance.py:26-48) on a caller-supplied dense FP64
// matrix B [n,n]: per row the k smallest entries, ascending, with their column indices.
// key_ws: float [n,n]; cand_ws: int [n, min(k+margin, n)]; count_ws: int [n].
extern "C" int v3s_knn_from_dense(const double* B, long long n, int k, int margin, float* key_ws, int* cand_ws,
                                  int* count_ws, double* S2, int* Nbr, void* stream) {
  V3S_REQUIRE(k >= 1 && margin >= 0 && n >= 1 && n < (1LL << 31), "v3s_knn_from_dense: bad arguments");
  V3S_REQUIRE(k <= n, "Number of nearest neighbors to preserve is too high.");
  cudaStream_t st = (cudaStream_t)stream;
  long long kk = (long long)k + margin;
  if (kk > n) kk = n;
  int kk_pad = 4;
  while (kk_pad < kk) kk_pad <<= 1;
  V3S_REQUIRE(kk_pad <= 4096, "v3s_knn_from_dense: k + margin too large (max 4096)");
  long long g = ceil_div64(n * n, 256);
  if (g > 16LL * kNumSMs) g = 16LL * kNumSMs;
  neg_key_kernel<<<(int)g, 256, 0, st>>>(B, n * n, key_ws);
  V3S_CHECK_LAUNCH();
  const int grid = (int)(n < 16LL * kNumSMs ? n : 16LL * kNumSMs);
  select_kernel<<<grid, kSelThreads, 0, st>>>(key_ws, n, (int)n, (int)n, 0, (int)kk, 0.f, (int)kk, nullptr, cand_ws,
                                              count_ws, nullptr);
  V3S_CHECK_LAUNCH();
  const size_t smem = (size_t)kk_pad * (sizeof(double) + sizeof(int));
  refine_sort_kernel<<<grid, kRefThreads, smem, st>>>(nullptr, nullptr, 0, 0, (int)n, cand_ws, count_ws, (int)kk, kk_pad, k,
                                                     nullptr, 0, B, n, S2, Nbr, nullptr, 0, nullptr);
  V3S_CHECK_LAUNCH();
  return 0;
}
