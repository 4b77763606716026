// K2b fused -- Gram tiles on tcgen05 with the per-row top-k selection INSIDE the tensor-core
// epilogue: the N x N Gram matrix is never written to memory.
//
// Replaces np.matmul(A, A.T) (distance.py:56) + the per-row argsort (distance.py:39-43) up to the
// candidate lists; the exact distances of the candidates are recomputed from the FP32 rows by
// select.cu (refine_sort_kernel), so the Gram values here only RANK candidates.
//
// Arithmetic: ONE FP16 product per element.  Operands are the centred unit rows scaled by 2^8 and
// rounded to FP16 (11 significant bits), accumulated in FP32 in TMEM.  For unit rows a, b:
//     |fl16(a).fl16(b) - a.b| <= (2^-10 + 2^-22) sum|a_i b_i| <= 2^-10      (Cauchy-Schwarz)
// plus the tensor core's FP32 accumulation error (<= 2^-23 per K=16 step).  With eps_g that bound,
// every entry of the exact top-k of a row is inside {g~ >= (k-th largest g~) - 2 eps_g}: the k-th
// largest g~ is at most t + eps_g (t = exact k-th largest) because fewer than k entries can exceed
// it, and every exact top-k entry has g~ >= t - eps_g.  The epilogue keeps exactly that band.
//
// Kernel shape (persistent, one CTA -- or one CTA PAIR with cta_group::2 -- per SM / SM pair):
//   warp 4 lane 0 : TMA producer (cp.async.bulk.tensor 2-D, SWIZZLE_128B, mbarrier ring)
//   warp 5 lane 0 : tcgen05.mma issuer, kind::f16 (FP16 in, FP32 accumulate), M = 128 * CG, N = 256, K = 16
//   warps 0-3     : epilogue: tcgen05.ld one accumulator row per thread (TMEM lane = query row), compare
//                   against the row's running threshold, append survivors (value, column) to the row's
//                   list; a nearly full list is compacted by its warp (exact k-th largest by a bitwise
//                   descent with warp votes), which also raises the threshold
//   TMEM          : 2 accumulators x 256 columns (the epilogue of tile i overlaps the MMAs of tile i+1)
// Work unit = (row block of 128*CG query rows, segment of the column tiles): a worker sweeps the
// segment's 256-column tiles with the row block's threshold state in registers.  Units are ordered
// so that the workers running at the same time cover few row blocks x few segments (operands
// shared through L2).  Each (row, segment) leaves one list; knn_merge_kernel unites them.
#include "common.cuh"
#include "v3s.h"
#include "tma.cuh"
#include <cuda_fp16.h>

namespace v3s {

constexpr int TK_BM = 128;          // query rows per CTA  (TMEM lanes)
constexpr int TK_BN = 256;          // candidate columns per tile (TMEM columns)
constexpr int TK_BK = 64;           // FP16 elements per K block = one 128-byte swizzle atom
constexpr int TK_UMMA_K = 16;
constexpr int TK_THREADS = 192;
constexpr float TK_OPERAND_SCALE = 256.0f;                 // rows are stored as fl16(2^8 a)
constexpr float TK_ACC_SCALE = 1.0f / 65536.0f;            // g~ = acc * 2^-16 (exact)

template <int CG> struct TopkCfg {
  static constexpr int A_BYTES = TK_BM * TK_BK * 2;                 // 16 kB: this CTA's 128 query rows
  static constexpr int B_BYTES = (TK_BN / CG) * TK_BK * 2;          // 32 kB, or this CTA's half (16 kB) of the pair's tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (CG == 1) ? 4 : 6;                  // 192 kB of operand staging either way
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tk_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tk_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// The four K = 16 MMAs of one 64-element K block in ONE asm statement: the issuing thread has ~128 cycles
// per MMA to keep the tensor pipe full, and every separate asm statement costs a divergent-path
// ELECT / R2UR sequence (registers -> uniform registers) of about that size (measured: one statement per
// MMA left the pipe 40 % idle with the issuing thread busy all the time).  Descriptor start addresses
// advance by 32 B (2 x 16-byte units) per K step inside the 128-byte swizzle atom.
template <int CG>
__device__ __forceinline__ void tk_umma_kblock(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 b1, %2, 2;\n\t"
        "add.u64 a2, %1, 4;\n\tadd.u64 b2, %2, 4;\n\t"
        "add.u64 a3, %1, 6;\n\tadd.u64 b3, %2, 6;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, q;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, q;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b64 a1, a2, a3, b1, b2, b3;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.eq.b32 q, 0, 0;\n\t"
        "add.u64 a1, %1, 2;\n\tadd.u64 b1, %2, 2;\n\t"
        "add.u64 a2, %1, 4;\n\tadd.u64 b2, %2, 4;\n\t"
        "add.u64 a3, %1, 6;\n\tadd.u64 b3, %2, 6;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, q;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, q;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, q;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  }
}
// completion of all MMAs issued so far -> mbarrier (CG = 2: the barrier at the same offset in BOTH CTAs)
template <int CG>
__device__ __forceinline__ void tk_commit(uint64_t* bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  } else {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
  }
}
// TMA load; CG = 2: the copy lands in THIS CTA's shared memory, its bytes are counted on the LEADER's barrier
template <int CG>
__device__ __forceinline__ void tk_tma_load(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  if constexpr (CG == 1) {
    tma_load_2d(dst, map, c0, c1, bar);
  } else {
    const uint32_t leader_bar = smem_u32(bar) & 0xFEFFFFFFu;       // clear the peer bit of the cluster-window address
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(leader_bar)
        : "memory");
  }
}
// arrive (no transaction bytes) on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// K-major SWIZZLE_128B shared-memory matrix descriptor: rows of 64 FP16 (128 B), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t tk_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void tk_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, N = 256, M = 128 * CG
template <int CG>
constexpr uint32_t tk_idesc() {
  return (1u << 4) | ((uint32_t)(TK_BN >> 3) << 17) | ((uint32_t)((TK_BM * CG) >> 4) << 24);
}

// ---- unit schedule (shared by host planning and the kernel) ---------------------------------
struct TopkPlan {
  int n_rb;        // row blocks of 128*CG query rows
  int S;           // column segments per row block
  int tps;         // 256-column tiles per segment
  int tiles_c;     // column tiles in all
  int rg;          // row blocks per group: units of a group are ordered segment-major, so the workers
                   // running together cover ~rg row blocks x S segments
  int units;
  int sync_lag;    // a worker starts tile step s once every worker has ARRIVED at step s - sync_lag (0: strict lockstep)
};
__host__ __device__ inline void topk_decode_unit(const TopkPlan& p, int u, int& rb, int& seg) {
  const int per_group = p.rg * p.S;
  const int g = u / per_group;
  const int first = g * p.rg;
  const int rows_here = (p.n_rb - first) < p.rg ? (p.n_rb - first) : p.rg;
  const int ul = u - g * per_group;
  seg = ul / rows_here;
  rb = first + ul % rows_here;
}

// Warp-cooperative compaction of one row's list L[0, n): find the k-th largest value (bitwise descent
// on the order-preserving key, counts by warp vote), keep every entry within `band` of it, at most
// `keep_cap` (the largest ones; ties at the cut in list order).  Returns the new count; *thr_out =
// the new lower limit for appends.  All 32 lanes call it with the same arguments.
template <int EPL>   // entries per lane: capacity C = 32 * EPL
__device__ __forceinline__ int warp_compact_list(unsigned long long* __restrict__ L, int n, int k, float band, int keep_cap,
                                                 float* thr_out, int* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned long long e[EPL];
  uint32_t key[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    const int idx = i * 32 + lane;
    e[i] = (idx < n) ? L[idx] : 0ull;
    key[i] = (idx < n) ? float_to_key(__uint_as_float((uint32_t)(e[i] >> 32))) : 0u;   // 0 sorts below every real key
  }
  __syncwarp();
  if (n <= k) {                       // fewer than k seen: nothing can be dropped yet
    *thr_out = -3.0e38f;
    return n;
  }
  uint32_t T = 0;
#pragma unroll 1
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int i = 0; i < EPL; ++i) c += (key[i] >= cand) ? 1 : 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (c >= k) T = cand;
  }
  float lim = key_to_float(T) - band;
  int total = 0;
#pragma unroll
  for (int i = 0; i < EPL; ++i) total += (key[i] != 0u && key_to_float(key[i]) >= lim) ? 1 : 0;
  total = __reduce_add_sync(0xffffffffu, total);
  uint32_t Tcut = 0;                  // keep key > Tcut, plus `need_eq` entries with key == Tcut (overflow only)
  int need_eq = 0;
  const bool over = total > keep_cap;
  if (over) {
    if (overflow && lane == 0) atomicAdd(overflow, 1);
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = Tcut | (1u << bit);
      int c = 0;
#pragma unroll
      for (int i = 0; i < EPL; ++i) c += (key[i] >= cand) ? 1 : 0;
      c = __reduce_add_sync(0xffffffffu, c);
      if (c >= keep_cap) Tcut = cand;
    }
    int gt = 0;
#pragma unroll
    for (int i = 0; i < EPL; ++i) gt += (key[i] > Tcut) ? 1 : 0;
    gt = __reduce_add_sync(0xffffffffu, gt);
    need_eq = keep_cap - gt;
    lim = key_to_float(Tcut);         // no band left: the list is saturated (reported through `overflow`)
  }
  int base = 0, eq_taken = 0;
#pragma unroll
  for (int i = 0; i < EPL; ++i) {
    bool keep;
    if (!over) {
      keep = (key[i] != 0u) && key_to_float(key[i]) >= lim;
    } else {
      const bool eq = key[i] == Tcut && key[i] != 0u;
      const unsigned eb = __ballot_sync(0xffffffffu, eq);
      keep = (key[i] > Tcut) || (eq && (eq_taken + __popc(eb & lt_mask)) < need_eq);
      eq_taken += __popc(eb);
    }
    const unsigned kb = __ballot_sync(0xffffffffu, keep);
    if (keep) L[base + __popc(kb & lt_mask)] = e[i];
    base += __popc(kb);
  }
  __syncwarp();
  *thr_out = lim;
  return base;
}

template <int CG, int EPL>
__global__ void __launch_bounds__(TK_THREADS, 1)
gram_topk_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_all, int n_rows,
                 int n_cols, int num_kb, const TopkPlan plan, int k, float band, unsigned long long* __restrict__ lists,
                 int* __restrict__ counts, int* __restrict__ overflow, unsigned int* __restrict__ tau_keys,
                 int* __restrict__ step_sync) {
  using Cfg = TopkCfg<CG>;
  constexpr int STAGES = Cfg::STAGES, A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int C = 32 * EPL;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;                      // [STAGES]  TMA -> MMA   (CG = 2: only the leader's are used)
  uint64_t* empty = bars + STAGES;            // [STAGES]  MMA -> TMA   (each CTA's own)
  uint64_t* tfull = bars + 2 * STAGES;        // [2]       MMA -> epilogue (each CTA's own)
  uint64_t* tempty = bars + 2 * STAGES + 2;   // [2]       epilogue -> MMA (CG = 2: the leader's, 4 arrivals per CTA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int worker = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_workers = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_rows) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_all) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], CG); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CG == 2) cluster_sync_all();            // both CTAs are resident before the pair allocates tensor memory
  if (warp == 5) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tk_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tk_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    {
      // ---------------- TMA producer (every CTA: its own query rows, its share of the column tile) ----------------
      // The whole warp runs the loop (uniform control flow), one elected lane issues the copies.
      // All workers start every tile together (one counter per (wave, tile) step, arrive + spin): workers that
      // share a row block or a column tile then stream the same K slices at the same time, so each slice is read
      // from HBM once and served to the others by L2.  Without it the workers drift apart by more than the time
      // a slice survives in L2 and every worker streams both operands from HBM (measured: 87x the operand bytes).
      int stage = 0;
      uint32_t phase = 0;
      const int target = num_workers * CG;
      int wave = 0;
      for (int u0 = 0; u0 < plan.units; u0 += num_workers, ++wave) {
        const int u = u0 + worker;
        const bool has_unit = u < plan.units;
        int rb = 0, seg = 0;
        if (has_unit) topk_decode_unit(plan, u, rb, seg);
        const int row0 = rb * (TK_BM * CG) + (int)cta_rank * TK_BM;
        const int t0 = seg * plan.tps, t1 = has_unit ? min(t0 + plan.tps, plan.tiles_c) : t0;
        for (int j = 0; j < plan.tps; ++j) {
          int* ctr = step_sync + wave * plan.tps + j;
          if (lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr) : "memory");
          const int t = t0 + j;
          if (t >= t1) continue;                              // nothing to do at this step: arrive only
          // (sync_lag > 0: wait for the arrivals at an EARLIER step -- the streams stay within sync_lag tiles of each
          // other, which L2 still bridges, while one worker's slow epilogue no longer holds up every other worker)
          const int widx = wave * plan.tps + j - plan.sync_lag;
          const int* wctr = step_sync + (widx > 0 ? widx : 0);
          int seen = widx < 0 ? target : 0;
          long long t_start = 0;
          if (widx >= 0) do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(wctr) : "memory");
            if (seen < target) {
              __nanosleep(64);
              // the barrier only aligns the workers' memory streams: after ~1 s (a worker that is not resident
              // yet, e.g. behind another stream's kernel) carry on alone rather than hang
              const long long now = clock64();
              if (t_start == 0) t_start = now;
              else if (now - t_start > 2000000000LL) break;
            }
          } while (seen < target);
          __syncwarp();
          const int col0 = t * TK_BN + (int)cta_rank * (TK_BN / CG);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (elect_one()) {
              const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
              if constexpr (CG == 1) {
                mbar_expect_tx(&full[stage], STAGE_BYTES);
              } else {
                if (leader) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);   // both CTAs' bytes land on the leader's barrier
                else mbar_arrive_remote(&full[stage], 0);
              }
              tk_tma_load<CG>(s, &tm_rows, kb * TK_BK, row0, &full[stage]);
              tk_tma_load<CG>(s + A_BYTES, &tm_all, kb * TK_BK, col0, &full[stage]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 5) {
    if (leader) {
      // ---------------- MMA issuer (CG = 2: the leader issues for the pair) ----------------
      // The whole warp runs the loop (uniform control flow: descriptors and barrier addresses stay in uniform
      // registers); one elected lane issues the tensor-core instructions.
      constexpr uint32_t idesc = tk_idesc<CG>();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = worker; u < plan.units; u += num_workers) {
        int rb, seg;
        topk_decode_unit(plan, u, rb, seg);
        const int t0 = seg * plan.tps, t1 = min(t0 + plan.tps, plan.tiles_c);
        for (int t = t0; t < t1; ++t) {
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          tk_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TK_BN);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full[stage], phase);
            tk_fence_after();
            const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t a_desc = tk_kmajor_desc(s);
            const uint64_t b_desc = tk_kmajor_desc(s + A_BYTES);
            static_assert(TK_BK / TK_UMMA_K == 4, "tk_umma_kblock issues four K steps");
            if (elect_one()) {
              tk_umma_kblock<CG>(d_tmem, a_desc, b_desc, idesc, kb > 0 ? 1u : 0u);
              tk_commit<CG>(&empty[stage]);                 // frees the stage (in both CTAs) once these MMAs retire
              if (kb == num_kb - 1) tk_commit<CG>(&tfull[acc]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      }
    }
  } else {
    // ---------------- epilogue: warps 0..3 <-> TMEM lane quadrants; one query row per thread ----------------
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = worker; u < plan.units; u += num_workers) {
      int rb, seg;
      topk_decode_unit(plan, u, rb, seg);
      const int row = rb * (TK_BM * CG) + (int)cta_rank * TK_BM + warp * 32 + lane;     // local query row (may be padding)
      const bool row_ok = row < n_rows;
      unsigned long long* L = lists + ((long long)row * plan.S + seg) * C;
      int cnt = 0;
      // lower limit of the row's band so far, shared by the row's segments through tau_keys (a limit found in
      // any subset of the columns is valid for all of them: the k-th largest of a subset is <= that of the whole)
      const uint32_t key0 = tau_keys[row];
      float thr = key0 ? key_to_float(key0) : -3.0e38f;
      const int t0 = seg * plan.tps, t1 = min(t0 + plan.tps, plan.tiles_c);
      for (int t = t0; t < t1; ++t) {
        mbar_wait(&tfull[acc], acc_phase);
        tk_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * TK_BN);
#pragma unroll 1
        for (int ch = 0; ch < TK_BN / 32; ++ch) {
          uint32_t v[32];
          tk_tmem_ld32(t_row + (uint32_t)(ch * 32), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int col0 = t * TK_BN + ch * 32;
          const int lim = row_ok ? (n_cols - col0) : 0;       // columns >= n_cols are TMA zero fill, not patterns
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float g = __uint_as_float(v[j]) * TK_ACC_SCALE;
            if (g >= thr && j < lim) {
              L[cnt] = ((unsigned long long)__float_as_uint(g) << 32) | (unsigned)(col0 + j);
              ++cnt;
            }
          }
          unsigned need = __ballot_sync(0xffffffffu, cnt > C - 32);
          while (need) {
            const int l = __ffs(need) - 1;
            need &= need - 1;
            unsigned long long* Ll = reinterpret_cast<unsigned long long*>(__shfl_sync(0xffffffffu, (unsigned long long)L, l));
            const int nl = __shfl_sync(0xffffffffu, cnt, l);
            __syncwarp();
            float new_thr;
            const int kept = warp_compact_list<EPL>(Ll, nl, k, band, C - 64, &new_thr, overflow);
            if (lane == l) {
              cnt = kept;
              if (new_thr > thr) { thr = new_thr; atomicMax(&tau_keys[row], float_to_key(thr)); }
            }
          }
        }
        // this accumulator is drained: hand it back to the MMA issuer (CG = 2: the leader's barrier)
        tk_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) mbar_arrive(&tempty[acc]);
          else mbar_arrive_remote(&tempty[acc], 0);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (row_ok) counts[(long long)row * plan.S + seg] = cnt;
    }
  }
  tk_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 5) {
    tk_fence_after();
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- symmetric form: half the tensor-core work -----------------------------------------------------------------
// G is symmetric, so a tile on or above the diagonal serves both its rows AND (mirrored) its columns.  With one
// list per pattern, appended to from any tile, the running threshold of the rectangular kernel has no single owner;
// instead every pattern's lower limit is FIXED beforehand: pass 1 runs the rectangular kernel above against every
// `sample_stride`-th pattern only (1/8 of the work) and the k-th largest g~ of that sample -- a valid lower bound
// of the k-th largest over all patterns -- minus the band becomes lim[i].  Pass 2 (this kernel) computes the tiles
// with col-tile >= row-block, and for every element (r, c), c >= r:
//     g~ >= lim[r]  ->  list r gets (g~, c);      c > r and g~ >= lim[c]  ->  list c gets (g~, r)
// (atomic slot counters; ~sample_stride * k entries per list).  Sharded jobs split the tile list over the ranks and
// append into the owner rank's lists through NVLink peer mappings -- compute and exchange in one kernel.
struct SymPeers {
  unsigned long long* lists[8];      // rank o's lists [cnt_o, C]
  int* counts[8];                    // rank o's slot counters [cnt_o]
  int world, rem;
  long long cnt_base;                // rank o owns cnt_base + (o < rem) consecutive patterns
};
__device__ __forceinline__ void sym_owner(const SymPeers& pr, long long x, int& o, long long& lo) {
  const long long big = pr.cnt_base + 1, split = big * pr.rem;
  if (x < split) { o = (int)(x / big); lo = o * big; }
  else { o = pr.rem + (int)((x - split) / pr.cnt_base); lo = split + (long long)(o - pr.rem) * pr.cnt_base; }
}
struct SymPlan {
  int n_rb, n_cb;      // row blocks of 128*CG patterns, column tiles of 256 patterns
  int PR, PC;          // a step = a patch of PR x PC tiles shared by all workers of all ranks
  int steps;           // patches that touch the upper triangle
  int rank, world, workers;
  int sync_lag;        // as TopkPlan::sync_lag
};
// patch `step` (in the order both sides enumerate) -> (pr, pc); returns false past the end
__host__ __device__ inline int sym_count_steps(const SymPlan& p, int rows_per_block) {
  int n = 0;
  const int gr = (p.n_rb + p.PR - 1) / p.PR, gc = (p.n_cb + p.PC - 1) / p.PC;
  for (int pr = 0; pr < gr; ++pr)
    for (int pc = 0; pc < gc; ++pc)
      if ((long long)(pc * p.PC + p.PC) * TK_BN - 1 >= (long long)pr * p.PR * rows_per_block) ++n;
  return n;
}

template <int CG>
__global__ void __launch_bounds__(TK_THREADS, 1)
gram_sym_topk_kernel(const __grid_constant__ CUtensorMap tm_rows, const __grid_constant__ CUtensorMap tm_cols, int n,
                     int num_kb, const SymPlan plan, const float* __restrict__ lim, int C, const SymPeers peers,
                     int* __restrict__ step_sync) {
  using Cfg = TopkCfg<CG>;
  constexpr int STAGES = Cfg::STAGES, A_BYTES = Cfg::A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int RB = TK_BM * CG;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  __shared__ float s_clim[2][TK_BN];          // lower limits of the tile's columns, per accumulator stage

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int worker = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int slot = plan.rank * plan.workers + worker;           // this worker's tile of every patch
  const int dr = slot / plan.PC, dc = slot % plan.PC;
  const int gr = (plan.n_rb + plan.PR - 1) / plan.PR, gc = (plan.n_cb + plan.PC - 1) / plan.PC;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_rows) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_cols) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], CG); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4 * CG); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (CG == 2) cluster_sync_all();
  if (warp == 5) {
    if constexpr (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  tk_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tk_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // every role walks the same patch sequence; `valid` = this worker has a tile (on or above the diagonal) in it
#define SYM_FOR_EACH_STEP(BODY)                                                                                      \
  {                                                                                                                  \
    int step = 0;                                                                                                    \
    for (int pr = 0; pr < gr; ++pr)                                                                                  \
      for (int pc = 0; pc < gc; ++pc) {                                                                              \
        if ((long long)(pc * plan.PC + plan.PC) * TK_BN - 1 < (long long)pr * plan.PR * RB) continue;                \
        const int rb = pr * plan.PR + dr, cb = pc * plan.PC + dc;                                                    \
        const bool valid = dr < plan.PR && rb < plan.n_rb && cb < plan.n_cb &&                                       \
                           (long long)cb * TK_BN + TK_BN - 1 >= (long long)rb * RB;                                  \
        BODY                                                                                                         \
        ++step;                                                                                                      \
      }                                                                                                              \
  }

  if (warp == 4) {
    int stage = 0;
    uint32_t phase = 0;
    const int target = (int)(gridDim.x);                       // every CTA of this rank arrives at every step
    SYM_FOR_EACH_STEP({
      int* ctr = step_sync + step;
      if (lane == 0) asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(ctr) : "memory");
      if (valid) {
        const int widx = step - plan.sync_lag;
        const int* wctr = step_sync + (widx > 0 ? widx : 0);
        int seen = widx < 0 ? target : 0;
        long long t_start = 0;
        if (widx >= 0) do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(wctr) : "memory");
          if (seen < target) {
            __nanosleep(64);
            const long long now = clock64();
            if (t_start == 0) t_start = now;
            else if (now - t_start > 2000000000LL) break;
          }
        } while (seen < target);
        __syncwarp();
        const int row0 = rb * RB + (int)cta_rank * TK_BM;
        const int col0 = cb * TK_BN + (int)cta_rank * (TK_BN / CG);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (elect_one()) {
            const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
            if constexpr (CG == 1) {
              mbar_expect_tx(&full[stage], STAGE_BYTES);
            } else {
              if (leader) mbar_expect_tx(&full[stage], 2 * STAGE_BYTES);
              else mbar_arrive_remote(&full[stage], 0);
            }
            tk_tma_load<CG>(s, &tm_rows, kb * TK_BK, row0, &full[stage]);
            tk_tma_load<CG>(s + A_BYTES, &tm_cols, kb * TK_BK, col0, &full[stage]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    })
  } else if (warp == 5) {
    if (leader) {
      constexpr uint32_t idesc = tk_idesc<CG>();
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      SYM_FOR_EACH_STEP({
        if (valid) {
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          tk_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * TK_BN);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full[stage], phase);
            tk_fence_after();
            const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
            const uint64_t a_desc = tk_kmajor_desc(s);
            const uint64_t b_desc = tk_kmajor_desc(s + A_BYTES);
            if (elect_one()) {
              tk_umma_kblock<CG>(d_tmem, a_desc, b_desc, idesc, kb > 0 ? 1u : 0u);
              tk_commit<CG>(&empty[stage]);
              if (kb == num_kb - 1) tk_commit<CG>(&tfull[acc]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        }
      })
    }
  } else {
    // ---------------- epilogue: one pattern row per thread, both directions of every element on / above the diagonal ----
    int acc = 0;
    uint32_t acc_phase = 0;
    SYM_FOR_EACH_STEP({
      if (valid) {
        const int row = rb * RB + (int)cta_rank * TK_BM + warp * 32 + lane;        // global pattern index
        const bool row_ok = row < n;
        const float rlim = row_ok ? lim[row] : 3.0e38f;
        int ro = 0;
        long long rlo = 0;
        if (row_ok) sym_owner(peers, row, ro, rlo);
        unsigned long long* Lr = peers.lists[ro] + (long long)(row - rlo) * C;
        int* Cr = peers.counts[ro] + (row - rlo);
        // the column limits of this tile -> shared memory (stage `acc`); all four epilogue warps, then a named barrier
        for (int j = threadIdx.x; j < TK_BN; j += 128) {
          const int c = cb * TK_BN + j;
          s_clim[acc][j] = c < n ? lim[c] : 3.0e38f;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        mbar_wait(&tfull[acc], acc_phase);
        tk_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * TK_BN);
        // A slot counter's atomicAdd takes an L2 round trip.  The entry of a hit is therefore stored only when the
        // thread's NEXT hit (or the end of the tile) comes: the round trip overlaps the compares in between
        // (stored right away, ~6 hits per thread and tile kept the epilogue as long as the tile's MMAs).
        int rpos = -1;
        int cpos = -1;
        unsigned long long rent = 0ull;
        unsigned long long cent = 0ull;
        unsigned long long* cL = nullptr;
#pragma unroll 1
        for (int ch = 0; ch < TK_BN / 32; ++ch) {
          uint32_t v[32];
          tk_tmem_ld32(t_row + (uint32_t)(ch * 32), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const int col0 = cb * TK_BN + ch * 32;
          if (col0 + 31 < row || !row_ok) continue;                   // this chunk lies entirely below the diagonal (per thread)
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = col0 + j;
            const float g = __uint_as_float(v[j]) * TK_ACC_SCALE;
            if (c >= row && c < n) {
              if (g >= rlim) {
                if (rpos >= 0 && rpos < C) Lr[rpos] = rent;
                rpos = atomicAdd(Cr, 1);
                rent = ((unsigned long long)__float_as_uint(g) << 32) | (unsigned)c;
              }
              if (c > row && g >= s_clim[acc][ch * 32 + j]) {
                if (cpos >= 0 && cpos < C) cL[cpos] = cent;
                int co;
                long long clo;
                sym_owner(peers, c, co, clo);
                cpos = atomicAdd(peers.counts[co] + (c - clo), 1);
                cL = peers.lists[co] + (long long)(c - clo) * C;
                cent = ((unsigned long long)__float_as_uint(g) << 32) | (unsigned)row;
              }
            }
          }
        }
        if (rpos >= 0 && rpos < C) Lr[rpos] = rent;
        if (cpos >= 0 && cpos < C) cL[cpos] = cent;
        tk_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 1) mbar_arrive(&tempty[acc]);
          else mbar_arrive_remote(&tempty[acc], 0);
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    })
  }
#undef SYM_FOR_EACH_STEP
  tk_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 5) {
    tk_fence_after();
    if constexpr (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// FP16 Gram operand of centred unit rows: fl16(2^8 a), zero padded to Ppad columns
__global__ void split_f16_kernel(const float* __restrict__ a, long long n_rows, int P, int Ppad, __half* __restrict__ out) {
  const long long total = n_rows * (long long)(Ppad / 2);
  const int half_pad = Ppad / 2;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / half_pad;
    const int c = (int)(e - r * half_pad) * 2;
    const float v0 = c < P ? a[r * P + c] * TK_OPERAND_SCALE : 0.f;
    const float v1 = c + 1 < P ? a[r * P + c + 1] * TK_OPERAND_SCALE : 0.f;
    reinterpret_cast<__half2*>(out)[e] = __floats2half2_rn(v0, v1);
  }
}

// One digit of the radix select, by warp 0: the largest digit d whose suffix count sum(hist[d .. 255]) reaches
// `remaining` (d = 0 if none does), and the count above it.  Lane l owns bins 8l .. 8l+7; the suffix sums over lanes
// come from five shuffles, only the lane that contains the crossing walks its eight bins.
__device__ __forceinline__ void radix_pick_digit(const int* __restrict__ hist, int remaining, int* d_out, int* above_out) {
  const int lane = threadIdx.x & 31;
  int t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t += hist[8 * lane + i];
  int sfx = t;                                            // inclusive suffix sum over lanes >= lane
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_down_sync(0xffffffffu, sfx, o);
    if (lane + o < 32) sfx += v;
  }
  const unsigned reach = __ballot_sync(0xffffffffu, sfx >= remaining);
  const int owner = reach ? 31 - __clz(reach) : 0;
  if (lane == owner) {
    int cum = sfx - t, d = 8 * lane + 7;
    for (; d > 8 * lane; --d) {
      if (cum + hist[d] >= remaining) break;
      cum += hist[d];
    }
    // (no break: d = 8 lane, which reaches `remaining` by the choice of the owner -- or, without any crossing,
    // owner = 0 and d = 0 with everything above it counted, like a serial scan from 255 down to 1)
    *d_out = d;
    *above_out = cum;
  }
}

// Unite the S per-segment lists of every row: k-th largest g~ over the union (4-pass radix select in
// shared memory), candidates = every entry within `band` of it (at most cap, the largest first on
// overflow).  Zero-norm patterns (0/0 -> NaN -> distance 0 in the reference, distance.py:58) have
// g~ = 0 here: a zero QUERY row gets the first columns (all its distances are 0, ties by index), and
// zero COLUMNS (listed in zero_cols) are appended to every row's candidates.
constexpr int kMergeThreads = 128;
// Segment table (n > 0): segment s of row r is lists[s] + r * C with its count at counts[s][r] -- the per-source-rank
// lists of the sharded symmetric Gram, read where they were written (the peers' memory, over NVLink).  n = 0: the
// segments of a row are contiguous in `lists` / `counts` (one GPU, rectangular kernel).
struct MergeSegs {
  const unsigned long long* lists[8];
  const int* counts[8];
  int n;
};
__global__ void __launch_bounds__(kMergeThreads)
knn_merge_kernel(const unsigned long long* __restrict__ lists, const int* __restrict__ counts, const MergeSegs segs, int S,
                 int C, int ent_cap, int n_rows,
                 int n_cols, long long row0, int k, float band, int cap, const int* __restrict__ zero_flag,
                 const int* __restrict__ zero_ws, int* __restrict__ cand, int* __restrict__ cand_count,
                 int* __restrict__ overflow, int anchor_stride, int* __restrict__ cluster_out) {
  // zero_ws (may be NULL): [0] = number of zero-norm patterns, [1 ...] = their indices (zero_cols_kernel)
  const int n_zero = zero_ws ? min(zero_ws[0], cap) : 0;
  const int* zero_cols = zero_ws ? zero_ws + 1 : nullptr;
  extern __shared__ unsigned long long s_ent[];             // S*C entries
  __shared__ int hist[256];
  __shared__ int s_n, s_slot, s_eq;
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining;
  __shared__ unsigned long long s_anchor;
  const int tid = threadIdx.x;
  // cluster_out (optional): the row's nearest ANCHOR (a pattern whose index is a multiple of anchor_stride) among
  // its listed columns -> rows that share an anchor are neighbours, so their candidate lists overlap; rows
  // without a listed anchor go to per-position fallback bins behind the anchor bins (refine_tiled_kernel)
  const int anchor_bins = cluster_out ? (n_cols + anchor_stride - 1) / anchor_stride : 0;
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    __syncthreads();
    int* out = cand + (long long)r * cap;
    if (zero_flag && zero_flag[row0 + r]) {
      const int m = min(cap, n_cols);
      for (int s = tid; s < m; s += kMergeThreads) out[s] = s;
      if (tid == 0) {
        cand_count[r] = m;
        if (cluster_out) cluster_out[r] = anchor_bins + r / anchor_stride;
      }
      continue;
    }
    if (tid == 0) s_anchor = 0ull;
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int sgm = 0; sgm < S; ++sgm) {
      int c = segs.n ? segs.counts[sgm][r] : counts[(long long)r * S + sgm];
      if (c > C) {                                           // symmetric form: a slot counter ran past its list
        if (tid == 0 && overflow) atomicAdd(overflow, 1);
        c = C;
      }
      const unsigned long long* L = segs.n ? segs.lists[sgm] + (long long)r * C : lists + ((long long)r * S + sgm) * C;
      __shared__ int s_base;
      if (tid == 0) {
        s_base = s_n;
        if (c > ent_cap - s_n && overflow) atomicAdd(overflow, 1);   // (segment table: the union is held in ent_cap entries)
        s_n += min(c, ent_cap - s_n);
      }
      __syncthreads();
      c = min(c, ent_cap - s_base);
      for (int i = tid; i < c; i += kMergeThreads) s_ent[s_base + i] = L[i];
      __syncthreads();
    }
    const int n = s_n;
    const int kk = min(k, n);
    uint32_t prefix = 0, mask = 0;
    int remaining = kk;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += kMergeThreads) hist[i] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += kMergeThreads) {
        const uint32_t key = float_to_key(__uint_as_float((uint32_t)(s_ent[i] >> 32)));
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
      }
      __syncthreads();
      if (tid < 32) {
        __shared__ int s_d, s_above;
        radix_pick_digit(hist, remaining, &s_d, &s_above);
        __syncwarp();
        if (tid == 0) {
          s_prefix = prefix | ((uint32_t)s_d << shift);
          s_remaining = remaining - s_above;
        }
      }
      __syncthreads();
      prefix = s_prefix;
      remaining = s_remaining;
      mask |= 255u << shift;
      __syncthreads();
    }
    const float tau = key_to_float(prefix);
    const float lim = tau - band;
    if (tid == 0) { s_slot = 0; s_eq = 0; }
    __syncthreads();
    for (int i = tid; i < n; i += kMergeThreads) {
      const float g = __uint_as_float((uint32_t)(s_ent[i] >> 32));
      const int col = (int)(uint32_t)s_ent[i];
      if (cluster_out && col % anchor_stride == 0)
        atomicMax(&s_anchor, ((unsigned long long)float_to_key(g) << 32) | (unsigned)col);
      if (g >= lim && !(zero_flag && zero_flag[col])) {
        const int slot = atomicAdd(&s_slot, 1);
        if (slot < cap) out[slot] = col;
      }
    }
    __syncthreads();
    int total = s_slot;
    if (total > cap) {
      // band larger than the buffer: keep the cap largest instead (approximate ranking; counted in `overflow`)
      if (tid == 0 && overflow) atomicAdd(overflow, 1);
      __syncthreads();
      uint32_t p2 = 0, m2 = 0;
      int rem2 = cap;
      for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += kMergeThreads) hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kMergeThreads) {
          const uint32_t key = float_to_key(__uint_as_float((uint32_t)(s_ent[i] >> 32)));
          if ((key & m2) == p2) atomicAdd(&hist[(key >> shift) & 255u], 1);
        }
        __syncthreads();
        if (tid < 32) {
          __shared__ int s_d2, s_above2;
          radix_pick_digit(hist, rem2, &s_d2, &s_above2);
          __syncwarp();
          if (tid == 0) {
            s_prefix = p2 | ((uint32_t)s_d2 << shift);
            s_remaining = rem2 - s_above2;
          }
        }
        __syncthreads();
        p2 = s_prefix;
        rem2 = s_remaining;
        m2 |= 255u << shift;
        __syncthreads();
      }
      if (tid == 0) { s_slot = 0; s_eq = 0; }
      __syncthreads();
      for (int i = tid; i < n; i += kMergeThreads) {
        const uint32_t key = float_to_key(__uint_as_float((uint32_t)(s_ent[i] >> 32)));
        bool keep = key > p2;
        if (key == p2) keep = atomicAdd(&s_eq, 1) < rem2;
        if (keep) {
          const int slot = atomicAdd(&s_slot, 1);
          if (slot < cap) out[slot] = (int)(uint32_t)s_ent[i];
        }
      }
      __syncthreads();
      total = min(s_slot, cap);
    }
    // zero-norm patterns are at distance 0 from every row
    if (n_zero > 0) {
      for (int z = tid; z < n_zero; z += kMergeThreads) {
        const int slot = total + z;
        if (slot < cap) out[slot] = zero_cols[z];
      }
      if (total + n_zero > cap && tid == 0 && overflow) atomicAdd(overflow, 1);
      total = min(total + n_zero, cap);
    }
    if (cluster_out) {
      // the tiled refinement looks rows up in each other's candidate lists (pairs listed from both sides are
      // evaluated once): sort the list by column index so that the look-up is a binary search
      __syncthreads();
      int* sc = reinterpret_cast<int*>(s_ent);
      int npad = 2;
      while (npad < total) npad <<= 1;
      for (int i = tid; i < npad; i += kMergeThreads) sc[i] = i < total ? out[i] : 0x7fffffff;
      __syncthreads();
      for (int size = 2; size <= npad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int t = tid; t < npad / 2; t += kMergeThreads) {
            const int lo = 2 * t - (t & (stride - 1));
            const int hi = lo + stride;
            const bool up = ((lo & size) == 0);
            const int a = sc[lo], b = sc[hi];
            if ((a > b) == up) { sc[lo] = b; sc[hi] = a; }
          }
          __syncthreads();
        }
      }
      for (int i = tid; i < total; i += kMergeThreads) out[i] = sc[i];
    }
    if (tid == 0) {
      cand_count[r] = total;
      if (cluster_out) cluster_out[r] = s_anchor ? (int)((uint32_t)s_anchor / (uint32_t)anchor_stride) : anchor_bins + r / anchor_stride;
    }
  }
}

// counting sort of the rows by cluster id -> order (rows of one cluster contiguous; order inside a cluster arbitrary)
__global__ void cluster_hist_kernel(const int* __restrict__ cluster, int n, int* __restrict__ bins) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) atomicAdd(&bins[cluster[r]], 1);
}
__global__ void __launch_bounds__(1024)
cluster_scan_kernel(int* __restrict__ bins, int n_bins) {      // exclusive scan in place, one block
  __shared__ int s_part[1024];
  __shared__ int s_carry;
  const int tid = threadIdx.x;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < n_bins; b0 += 1024) {
    const int i = b0 + tid;
    const int v = i < n_bins ? bins[i] : 0;
    s_part[tid] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = tid >= o ? s_part[tid - o] : 0;
      __syncthreads();
      s_part[tid] += t;
      __syncthreads();
    }
    if (i < n_bins) bins[i] = s_carry + s_part[tid] - v;
    __syncthreads();
    if (tid == 1023) s_carry += s_part[1023];
    __syncthreads();
  }
}
__global__ void cluster_scatter_kernel(const int* __restrict__ cluster, int n, int* __restrict__ cursor, int* __restrict__ order) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x)
    order[atomicAdd(&cursor[cluster[r]], 1)] = r;
}
// Tiles of at most qt consecutive rows of `order` that do not straddle clusters needlessly: a cluster larger than
// qt is cut into pieces, smaller consecutive ones are packed while they fit.  ends[c] = end of cluster c in
// `order` (the scatter's cursors).  tiles[0] = number of tiles, tiles[1 + 2 t] = first position, tiles[2 + 2 t] = rows.
// (A serial pass over a few thousand clusters: one thread.)
__global__ void cluster_tiles_kernel(const int* __restrict__ ends, int n_bins, int qt, int max_tiles, int* __restrict__ tiles) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int nt = 0, start = 0, cur0 = 0, cur = 0;
  for (int c = 0; c < n_bins; ++c) {
    const int end = ends[c];
    int sz = end - start;
    if (sz > 0 && cur + sz > qt && cur > 0) {            // flush what is packed so far
      if (nt < max_tiles) { tiles[1 + 2 * nt] = cur0; tiles[2 + 2 * nt] = cur; }
      ++nt;
      cur = 0;
    }
    if (cur == 0) cur0 = start;
    while (sz > qt) {                                    // a big cluster: whole tiles of its own
      if (nt < max_tiles) { tiles[1 + 2 * nt] = cur0; tiles[2 + 2 * nt] = qt; }
      ++nt;
      cur0 += qt;
      sz -= qt;
    }
    cur += sz;
    start = end;
  }
  if (cur > 0) {
    if (nt < max_tiles) { tiles[1 + 2 * nt] = cur0; tiles[2 + 2 * nt] = cur; }
    ++nt;
  }
  tiles[0] = nt < max_tiles ? nt : max_tiles;
}

// lower limit of every row's candidate band from the lists of a SAMPLE of the columns (pass 1 of the symmetric
// form): lim = (k-th largest g~ of the union of the row's lists) - band, or -inf while fewer than k were seen
__global__ void __launch_bounds__(kMergeThreads)
knn_limits_kernel(const unsigned long long* __restrict__ lists, const int* __restrict__ counts, int S, int C, int n_rows,
                  int k, float band, float* __restrict__ lim_out) {
  extern __shared__ unsigned long long s_ent[];
  __shared__ int hist[256];
  __shared__ int s_n, s_base;
  __shared__ uint32_t s_prefix;
  __shared__ int s_remaining;
  const int tid = threadIdx.x;
  for (int r = blockIdx.x; r < n_rows; r += gridDim.x) {
    __syncthreads();
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int sgm = 0; sgm < S; ++sgm) {
      const int c = min(counts[(long long)r * S + sgm], C);
      const unsigned long long* L = lists + ((long long)r * S + sgm) * C;
      if (tid == 0) { s_base = s_n; s_n += c; }
      __syncthreads();
      for (int i = tid; i < c; i += kMergeThreads) s_ent[s_base + i] = L[i];
      __syncthreads();
    }
    const int n = s_n;
    if (n < k) {
      if (tid == 0) lim_out[r] = -3.0e38f;
      continue;
    }
    uint32_t prefix = 0, mask = 0;
    int remaining = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (int i = tid; i < 256; i += kMergeThreads) hist[i] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += kMergeThreads) {
        const uint32_t key = float_to_key(__uint_as_float((uint32_t)(s_ent[i] >> 32)));
        if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
      }
      __syncthreads();
      if (tid < 32) {
        __shared__ int s_d, s_above;
        radix_pick_digit(hist, remaining, &s_d, &s_above);
        __syncwarp();
        if (tid == 0) {
          s_prefix = prefix | ((uint32_t)s_d << shift);
          s_remaining = remaining - s_above;
        }
      }
      __syncthreads();
      prefix = s_prefix;
      remaining = s_remaining;
      mask |= 255u << shift;
      __syncthreads();
    }
    if (tid == 0) lim_out[r] = key_to_float(prefix) - band;
  }
}

// indices of the zero-norm patterns: ws[0] = count (may exceed cap), ws[1 + i] = index (first cap only)
__global__ void zero_cols_kernel(const int* __restrict__ zero_flag, int n, int cap, int* __restrict__ ws) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
    if (zero_flag[c]) {
      const int slot = atomicAdd(&ws[0], 1);
      if (slot < cap) ws[1 + slot] = c;
    }
}

}  // namespace v3s

using namespace v3s;

// Persistent workers (CTAs, or CTA pairs) that are co-resident: a static schedule over more workers
// than fit would leave the late ones a whole extra share.  CTA pairs need two free SMs of one TPC.
template <int CG>
static int topk_query_workers() {
  auto kern = gram_topk_kernel<CG, 16>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TopkCfg<CG>::SMEM) != cudaSuccess) {
    (void)cudaGetLastError();
    return kNumSMs / CG;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)kNumSMs);
  cfg.blockDim = dim3(TK_THREADS);
  cfg.dynamicSmemBytes = TopkCfg<CG>::SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc < 1) { (void)cudaGetLastError(); nc = kNumSMs / CG; }
  return nc < kNumSMs / CG ? nc : kNumSMs / CG;
}
static int g_reserved_sms = 0;
static int g_sync_lag = 0;
static int topk_workers(int cg) {
  static int w[3] = {0, 0, 0};
  if (w[cg] == 0) w[cg] = (cg == 1) ? topk_query_workers<1>() : topk_query_workers<2>();
  const int r = w[cg] - (g_reserved_sms + cg - 1) / cg;
  return r < 1 ? 1 : r;
}
// Slack of the workers' per-tile lockstep (TopkPlan::sync_lag): 0 = every worker starts every tile together.
extern "C" int v3s_set_gram_sync_lag(int lag) {
  V3S_REQUIRE(lag >= 0 && lag <= 8, "v3s_set_gram_sync_lag: 0..8");
  g_sync_lag = lag;
  return 0;
}
// SMs the persistent Gram kernels leave free (0 = none): a collective that runs BESIDE the kernel needs a few, and
// the kernel's workers must all be resident at once (they start every tile together)
extern "C" int v3s_set_gram_reserved_sms(int sms) {
  V3S_REQUIRE(sms >= 0 && sms <= 64, "v3s_set_gram_reserved_sms: 0..64");
  g_reserved_sms = sms;
  return 0;
}

static TopkPlan make_topk_plan(long long n_rows, long long n_cols, int cg, int workers) {
  TopkPlan p{};
  const int rows_per_block = TK_BM * cg;
  p.n_rb = (int)ceil_div64(n_rows, rows_per_block);
  p.tiles_c = (int)ceil_div64(n_cols, TK_BN);
  // segments per row block: fewest waves x tiles per unit (+ half a tile of per-unit overhead), at most 32
  long long best = -1;
  int best_s = 1;
  for (int s = 1; s <= 32 && s <= p.tiles_c; ++s) {
    const long long units = (long long)p.n_rb * s;
    const long long waves = ceil_div64(units, workers);
    const long long tps = ceil_div64(p.tiles_c, s);
    const long long cost = waves * (2 * tps + 1);
    if (best < 0 || cost < best) { best = cost; best_s = s; }
  }
  p.S = best_s;
  p.tps = (int)ceil_div64(p.tiles_c, p.S);
  p.S = (int)ceil_div64(p.tiles_c, p.tps);               // drop empty trailing segments
  p.rg = workers / p.S;
  if (p.rg < 1) p.rg = 1;
  p.units = p.n_rb * p.S;
  p.sync_lag = g_sync_lag;
  return p;
}

static int g_topk_cg = 0;

// cta_group of the fused Gram/top-k kernel: 1 = one CTA per SM (M = 128), 2 = CTA pairs (M = 256),
// 0 = by problem size (default): pairs from 2.5e8 Gram entries on -- below that the per-unit start-up of
// the selection (every list fills once before its first threshold exists) outweighs the operand sharing
extern "C" int v3s_set_gram_topk_cta_group(int cg) {
  V3S_REQUIRE(cg >= 0 && cg <= 2, "v3s_set_gram_topk_cta_group: 0 (auto), 1 or 2");
  g_topk_cg = cg;
  return 0;
}
static int topk_cta_group(long long n_rows, long long n_cols) {
  if (g_topk_cg) return g_topk_cg;
  return (double)n_rows * (double)n_cols >= 2.5e8 ? 2 : 1;
}

// Sizes of the list buffers of v3s_gram_topk for a problem: out[0] = S (lists per row), out[1] = C
// (entries per list), out[2] = padded row count (rows of `lists` / `counts`), out[3] = ints of the workspace `ws`.
extern "C" int v3s_gram_topk_plan(long long n_rows, long long n_cols, int k, int* out /* host [4] */) {
  V3S_REQUIRE(n_rows >= 0 && n_cols >= 1 && k >= 1 && out, "v3s_gram_topk_plan: bad arguments");
  const int cg = topk_cta_group(n_rows, n_cols);
  const TopkPlan p = make_topk_plan(n_rows, n_cols, cg, topk_workers(cg));
  int C = 256;
  while (C < k + 128 && C < 1024) C <<= 1;               // room for k + band + the 32 appends between two checks
  V3S_REQUIRE(C >= k + 96, "v3s_gram_topk: k = %d too large for the fused selection (max 928)", k);
  const int workers = p.units < topk_workers(cg) ? p.units : topk_workers(cg);
  const long long waves = ceil_div64(p.units, workers > 0 ? workers : 1);
  out[0] = p.S;
  out[1] = C;
  out[2] = p.n_rb * TK_BM * cg;
  out[3] = out[2] + (int)(waves * p.tps);     // ints of `ws`: one threshold key per row + one counter per (wave, tile) step
  return 0;
}

extern "C" int v3s_split_f16(const float* a_f32, long long n_rows, int P, int Ppad, void* a16, void* stream) {
  V3S_REQUIRE(n_rows >= 0 && P > 0 && Ppad >= P && Ppad % 8 == 0 && a16, "v3s_split_f16: bad arguments");
  if (n_rows == 0) return 0;
  long long g = ceil_div64(n_rows * (long long)(Ppad / 2), 256);
  if (g > 16LL * kNumSMs) g = 16LL * kNumSMs;
  split_f16_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(a_f32, n_rows, P, Ppad, (__half*)a16);
  V3S_CHECK_LAUNCH();
  return 0;
}

template <int CG, int EPL>
static int launch_topk(const CUtensorMap& mr, const CUtensorMap& ma, long long n_rows, long long n_cols, int num_kb,
                       const TopkPlan& plan, int max_workers, int k, float band, unsigned long long* lists, int* counts,
                       int* overflow, unsigned int* tau_keys, int* step_sync, cudaStream_t st) {
  auto kern = gram_topk_kernel<CG, EPL>;
  static bool attr_set = false;
  if (!attr_set) {
    V3S_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TopkCfg<CG>::SMEM));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TK_THREADS);
  cfg.dynamicSmemBytes = TopkCfg<CG>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int workers = plan.units < max_workers ? plan.units : max_workers;
  cfg.gridDim = dim3((unsigned)(workers * CG));
  V3S_CUDA(cudaLaunchKernelEx(&cfg, kern, mr, ma, (int)n_rows, (int)n_cols, num_kb, plan, k, band, lists, counts, overflow,
                              tau_keys, step_sync));
  ++v3s::g_kernel_launches;
  return 0;
}

// Fused Gram + per-row top-k candidate lists (distance.py:56 and :39-43 up to the candidate set).
//   a16_all [n_cols, Ppad], a16_rows [n_rows, Ppad]: FP16 operands from v3s_split_f16
//   lists  [rows_pad * S * C] 64-bit entries (g~ bits << 32 | column), counts [rows_pad * S]
//   (S, C, rows_pad from v3s_gram_topk_plan); eps_g = bound on |g~ - g|; *overflow counts saturated lists
//   ws int [plan[3]]: per-row threshold keys + per-step arrival counters (zeroed here)
extern "C" int v3s_gram_topk(const void* a16_all, long long n_cols, const void* a16_rows, long long n_rows, int Ppad, int k,
                             float eps_g, unsigned long long* lists, int* counts, int* overflow, int* ws, void* stream) {
  V3S_REQUIRE(Ppad > 0 && Ppad % 64 == 0, "v3s_gram_topk: Ppad=%d must be a positive multiple of 64", Ppad);
  V3S_REQUIRE(n_cols >= 1 && n_rows >= 0 && n_cols < (1LL << 31) && n_rows < (1LL << 31), "v3s_gram_topk: bad shape");
  V3S_REQUIRE(k >= 1 && k <= n_cols, "Number of nearest neighbors to preserve is too high.");   // distance.py:33-34
  if (n_rows == 0) return 0;
  const int cg = topk_cta_group(n_rows, n_cols);
  V3S_REQUIRE(lists && counts && ws, "v3s_gram_topk: lists, counts and ws are required");
  int pl[4];
  if (v3s_gram_topk_plan(n_rows, n_cols, k, pl)) return 1;
  V3S_CUDA(cudaMemsetAsync(ws, 0, sizeof(int) * (size_t)pl[3], (cudaStream_t)stream));
  unsigned int* tau_keys = reinterpret_cast<unsigned int*>(ws);
  int* step_sync = ws + pl[2];
  const int workers = topk_workers(cg);
  const TopkPlan plan = make_topk_plan(n_rows, n_cols, cg, workers);
  const int C = pl[1];
  CUtensorMap mr, ma;
  if (make_tmap_2d(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a16_rows, n_rows, Ppad, (long long)Ppad * 2, TK_BK, TK_BM,
                   CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  if (make_tmap_2d(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a16_all, n_cols, Ppad, (long long)Ppad * 2, TK_BK, TK_BN / cg,
                   CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  const float band = 2.0f * eps_g;
  cudaStream_t st = (cudaStream_t)stream;
  const int num_kb = Ppad / TK_BK;
#define V3S_TOPK_LAUNCH(CGV, EPLV) return launch_topk<CGV, EPLV>(mr, ma, n_rows, n_cols, num_kb, plan, workers, k, band, lists, counts, overflow, tau_keys, step_sync, st)
  if (cg == 1) {
    if (C == 256) V3S_TOPK_LAUNCH(1, 8);
    if (C == 512) V3S_TOPK_LAUNCH(1, 16);
    V3S_TOPK_LAUNCH(1, 32);
  } else {
    if (C == 256) V3S_TOPK_LAUNCH(2, 8);
    if (C == 512) V3S_TOPK_LAUNCH(2, 16);
    V3S_TOPK_LAUNCH(2, 32);
  }
#undef V3S_TOPK_LAUNCH
}

static int merge_lists_impl(const unsigned long long* lists, const int* counts, const MergeSegs& segs, int S, int C,
                            long long n_rows, long long n_cols, long long row0, int k, float eps_g, int cap,
                            const int* zero_flag, int* zero_ws, int* cand_ws, int* count_ws, int* overflow, int anchor_stride,
                            int* order_ws, int tile_rows, int max_tiles, int* tiles_ws, void* stream) {
  V3S_REQUIRE((zero_flag == nullptr) || (zero_ws != nullptr), "v3s_knn_merge_lists: zero_flag needs zero_ws");
  V3S_REQUIRE(S >= 1 && C >= 32 && n_rows >= 0 && k >= 1 && cap >= 1, "v3s_knn_merge_lists: bad arguments");
  // segment table: the S lists of a row hold ONE pattern's candidates split by source rank -- their union is sized
  // like one list (C entries, at most 24576); contiguous segments: every segment can be full
  const int ent_cap = segs.n ? (C < 24576 ? C : 24576) : S * C;
  V3S_REQUIRE((size_t)ent_cap * 8 <= 200 * 1024, "v3s_knn_merge_lists: S*C = %d entries exceed shared memory", ent_cap);
  if (n_rows == 0) return 0;
  const size_t smem = (size_t)ent_cap * 8;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    V3S_CUDA(cudaFuncSetAttribute(knn_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  if (zero_flag) {
    V3S_CUDA(cudaMemsetAsync(zero_ws, 0, sizeof(int), (cudaStream_t)stream));
    zero_cols_kernel<<<2 * kNumSMs, 256, 0, (cudaStream_t)stream>>>(zero_flag, (int)n_cols, cap, zero_ws);
    V3S_CHECK_LAUNCH();
  }
  const int grid = (int)(n_rows < 16LL * kNumSMs ? n_rows : 16LL * kNumSMs);
  V3S_REQUIRE(order_ws == nullptr || anchor_stride >= 1, "v3s_knn_merge_lists: anchor_stride >= 1 with order_ws");
  // order_ws layout: [0, n_rows) order, [n_rows, 2 n_rows) cluster ids, then the bins
  int* cluster = order_ws ? order_ws + n_rows : nullptr;
  knn_merge_kernel<<<grid, kMergeThreads, smem, (cudaStream_t)stream>>>(lists, counts, segs, S, C, ent_cap, (int)n_rows,
                                                                        (int)n_cols, row0, k, 2.0f * eps_g, cap, zero_flag,
                                                                        zero_flag ? zero_ws : nullptr, cand_ws, count_ws, overflow,
                                                                        anchor_stride, cluster);
  V3S_CHECK_LAUNCH();
  if (order_ws) {
    cudaStream_t st = (cudaStream_t)stream;
    const int n_bins = (int)(ceil_div64(n_cols, anchor_stride) + ceil_div64(n_rows, anchor_stride));
    int* bins = order_ws + 2 * n_rows;
    V3S_CUDA(cudaMemsetAsync(bins, 0, sizeof(int) * (size_t)n_bins, st));
    const int g = (int)(ceil_div64(n_rows, 256) < 4LL * kNumSMs ? ceil_div64(n_rows, 256) : 4LL * kNumSMs);
    cluster_hist_kernel<<<g, 256, 0, st>>>(cluster, (int)n_rows, bins);
    V3S_CHECK_LAUNCH();
    cluster_scan_kernel<<<1, 1024, 0, st>>>(bins, n_bins);
    V3S_CHECK_LAUNCH();
    cluster_scatter_kernel<<<g, 256, 0, st>>>(cluster, (int)n_rows, bins, order_ws);
    V3S_CHECK_LAUNCH();
    if (tiles_ws) {
      V3S_REQUIRE(tile_rows >= 1 && max_tiles >= 1, "v3s_knn_merge_lists: tile_rows / max_tiles with tiles_ws");
      cluster_tiles_kernel<<<1, 32, 0, st>>>(bins, n_bins, tile_rows, max_tiles, tiles_ws);
      V3S_CHECK_LAUNCH();
    }
  }
  return 0;
}

// Candidate index lists from the per-segment lists of v3s_gram_topk (see knn_merge_kernel).
extern "C" int v3s_knn_merge_lists(const unsigned long long* lists, const int* counts, int S, int C, long long n_rows,
                                   long long n_cols, long long row0, int k, float eps_g, int cap, const int* zero_flag,
                                   int* zero_ws /* int [cap + 1], required with zero_flag */, int* cand_ws, int* count_ws,
                                   int* overflow, int anchor_stride, int* order_ws /* int [2 n_rows + bins], see header */,
                                   int tile_rows, int max_tiles, int* tiles_ws /* int [1 + 2 max_tiles] */, void* stream) {
  V3S_REQUIRE(lists && counts, "v3s_knn_merge_lists: null lists");
  MergeSegs segs{};
  return merge_lists_impl(lists, counts, segs, S, C, n_rows, n_cols, row0, k, eps_g, cap, zero_flag, zero_ws, cand_ws, count_ws,
                          overflow, anchor_stride, order_ws, tile_rows, max_tiles, tiles_ws, stream);
}

// The same with the S segments of a row given by a table: seg_lists[s] -> [n_rows, C] entries, seg_counts[s] ->
// [n_rows] counters (host arrays of device addresses, S <= 8) -- the sharded symmetric Gram keeps one segment per
// SOURCE rank in that rank's memory; the owner of the rows reads them over the NVLink peer mappings.
extern "C" int v3s_knn_merge_segment_lists(const unsigned long long* seg_lists /* host [S] */,
                                           const unsigned long long* seg_counts /* host [S] */, int S, int C, long long n_rows,
                                           long long n_cols, long long row0, int k, float eps_g, int cap, const int* zero_flag,
                                           int* zero_ws, int* cand_ws, int* count_ws, int* overflow, int anchor_stride,
                                           int* order_ws, int tile_rows, int max_tiles, int* tiles_ws, void* stream) {
  V3S_REQUIRE(seg_lists && seg_counts && S >= 1 && S <= 8, "v3s_knn_merge_segment_lists: 1..8 segments");
  MergeSegs segs{};
  for (int i = 0; i < S; ++i) {
    segs.lists[i] = reinterpret_cast<const unsigned long long*>(seg_lists[i]);
    segs.counts[i] = reinterpret_cast<const int*>(seg_counts[i]);
    V3S_REQUIRE(segs.lists[i] && segs.counts[i], "v3s_knn_merge_segment_lists: null segment %d", i);
  }
  segs.n = S;
  return merge_lists_impl(nullptr, nullptr, segs, S, C, n_rows, n_cols, row0, k, eps_g, cap, zero_flag, zero_ws, cand_ws,
                          count_ws, overflow, anchor_stride, order_ws, tile_rows, max_tiles, tiles_ws, stream);
}


// ---- symmetric form, host side ---------------------------------------------------------------------------------
static int g_sym_cg = 2;

// Pass 1: lower limits lim [n_rows] of the rows' candidate bands from every `sample_stride`-th pattern (the
// rectangular fused kernel on a strided view of a16_all; lists / counts / ws sized by v3s_gram_topk_plan for
// (n_rows, ceil(n_cols / sample_stride), k)).
extern "C" int v3s_gram_topk_limits(const void* a16_all, long long n_cols, const void* a16_rows, long long n_rows, int Ppad,
                                    int k, float eps_g, int sample_stride, unsigned long long* lists, int* counts, int* overflow,
                                    int* ws, float* lim_rows, void* stream) {
  V3S_REQUIRE(sample_stride >= 1 && lim_rows, "v3s_gram_topk_limits: bad arguments");
  const long long ns = ceil_div64(n_cols, sample_stride);
  V3S_REQUIRE(k <= ns, "v3s_gram_topk_limits: the column sample (%lld) is smaller than k", ns);
  if (n_rows == 0) return 0;
  const int cg = topk_cta_group(n_rows, ns);
  int pl[4];
  if (v3s_gram_topk_plan(n_rows, ns, k, pl)) return 1;
  V3S_CUDA(cudaMemsetAsync(ws, 0, sizeof(int) * (size_t)pl[3], (cudaStream_t)stream));
  const int workers = topk_workers(cg);
  const TopkPlan plan = make_topk_plan(n_rows, ns, cg, workers);
  const int C = pl[1];
  CUtensorMap mr, ma;
  if (make_tmap_2d(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a16_rows, n_rows, Ppad, (long long)Ppad * 2, TK_BK, TK_BM,
                   CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  if (make_tmap_2d(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a16_all, ns, Ppad, (long long)Ppad * 2 * sample_stride, TK_BK, TK_BN / cg,
                   CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  const float band = 2.0f * eps_g;
  cudaStream_t st = (cudaStream_t)stream;
  const int num_kb = Ppad / TK_BK;
  unsigned int* tau_keys = reinterpret_cast<unsigned int*>(ws);
  int* step_sync = ws + pl[2];
  int rc;
#define V3S_TOPK_LAUNCH(CGV, EPLV) rc = launch_topk<CGV, EPLV>(mr, ma, n_rows, ns, num_kb, plan, workers, k, band, lists, counts, overflow, tau_keys, step_sync, st)
  if (cg == 1) {
    if (C == 256) V3S_TOPK_LAUNCH(1, 8); else if (C == 512) V3S_TOPK_LAUNCH(1, 16); else V3S_TOPK_LAUNCH(1, 32);
  } else {
    if (C == 256) V3S_TOPK_LAUNCH(2, 8); else if (C == 512) V3S_TOPK_LAUNCH(2, 16); else V3S_TOPK_LAUNCH(2, 32);
  }
#undef V3S_TOPK_LAUNCH
  if (rc) return rc;
  const size_t smem = (size_t)pl[0] * C * 8;
  V3S_REQUIRE(smem <= 200 * 1024, "v3s_gram_topk_limits: S*C = %d entries exceed shared memory", pl[0] * C);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    V3S_CUDA(cudaFuncSetAttribute(knn_limits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  const int grid = (int)(n_rows < 16LL * kNumSMs ? n_rows : 16LL * kNumSMs);
  knn_limits_kernel<<<grid, kMergeThreads, smem, st>>>(lists, counts, pl[0], C, (int)n_rows, k, band, lim_rows);
  V3S_CHECK_LAUNCH();
  return 0;
}

template <int CG>
static int sym_workers() {
  auto kern = gram_sym_topk_kernel<CG>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, TopkCfg<CG>::SMEM) != cudaSuccess) {
    (void)cudaGetLastError();
    return kNumSMs / CG;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)kNumSMs);
  cfg.blockDim = dim3(TK_THREADS);
  cfg.dynamicSmemBytes = TopkCfg<CG>::SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nc = 0;
  if (cudaOccupancyMaxActiveClusters(&nc, kern, &cfg) != cudaSuccess || nc < 1) { (void)cudaGetLastError(); nc = kNumSMs / CG; }
  return nc < kNumSMs / CG ? nc : kNumSMs / CG;
}
static SymPlan make_sym_plan(long long n, int cg, int workers, int rank, int world) {
  SymPlan p{};
  p.n_rb = (int)ceil_div64(n, TK_BM * cg);
  p.n_cb = (int)ceil_div64(n, TK_BN);
  const int total = workers * world;
  int pc = 1;
  while ((pc + 1) * (pc + 1) <= total) ++pc;               // floor(sqrt(total))
  if (pc * (pc + 1) <= total) ++pc;                         // a slightly wide patch wastes fewer workers
  p.PC = pc;
  p.PR = total / pc;
  p.rank = rank; p.world = world; p.workers = workers;
  p.sync_lag = g_sync_lag;
  p.steps = sym_count_steps(p, TK_BM * cg);
  return p;
}
// ints of the step-counter workspace of v3s_gram_topk_sym for n patterns on `world` ranks
// CTA pairs of the symmetric kernel: all that fit, minus the SMs reserved for a collective running beside it (every
// worker must be resident: they start each patch together)
static int sym_workers_now() {
  static int w2 = 0;
  if (!w2) w2 = sym_workers<2>();
  const int r = w2 - (g_reserved_sms + 1) / 2;
  return r < 1 ? 1 : r;
}
extern "C" int v3s_gram_topk_sym_steps(long long n, int world) {
  if (n < 1 || world < 1 || world > 8) return 0;
  // (largest over the worker counts a reservation can leave: fewer workers -> smaller patches -> more steps)
  int most = 0;
  static int w2 = 0;
  if (!w2) w2 = sym_workers<2>();
  for (int w = w2 > 32 ? w2 - 32 : 1; w <= w2; ++w) {
    const int st = make_sym_plan(n, g_sym_cg, w, 0, world).steps + 1;
    most = st > most ? st : most;
  }
  return most;
}

// Pass 2: the tiles on / above the diagonal of the n x n Gram matrix, this rank's share; appends (g~, index) to the
// lists of BOTH patterns of every element that reaches their limit lim [n] (complete on every rank).  peer_lists /
// peer_counts (host arrays [world]): rank o's lists [cnt_o, C] (64-bit entries) and slot counters [cnt_o] (zeroed by
// the caller) as mapped in THIS process; rank o owns cnt_base + (o < rem) consecutive patterns.  step_ws: int
// [v3s_gram_topk_sym_steps(n, world)] (zeroed here).  The caller synchronises the ranks before reading its lists.
extern "C" int v3s_gram_topk_sym(const void* a16_all, long long n, int Ppad, const float* lim, int C,
                                 const unsigned long long* peer_lists /* host [world] */,
                                 const unsigned long long* peer_counts /* host [world] */, int world, int rank,
                                 long long cnt_base, int rem, int* step_ws, void* stream) {
  V3S_REQUIRE(Ppad > 0 && Ppad % 64 == 0 && n >= 1 && n < (1LL << 31), "v3s_gram_topk_sym: bad shape");
  V3S_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world && cnt_base >= 1 && rem >= 0 && rem < world &&
                  cnt_base * world + rem == n, "v3s_gram_topk_sym: bad sharding");
  V3S_REQUIRE(lim && peer_lists && peer_counts && step_ws && C >= 32, "v3s_gram_topk_sym: null argument");
  const int workers = sym_workers_now();
  const SymPlan plan = make_sym_plan(n, 2, workers, rank, world);
  cudaStream_t st = (cudaStream_t)stream;
  V3S_CUDA(cudaMemsetAsync(step_ws, 0, sizeof(int) * (size_t)(plan.steps + 1), st));
  SymPeers pr{};
  for (int i = 0; i < world; ++i) {
    pr.lists[i] = reinterpret_cast<unsigned long long*>(peer_lists[i]);
    pr.counts[i] = reinterpret_cast<int*>(peer_counts[i]);
  }
  pr.world = world; pr.rem = rem; pr.cnt_base = cnt_base;
  CUtensorMap mr, mc;
  if (make_tmap_2d(&mr, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a16_all, n, Ppad, (long long)Ppad * 2, TK_BK, TK_BM,
                   CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  if (make_tmap_2d(&mc, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, a16_all, n, Ppad, (long long)Ppad * 2, TK_BK, TK_BN / 2,
                   CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
  auto kern = gram_sym_topk_kernel<2>;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(TK_THREADS);
  cfg.dynamicSmemBytes = TopkCfg<2>::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3((unsigned)(workers * 2));
  V3S_CUDA(cudaLaunchKernelEx(&cfg, kern, mr, mc, (int)n, Ppad / TK_BK, plan, lim, C, pr, step_ws));
  ++v3s::g_kernel_launches;
  return 0;
}
