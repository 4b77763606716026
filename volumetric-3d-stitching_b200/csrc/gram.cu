// K2b -- Gram row panel  G[r, c] = <a_r, a_c>  on tcgen05 tensor cores (sm_100a).
//
// Replaces np.matmul(A, A.T) in distance.py:56.  FP32-equivalent accuracy from BF16 tensor
// cores by operand splitting a = hi + lo (prep.cu):  G = hi.hi^T + hi.lo^T + lo.hi^T  (the
// dropped lo.lo^T term is 2^-18 relative).  All three products accumulate into the same TMEM
// accumulator, so the split is a K-extension of one GEMM, not three GEMMs.
//
// Kernel shape (one CTA per SM, persistent over tiles):
//   warp 4 lane 0 : TMA producer (cp.async.bulk.tensor, SWIZZLE_128B, 2-stage mbarrier ring)
//   warp 5 lane 0 : tcgen05.mma issuer (cta_group::1, M=128 x N=256 x K=16, kind::f16/BF16)
//   warps 0-3     : epilogue, tcgen05.ld TMEM -> registers -> coalesced global stores
//   TMEM          : 2 accumulators x 256 columns (double buffered: epilogue overlaps next tile)
// The MMA "M" side (TMEM lanes) carries the column index c of G and the "N" side (TMEM
// columns) the panel row r, so for a fixed r the 32 lanes of a warp store 128 contiguous bytes.
#include "common.cuh"
#include "v3s.h"
#include "tma.cuh"

namespace v3s {

constexpr int BM = 128;        // TMEM lanes   <-> c (all patterns)
constexpr int BN = 256;        // TMEM columns <-> r (panel rows)
constexpr int UMMA_K = 16;
constexpr int GRAM_THREADS = 192;
// Two pipeline shapes with the same 192 KB of operand staging:
//   BK = 64: one 128-byte swizzle atom per row, 2 stages of 96 KB
//   BK = 32: one  64-byte swizzle atom per row, 4 stages of 48 KB (finer refills hide TMA latency better)
template <int BK> struct GramCfg {
  static constexpr int STAGES = (BK == 64) ? 2 : 4;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // hi+lo for both operands
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// K-major swizzled shared-memory matrix descriptor: rows of BK bf16 (128 B -> SWIZZLE_128B, 64 B ->
// SWIZZLE_64B), 8-row groups 8 * row-bytes apart
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);       // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * BK * 2) >> 4) << 32;         // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
  d |= (uint64_t)(BK == 64 ? 2 : 4) << 61;          // layout type: SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return d;
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Tile order: groups of GRAM_GROUP column blocks, row blocks innermost-but-one, so that the ~148
// tiles in flight at any time form a compact (rows x cols) patch whose operands are shared through
// L2 instead of being re-read from HBM for every tile.
constexpr int GRAM_GROUP = 16;
__device__ __forceinline__ void gram_decode_tile(int tile, int tiles_c, int tiles_r, int& tc, int& tr) {
  const int per_group = GRAM_GROUP * tiles_r;
  const int g = tile / per_group;
  const int first_c = g * GRAM_GROUP;
  const int gw = min(GRAM_GROUP, tiles_c - first_c);
  const int rem = tile - g * per_group;
  tc = first_c + rem % gw;
  tr = rem / gw;
}

// instruction descriptor: D=F32, A=B=BF16, both K-major, N=256, M=128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// Row-sharded output: rank o owns rows [lo_o, lo_o + cnt_o) of G in its own buffer base[o]
// (peer-mapped over NVLink for o != this rank).  world == 0: one local buffer (base[0], all rows).
struct GramPeers {
  float* base[8];
  int world;
  int rem;                 // the first `rem` ranks own cnt_base + 1 rows
  long long cnt_base;
};
__device__ __forceinline__ void gram_owner(const GramPeers& pr, long long x, int& o, long long& lo, long long& hi) {
  const long long big = pr.cnt_base + 1, split = big * pr.rem;
  if (x < split) { o = (int)(x / big); lo = o * big; hi = lo + big; }
  else { o = pr.rem + (int)((x - split) / pr.cnt_base); lo = split + (long long)(o - pr.rem) * pr.cnt_base; hi = lo + pr.cnt_base; }
}

template <int BK>
__global__ void __launch_bounds__(GRAM_THREADS, 1)
gram_tc_kernel(const __grid_constant__ CUtensorMap tm_c_hi, const __grid_constant__ CUtensorMap tm_c_lo,
               const __grid_constant__ CUtensorMap tm_r_hi, const __grid_constant__ CUtensorMap tm_r_lo,
               float* __restrict__ G, long long ldg, int n_cols, int n_rows, int num_kb, int tiles_c, int tiles_r,
               int num_tiles, const int2* __restrict__ tile_list, int mirror, const GramPeers peers) {
  constexpr int STAGES = GramCfg<BK>::STAGES, A_BYTES = GramCfg<BK>::A_BYTES, B_BYTES = GramCfg<BK>::B_BYTES,
                STAGE_BYTES = GramCfg<BK>::STAGE_BYTES;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* tfull = bars + 2 * STAGES;
  uint64_t* tempty = bars + 2 * STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_c_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_r_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_r_lo) : "memory");
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // ---------------- TMA producer ----------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int tc, tr;
        if (tile_list) { const int2 t = tile_list[tile]; tc = t.x; tr = t.y; }
        else gram_decode_tile(tile, tiles_c, tiles_r, tc, tr);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], STAGE_BYTES);
          const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
          tma_load_2d(s, &tm_c_hi, kb * BK, tc * BM, &full[stage]);
          tma_load_2d(s + A_BYTES, &tm_c_lo, kb * BK, tc * BM, &full[stage]);
          tma_load_2d(s + 2 * A_BYTES, &tm_r_hi, kb * BK, tr * BN, &full[stage]);
          tma_load_2d(s + 2 * A_BYTES + B_BYTES, &tm_r_lo, kb * BK, tr * BN, &full[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ---------------- MMA issuer ----------------
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t s = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_hi = make_kmajor_desc<BK>(s);
          const uint64_t a_lo = make_kmajor_desc<BK>(s + A_BYTES);
          const uint64_t b_hi = make_kmajor_desc<BK>(s + 2 * A_BYTES);
          const uint64_t b_lo = make_kmajor_desc<BK>(s + 2 * A_BYTES + B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t ko = (uint64_t)((k * UMMA_K * 2) >> 4);   // 32 B per K step, in 16-byte units
            umma_bf16(d_tmem, a_hi + ko, b_hi + ko, kIdesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_bf16(d_tmem, a_hi + ko, b_lo + ko, kIdesc, 1u);
            umma_bf16(d_tmem, a_lo + ko, b_hi + ko, kIdesc, 1u);
          }
          umma_commit(&empty[stage]);                 // frees the smem stage once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ---------------- epilogue (warps 0..3 <-> TMEM lane quadrants) ----------------
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int tc, tr;
      if (tile_list) { const int2 t = tile_list[tile]; tc = t.x; tr = t.y; }
      else gram_decode_tile(tile, tiles_c, tiles_r, tc, tr);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int c = tc * BM + warp * 32 + lane;
      const bool c_ok = c < n_cols;
      const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int ch = 0; ch < BN / 32; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_row + (uint32_t)(ch * 32), v);
        tmem_ld_wait();
        const int r0 = tr * BN + ch * 32;
        if (peers.world == 0) {
          float* dst = G + (long long)r0 * ldg + c;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c_ok && r0 + j < n_rows) dst[(long long)j * ldg] = __uint_as_float(v[j]);
          if (mirror && c_ok) {
            // symmetric mode: only tiles on / above the diagonal are computed; G[c][r] = G[r][c]
            float* mdst = G + (long long)c * ldg + r0;
            if (r0 + 32 <= n_rows) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(mdst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                   __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (r0 + j < n_rows) mdst[j] = __uint_as_float(v[j]);
            }
          }
        } else if (c_ok) {
          // sharded symmetric mode: every store goes to the buffer of the rank that owns the row --
          // local memory or a peer GPU's HBM over NVLink (the transfer overlaps the next tile's MMAs)
          int o; long long lo, hi;
          gram_owner(peers, r0, o, lo, hi);
          float* dst = peers.base[o] + (r0 - lo) * ldg + c;          // row r0 in its owner's block
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const long long r = r0 + j;
            if (r < n_rows) {
              if (r >= hi) { gram_owner(peers, r, o, lo, hi); dst = peers.base[o] + (r - lo) * ldg + c - (long long)j * ldg; }
              dst[(long long)j * ldg] = __uint_as_float(v[j]);
            }
          }
          int oc; long long loc, hic;
          gram_owner(peers, c, oc, loc, hic);
          float* mdst = peers.base[oc] + (c - loc) * ldg + r0;       // mirror: row c in its owner's block
          if (r0 + 32 <= n_rows) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(mdst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                 __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (r0 + j < n_rows) mdst[j] = __uint_as_float(v[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---- plain FP32 SIMT Gram: validation path for the tensor-core kernel (tests, debugging) ----
__global__ void __launch_bounds__(256)
gram_simt_kernel(const float* __restrict__ a_cols, const float* __restrict__ a_rows, int P, float* __restrict__ G,
                 long long ldg, int n_cols, int n_rows) {
  __shared__ float sc[16][65], sr[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < P; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int rr = i >> 4, kk = i & 15;
      const int k = k0 + kk;
      sc[kk][rr] = (c0 + rr < n_cols && k < P) ? a_cols[(long long)(c0 + rr) * P + k] : 0.f;
      sr[kk][rr] = (r0 + rr < n_rows && k < P) ? a_rows[(long long)(r0 + rr) * P + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float cv[4], rv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { cv[i] = sc[kk][tx * 4 + i]; rv[i] = sr[kk][ty * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(rv[i], cv[j], acc[i][j]);
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + ty * 4 + i, c = c0 + tx * 4 + j;
      if (r < n_rows && c < n_cols) G[(long long)r * ldg + c] = acc[i][j];
    }
}

// Exact all-pairs round distance d = 2 asin(|a_r - a_c| / 2) (== arccos(<a_r,a_c>) for unit rows),
// FP32 differences, FP64 result: the dense form of convert_to_round_distance (distance.py:51-63).
__global__ void __launch_bounds__(256)
round_distance_kernel(const float* __restrict__ a, int P, const int* __restrict__ zero_flag, double* __restrict__ B,
                      long long n) {
  __shared__ float sc[16][65], sr[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long c0 = (long long)blockIdx.x * 64, r0 = (long long)blockIdx.y * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < P; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int rr = i >> 4, kk = i & 15;
      const int k = k0 + kk;
      sc[kk][rr] = (c0 + rr < n && k < P) ? a[(c0 + rr) * P + k] : 0.f;
      sr[kk][rr] = (r0 + rr < n && k < P) ? a[(r0 + rr) * P + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float cv[4], rv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { cv[i] = sc[kk][tx * 4 + i]; rv[i] = sr[kk][ty * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float d = rv[i] - cv[j]; acc[i][j] = fmaf(d, d, acc[i][j]); }
    }
    __syncthreads();
  }
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      const long long r = r0 + ty * 4 + i, c = c0 + tx * 4 + j;
      if (r < n && c < n) {
        double h = 0.5 * sqrt((double)acc[i][j]);
        if (h > 1.0) h = 1.0;
        double d = 2.0 * asin(h);
        if (zero_flag && (zero_flag[r] || zero_flag[c])) d = 0.0;   // NaN -> 0, distance.py:58
        B[r * n + c] = d;
      }
    }
}

int make_tmap_2d(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, long long rows, long long cols,
                 long long pitch_bytes, int box_cols, int box_rows, CUtensorMapSwizzle swz) {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    V3S_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    V3S_REQUIRE(p != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available from the driver");
    fn = (PFN_tmapEncodeTiled)p;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)pitch_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  (void)elem_bytes;
  CUresult r = fn(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  V3S_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

static int g_gram_bk = 32;

static int make_bf16_map(CUtensorMap* map, const void* base, long long rows, int Ppad, int box_rows) {
  return make_tmap_2d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, rows, Ppad, (long long)Ppad * 2, g_gram_bk, box_rows,
                      g_gram_bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

}  // namespace v3s

using namespace v3s;

static int gram_launch(const void* all_hi, const void* all_lo, long long n_cols, const void* rows_hi, const void* rows_lo,
                       long long n_rows, int Ppad, float* G, long long ldg, const int2* tile_list, int num_list_tiles,
                       int mirror, const GramPeers& peers, void* stream) {
  V3S_REQUIRE(Ppad > 0 && Ppad % 64 == 0, "v3s_gram_rows: Ppad=%d must be a positive multiple of 64", Ppad);
  V3S_REQUIRE(n_cols >= 0 && n_rows >= 0 && ldg >= n_cols, "v3s_gram_rows: bad shape");
  V3S_REQUIRE(n_cols < (1LL << 31) && n_rows < (1LL << 31), "v3s_gram_rows: extents must fit int32");
  if (n_cols == 0 || n_rows == 0) return 0;
  CUtensorMap mc_hi, mc_lo, mr_hi, mr_lo;
  if (make_bf16_map(&mc_hi, all_hi, n_cols, Ppad, BM)) return 1;
  if (make_bf16_map(&mc_lo, all_lo, n_cols, Ppad, BM)) return 1;
  if (make_bf16_map(&mr_hi, rows_hi, n_rows, Ppad, BN)) return 1;
  if (make_bf16_map(&mr_lo, rows_lo, n_rows, Ppad, BN)) return 1;
  const int tiles_c = (int)ceil_div64(n_cols, BM), tiles_r = (int)ceil_div64(n_rows, BN);
  const long long nt = (long long)tiles_c * tiles_r;
  V3S_REQUIRE(nt < (1LL << 31), "v3s_gram_rows: too many tiles");
  static bool attr_set = false;
  if (!attr_set) {
    V3S_CUDA(cudaFuncSetAttribute(gram_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, GramCfg<64>::SMEM));
    V3S_CUDA(cudaFuncSetAttribute(gram_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GramCfg<32>::SMEM));
    attr_set = true;
  }
  const long long ntl = tile_list ? (long long)num_list_tiles : nt;
  if (ntl == 0) return 0;
  const int grid = (int)(ntl < kNumSMs ? ntl : kNumSMs);
  if (g_gram_bk == 64)
    gram_tc_kernel<64><<<grid, GRAM_THREADS, GramCfg<64>::SMEM, (cudaStream_t)stream>>>(
        mc_hi, mc_lo, mr_hi, mr_lo, G, ldg, (int)n_cols, (int)n_rows, Ppad / 64, tiles_c, tiles_r, (int)ntl, tile_list, mirror,
        peers);
  else
    gram_tc_kernel<32><<<grid, GRAM_THREADS, GramCfg<32>::SMEM, (cudaStream_t)stream>>>(
        mc_hi, mc_lo, mr_hi, mr_lo, G, ldg, (int)n_cols, (int)n_rows, Ppad / 32, tiles_c, tiles_r, (int)ntl, tile_list, mirror,
        peers);
  V3S_CHECK_LAUNCH();
  return 0;
}

// Pipeline shape of the Gram kernel: 64 (2 stages, SWIZZLE_128B) or 32 (4 stages, SWIZZLE_64B; default).
extern "C" int v3s_set_gram_kblock(int bk) {
  V3S_REQUIRE(bk == 32 || bk == 64, "v3s_set_gram_kblock: 32 or 64");
  g_gram_bk = bk;
  return 0;
}

extern "C" int v3s_gram_rows(const void* all_hi, const void* all_lo, long long n_cols, const void* rows_hi,
                             const void* rows_lo, long long n_rows, int Ppad, float* G, long long ldg, void* stream) {
  GramPeers none{};
  return gram_launch(all_hi, all_lo, n_cols, rows_hi, rows_lo, n_rows, Ppad, G, ldg, nullptr, 0, 0, none, stream);
}

// Whole symmetric Gram matrix G [n, ldg] of one operand set: only the (column block 128 x row block
// 256) tiles listed in tile_list (x = column block, y = row block; the caller lists those with
// 128*x + 127 >= 256*y, i.e. on or above the diagonal, in an L2-friendly order) are computed, and
// every tile is stored twice, as G[r][c] and as its mirror G[c][r].  Halves the tensor-core work.
extern "C" int v3s_gram_symmetric(const void* a_hi, const void* a_lo, long long n, int Ppad, float* G, long long ldg,
                                  const int* tile_list_xy, int num_tiles, void* stream) {
  V3S_REQUIRE(tile_list_xy != nullptr && num_tiles >= 0, "v3s_gram_symmetric: tile list required");
  GramPeers none{};
  return gram_launch(a_hi, a_lo, n, a_hi, a_lo, n, Ppad, G, ldg, reinterpret_cast<const int2*>(tile_list_xy), num_tiles, 1,
                     none, stream);
}

// Row-sharded symmetric Gram, fused with its exchange: this rank computes the tiles of its share of
// the global upper-triangle tile list (any rank can compute any tile: the operands are all-gathered)
// and stores every value twice -- G[r][c] into the buffer of the rank owning row r and the mirror
// G[c][r] into the buffer of the rank owning row c -- straight from the tensor-core epilogue through
// NVLink peer mappings (peer_bases[o] = base of rank o's [cnt_o, ldg] row block, e.g. from
// torch.distributed._symmetric_memory).  Rank o owns cnt_base + (o < rem) consecutive rows.
// The caller synchronises the ranks (a barrier) before reading its block.
extern "C" int v3s_gram_symmetric_sharded(const void* a_hi, const void* a_lo, long long n, int Ppad,
                                          const unsigned long long* peer_bases /* host [world] */, int world,
                                          long long cnt_base, int rem, long long ldg, const int* tile_list_xy,
                                          int num_tiles, void* stream) {
  V3S_REQUIRE(world >= 1 && world <= 8 && cnt_base >= 1 && rem >= 0 && rem < world, "v3s_gram_symmetric_sharded: bad sharding");
  V3S_REQUIRE(cnt_base * world + rem == n, "v3s_gram_symmetric_sharded: counts do not add up to n");
  V3S_REQUIRE(tile_list_xy != nullptr && num_tiles >= 0 && ldg >= n, "v3s_gram_symmetric_sharded: bad arguments");
  GramPeers pr{};
  for (int i = 0; i < world; ++i) pr.base[i] = reinterpret_cast<float*>(peer_bases[i]);
  pr.world = world; pr.rem = rem; pr.cnt_base = cnt_base;
  return gram_launch(a_hi, a_lo, n, a_hi, a_lo, n, Ppad, pr.base[0], ldg, reinterpret_cast<const int2*>(tile_list_xy),
                     num_tiles, 1, pr, stream);
}

extern "C" int v3s_gram_rows_simt(const float* a_all, long long n_cols, const float* a_rows, long long n_rows, int P,
                                  float* G, long long ldg, void* stream) {
  V3S_REQUIRE(P > 0 && n_cols >= 0 && n_rows >= 0 && ldg >= n_cols, "v3s_gram_rows_simt: bad shape");
  if (n_cols == 0 || n_rows == 0) return 0;
  dim3 grid((unsigned)ceil_div64(n_cols, 64), (unsigned)ceil_div64(n_rows, 64));
  gram_simt_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a_all, a_rows, P, G, ldg, (int)n_cols, (int)n_rows);
  V3S_CHECK_LAUNCH();
  return 0;
}


extern "C" int v3s_round_distance_dense(const float* a_f32, long long n, int P, const int* zero_flag, double* B,
                                        void* stream) {
  V3S_REQUIRE(n >= 0 && P > 0, "v3s_round_distance_dense: bad shape");
  if (n == 0) return 0;
  dim3 grid((unsigned)ceil_div64(n, 64), (unsigned)ceil_div64(n, 64));
  round_distance_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a_f32, P, zero_flag, B, n);
  V3S_CHECK_LAUNCH();
  return 0;
}
