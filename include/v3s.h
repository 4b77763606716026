/* v3s -- C ABI of the B200-native hot path of nurlicht/Volumetric-3D-Stitching.
 *
 * The reference has no FFI; its boundary is the Python stage API.
 * Each entry point below names the reference function(s) it replaces (paths relative to the
 * reference's src/python/).  Conventions:
 *   - every pointer is a DEVICE pointer unless marked "host"; matrices are row-major;
 *   - the caller owns every buffer (including workspaces); nothing persistent is allocated;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream and do
 *     not synchronise the host (v3s_synth_patterns does one small pageable host-to-device copy);
 *   - return value 0 = ok, non-zero = error, message from v3s_last_error() (thread-local);
 *   - there is no CPU fallback: without an sm_100a device the calls fail.
 */
#ifndef V3S_H_
#define V3S_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* v3s_last_error(void);
int v3s_version(void);
/* CUDA kernels launched by the library since it was loaded (bench.py reports launches per step) */
long long v3s_kernel_launches(void);
int v3s_check_device(void);

/* fill n floats at dst, which may be a peer-mapped address on another GPU (peer-mapping check) */
int v3s_peer_fill(float* dst, float value, long long n, void* stream);

/* ---- stage 1: pattern synthesis ------------------------------------------------------- */

/* experiment constants of projection.py:96-107 (Projector.experiment_parameters) */
typedef struct v3s_geometry {
  double pixel;      /* 75e-6 m  */
  double zd;         /* 0.5 m    */
  double lambda;     /* 2e-9 m   */
  int    n_p_nobin;  /* 1024     */
  double grid_scale; /* 2e-7 m: Grid_3D = grid_scale * U, projection.py:91 */
} v3s_geometry;

void v3s_default_geometry(v3s_geometry* g /* host */);

/* Geometry of Projector.q_and_mask / camera_x_y / fourier_scaled_axis (projection.py:121-190),
 * once per (n, n_p, geometry): writes the per-pixel interpolation taps into the caller's device
 * workspace (v3s_synth_workspace_bytes(n_p) bytes) and the k_z plane range into meta_host[2].
 * geom: host pointer or NULL for the reference's constants.                                   */
long long v3s_synth_workspace_bytes(int n_p);
int v3s_synth_prepare(int n, int n_p, const v3s_geometry* geom /* host */, void* ws_dev, int* meta_host /* host [2] */,
                      void* stream);

/* Projector.generate's hot loop (projection.py:35-43) = N x Projector.generate_single_image
 * (projection.py:55-71) incl. Rotation.rotate_structure_index (rotation.py:143-162).
 *   ed    [n,n,n] FP32 object, array axes as in Projector.synsthesize_3d (projection.py:75-91)
 *   rot9  [N,9]   FP64, row-major 3x3 per pattern = R[:,c].reshape(3,3).T of projection.py:38
 *   ws_dev, z0, nz  from v3s_synth_prepare
 *   out   [N, n_p*n_p] FP32 patterns, row stride out_stride elements                        */
int v3s_synth_patterns(const float* ed, int n, const double* rot9, long long N, int n_p, const void* ws_dev, int z0,
                       int nz, float* out, long long out_stride, void* stream);

/* Direct mode (BASELINE north_star kernel 1; not the reference's numbers): the
 * analytical amplitude | sum_ijk F[i,j,k] exp(-2 pi i (R Pm kappa).(a_i,a_j,a_k)) | evaluated on
 * each orientation's rotated Ewald-sphere q-grid, without the reference's two trilinear
 * interpolations.  Same arguments as v3s_synth_patterns; ws_dev: v3s_synth_workspace_bytes(n_p). */
int v3s_synth_patterns_direct(const float* ed, int n, const double* rot9, long long N, int n_p,
                              const v3s_geometry* geom /* host */, void* ws_dev, float* out, long long out_stride,
                              void* stream);

/* Poisson photon noise, in place (BASELINE config 5; the reference has no noise model -- a
 * builder-defined input transform applied before stage 2).  Each pattern is scaled to
 * photons_per_pattern expected counts and every pixel replaced by a Poisson draw (Philox-4x32-10
 * keyed by seed, global pattern index row0_global + r, pixel): independent of the sharding.       */
int v3s_poisson_noise(float* img, long long n_rows, int P, long long stride, long long row0_global,
                      double photons_per_pattern, unsigned long long seed, void* stream);

/* ---- stage 2: distances + k nearest neighbours ----------------------------------------- */

/* DistanceMatrix.remove_offset (distance.py:70-71), first half: per-pixel sums over the local
 * rows (FP64).  The caller divides by the global N (after an all-reduce when sharded).      */
int v3s_column_sums(const float* img, long long n_rows, int P, long long stride, double* sums /* [P] */, void* stream);

/* remove_offset + DistanceMatrix.set_norm_to_one (distance.py:66-71) and the BF16 hi/lo split
 * of the Gram operands.  a_f32 [n_rows,P]; a_hi/a_lo [n_rows,Ppad] bf16 (NULL to skip);
 * norms [n_rows] FP64 and zero_flag [n_rows] (1 for an all-zero centred row) are optional.   */
int v3s_center_normalize(const float* img, long long n_rows, int P, long long stride, const double* mean /* [P] */,
                         float* a_f32, void* a_hi, void* a_lo, int Ppad, double* norms, int* zero_flag, void* stream);

/* v3s_center_normalize that also writes the FP16 operand of the fused Gram / top-k kernel (what v3s_split_f16
 * makes of a_f32: fl16(2^8 a), zero padded to Ppad columns, Ppad a multiple of 64 for the kernel) in the same pass. */
int v3s_center_normalize_f16(const float* img, long long n_rows, int P, long long stride, const double* mean,
                             float* a_f32, void* a_f16, int Ppad, int* zero_flag, void* stream);

/* BF16 hi/lo split of already centred unit rows a_f32 [n_rows,P] -> a_hi, a_lo [n_rows,Ppad]
 * (identical to what v3s_center_normalize emits; used on the all-gathered rows of a sharded job). */
int v3s_split_bf16(const float* a_f32, long long n_rows, int P, int Ppad, void* a_hi, void* a_lo, void* stream);

/* np.matmul(A, A.T) of DistanceMatrix.convert_to_round_distance (distance.py:56) for a row
 * panel: G[r, c] = <a_rows[r], a_all[c]>, r < n_rows, c < n_cols, tcgen05/TMEM/TMA kernel on
 * the split BF16 operands (FP32-equivalent).  Ppad must be a multiple of 64.                */
int v3s_gram_rows(const void* all_hi, const void* all_lo, long long n_cols, const void* rows_hi, const void* rows_lo,
                  long long n_rows, int Ppad, float* G, long long ldg, void* stream);

/* Pipeline shape of the tcgen05 Gram kernel: 32 = 4 stages x 48 kB, SWIZZLE_64B (default);
 * 64 = 2 stages x 96 kB, SWIZZLE_128B.  Same results to rounding.                                */
int v3s_set_gram_kblock(int bk);

/* The whole symmetric Gram matrix G [n, ldg] with half the tensor-core work: only the tiles
 * (column block of 128, row block of 256) listed in tile_list_xy (int pairs x = column block,
 * y = row block, those on or above the diagonal, in the caller's L2-friendly order) are computed;
 * each is stored as G[r][c] and mirrored to G[c][r].                                            */
int v3s_gram_symmetric(const void* a_hi, const void* a_lo, long long n, int Ppad, float* G, long long ldg,
                       const int* tile_list_xy, int num_tiles, void* stream);

/* Row-sharded symmetric Gram fused with its exchange (one kernel: tcgen05 tiles + NVLink peer
 * stores).  This rank computes the tiles in tile_list_xy (its share of the global upper-triangle
 * list; operands a_hi/a_lo are the all-gathered [n, Ppad] matrices) and stores every value twice
 * from the epilogue: G[r][c] into the row block of the rank owning r and the mirror G[c][r] into
 * the row block of the rank owning c.  peer_bases (host array [world]): base address of each
 * rank's [cnt_o, ldg] FP32 row block as mapped in THIS process (peer memory for o != rank, e.g.
 * torch.distributed._symmetric_memory buffer_ptrs); rank o owns cnt_base + (o < rem) rows.  The
 * caller synchronises the ranks before reading its block.                                       */
int v3s_gram_symmetric_sharded(const void* a_hi, const void* a_lo, long long n, int Ppad,
                               const unsigned long long* peer_bases /* host [world] */, int world, long long cnt_base,
                               int rem, long long ldg, const int* tile_list_xy, int num_tiles, void* stream);

/* Same contraction on CUDA cores in plain FP32 -- the validation path of v3s_gram_rows.       */
int v3s_gram_rows_simt(const float* a_all, long long n_cols, const float* a_rows, long long n_rows, int P, float* G,
                       long long ldg, void* stream);

/* DistanceMatrix.convert_to_round_distance (distance.py:51-63) as a dense FP64 matrix B [n,n]
 * from the centred unit rows a_f32 [n,P]: d = 2 asin(|a_r - a_c| / 2), exact all-pairs form.   */
int v3s_round_distance_dense(const float* a_f32, long long n, int P, const int* zero_flag, double* B, void* stream);

/* DistanceMatrix.knn_from_round_distance (distance.py:26-48) on a caller-supplied dense FP64
 * B [n,n]: k smallest entries per row, ascending.  key_ws float [n,n]; cand_ws int [n,min(k+margin,n)];
 * count_ws int [n].                                                                            */
int v3s_knn_from_dense(const double* B, long long n, int k, int margin, float* key_ws, int* cand_ws, int* count_ws,
                       double* S2, int* Nbr, void* stream);

/* arccos epilogue + DistanceMatrix.knn_from_round_distance (distance.py:56-59, 26-48) for the
 * panel rows [row0, row0+n_rows): S2 [n_rows,k] FP64 ascending distances, Nbr [n_rows,k] int32.
 * Candidates = every entry of the row within eps_g of its (k+margin)-th largest Gram value (a
 * value margin: with |G - G_exact| <= eps_g/2 the exact k nearest are provably among them), at
 * most `cap` per row (cand_ws int [n_rows,cap], count_ws int [n_rows]); *overflow_flag (device,
 * may be NULL) counts rows whose band exceeded cap and fell back to exactly k+margin candidates.
 * Their distances are recomputed from the FP32 rows as 2 asin(|a_r - a_c|/2) and sorted.
 * dist_ws (may be NULL): FP64 [n_rows,cap]; with it every unordered pair of rows of this call that
 * list each other is evaluated once instead of twice (same values bit for bit, ~1/3 fewer row reads).
 * Fails with the reference's message "Number of nearest neighbors to preserve is too high."
 * when k > n_cols.                                                                              */
int v3s_knn_from_gram(const float* G, long long ldg, long long n_rows, long long n_cols, long long row0,
                      const float* a_rows, const float* a_all, int P, int k, int margin, float eps_g, int cap,
                      const int* zero_flag, int* cand_ws, int* count_ws, int* overflow_flag, double* S2, int* Nbr,
                      double* dist_ws, void* stream);

/* ---- stage 2, fused form: the Gram matrix is never materialised --------------------------- */

/* FP16 Gram operand of centred unit rows a_f32 [n_rows,P]: a16 [n_rows,Ppad] = fl16(2^8 a), zero
 * padded (Ppad a multiple of 64).  One FP16 product ranks candidates; distances come from a_f32.  */
int v3s_split_f16(const float* a_f32, long long n_rows, int P, int Ppad, void* a16, void* stream);

/* cta_group of the fused kernel: 2 = CTA pairs, tcgen05.mma.cta_group::2 M=256 N=256 with each CTA
 * staging half of the shared column tile; 1 = one CTA per SM, M=128 N=256; 0 (default) = by size.   */
int v3s_set_gram_topk_cta_group(int cg);

/* SMs the persistent fused Gram kernels leave free (default 0): a collective running beside them (the all-gather of
 * the FP32 rows in a sharded job) needs a few, and the kernels' workers must all be resident at once.             */
int v3s_set_gram_reserved_sms(int sms);

/* Slack of the persistent Gram workers' per-tile lockstep (default 0: all workers start every tile together, so
 * that every K slice is read from HBM once; lag > 0: a worker may run up to `lag` tiles ahead of the slowest).   */
int v3s_set_gram_sync_lag(int lag);

/* buffer sizes of v3s_gram_topk for a problem: out[0] = S lists per row (column segments),
 * out[1] = C entries per list, out[2] = padded row count of `lists` / `counts`, out[3] = ints of `ws`. */
int v3s_gram_topk_plan(long long n_rows, long long n_cols, int k, int* out /* host [4] */);

/* np.matmul(A, A.T) (distance.py:56) FUSED with the selection half of
 * DistanceMatrix.knn_from_round_distance (distance.py:39-43): tcgen05/TMEM/TMA Gram tiles whose
 * epilogue keeps, per query row, a running threshold and a candidate list -- G [n_rows, n_cols] is
 * never written.  a16_all [n_cols,Ppad], a16_rows [n_rows,Ppad] from v3s_split_f16.  Output:
 * lists [rows_pad*S*C] 64-bit entries (FP32 bits of g~ << 32 | column) and counts [rows_pad*S];
 * the union of a row's S lists contains every column with g~ >= (k-th largest g~ of the row) -
 * 2 eps_g, eps_g = the caller's bound on |g~ - g| (2^-10 + accumulation for one FP16 product),
 * hence the row's exact k nearest.  *overflow (device, may be NULL) counts lists that saturated.
 * ws int [plan[3]] (zeroed by the call): the rows' shared thresholds and the per-tile arrival counters
 * that keep the persistent workers' operand streams aligned (each K slice read from HBM once).
 * Fails with "Number of nearest neighbors to preserve is too high." when k > n_cols.            */
int v3s_gram_topk(const void* a16_all, long long n_cols, const void* a16_rows, long long n_rows, int Ppad, int k,
                  float eps_g, unsigned long long* lists, int* counts, int* overflow, int* ws, void* stream);

/* The same selection with HALF the tensor-core work (G is symmetric: a tile on / above the diagonal serves its rows
 * and, mirrored, its columns).  Two calls:
 *  v3s_gram_topk_limits -- pass 1: v3s_gram_topk for the rank's rows against every `sample_stride`-th pattern only;
 *    lim_rows [n_rows] = (k-th largest g~ of that sample) - 2 eps_g, a valid lower limit of the row's candidate band
 *    (the k-th largest over a subset never exceeds the one over all patterns).  lists / counts / ws as for
 *    v3s_gram_topk with n_cols = ceil(n_cols / sample_stride) in v3s_gram_topk_plan.
 *  v3s_gram_topk_sym -- pass 2: the tiles with column tile >= row block of the n x n problem (this rank's share of
 *    them), one list per pattern: element (r, c), c >= r, is appended to list r when g~ >= lim[r] and to list c when
 *    c > r and g~ >= lim[c] (atomic slot counters; about sample_stride * k entries per list, capacity C; a counter
 *    past C is reported by v3s_knn_merge_lists as overflow).  lim [n] must be complete on every rank.  peer_lists /
 *    peer_counts (host arrays [world]): the lists [cnt_o, C] and zero-initialised counters [cnt_o] that receive the
 *    entries of rank o's patterns -- either rank o's own buffers as mapped in THIS process (NVLink peer memory for
 *    o != rank: compute and exchange are one kernel, then v3s_knn_merge_lists with S = 1), or LOCAL per-owner buffers
 *    (no remote atomics; every owner then reads its segments from the peers: v3s_knn_merge_segment_lists, the
 *    default); the caller synchronises the ranks in between.  Rank o owns cnt_base + (o < rem) consecutive patterns.
 *    step_ws: int [v3s_gram_topk_sym_steps(n, world)].                                                          */
int v3s_gram_topk_limits(const void* a16_all, long long n_cols, const void* a16_rows, long long n_rows, int Ppad, int k,
                         float eps_g, int sample_stride, unsigned long long* lists, int* counts, int* overflow, int* ws,
                         float* lim_rows, void* stream);
int v3s_gram_topk_sym_steps(long long n, int world);
int v3s_gram_topk_sym(const void* a16_all, long long n, int Ppad, const float* lim, int C,
                      const unsigned long long* peer_lists /* host [world] */,
                      const unsigned long long* peer_counts /* host [world] */, int world, int rank, long long cnt_base,
                      int rem, int* step_ws, void* stream);

/* unite the S lists of every row: candidates = entries within 2 eps_g of the k-th largest g~ of the
 * union -> cand_ws int [n_rows,cap], count_ws int [n_rows] (input of v3s_knn_refine).  Zero-norm
 * patterns (NaN -> distance 0, distance.py:58; zero_flag [n_cols] may be NULL): a zero query row
 * gets columns 0..cap-1, and the zero columns (compacted into zero_ws, int [cap + 1]) are appended to
 * every row.  *overflow counts rows cut at cap.
 * order_ws (may be NULL): int [2 n_rows + ceil(n_cols/anchor_stride) + ceil(n_rows/anchor_stride)]; its first
 * n_rows ints receive a processing ORDER of the rows in which rows sharing their nearest anchor pattern
 * (index a multiple of anchor_stride, nearest by g~ among the row's lists) are contiguous -- neighbouring
 * patterns, whose candidate lists overlap (input of v3s_knn_refine_tiled); every candidate list is then also
 * sorted by column index (v3s_knn_refine_tiled binary-searches them).  tiles_ws (may be NULL, with
 * order_ws): int [1 + 2 max_tiles] = the table of row tiles for v3s_knn_refine_tiled: [0] = number of tiles,
 * then (first position in `order`, rows <= tile_rows) per tile; a neighbourhood larger than tile_rows is cut,
 * smaller consecutive ones are packed.  max_tiles >= (bins of order_ws) + n_rows / tile_rows + 2 suffices.  */
int v3s_knn_merge_lists(const unsigned long long* lists, const int* counts, int S, int C, long long n_rows,
                        long long n_cols, long long row0, int k, float eps_g, int cap, const int* zero_flag,
                        int* zero_ws, int* cand_ws, int* count_ws, int* overflow, int anchor_stride, int* order_ws,
                        int tile_rows, int max_tiles, int* tiles_ws, void* stream);

/* v3s_knn_merge_lists with the S (<= 8) lists of a row given by a table instead of one contiguous array:
 * seg_lists[s] -> [n_rows, C] entries, seg_counts[s] -> [n_rows] counters (host arrays of device addresses).  The
 * sharded symmetric Gram (v3s_gram_topk_sym with peer_lists / peer_counts pointing at LOCAL per-owner buffers) leaves
 * one segment per SOURCE rank in that rank's memory; the owner of the rows merges them straight from the peers'
 * memory over NVLink (the only exchange of that stage, a few kB per row).  The union of a row's segments is held in
 * min(C, 24576) entries; more counts as overflow.                                                              */
int v3s_knn_merge_segment_lists(const unsigned long long* seg_lists /* host [S] */,
                                const unsigned long long* seg_counts /* host [S] */, int S, int C, long long n_rows,
                                long long n_cols, long long row0, int k, float eps_g, int cap, const int* zero_flag,
                                int* zero_ws, int* cand_ws, int* count_ws, int* overflow, int anchor_stride, int* order_ws,
                                int tile_rows, int max_tiles, int* tiles_ws, void* stream);

/* arccos epilogue (distance.py:56-59) + final ordering (distance.py:39-43) on selected candidates:
 * d = 2 asin(|a_r - a_c| / 2) from the FP32 rows, sorted by (d, index), first k kept -> S2 [n_rows,k]
 * FP64, Nbr [n_rows,k] int32.  dist_ws FP64 [n_rows,cap] (may be NULL): mutual pairs evaluated once.  order
 * (may be NULL, int [n_rows]): processing order of the rows, a cache-locality hint only.            */
int v3s_knn_refine(const float* a_rows, const float* a_all, int P, long long row0, long long n_rows, long long n_cols,
                   int k, int cap, const int* zero_flag, const int* cand_ws, const int* count_ws, double* S2, int* Nbr,
                   double* dist_ws, const int* order, void* stream);

/* v3s_knn_refine with the pair evaluation TILED over neighbouring rows: a tile of
 * v3s_knn_refine_tile_rows(cap) rows (consecutive in `order`) keeps its rows in shared memory and reads every
 * candidate row once for all the tile's rows that list it (the untiled form reads one candidate row per pair).
 * Same distances up to FP32 summation order (the pixel axis is summed in chunks of 512).  P % 4 == 0.
 * `tiles` / max_tiles: the tile table of v3s_knn_merge_lists.  Workspaces: dist_ws FP64 [n_rows,cap],
 * pref_ws int [n_rows,cap], keys_ws 64-bit [max_tiles * 16384], np_ws int [max_tiles].  dedup != 0: a pair
 * listed from both sides by rows of this call is evaluated once.                                    */
int v3s_knn_refine_tile_rows(int cap);
/* The two halves of v3s_knn_refine_tiled, for sharded jobs (the ranks synchronise in between):
 *  _pairs: |a_r - a_c|^2 of every pair THIS rank owns -> dist_ws (FP64 sums), pref_ws = slot of the partner row's
 *    list for pairs the partner owns.  cand_all [n_cols,cap] / count_all [n_cols] (may be NULL): the candidate lists
 *    of all patterns (all-gathered) -- a pair listed from both sides is then evaluated once even when its two rows
 *    belong to different ranks (owner: the lower index for indices of equal parity, else the higher).
 *  _emit: distances, ordering, first k -> S2 / Nbr; a sum owned by a partner row is read from that row's dist_ws:
 *    peer_dss (host array [world], may be NULL for one rank) = the ranks' dist_ws as mapped in this process
 *    (NVLink peer memory); rank o owns cnt_base + (o < rem) consecutive patterns.                              */
int v3s_knn_refine_tiled_pairs(const float* a_rows, const float* a_all, int P, long long row0, long long n_rows,
                               long long n_cols, int cap, const int* cand_ws, const int* count_ws, double* dist_ws,
                               const int* order, const int* tiles, int max_tiles, int* pref_ws,
                               unsigned long long* keys_ws, int* np_ws, int dedup, const int* cand_all,
                               const int* count_all, void* stream);
int v3s_knn_refine_tiled_emit(long long row0, long long n_rows, long long n_cols, int k, int cap, const int* zero_flag,
                              const int* cand_ws, const int* count_ws, double* S2, int* Nbr, double* dist_ws,
                              const int* pref_ws, const unsigned long long* peer_dss /* host [world] */, int world,
                              long long cnt_base, int rem, void* stream);
int v3s_knn_refine_tiled(const float* a_rows, const float* a_all, int P, long long row0, long long n_rows,
                         long long n_cols, int k, int cap, const int* zero_flag, const int* cand_ws,
                         const int* count_ws, double* S2, int* Nbr, double* dist_ws, const int* order, const int* tiles,
                         int max_tiles, int* pref_ws, unsigned long long* keys_ws, int* np_ws, int dedup, void* stream);

/* ---- stage 3: diffusion map ---------------------------------------------------------------- */

/* scalar part of DiffusionMapSolver.s2_to_w_matrix (diffusion.py:297-313), bug-compatible row
 * indexing.  scalars5 = {S2_Max, Dist_Thr, Epsilon*S2_Max, status, Index}; status 1 -> the
 * reference's 'S2_Max = 0', 2 -> index out of range.                                          */
int v3s_s2_scalars(const double* S2, long long N, int k, double eps, int med_row, int idx_off, double* scalars5,
                   void* stream);
/* in-place S2[S2 > Dist_Thr] = inf (diffusion.py:311) */
int v3s_s2_threshold(double* S2, long long N, int k, const double* scalars5, void* stream);

/* s2_to_w_matrix + anisotropic_norm + dm_cov (diffusion.py:313-325, 89-105, 244-260) fused on
 * the kNN lists: row block [row0,row0+n_rows) of P_ep as dense FP32 [n_rows, ldp] and
 * dinv_sqrt [N] (the diagonal of D).  ws_2N: FP64 workspace of 2N.                            */
int v3s_build_markov(const double* S2_thresholded, const int* Nbr, long long N, int k, const double* scalars5,
                     long long row0, long long n_rows, float* P_rows, long long ldp, double* dinv_sqrt, double* ws_2N,
                     void* stream);

/* dense FP64 forms of the same three functions, for callers of the individual reference APIs */
int v3s_w_dense(const double* S2_thresholded, const int* Nbr, long long N, int k, double eps_eff, double* W, void* stream);
int v3s_anisotropic_norm_dense(const double* W, long long N, double* out, double* ws_N, void* stream);
int v3s_dm_cov_dense(const double* W, long long N, double* P_out, double* dinv_sqrt, double* ws_N, void* stream);

/* dense P_ep @ X of the eigen-solver behind scipy.sparse.linalg.eigs (diffusion.py:61):
 * Y_rows [n_rows,b] = P_rows [n_rows,N] (FP32, ld ldp) . X [N,b] (FP64), b <= 8.               */
long long v3s_block_matvec_workspace_bytes(long long n_rows, long long N);
int v3s_block_matvec(const float* P_rows, long long ldp, long long n_rows, long long N, const double* X, int b,
                     double* Y_rows, void* workspace, void* stream);

/* the same product with the block already in FP32 "strip layout" (the form v3s_lanczos_step and
 * v3s_cheb_combine emit): v3s_xf_floats(N) floats, columns in strips of 128, each strip stored as
 * float4 [j/4][col%4][(col%128)/4]; padding columns of the last strip must be zero.            */
long long v3s_xf_floats(long long N);
int v3s_block_matvec_f32(const float* P_rows, long long ldp, long long n_rows, long long N, const float* Xf_strips,
                         double* Y_rows, void* workspace, void* stream);

/* Device-side block-Lanczos iteration, block size 8 -- the per-iteration work ARPACK does on the
 * host behind scipy.sparse.linalg.eigs (diffusion.py:61).  V [max_dim, ldv] FP64 basis (vectors as
 * rows), W [N,8] FP64 block, Xf FP32 strip-layout copy for the mat-vec (v3s_xf_floats(N) floats,
 * zero-initialised by the caller), ws: v3s_lanczos_workspace_doubles(N, max_dim) doubles.
 *   v3s_lanczos_start: orthonormalise the start block W in place; V[0..8) <- Q^T, Xf <- float(Q).
 *   v3s_lanczos_step : W = P Q_j on entry (m = basis size, Q_j = V[m-8..m)); two-pass block
 *       Gram-Schmidt against V[0..m), Cholesky-QR twice; on return W = Q_{j+1}, appended to V when
 *       m + 8 <= max_dim, Xf its FP32 copy, Tblk[128] = {A_j (8x8), B_j (8x8)} of the projected
 *       block-tridiagonal matrix; *status_out_dev != 0 after a Cholesky breakdown.                */
long long v3s_lanczos_workspace_doubles(long long N, int max_dim);
int v3s_lanczos_start(double* V, long long ldv, long long N, int max_dim, double* W, float* Xf, double* ws, void* stream);
int v3s_lanczos_step(double* V, long long ldv, long long N, int m, int max_dim, double* W, float* Xf, double* Tblk,
                     double* ws, int* status_out_dev, void* stream);

/* Chebyshev three-term recurrence on a block (polynomial acceleration of the Lanczos operator):
 * Y_out = alpha * PY + beta * Yk + gamma * Ykm1 (Ykm1 may be NULL; Y_out may alias it), FP64 [count],
 * count = 8 N, and Xf_out = float(Y_out) in strip layout for the next mat-vec (may be NULL).      */
int v3s_cheb_combine(const double* PY, const double* Yk, const double* Ykm1, double alpha, double beta, double gamma,
                     double* Y_out, float* Xf_out, long long count, void* stream);

/* Fused forms of the mat-vec inside the eigen-solver recurrences (diffusion.py:61), for one GPU or a
 * row-sharded job of up to 8 ranks on one NVLink domain.  The rank owns rows [row0,row0+n_rows) of
 * P.  After the mat-vec kernel, ONE kernel reduces the rank's rows, forms
 *     v = alpha * (P X) + beta * Yk + gamma * Ykm1        (FP64; Yk / Ykm1 [N,8] may be NULL)
 * and stores v into the [N,8] result block of EVERY rank (yout_ptrs[r], peer-mapped addresses) and,
 * when xfout_ptrs != NULL, float(v) into every rank's strip-layout operand of the next mat-vec:
 * the all-gather of the Lanczos block is part of the kernel's stores, not a collective.  When
 * sharded, the block of that kernel that finishes last signals every peer's flag word
 * (flag_ptrs[r] = rank r's int[world] array in peer-mapped memory, zero-initialised; epochs grow
 * monotonically) and waits for every peer's signal, so the stream only moves on when the whole
 * block has arrived; ticket_dev = a zero-initialised word of the rank's own memory;
 * *status_dev becomes 3 if a peer never arrives.  v3s_peer_barrier is the same barrier stand-alone. */
int v3s_matvec_combine(const float* P_rows, long long ldp, long long n_rows, long long N, long long row0,
                       const float* Xf_strips, const double* Yk, const double* Ykm1, double alpha, double beta,
                       double gamma, const unsigned long long* yout_ptrs, const unsigned long long* xfout_ptrs,
                       const unsigned long long* flag_ptrs, int rank, int world, int epoch, unsigned* ticket_dev,
                       int* status_dev, void* workspace, void* stream);
int v3s_peer_barrier(const unsigned long long* flag_ptrs, int rank, int world, int epoch, int* status_dev, void* stream);
/* T_degree((P - c)/e) applied to the block Q [N,8] by one host call (degree x {mat-vec, fused
 * combine / scatter / barrier}).  ybuf_ptrs[b*world + r]: rotating FP64 result buffer b = 0..2 on rank
 * r; xf_ptrs[p*world + r]: ping-pong FP32 strip operand p = 0..1 on rank r; Xf0 = float(Q) in strip
 * layout; q_index = which rotating buffer Q is (-1: a separate block); the barriers use epochs
 * epoch0+1 .. epoch0+degree.  *result_index (host) = rotating buffer holding the result.       */
int v3s_cheb_filter(const float* P_rows, long long ldp, long long n_rows, long long N, long long row0, const float* Xf0,
                    const double* Q, int q_index, int degree, double c, double e, const unsigned long long* ybuf_ptrs,
                    const unsigned long long* xf_ptrs, const unsigned long long* flag_ptrs, int rank, int world,
                    int epoch0, unsigned* ticket_dev, int* status_dev, void* workspace, int* result_index, void* stream);
/* ---- sparse form of the same operator -----------------------------------------------------------
 * P_ep has at most k + in-degree(i) non-zeros in row i (the reference itself assembles W through
 * scipy.sparse.csr_array, diffusion.py:317, and only then densifies it): P = A + A^T, A = the kNN
 * lists with the values h_ij of v3s_build_markov.  FP64 values and FP64 operand, so the operator
 * carries no FP32 rounding at all; per mat-vec it moves 76 B per non-zero instead of 4 N B per row,
 * which is what makes BASELINE config 5 (N = 2 10^5: 160 GB dense) affordable.  Device pointers;
 * the descriptor itself lives on the host.                                                       */
typedef struct v3s_sparse_markov {
  const int* nbr;        /* [N,k]   column of entry (i,t): the kNN list                             */
  const double* a_val;   /* [N,k]   h_it (0 for thresholded entries)                                */
  const int* t_ptr;      /* [N+1]   row pointers of the transposed lists                            */
  const int* t_col;      /* [nnz]   source rows, ascending inside a row                             */
  const double* t_val;   /* [nnz]                                                                   */
  int k;
  /* optional locality hint (NULL / 0: natural order): a permutation of the rows order_row0 .. order_row0 +
   * order_rows - 1 (as offsets 0 .. order_rows-1) in which neighbouring patterns are contiguous (the `order` of
   * v3s_knn_merge_lists).  A mat-vec over exactly that row block walks its rows in this order, one contiguous
   * chunk per CTA, so that rows processed together gather the same rows of X (their neighbour lists overlap) and
   * hit in L1.  Results do not depend on it (fixed summation order per row).                                */
  const int* row_order;
  long long order_row0, order_rows;
} v3s_sparse_markov;
/* s2_to_w_matrix + anisotropic_norm + dm_cov on the lists (as v3s_build_markov) -> the arrays of the
 * descriptor (caller-allocated: a_val [N k], t_ptr [N+1], t_col / t_val [N k]) and dinv_sqrt [N].
 * ws_2N: 2N doubles; iws: N + N k ints.  Replicated on every rank of a sharded job (O(N k)).     */
int v3s_build_markov_sparse(const double* S2_thresholded, const int* Nbr, long long N, int k, const double* scalars5,
                            double* a_val, int* t_ptr, int* t_col, double* t_val, double* dinv_sqrt, double* ws_2N,
                            int* iws, void* stream);
/* Y_rows [n_rows,b] = P[row0 .. row0+n_rows, :] . X [N,b] (FP64), b <= 8 */
int v3s_sparse_block_matvec(const v3s_sparse_markov* op, long long row0, long long n_rows, long long N, const double* X,
                            int b, double* Y_rows, void* stream);
/* v3s_matvec_combine / v3s_cheb_filter with the sparse operator: X (Q) is the FP64 block [N,8] itself,
 * no FP32 strip operands are read or (unless xfout_ptrs is given) written; workspace: 8 n_rows doubles. */
int v3s_sparse_matvec_combine(const v3s_sparse_markov* op, long long n_rows, long long N, long long row0, const double* X,
                              const double* Yk, const double* Ykm1, double alpha, double beta, double gamma,
                              const unsigned long long* yout_ptrs, const unsigned long long* xfout_ptrs,
                              const unsigned long long* flag_ptrs, int rank, int world, int epoch, unsigned* ticket_dev,
                              int* status_dev, void* workspace, void* stream);
int v3s_sparse_cheb_filter(const v3s_sparse_markov* op, long long n_rows, long long N, long long row0, const double* Q,
                           int q_index, int degree, double c, double e, const unsigned long long* ybuf_ptrs,
                           const unsigned long long* flag_ptrs, int rank, int world, int epoch0, unsigned* ticket_dev,
                           int* status_dev, void* workspace, int* result_index, void* stream);
/* ---- the whole eigen-solver behind one call -----------------------------------------------------
 * DiffusionMapSolver.diffusion_coordinate's `eigs(P_ep, k)` (diffusion.py:61; ARPACK, 'LM'): the nev
 * largest-magnitude eigenpairs of the symmetric P_ep by Chebyshev-filtered block Lanczos on the kernels
 * above, driven from the library (no Python, no LAPACK).  UNLIKE every other entry point it synchronises
 * the stream (a few small device-to-host copies per Ritz check) and allocates transient host memory.
 * The operator is the rank's row block in either form plus the buffers the fused mat-vecs store into
 * (one arena per job, as for v3s_cheb_filter): ybuf_ptrs [3*world] rotating [N,8] FP64 result blocks,
 * xf_ptrs [2*world] FP32 strip operands (dense form only), flag_ptrs [world] (sharded only), ticket /
 * status words, the mat-vec workspace (dense: v3s_block_matvec_workspace_bytes; sparse: 8 n_rows doubles);
 * `epoch` is the arena's barrier counter, advanced by the call.  Host struct, device pointers inside.   */
typedef struct v3s_eig_problem {
  const float* P_rows;                  /* dense FP32 rows [n_rows, ldp], or NULL                     */
  long long ldp;
  const v3s_sparse_markov* sparse;      /* or the sparse operator (host descriptor), or NULL          */
  long long n_rows, N, row0;
  int rank, world;
  const unsigned long long* ybuf_ptrs;  /* host [3*world]                                             */
  const unsigned long long* xf_ptrs;    /* host [2*world], dense form                                 */
  const unsigned long long* flag_ptrs;  /* host [world], world > 1                                    */
  unsigned* ticket_dev;
  int* status_dev;
  void* mv_workspace;
  int epoch;                            /* in / out                                                   */
} v3s_eig_problem;
/* ws: v3s_eig_topk_workspace_doubles(N, max_dim) doubles (device).  max_dim = Krylov cap (multiple of 8,
 * e.g. 320), degree = Chebyshev degree (6; 0 = plain Lanczos; the filter is used from N = 2048 on),
 * probe_steps = plain steps before the filter (10), seed = start block.  Outputs: lambda_host [nev]
 * (host, by decreasing magnitude), Y_dev [N,nev] FP64 row-major orthonormal eigenvectors (device, complete
 * on every rank), resid_host [nev] (may be NULL) residual norms, info_host [8] = {converged, mat-vecs,
 * Krylov dimension, 1 if filtered, Ritz checks, 1 if an eigenvalue cluster of multiplicity >= 8 may hide
 * further copies (disconnected graph: deflate and rerun, as v3s/eig.py's topk_with_locking does), 0, 0}.
 * Fails with the reference's "No spare-matrix for eigs-calculation" when nev >= N - 1.                  */
long long v3s_eig_topk_workspace_doubles(long long N, int max_dim);
int v3s_eig_topk(v3s_eig_problem* problem /* host */, int nev, double tol, int max_dim, int degree, int probe_steps,
                 unsigned long long seed, double* ws, double* lambda_host, double* Y_dev, double* resid_host,
                 int* info_host, void* stream);
/* the host eigen-solver of the small projected matrices (Householder tridiagonalisation + implicit QL): a [n,n]
 * symmetric row-major, overwritten with the eigenvectors in its columns; w [n] ascending.  Host pointers.    */
int v3s_sym_eig_host(double* a, int n, double* w);

/* Small device -> host read (<= 256 kB) that uses no copy engine: a one-block kernel stores the bytes into a pinned,
 * device-mapped mailbox, the stream is synchronised, the bytes are copied to host_dst.  A cudaMemcpy D2H would queue
 * behind any bulk transfer running on the device-to-host engine (solve() ships 13 GB of patterns while stages 2-4
 * run); the eigen-solver's Ritz checks and the stage API's validation scalars read through this instead.        */
int v3s_read_small(const void* dev_src, void* host_dst, long long bytes, void* stream);

/* per-launch device time of the mat-vec kernel, dense or sparse (CUDA events owned by the library, recorded on the
 * launching stream): enable/reset, then read the elapsed milliseconds (returns the launch count). */
int v3s_profile_matvec(int enable);
int v3s_profile_matvec_read(float* ms_out, int cap);
/* milliseconds between the end of recorded launch i and the start of launch i+1 (the fused reduction / exchange kernel
 * with its cross-rank barrier, Lanczos steps, host gaps); call before v3s_profile_matvec_read; returns the gap count. */
int v3s_profile_matvec_gaps(float* ms_out, int cap);

/* ---- stage 4: orientation retrieval ---------------------------------------------------------- */

/* ComputeUtils.lsq_linear (compute.py:103-112): out162 = {X^T X (81), X^T Y (81)} partial sums
 * over the local rows (all-reduce when sharded), X = Ps[:,1:10], Y = Rotation_R.T [n_rows,9].  */
int v3s_fit_gram(const double* Ps, int ldx, const double* Y, long long n_rows, double* out162, void* stream);
/* c [9,9] = pinv(X^T X) (X^T Y) */
int v3s_fit_solve(const double* in162, double* c81, void* stream);
/* Recon = X c, ComputeUtils.polar_decompose per row (compute.py:51-55), Rotation.rot_mat_to_quat
 * for estimate and truth (diffusion.py:144-167, rotation.py:165-210).  Outputs (optional ones may
 * be NULL): recon_pre [n,9], rproj [n,9], quat_est [n,4], quat_true [n,4].
 * polar_mode 0: U.Vh^T (reference form), 1: U.Vh, 2: nearest rotation.                          */
int v3s_fit_project(const double* Ps, int ldx, const double* Y, long long n_rows, const double* c81, int polar_mode,
                    double* recon_pre, double* rproj, double* quat_est, double* quat_true, void* stream);
/* DiffusionMapSolver.or_func (diffusion.py:265-287; src/octave/CalculateDiffusion.m:136-157): the
 * rotation-matrix constraint residuals of the fit without known orientations, G [n_rows], for
 * C = c81 (row-major 9x9) and psi_l = Ps[l,1:10]; r_total = the r of the 1/sqrt(r) factor (N when
 * sharded); form 0 = the Python expression, 1 = the Octave expression.  J (may be NULL) [n_rows,81]
 * = forward-difference Jacobian with steps rel_step * max(1,|c_p|) (scipy least_squares '2-point'). */
int v3s_or_func(const double* c81, const double* Ps, int ldx, long long n_rows, long long r_total, int form,
                double rel_step, double* G, double* J, void* stream);
/* Rotation.mean_geodesic_distance (rotation.py:228-244) inner sum over rows [row0,row0+n_rows)
 * against all N columns; partials_ws: 65536 doubles; out_sum[0] = sum |acos - acos|.  FP64 like the
 * reference up to N = 20000; beyond, the N^2 pair terms are FP32 (acosf) summed in FP64 (~1e-7 relative). */
int v3s_sigma_pairs(const double* quat_est, const double* quat_true, long long N, long long row0, long long n_rows,
                    double* partials_ws, double* out_sum, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V3S_H_ */
