"""Device operations of the hot path: thin torch-tensor wrappers over the C-ABI (include/v3s.h).

Every method takes and returns CUDA tensors on the current device/stream.  This is the only
compute backend of the product; `CudaOps()` raises without a B200 (no CPU fallback).  The host
logic above it (pipeline.py, eig.py) is written against this small interface so that the sharding
and solver logic can be exercised on CPU in tests with a stand-in backend.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

GRAM_KPAD = 64          # K-block of the tcgen05 Gram kernel (bf16 elements)
KNN_MARGIN = 8          # extra candidates (by count) kept through the exact refinement
KNN_EXTRA_CAP = 160     # room per row for the value margin (entries within eps_g of the (k+margin)-th)


def gram_error_bound(P):
    """|G_tcgen05 - G_exact| for unit rows of length P: dropped lo.lo^T term (2^-18) + the tensor
    core's truncating FP32 accumulate over 3P/16 steps (validated in test_gram_tcgen05_vs_fp64)."""
    return 1e-5 + 3.0e-8 * (3.0 * P / 16.0)


def fused_gram_error_bound(P):
    """|g~ - g| of the fused Gram / top-k kernel for unit rows of length P (gram_topk.cu): one FP16
    product, (2^-10 + 2^-22) sum|a_i b_i| <= 2^-10 by Cauchy-Schwarz, + the FP32 accumulation of the
    P/16 tensor-core steps (<= 2^-23 each) + FP16 subnormal flushes of 2^8-scaled entries."""
    return 2.0 ** -10 * (1.0 + 2.0 ** -11) + 1.2e-7 * (P / 16.0) + 2e-6


GRAM_SYNC_LAG = 0       # slack (in tiles) of the persistent Gram workers' lockstep (v3s_set_gram_sync_lag)
SYM_SAMPLE_STRIDE = 8   # symmetric fused Gram: pass 1 looks at every 8th pattern to fix the rows' lower limits
PANEL_BYTES = 4 << 30   # Gram row panel budget when G is streamed in panels
SYM_BYTES = 48 << 30    # the whole symmetric G is materialised (half the MMA work) up to this size


DENSE_OPERATOR_BYTES = 12 << 30   # operator "auto": dense FP32 row block up to this size per rank, sparse beyond


def round_up(x, m):
    return (x + m - 1) // m * m


class SparseMarkov:
    """Sparse form of P_ep (v3s_sparse_markov): P = A + A^T on the kNN lists, FP64.  Holds every
    rank's copy of the whole operator (O(N k)); a rank multiplies its own rows [r0, r1)."""

    def __init__(self, nbr, a_val, t_ptr, t_col, t_val, r0, r1, row_order=None):
        self.nbr, self.a_val, self.t_ptr, self.t_col, self.t_val = nbr, a_val, t_ptr, t_col, t_val
        self.N, self.k = int(nbr.shape[0]), int(nbr.shape[1])
        self.r0, self.r1 = int(r0), int(r1)
        self.device = nbr.device
        # row_order (optional): (row0, rows, int32 permutation) -- the neighbourhood order of the kNN stage for this
        # rank's row block; a locality hint of the mat-vec kernel, never part of the result
        self.row_order = row_order
        ro = row_order if (row_order is not None and row_order[0] == self.r0 and row_order[1] == self.r1 - self.r0) else None
        self.desc = _lib.SparseMarkovDesc(nbr.data_ptr(), a_val.data_ptr(), t_ptr.data_ptr(), t_col.data_ptr(),
                                          t_val.data_ptr(), self.k, ro[2].data_ptr() if ro else None,
                                          ro[0] if ro else 0, ro[1] if ro else 0)

    @property
    def n_rows(self):
        return self.r1 - self.r0

    def nnz(self):
        """Stored entries of the rank's rows (list part incl. thresholded zeros + transposed part)."""
        tp = self.t_ptr[[self.r0, self.r1]].cpu()
        return self.n_rows * self.k + int(tp[1] - tp[0])

    def rows(self, r0, r1):
        """The same operator restricted to another row block (shares the arrays)."""
        return SparseMarkov(self.nbr, self.a_val, self.t_ptr, self.t_col, self.t_val, r0, r1, self.row_order)

    def to_dense(self):
        """[n_rows, N] FP64 (tests / tiny problems)."""
        N, k = self.N, self.k
        P = torch.zeros((N, N), dtype=torch.float64, device=self.device)
        rows = torch.arange(N, device=self.device).repeat_interleave(k)
        P.index_put_((rows, self.nbr.reshape(-1).long()), self.a_val.reshape(-1), accumulate=True)
        nnz = int(self.t_ptr[-1].item())
        cnt = (self.t_ptr[1:] - self.t_ptr[:-1]).long()
        trow = torch.arange(N, device=self.device).repeat_interleave(cnt)
        P.index_put_((trow, self.t_col[:nnz].long()), self.t_val[:nnz], accumulate=True)
        return P[self.r0:self.r1]


class _StreamEvents:
    """wait(): the current stream waits for the recorded events (transfers issued on side streams)."""
    uses_sms = False

    def __init__(self, events):
        self.events = events

    def wait(self):
        cur = torch.cuda.current_stream()
        for ev in self.events:
            cur.wait_event(ev)


class CudaOps:
    name = "cuda-sm100a"

    def __init__(self, device=None):
        _lib.require_device()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._mv_ws = None
        self._sigma_ws = None
        # "fused" = tcgen05 Gram tiles with the top-k selection in their epilogue (G never materialised; default),
        # "tcgen05" = materialised BF16x3 Gram panels + radix select, "simt" = FP32 CUDA-core validation path
        self.gram_impl = "fused"
        self.knn_overflow = None
        self._profile = False        # when True, CUDA events bracket the named kernels (bench.py)
        self._events = []
        self._synth_prep = {}
        self._sym_tile_cache = {}
        self._sym_fits = {}
        self._snap_events = []
        self._symm_G = {}
        self.dedup_refine = True     # exact-distance refinement: mutual candidate pairs once (select.cu)
        self.tiled_refine = True     # ... tiled over neighbouring rows for long rows (refine_tiled_kernel)
        self.symmetric_gram = True   # fused Gram / top-k on the upper-triangle tiles only (gram_sym_topk_kernel) for large N
        # ... in a sharded job too: each rank a share of the tiles, appends into LOCAL per-owner lists, every owner merges
        # its patterns' segments from the peers' memory over NVLink (remote slot-counter atomics from inside the
        # kernel's epilogue cost more than the halved MMA work saved)
        self.symmetric_gram_sharded = True
        import os as _os
        # sharded jobs: operand all-gathers as copy-engine pulls from symmetric memory (V3S_CE_GATHER=0: NCCL all-gathers)
        self.copy_engine_gather = _os.environ.get("V3S_CE_GATHER", "1") != "0"
        lag = int(_os.environ.get("V3S_GRAM_SYNC_LAG", str(GRAM_SYNC_LAG)))
        if lag:
            call("v3s_set_gram_sync_lag", lag)
        self._row_order = None       # (row0, rows, order int32 [rows], N) of the last fused kNN call with a tile order
        self.sparse_row_order = True # sparse mat-vec: walk the rows in the kNN stage's neighbourhood order (L1 reuse)
        self.share_pairs_across_ranks = True   # sharded tiled refinement: mutual pairs once per JOB, not once per rank
        self.overlap_reserved_sms = 16   # SMs left to NCCL while the all-gather of the FP32 rows overlaps the Gram kernel
        self.use_fused_matvec = True # eigen-solver mat-vecs with the reduction / recurrence / exchange fused (FusedOperator)
        self.use_peer_gram = True    # sharded jobs: symmetric Gram with NVLink peer stores (else rectangular panels)
        self.operator = "auto"       # Markov operator of the eigen-solver: "dense" (FP32 rows), "sparse" (FP64 lists), "auto"
        # "cabi" (default): the whole eigen-solver through ONE C-ABI call, v3s_eig_topk (library-driven: no Python, no
        # LAPACK in the loop); "python": the host-driven loop of v3s/eig.py (also the fallback when the C-ABI solver
        # reaches its Krylov cap, and the driver of the deflation rounds of disconnected graphs)
        self.eig_driver = "cabi"
        self.symm_error = None
        self._pin = None

    # ---- per-kernel device timing (CUDA events on the launching stream) ------------------------
    @property
    def profile(self):
        return self._profile

    @profile.setter
    def profile(self, on):
        # the mat-vec kernel is launched from inside multi-kernel C calls: its events are recorded by the library
        self._profile = bool(on)
        _lib.lib.v3s_profile_matvec(1 if on else 0)

    def _tic(self, name):
        if not self.profile:
            return None
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        return (name, e0, e1)

    def _toc(self, h):
        if h is not None:
            h[2].record()
            self._events.append(h)

    def pop_timings(self):
        """-> {name: [ms per launch, ...]} for the launches since the last call (synchronises)."""
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self._events:
            out.setdefault(name, []).append(e0.elapsed_time(e1))
        self._events = []
        if self._profile:
            cap = 1 << 16
            buf = (C.c_float * cap)()
            ng = min(int(_lib.lib.v3s_profile_matvec_gaps(buf, cap)), cap)
            if ng:
                out["matvec_gap"] = [float(buf[i]) for i in range(ng)]
            n = min(int(_lib.lib.v3s_profile_matvec_read(buf, cap)), cap)
            if n:
                out["matvec"] = [float(buf[i]) for i in range(n)]
        return out

    def read_small(self, t):
        """Small device tensor -> numpy WITHOUT the device-to-host copy engine (v3s_read_small: pinned mapped mailbox).
        `.cpu()` / `.item()` would queue behind a bulk D2H transfer in flight (solve() ships the pattern matrix while
        the later stages run) and stall the launching thread for its whole remainder."""
        import numpy as np
        t = t.contiguous()
        out = np.empty(tuple(t.shape), dtype=torch.empty(0, dtype=t.dtype).numpy().dtype)
        call("v3s_read_small", ptr(t), C.c_void_p(out.ctypes.data), t.numel() * t.element_size(), stream_ptr())
        return out

    # ---- allocation helpers ---------------------------------------------------------------
    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype, device=self.device)

    def to_device(self, x, dtype=None):
        t = torch.as_tensor(x)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=True).contiguous()

    # ---- stage 1 -----------------------------------------------------------------------------
    def synth_patterns(self, ed, rot9, n_p, geom=None, mode="exact"):
        """ed [n,n,n] f32, rot9 [N,9] f64 (row-major Rm) -> patterns [N, n_p^2] f32.
        mode "exact" = the reference's rotate -> FFT -> interpolate chain; "direct" = analytical
        amplitude on the rotated q-grid (north_star kernel 1; proper rotations only)."""
        n = ed.shape[0]
        N = rot9.shape[0]
        out = self.empty((N, n_p * n_p), torch.float32)
        if mode == "direct":
            ws = self.empty((int(_lib.lib.v3s_synth_workspace_bytes(n_p)),), torch.uint8)
            h = self._tic("synth_direct")
            call("v3s_synth_patterns_direct", ptr(ed), n, ptr(rot9), N, n_p, None, ptr(ws), ptr(out), n_p * n_p, stream_ptr())
            self._toc(h)
            return out
        key = (n, n_p, tuple(sorted(geom.items())) if geom else None)
        prep = self._synth_prep.get(key)
        if prep is None:                       # geometry set-up once per (n, n_p, geometry)
            g = None
            if geom is not None:
                g = _lib.Geometry()
                _lib.lib.v3s_default_geometry(C.byref(g))
                for k_, v_ in geom.items():
                    setattr(g, "lambda_" if k_ == "lambda" else k_, v_)
            ws = self.empty((int(_lib.lib.v3s_synth_workspace_bytes(n_p)),), torch.uint8)
            meta = (C.c_int * 2)()
            call("v3s_synth_prepare", n, n_p, C.byref(g) if g is not None else None, ptr(ws), meta, stream_ptr())
            prep = (ws, int(meta[0]), int(meta[1]))
            self._synth_prep[key] = prep
        h = self._tic("synth")
        call("v3s_synth_patterns", ptr(ed), n, ptr(rot9), N, n_p, ptr(prep[0]), prep[1], prep[2], ptr(out), n_p * n_p,
             stream_ptr())
        self._toc(h)
        return out

    def poisson_noise_(self, img, photons_per_pattern, seed=0, row0_global=0):
        """In-place Poisson photon noise on patterns [n,P] f32 (BASELINE config 5)."""
        call("v3s_poisson_noise", ptr(img), img.shape[0], img.shape[1], img.stride(0), int(row0_global),
             float(photons_per_pattern), int(seed), stream_ptr())
        return img

    # ---- stage 2 -----------------------------------------------------------------------------
    def column_sums(self, img):
        sums = self.empty((img.shape[1],), torch.float64)
        call("v3s_column_sums", ptr(img), img.shape[0], img.shape[1], img.stride(0), ptr(sums), stream_ptr())
        return sums

    def center_normalize(self, img, mean, want_split=True, out_a32=None):
        """-> a32 [n,P] f32, hi, lo [n,Ppad] bf16 (or None), zero_flag [n] i32."""
        n, P = img.shape
        Ppad = round_up(P, GRAM_KPAD)
        a32 = out_a32 if out_a32 is not None else self.empty((n, P), torch.float32)
        hi = self.empty((n, Ppad), torch.bfloat16) if want_split else None
        lo = self.empty((n, Ppad), torch.bfloat16) if want_split else None
        zf = self.empty((n,), torch.int32)
        h = self._tic("center_normalize")
        call("v3s_center_normalize", ptr(img), n, P, img.stride(0), ptr(mean), ptr(a32), ptr(hi), ptr(lo), Ppad, None,
             ptr(zf), stream_ptr())
        self._toc(h)
        return a32, hi, lo, zf

    def center_normalize_f16(self, img, mean, out_a32=None, out_a16=None):
        """center_normalize + split_f16 in one pass over the patterns -> a32 [n,P] f32, a16 [n,Ppad] f16, zero_flag [n]."""
        n, P = img.shape
        Ppad = round_up(P, GRAM_KPAD)
        a32 = out_a32 if out_a32 is not None else self.empty((n, P), torch.float32)
        a16 = out_a16 if out_a16 is not None else self.empty((n, Ppad), torch.float16)
        zf = self.empty((n,), torch.int32)
        h = self._tic("center_normalize")
        call("v3s_center_normalize_f16", ptr(img), n, P, img.stride(0), ptr(mean), ptr(a32), ptr(a16), Ppad, ptr(zf), stream_ptr())
        self._toc(h)
        return a32, a16, zf

    def split_bf16(self, a32):
        """BF16 hi/lo Gram operands of centred unit rows a32 [n,P] (same values as center_normalize's)."""
        n, P = a32.shape
        Ppad = round_up(P, GRAM_KPAD)
        hi = self.empty((n, Ppad), torch.bfloat16)
        lo = self.empty((n, Ppad), torch.bfloat16)
        call("v3s_split_bf16", ptr(a32), n, P, Ppad, ptr(hi), ptr(lo), stream_ptr())
        return hi, lo

    def split_f16(self, a32, out=None):
        """FP16 operand fl16(2^8 a) [n,Ppad] of the fused Gram / top-k kernel."""
        n, P = a32.shape
        Ppad = round_up(P, GRAM_KPAD)
        a16 = out if out is not None else self.empty((n, Ppad), torch.float16)
        call("v3s_split_f16", ptr(a32), n, P, Ppad, ptr(a16), stream_ptr())
        return a16

    def sym_stride(self):
        import os as _os
        return int(_os.environ.get('V3S_SYM_STRIDE', SYM_SAMPLE_STRIDE))

    def wants_symmetric_gram(self, N, nr, k, plan=None, whole=True):
        """Whether knn_rows_fused takes the symmetric form (half the MMA work + a sampling pass) for this problem."""
        stride = self.sym_stride()
        world = 1 if plan is None else plan.world
        return bool(self.symmetric_gram and whole and N >= 16384 and (N + stride - 1) // stride >= 4 * k and world <= 8
                    and (world == 1 or (self.symmetric_gram_sharded and self.use_peer_gram and min(plan.counts) >= 1)))

    def knn_rows_fused(self, a16_all, a32_all, zero_flag, r0, r1, k, order=None, plan=None, a32_ready=None,
                       a16_ready=None, a16_rows=None, a16_sample=None):
        """k nearest neighbours of rows [r0,r1) with the selection fused into the Gram kernel's epilogue:
        v3s_gram_topk (candidate lists per row and column segment) -> v3s_knn_merge_lists (value band around
        the k-th largest) -> v3s_knn_refine (exact distances from the FP32 rows, sort)."""
        N, P = a32_all.shape
        nr = r1 - r0
        S2 = self.empty((nr, k), torch.float64)
        Nbr = self.empty((nr, k), torch.int32)
        if nr == 0:
            return S2, Nbr
        self.knn_overflow = self.zeros((1,), torch.int32)
        eps_g = fused_gram_error_bound(P)
        if a32_ready is not None and getattr(a32_ready, "uses_sms", True):
            call("v3s_set_gram_reserved_sms", self.overlap_reserved_sms)   # room for the NCCL kernels running beside the Gram kernel
        stride = self.sym_stride()
        # symmetric form (half the MMA work + a 1/stride sampling pass) for large problems: the whole matrix on one
        # GPU, or the rank's share of the upper-triangle tiles with per-owner lists merged over NVLink
        world = 1 if plan is None else plan.world
        whole = (r0 == 0 and nr == N) if world == 1 else (r0 == plan.lo and r1 == plan.hi)
        sym = self.wants_symmetric_gram(N, nr, k, plan, whole)
        if a16_rows is None:
            a16_rows = a16_all[r0:r1]
        if a16_ready is not None and not (sym and a16_sample is not None):
            a16_ready.wait()
            a16_ready = None
        Csym = 2048
        while Csym < 1.5 * stride * k + 256:
            Csym *= 2
        peer = None
        if sym and world > 1:
            peer = self._sym_lists_arena(plan, Csym)
            sym = peer is not None
            if not sym and a16_ready is not None:      # rectangular kernel after all: it needs the whole operand now
                a16_ready.wait()
                a16_ready = None
        pl = (C.c_int * 4)()
        h = self._tic("gram")
        if sym:
            # pass 1 against a sample of the patterns: every stride-th row of a16_all, or (sharded) the already gathered
            # compact sample a16_sample while the all-gather of the full operand is still running
            ns = (N + stride - 1) // stride if a16_sample is None else int(a16_sample.shape[0])
            call("v3s_gram_topk_plan", nr, ns, k, pl)
            S1, C1, rows_pad1 = int(pl[0]), int(pl[1]), int(pl[2])
            lists1 = self.empty((rows_pad1 * S1 * C1,), torch.int64)
            counts1 = self.empty((rows_pad1 * S1,), torch.int32)
            tws1 = self.empty((int(pl[3]),), torch.int32)
            lim_loc = self.empty((nr,), torch.float32)
            h1 = self._tic("gram_limits_pass")
            if a16_sample is None:
                call("v3s_gram_topk_limits", ptr(a16_all), N, ptr(a16_rows), nr, a16_all.shape[1], k, eps_g, stride, ptr(lists1),
                     ptr(counts1), ptr(self.knn_overflow), ptr(tws1), ptr(lim_loc), stream_ptr())
            else:
                call("v3s_gram_topk_limits", ptr(a16_sample), ns, ptr(a16_rows), nr, a16_sample.shape[1], k, eps_g, 1, ptr(lists1),
                     ptr(counts1), ptr(self.knn_overflow), ptr(tws1), ptr(lim_loc), stream_ptr())
            self._toc(h1)
            if a16_ready is not None:
                a16_ready.wait()                       # the full FP16 operand (gathered beside pass 1) is needed from here on
            S, Cc = 1, Csym
            step_ws = self.empty((int(_lib.lib.v3s_gram_topk_sym_steps(N, world)),), torch.int32)
            if world == 1:
                lists = self.empty((N * Cc,), torch.int64)
                counts = self.zeros((N,), torch.int32)
                pls = (C.c_ulonglong * 1)(lists.data_ptr())
                pcs = (C.c_ulonglong * 1)(counts.data_ptr())
                call("v3s_gram_topk_sym", ptr(a16_all), N, a16_all.shape[1], ptr(lim_loc), Cc, pls, pcs, 1, 0, N, 0, ptr(step_ws),
                     stream_ptr())
            else:
                import torch.distributed as dist
                lim = plan.all_gather_rows(lim_loc)
                counts_loc, pls, pcs, seg_lists, seg_counts, hdl_lists = peer
                # (the all-gather above is also the barrier that orders the peers' merge of the previous pass's
                # segments before the counters are cleared)
                counts_loc.zero_()
                base, rem = divmod(N, world)
                # appends go into LOCAL per-owner buffers (no remote atomics); each owner then merges its patterns'
                # segments straight from the ranks that produced them
                call("v3s_gram_topk_sym", ptr(a16_all), N, a16_all.shape[1], ptr(lim), Cc, pls, pcs, world, plan.rank, base, rem,
                     ptr(step_ws), stream_ptr())
                hdl_lists.barrier()                      # every rank's segments are complete (device-side, stream-ordered)
                lists = counts = None
        else:
            call("v3s_gram_topk_plan", nr, N, k, pl)
            S, Cc, rows_pad = int(pl[0]), int(pl[1]), int(pl[2])
            lists = self.empty((rows_pad * S * Cc,), torch.int64)
            counts = self.empty((rows_pad * S,), torch.int32)
            tws = self.empty((int(pl[3]),), torch.int32)
            call("v3s_gram_topk", ptr(a16_all), N, ptr(a16_rows), nr, a16_all.shape[1], k, eps_g, ptr(lists), ptr(counts),
                 ptr(self.knn_overflow), ptr(tws), stream_ptr())
        self._toc(h)
        if a32_ready is not None:
            call("v3s_set_gram_reserved_sms", 0)
            a32_ready.wait()                           # the all-gather of the FP32 rows ran beside the Gram kernel
        cap = min(k + KNN_MARGIN + KNN_EXTRA_CAP, N)
        cand = self.empty((nr, cap), torch.int32)
        cnt = self.empty((nr,), torch.int32)
        zws = self.empty((cap + 1,), torch.int32) if zero_flag is not None else None
        # tiled refinement (a candidate row is read once per tile of neighbouring rows) where it pays: long rows
        qt = int(_lib.lib.v3s_knn_refine_tile_rows(cap))
        tiled = self.tiled_refine and qt >= 8 and P % 4 == 0 and P >= 1024 and nr >= 2048
        order_ws = tiles_ws = None
        max_tiles = 0
        if tiled:
            stride = qt                                    # one anchor per tile's worth of patterns
            n_bins = (N + stride - 1) // stride + (nr + stride - 1) // stride
            order_ws = self.empty((2 * nr + n_bins,), torch.int32)
            max_tiles = n_bins + nr // qt + 2             # every neighbourhood: its whole tiles + at most one packed flush
            tiles_ws = self.empty((1 + 2 * max_tiles,), torch.int32)
        self._row_order = None
        h = self._tic("select_refine")
        if sym and world > 1:
            call("v3s_knn_merge_segment_lists", seg_lists, seg_counts, world, Cc, nr, N, r0, k, eps_g, cap, ptr(zero_flag),
                 ptr(zws), ptr(cand), ptr(cnt), ptr(self.knn_overflow), qt if tiled else 1, ptr(order_ws), qt, max_tiles,
                 ptr(tiles_ws), stream_ptr())
        else:
            call("v3s_knn_merge_lists", ptr(lists), ptr(counts), S, Cc, nr, N, r0, k, eps_g, cap, ptr(zero_flag), ptr(zws),
                 ptr(cand), ptr(cnt), ptr(self.knn_overflow), qt if tiled else 1, ptr(order_ws), qt, max_tiles, ptr(tiles_ws),
                 stream_ptr())
        # sharded job: pairs listed from both sides are shared ACROSS ranks too -- the candidate lists of all patterns
        # are all-gathered, every rank evaluates the pairs it owns, and the other side reads the sum from the owner's
        # workspace over NVLink (without this 7/8 of an 8-rank job's mutual pairs are evaluated twice)
        share = None
        if (tiled and self.dedup_refine and self.share_pairs_across_ranks and world > 1 and plan is not None
                and r0 == plan.lo and r1 == plan.hi and len(set(plan.counts)) == 1):
            share = self._sym_dss_arena(plan, cap)
        dws = None
        if share is None and (self.dedup_refine or tiled):
            dws = self.empty((nr, cap), torch.float64)
        if tiled:
            pref = self.empty((nr, cap), torch.int32)
            keys = self.empty((max_tiles * 16384,), torch.int64)
            npw = self.empty((max_tiles,), torch.int32)
            if share is None:
                call("v3s_knn_refine_tiled", ptr(a32_all[r0:r1]), ptr(a32_all), P, r0, nr, N, k, cap, ptr(zero_flag), ptr(cand),
                     ptr(cnt), ptr(S2), ptr(Nbr), ptr(dws), ptr(order_ws), ptr(tiles_ws), max_tiles, ptr(pref), ptr(keys),
                     ptr(npw), 1 if self.dedup_refine else 0, stream_ptr())
            else:
                import torch.distributed as dist
                dws, peer_dss, hdl_dss = share
                cand_all = plan.all_gather_rows(cand)
                cnt_all = plan.all_gather_rows(cnt)        # (also orders the peers' reads of the previous pass before the memset)
                call("v3s_knn_refine_tiled_pairs", ptr(a32_all[r0:r1]), ptr(a32_all), P, r0, nr, N, cap, ptr(cand), ptr(cnt),
                     ptr(dws), ptr(order_ws), ptr(tiles_ws), max_tiles, ptr(pref), ptr(keys), ptr(npw), 1, ptr(cand_all),
                     ptr(cnt_all), stream_ptr())
                hdl_dss.barrier()                          # every rank's pair sums are complete (device-side, stream-ordered)
                base, rem = divmod(N, world)
                call("v3s_knn_refine_tiled_emit", r0, nr, N, k, cap, ptr(zero_flag), ptr(cand), ptr(cnt), ptr(S2), ptr(Nbr),
                     ptr(dws), ptr(pref), peer_dss, world, base, rem, stream_ptr())
        else:
            call("v3s_knn_refine", ptr(a32_all[r0:r1]), ptr(a32_all), P, r0, nr, N, k, cap, ptr(zero_flag), ptr(cand), ptr(cnt),
                 ptr(S2), ptr(Nbr), ptr(dws), None, stream_ptr())
        self._toc(h)
        if order_ws is not None:
            self._row_order = (r0, nr, order_ws[:nr], N)
        return S2, Nbr

    def rows_arena(self, plan, P):
        """The rank's centred unit rows (FP32 [cmax, P]) and Gram operand (FP16 [cmax, Ppad]) in torch symmetric memory,
        so that the peers can PULL them with copy-engine transfers (`gather_rows_by_copy_engines`) instead of an NCCL
        all-gather whose kernels would take SMs from the Gram kernels running beside it.
        -> dict(a32, a16: local tensors; a32_peer, a16_peer: every rank's buffer as mapped here; hdl) or None."""
        import torch.distributed as dist
        if not self.copy_engine_gather:
            return None
        Ppad = round_up(P, GRAM_KPAD)
        key = ("rows", plan.N, plan.world, plan.rank, P)
        ent = self._symm_G.get(key)
        if ent is None:
            try:
                import torch.distributed._symmetric_memory as symm
                cmax = max(plan.counts)
                b32 = round_up(cmax * P * 4, 256)
                buf = symm.empty((b32 + cmax * Ppad * 2,), dtype=torch.uint8, device=self.device)
                hdl = symm.rendezvous(buf, plan.group if plan.group is not None else dist.group.WORLD)
                ent = {"buf": buf, "hdl": hdl,
                       "a32": buf[: cmax * P * 4].view(torch.float32).view(cmax, P),
                       "a16": buf[b32: b32 + cmax * Ppad * 2].view(torch.float16).view(cmax, Ppad),
                       "a32_peer": [hdl.get_buffer(r, (cmax, P), torch.float32, 0) for r in range(plan.world)],
                       "a16_peer": [hdl.get_buffer(r, (cmax, Ppad), torch.float16, b32 // 2) for r in range(plan.world)],
                       "streams": [torch.cuda.Stream() for _ in range(min(4, max(1, plan.world - 1)))]}
            except Exception as e:                      # no peer access / API missing: NCCL all-gathers instead
                ent = False
                self.symm_error = repr(e)
            self._symm_G[key] = ent
        return None if ent is False else ent

    def gather_rows_by_copy_engines(self, arena, plan, which, out):
        """out [N, ...] <- every rank's block of arena[which] ("a16" / "a32"), pulled over NVLink by the copy engines on
        side streams (no SMs).  The caller has made sure every rank's block is complete (a collective after the kernels
        that wrote it).  -> an object whose wait() makes the current stream wait for the transfers."""
        ready = torch.cuda.Event()
        ready.record()
        events = []
        streams = arena["streams"]
        for i in range(plan.world):
            q = (plan.rank + i) % plan.world                   # own block first, then the peers in rotating order
            st = streams[i % len(streams)]
            st.wait_event(ready)
            with torch.cuda.stream(st):
                out[plan.offsets[q]: plan.offsets[q] + plan.counts[q]].copy_(arena[which + "_peer"][q][: plan.counts[q]],
                                                                              non_blocking=True)
        for st in streams:
            ev = torch.cuda.Event()
            ev.record(st)
            events.append(ev)
        return _StreamEvents(events)

    def _sym_dss_arena(self, plan, cap):
        """The rank's pair-sum workspace [cmax, cap] FP64 of the tiled refinement in torch symmetric memory (peers read
        the sums of pairs this rank owns).  -> (local tensor, table of every rank's workspace) or None."""
        import torch.distributed as dist
        key = ("dss", plan.N, plan.world, plan.rank, cap)
        ent = self._symm_G.get(key)
        if ent is None:
            try:
                import torch.distributed._symmetric_memory as symm
                cmax = max(plan.counts)
                buf = symm.empty((cmax, cap), dtype=torch.float64, device=self.device)
                hdl = symm.rendezvous(buf, plan.group if plan.group is not None else dist.group.WORLD)
                ent = (buf, (C.c_ulonglong * plan.world)(*[int(x) for x in hdl.buffer_ptrs]), hdl)
            except Exception as e:                      # no peer access / API missing: every rank evaluates its own pairs
                ent = False
                self.symm_error = repr(e)
            self._symm_G[key] = ent
        return None if ent is False else ent

    def _sym_lists_arena(self, plan, Cc):
        """Buffers of the sharded symmetric Gram in torch symmetric memory: on every rank, one (lists [cmax, Cc] int64,
        counters [cmax] int32) region PER OWNER rank, filled by the local kernel with local atomics; region `o` of rank
        q holds what q found for o's patterns, and o reads it over NVLink when it merges.
        -> (local counters [world * cmax] int32, append tables (lists, counters: local, per owner), merge tables
        (lists, counters: this rank's region on every source rank)) or None."""
        import torch.distributed as dist
        key = ("symlists", plan.N, plan.world, plan.rank, Cc)
        ent = self._symm_G.get(key)
        if ent is None:
            try:
                import torch.distributed._symmetric_memory as symm
                cmax = max(plan.counts)
                w = plan.world
                seg = cmax * Cc * 8                          # one owner's lists
                lbytes = w * seg
                buf = symm.empty((lbytes + w * cmax * 4 + 256,), dtype=torch.uint8, device=self.device)
                hdl = symm.rendezvous(buf, plan.group if plan.group is not None else dist.group.WORLD)
                bases = [int(x) for x in hdl.buffer_ptrs]
                mine = bases[plan.rank]
                counts = buf[lbytes: lbytes + w * cmax * 4].view(torch.int32)
                u64 = C.c_ulonglong * w
                ent = (counts,
                       u64(*[mine + o * seg for o in range(w)]), u64(*[mine + lbytes + o * cmax * 4 for o in range(w)]),
                       u64(*[b + plan.rank * seg for b in bases]), u64(*[b + lbytes + plan.rank * cmax * 4 for b in bases]), hdl, buf)
            except Exception as e:                      # no peer access / API missing: rectangular kernel instead
                ent = False
                self.symm_error = repr(e)
            self._symm_G[key] = ent
        return None if ent is False else ent[:6]

    def gram_rows(self, hi_all, lo_all, a32_all, r0, r1, out=None):
        """G[r0:r1, :] as f32 [r1-r0, ldg]."""
        N = a32_all.shape[0]
        nr = r1 - r0
        ldg = round_up(N, 4)
        G = out if out is not None else self.empty((nr, ldg), torch.float32)
        h = self._tic("gram")
        if self.gram_impl == "simt":
            call("v3s_gram_rows_simt", ptr(a32_all), N, ptr(a32_all[r0:r1]), nr, a32_all.shape[1], ptr(G), ldg, stream_ptr())
        else:
            call("v3s_gram_rows", ptr(hi_all), ptr(lo_all), N, ptr(hi_all[r0:r1]), ptr(lo_all[r0:r1]), nr,
                 hi_all.shape[1], ptr(G), ldg, stream_ptr())
        self._toc(h)
        return G

    def _sym_tiles(self, N):
        """Tile list of the symmetric Gram: (column block 128, row block 256) pairs on or above the
        diagonal, ordered in groups of 16 column blocks so concurrent CTAs share operands in L2."""
        t = self._sym_tile_cache.get(N)
        if t is None:
            import numpy as np
            tiles_c, tiles_r = (N + 127) // 128, (N + 255) // 256
            out = []
            for g0 in range(0, tiles_c, 16):
                cols = np.arange(g0, min(g0 + 16, tiles_c))
                tr, tc = np.meshgrid(np.arange(tiles_r), cols, indexing="ij")
                keep = (128 * tc + 127) >= (256 * tr)
                out.append(np.stack([tc[keep], tr[keep]], axis=1))
            lst = np.ascontiguousarray(np.concatenate(out, axis=0).astype(np.int32))
            t = (self.to_device(lst, torch.int32), int(lst.shape[0]))
            self._sym_tile_cache[N] = t
        return t

    def gram_symmetric(self, hi_all, lo_all, N, out=None):
        """Whole G [N, ldg] f32 with half the tensor-core work (upper-triangle tiles, mirrored stores)."""
        ldg = round_up(N, 4)
        G = out if out is not None else self.empty((N, ldg), torch.float32)
        tiles, nt = self._sym_tiles(N)
        h = self._tic("gram")
        call("v3s_gram_symmetric", ptr(hi_all), ptr(lo_all), N, hi_all.shape[1], ptr(G), ldg, ptr(tiles), nt, stream_ptr())
        self._toc(h)
        return G

    def gram_symmetric_sharded(self, hi_all, lo_all, plan):
        """Row block [n_loc, ldg] of the symmetric Gram in a sharded job, with half the tensor-core
        work per rank: each rank computes 1/world of the global upper-triangle tiles and stores every
        tile (and its mirror) straight into the owning ranks' buffers over NVLink from the kernel's
        epilogue (torch symmetric memory provides the peer mappings).  Returns None when symmetric
        memory is unavailable (the caller then uses rectangular panels)."""
        import torch.distributed as dist
        N = plan.N
        ldg = round_up(N, 4)
        cmax = max(plan.counts)
        key = (N, plan.world, plan.rank)
        ent = self._symm_G.get(key)
        if ent is None:
            try:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty((cmax, ldg), dtype=torch.float32, device=self.device)
                hdl = symm.rendezvous(buf, plan.group if plan.group is not None else dist.group.WORLD)
                ptrs = (C.c_ulonglong * plan.world)(*[int(x) for x in hdl.buffer_ptrs])
            except Exception as e:                      # no peer access / API missing: not an error
                self._symm_G[key] = False
                self.symm_error = repr(e)
                return None
            tiles, nt = self._sym_tiles(N)
            per = (nt + plan.world - 1) // plan.world
            t0, t1 = min(plan.rank * per, nt), min((plan.rank + 1) * per, nt)
            ent = (buf, ptrs, tiles[t0:t1].contiguous(), t1 - t0)
            self._symm_G[key] = ent
        if ent is False:
            return None
        buf, ptrs, my_tiles, my_nt = ent
        base, rem = divmod(N, plan.world)
        dist.barrier(group=plan.group)                   # nobody is still reading its block from the previous pass
        h = self._tic("gram")
        call("v3s_gram_symmetric_sharded", ptr(hi_all), ptr(lo_all), N, hi_all.shape[1], ptrs, plan.world, base, rem, ldg,
             ptr(my_tiles), my_nt, stream_ptr())
        self._toc(h)
        dist.barrier(group=plan.group)                   # every rank's tiles (incl. peer stores into this block) are done
        return buf[: plan.hi - plan.lo]

    def knn_rows(self, hi_all, lo_all, a32_all, zero_flag, r0, r1, k, plan=None, a16_all=None):
        """k nearest neighbours (round distance) of rows [r0,r1) against all rows.
        -> S2 [r1-r0,k] f64 ascending, Nbr [r1-r0,k] i32.  Gram panels are streamed."""
        N, P = a32_all.shape
        if k > N:
            raise Exception("Number of nearest neighbors to preserve is too high.")   # distance.py:33-34
        nr = r1 - r0
        if self.gram_impl == "fused":
            return self.knn_rows_fused(a16_all if a16_all is not None else self.split_f16(a32_all), a32_all, zero_flag, r0, r1, k,
                                       plan=plan)
        if hi_all is None:
            hi_all, lo_all = self.split_bf16(a32_all)
        S2 = self.empty((nr, k), torch.float64)
        Nbr = self.empty((nr, k), torch.int32)
        if nr == 0:
            return S2, Nbr
        ldg = round_up(N, 4)
        sym = self.gram_impl == "tcgen05" and r0 == 0 and r1 == N and 4 * ldg * N <= SYM_BYTES
        if sym:
            # cudaMemGetInfo is a kernel-mode driver call (it can queue behind any other RM client of the box, e.g. a
            # monitoring nvidia-smi, for tens of milliseconds): ask once per problem size, not once per pass
            fits = self._sym_fits.get(N)
            if fits is None:
                fits = self._sym_fits[N] = bool(4 * ldg * N <= 0.6 * torch.cuda.mem_get_info()[0])
            sym = fits
        G_shard = None
        if (self.gram_impl == "tcgen05" and plan is not None and plan.world > 1 and plan.world <= 8 and self.use_peer_gram
                and r0 == plan.lo and r1 == plan.hi and N // plan.world >= 1 and 4 * ldg * max(plan.counts) <= SYM_BYTES):
            G_shard = self.gram_symmetric_sharded(hi_all, lo_all, plan)
        panel = nr if (sym or G_shard is not None) else max(256, min(nr, (PANEL_BYTES // (4 * ldg)) // 256 * 256))
        cap = min(k + KNN_MARGIN + KNN_EXTRA_CAP, N)
        eps_g = 2.0 * gram_error_bound(P) if self.gram_impl == "tcgen05" else 2e-5
        G = G_shard if G_shard is not None else self.empty((min(panel, nr), ldg), torch.float32)
        cand = self.empty((min(panel, nr), cap), torch.int32)
        cnt = self.empty((min(panel, nr),), torch.int32)
        dws = self.empty((min(panel, nr), cap), torch.float64) if self.dedup_refine else None   # mutual pairs evaluated once
        self.knn_overflow = self.zeros((1,), torch.int32)
        zf = zero_flag if (zero_flag is not None) else None
        for p0 in range(r0, r1, panel):
            p1 = min(p0 + panel, r1)
            if G_shard is not None:
                pass                                     # already complete (computed by all ranks, stored over NVLink)
            elif sym:
                self.gram_symmetric(hi_all, lo_all, N, out=G)
            else:
                self.gram_rows(hi_all, lo_all, a32_all, p0, p1, out=G)
            h = self._tic("select_refine")
            call("v3s_knn_from_gram", ptr(G), ldg, p1 - p0, N, p0, ptr(a32_all[p0:p1]), ptr(a32_all), P, k, KNN_MARGIN,
                 eps_g, cap, ptr(zf), ptr(cand), ptr(cnt), ptr(self.knn_overflow), ptr(S2[p0 - r0:p1 - r0]),
                 ptr(Nbr[p0 - r0:p1 - r0]), ptr(dws), stream_ptr())
            self._toc(h)
        return S2, Nbr

    def round_distance_dense(self, a32, zero_flag):
        n, P = a32.shape
        B = self.empty((n, n), torch.float64)
        call("v3s_round_distance_dense", ptr(a32), n, P, ptr(zero_flag), ptr(B), stream_ptr())
        return B

    def knn_from_dense(self, B, k):
        n = B.shape[0]
        if k > n:
            raise Exception("Number of nearest neighbors to preserve is too high.")
        kk = min(k + KNN_MARGIN, n)
        key = self.empty((n, n), torch.float32)
        cand = self.empty((n, kk), torch.int32)
        cnt = self.empty((n,), torch.int32)
        S2 = self.empty((n, k), torch.float64)
        Nbr = self.empty((n, k), torch.int32)
        call("v3s_knn_from_dense", ptr(B), n, k, KNN_MARGIN, ptr(key), ptr(cand), ptr(cnt), ptr(S2), ptr(Nbr), stream_ptr())
        return S2, Nbr

    # ---- stage 3 -----------------------------------------------------------------------------
    def s2_scalars(self, S2, eps, med_row=4, idx_off=1):
        sc = self.empty((5,), torch.float64)
        call("v3s_s2_scalars", ptr(S2), S2.shape[0], S2.shape[1], float(eps), med_row, idx_off, ptr(sc), stream_ptr())
        return sc

    def s2_threshold_(self, S2, scalars):
        call("v3s_s2_threshold", ptr(S2), S2.shape[0], S2.shape[1], ptr(scalars), stream_ptr())
        return S2

    def build_markov(self, S2t, Nbr, scalars, r0, r1):
        N, k = S2t.shape
        ldp = round_up(N, 4)
        P_rows = self.empty((r1 - r0, ldp), torch.float32)
        dis = self.empty((N,), torch.float64)
        ws = self.empty((2 * N,), torch.float64)
        h = self._tic("markov")
        call("v3s_build_markov", ptr(S2t), ptr(Nbr), N, k, ptr(scalars), r0, r1 - r0, ptr(P_rows), ldp, ptr(dis), ptr(ws),
             stream_ptr())
        self._toc(h)
        return P_rows, dis

    def wants_sparse_operator(self, n_rows, N, k=150):
        """operator "auto": the form a cost model predicts to be faster for one block mat-vec (measured, round 1:
        dense = 11 us + 4 rows N bytes at ~7 TB/s; sparse = ~10 us + ~12 ps per stored entry, ~2 k per row), with
        the dense row block also costing its N^2 build and at most DENSE_OPERATOR_BYTES of memory."""
        if self.operator == "sparse":
            return True
        if self.operator != "auto":
            return False
        dense_bytes = 4.0 * n_rows * round_up(N, 4)
        if dense_bytes > DENSE_OPERATOR_BYTES:
            return True
        t_dense = 11e-6 + dense_bytes / 7.0e12
        t_sparse = 10e-6 + 2.0 * k * n_rows * 1.2e-11
        return t_sparse < t_dense

    def build_markov_sparse(self, S2t, Nbr, scalars, r0, r1):
        """-> SparseMarkov (rows [r0,r1) of P_ep = A + A^T on the lists, FP64), dinv_sqrt [N]."""
        N, k = S2t.shape
        a_val = self.empty((N, k), torch.float64)
        t_ptr = self.empty((N + 1,), torch.int32)
        t_col = self.empty((N * k,), torch.int32)
        t_val = self.empty((N * k,), torch.float64)
        dis = self.empty((N,), torch.float64)
        ws = self.empty((2 * N,), torch.float64)
        iws = self.empty((N + N * k,), torch.int32)
        h = self._tic("markov")
        call("v3s_build_markov_sparse", ptr(S2t), ptr(Nbr), N, k, ptr(scalars), ptr(a_val), ptr(t_ptr), ptr(t_col), ptr(t_val),
             ptr(dis), ptr(ws), ptr(iws), stream_ptr())
        self._toc(h)
        # the neighbourhood order the kNN stage found for exactly these rows (if it ran in this process for this N)
        ro = self._row_order if (self._row_order is not None and self._row_order[0] == r0 and self._row_order[1] == r1 - r0
                                 and self._row_order[3] == N and self.sparse_row_order) else None
        return SparseMarkov(Nbr, a_val, t_ptr, t_col, t_val, r0, r1, ro[:3] if ro else None), dis

    def sparse_block_matvec(self, sp, X, out=None):
        """Y_rows [n_rows,b] f64 = P[r0:r1, :] @ X [N,b] (f64), sparse operator."""
        b = X.shape[1]
        Y = out if out is not None else self.empty((sp.n_rows, b), torch.float64)
        call("v3s_sparse_block_matvec", C.byref(sp.desc), sp.r0, sp.n_rows, sp.N, ptr(X), b, ptr(Y), stream_ptr())
        return Y

    def w_dense(self, S2t, Nbr, eps_eff):
        N, k = S2t.shape
        W = self.empty((N, N), torch.float64)
        call("v3s_w_dense", ptr(S2t), ptr(Nbr), N, k, float(eps_eff), ptr(W), stream_ptr())
        return W

    def anisotropic_norm_dense(self, W):
        N = W.shape[0]
        out = self.empty((N, N), torch.float64)
        ws = self.empty((N,), torch.float64)
        call("v3s_anisotropic_norm_dense", ptr(W), N, ptr(out), ptr(ws), stream_ptr())
        return out

    def dm_cov_dense(self, W):
        N = W.shape[0]
        P = self.empty((N, N), torch.float64)
        dis = self.empty((N,), torch.float64)
        ws = self.empty((N,), torch.float64)
        call("v3s_dm_cov_dense", ptr(W), N, ptr(P), ptr(dis), ptr(ws), stream_ptr())
        return P, dis

    def block_matvec(self, P_rows, N, X):
        """Y_rows [n_rows,b] f64 = P_rows[:, :N] (f32) @ X [N,b] (f64)."""
        n_rows, ldp = P_rows.shape
        b = X.shape[1]
        ws = self._mv_workspace(n_rows, N)
        Y = self.empty((n_rows, b), torch.float64)
        call("v3s_block_matvec", ptr(P_rows), ldp, n_rows, N, ptr(X), b, ptr(Y), ptr(ws), stream_ptr())
        return Y

    def _mv_workspace(self, n_rows, N):
        need = _lib.lib.v3s_block_matvec_workspace_bytes(n_rows, N)
        if self._mv_ws is None or self._mv_ws.numel() < need:
            self._mv_ws = self.empty((need,), torch.uint8)
        return self._mv_ws

    def alloc_xf(self, N):
        """FP32 operand block of the mat-vec in strip layout (opaque to the host logic), zeroed."""
        return self.zeros((int(_lib.lib.v3s_xf_floats(N)),), torch.float32)

    def block_matvec_f32(self, P_rows, N, Xf, out=None):
        """Y_rows [n_rows,8] f64 = P_rows[:, :N] (f32) @ X, X given as the FP32 strip-layout block."""
        n_rows, ldp = P_rows.shape
        Y = out if out is not None else self.empty((n_rows, 8), torch.float64)
        ws = self._mv_workspace(n_rows, N)
        call("v3s_block_matvec_f32", ptr(P_rows), ldp, n_rows, N, ptr(Xf), ptr(Y), ptr(ws), stream_ptr())
        return Y

    def fused_operator(self, P_rows, N, plan=None):
        """The eigen-solver's operator with the result exchange fused into the kernels (`FusedOperator`)."""
        try:
            return FusedOperator(self, P_rows, N, plan)
        except Exception as e:                      # symmetric memory / peer access unavailable: NCCL path instead
            if plan is None or plan.world == 1:
                raise
            self.symm_error = repr(e)
            self._symm_G[("fused", int(N), plan.world, plan.rank)] = False
            return None

    # device-side block-Lanczos state and steps (lanczos.cu)
    def lanczos_alloc(self, N, max_dim, max_steps):
        st = {"N": N, "max_dim": max_dim,
              "V": self.empty((max_dim, N), torch.float64), "Xf": self.alloc_xf(N),
              "T": self.zeros((max_steps, 128), torch.float64), "status": self.zeros((1,), torch.int32),
              "ws": self.empty((int(_lib.lib.v3s_lanczos_workspace_doubles(N, max_dim)),), torch.float64)}
        return st

    def lanczos_snapshot(self, st, s0, s1):
        """Async D2H of the (A_j, B_j) blocks of steps [s0, s1) and the breakdown flag into pinned
        host memory; returns a handle for `lanczos_snapshot_wait`."""
        if self._pin is None or self._pin.numel() < st["T"].numel() + 1:
            self._pin = torch.empty((st["T"].numel() + 1,), dtype=torch.float64, pin_memory=True)
        n = (s1 - s0) * 128
        self._pin[:n].copy_(st["T"][s0:s1].reshape(-1), non_blocking=True)
        self._pin[n:n + 1].copy_(st["status"].to(torch.float64), non_blocking=True)
        ev = self._snap_events.pop() if self._snap_events else torch.cuda.Event()     # reused: no event creation per check
        ev.record()
        return (ev, n)

    def lanczos_snapshot_wait(self, snap):
        ev, n = snap
        ev.synchronize()
        self._snap_events.append(ev)
        host = self._pin[:n + 1].numpy()
        return host[:n].copy(), int(host[n])

    def lanczos_start(self, st, W0, row0=0):
        """Orthonormalise W0 [N,8] (FP64, overwritten) into the basis block at rows [row0, row0+8)."""
        call("v3s_lanczos_start", ptr(st["V"][row0:]), st["N"], st["N"], st["max_dim"], ptr(W0), ptr(st["Xf"]),
             ptr(st["ws"]), stream_ptr())

    def lanczos_step(self, st, m, W, step):
        h = self._tic("lanczos_orth")
        call("v3s_lanczos_step", ptr(st["V"]), st["N"], st["N"], m, st["max_dim"], ptr(W), ptr(st["Xf"]),
             ptr(st["T"][step]), ptr(st["ws"]), ptr(st["status"]), stream_ptr())
        self._toc(h)

    def cheb_combine(self, PY, Yk, Ykm1, alpha, beta, gamma, Y_out, Xf_out):
        call("v3s_cheb_combine", ptr(PY), ptr(Yk), ptr(Ykm1), float(alpha), float(beta), float(gamma), ptr(Y_out),
             ptr(Xf_out), PY.numel(), stream_ptr())

    def sym_eig_host(self, A):
        """Eigenpairs of a small symmetric matrix (numpy, host) by v3s_sym_eig_host: w ascending, vectors in columns."""
        import numpy as np
        a = np.ascontiguousarray(np.asarray(A, dtype=np.float64)).copy()
        n = a.shape[0]
        w = np.empty((n,), dtype=np.float64)
        call("v3s_sym_eig_host", a.ctypes.data_as(C.POINTER(C.c_double)), n, w.ctypes.data_as(C.POINTER(C.c_double)))
        return w, a

    # ---- stage 4 -----------------------------------------------------------------------------
    def fit_gram(self, Ps, Y):
        out = self.empty((162,), torch.float64)
        call("v3s_fit_gram", ptr(Ps), Ps.stride(0), ptr(Y), Ps.shape[0], ptr(out), stream_ptr())
        return out

    def fit_solve(self, g162):
        c = self.empty((9, 9), torch.float64)
        call("v3s_fit_solve", ptr(g162), ptr(c), stream_ptr())
        return c

    def fit_project(self, Ps, Y, c, polar_mode):
        n = Ps.shape[0]
        recon = self.empty((n, 9), torch.float64)
        rproj = self.empty((n, 9), torch.float64)
        qe = self.empty((n, 4), torch.float64)
        qt = self.empty((n, 4), torch.float64)
        call("v3s_fit_project", ptr(Ps), Ps.stride(0), ptr(Y), n, ptr(c), int(polar_mode), ptr(recon), ptr(rproj), ptr(qe),
             ptr(qt), stream_ptr())
        return recon, rproj, qe, qt

    def or_func(self, c81, Ps, r_total, form=0, want_jac=True, rel_step=1.4901161193847656e-08):
        """Residuals G [n] of DiffusionMapSolver.or_func and (optionally) its forward-difference
        Jacobian J [n,81] (rel_step = sqrt(eps), scipy's '2-point' default)."""
        n = Ps.shape[0]
        G = self.empty((n,), torch.float64)
        J = self.empty((n, 81), torch.float64) if want_jac else None
        call("v3s_or_func", ptr(c81), ptr(Ps), Ps.stride(0), n, int(r_total), int(form), float(rel_step), ptr(G), ptr(J),
             stream_ptr())
        return G, J

    def sigma_pairs(self, qe, qt, r0, r1):
        if self._sigma_ws is None:
            self._sigma_ws = self.empty((65536,), torch.float64)
        out = self.empty((1,), torch.float64)
        h = self._tic("sigma_pairs")
        call("v3s_sigma_pairs", ptr(qe), ptr(qt), qe.shape[0], r0, r1 - r0, ptr(self._sigma_ws), ptr(out), stream_ptr())
        self._toc(h)
        return out


class FusedOperator:
    """P (and Chebyshev polynomials of P) applied to a Lanczos block with no collective and no
    separate reduction / combination launches: `v3s_matvec_combine` / `v3s_cheb_filter` store each
    rank's rows of the result -- FP64 block and FP32 operand of the next mat-vec -- straight into every
    rank's buffers (NVLink peer mappings from torch symmetric memory when sharded) and order the steps
    with `v3s_peer_barrier`.  One arena per (N, world) is cached on the ops object: three rotating
    [N,8] FP64 result blocks, two ping-pong FP32 strip operands, the barrier flag words.

    plain(Xf, Q) -> P X;  cheb(Xf, Q, degree, c, e) -> T_degree((P - c)/e) X;  both return one of the
    rotating blocks (complete on every rank), never the block Q itself."""

    def __init__(self, ops, P_rows, N, plan=None):
        import torch.distributed as dist
        self.ops, self.P_rows, self.N = ops, P_rows, int(N)
        self.sparse = isinstance(P_rows, SparseMarkov)
        self.world = 1 if plan is None else plan.world
        self.rank = 0 if plan is None else plan.rank
        self.row0 = 0 if plan is None else plan.lo
        self.group = None if plan is None else plan.group
        if self.world > 8:
            raise Exception("FusedOperator: at most 8 ranks (one NVLink domain)")
        key = ("fused", self.N, self.world, self.rank)
        arena = ops._symm_G.get(key)
        if arena is False:
            raise Exception("symmetric memory unavailable (earlier attempt failed)")
        if arena is None:
            xf_floats = int(_lib.lib.v3s_xf_floats(self.N))
            y_bytes = round_up(self.N * 8 * 8, 256)
            x_bytes = round_up(xf_floats * 4, 256)
            total = 3 * y_bytes + 2 * x_bytes + 256
            if self.world > 1:
                import torch.distributed._symmetric_memory as symm
                buf = symm.empty((total,), dtype=torch.uint8, device=ops.device)
                hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
                bases = [int(x) for x in hdl.buffer_ptrs]
            else:
                buf = torch.empty((total,), dtype=torch.uint8, device=ops.device)
                bases = [buf.data_ptr()]
            buf.zero_()                                   # strip padding columns and flag words start at zero
            if self.world > 1:
                torch.cuda.synchronize()
                dist.barrier(group=self.group)            # nobody stores into a buffer that is still being zeroed
            ybufs = [buf[b * y_bytes: b * y_bytes + self.N * 64].view(torch.float64).view(self.N, 8) for b in range(3)]
            xfs = [buf[3 * y_bytes + q * x_bytes: 3 * y_bytes + q * x_bytes + xf_floats * 4].view(torch.float32) for q in range(2)]
            flag_off = 3 * y_bytes + 2 * x_bytes
            arena = {
                "buf": buf, "ybufs": ybufs, "xfs": xfs,
                "yptrs": (C.c_ulonglong * (3 * self.world))(*[bases[r] + b * y_bytes for b in range(3) for r in range(self.world)]),
                "xptrs": (C.c_ulonglong * (2 * self.world))(*[bases[r] + 3 * y_bytes + q * x_bytes for q in range(2) for r in range(self.world)]),
                "fptrs": (C.c_ulonglong * self.world)(*[bases[r] + flag_off for r in range(self.world)]),
                "ticket": buf[flag_off + 128: flag_off + 132],            # last-block ticket of the fused combine kernel
                "status": torch.zeros((1,), dtype=torch.int32, device=ops.device), "epoch": 0,
            }
            ops._symm_G[key] = arena
        self.a = arena
        self.n_matvec = 0
        if self.sparse:
            if (P_rows.r0, P_rows.N) != (self.row0, self.N):
                raise Exception("FusedOperator: the sparse operator's row block does not start at the rank's first row")
            self._sp_ws = ops.empty((max(P_rows.n_rows, 1) * 8,), torch.float64)

    def _index_of(self, Q):
        for b, t in enumerate(self.a["ybufs"]):
            if t.data_ptr() == Q.data_ptr():
                return b
        return -1

    def sync(self):
        """Cross-rank barrier on the stream: everything this rank enqueued so far is complete before any
        peer's later stores into this rank's buffers (phase changes of the solver)."""
        if self.world > 1:
            self.a["epoch"] += 1
            call("v3s_peer_barrier", self.a["fptrs"], self.rank, self.world, self.a["epoch"], ptr(self.a["status"]), stream_ptr())

    def plain(self, Xf, Q):
        a, w = self.a, self.world
        qi = self._index_of(Q)
        out = min(b for b in range(3) if b != qi)
        yout = (C.c_ulonglong * w)(*[a["yptrs"][out * w + r] for r in range(w)])
        a["epoch"] += 1 if w > 1 else 0
        if self.sparse:
            sp = self.P_rows
            call("v3s_sparse_matvec_combine", C.byref(sp.desc), sp.n_rows, self.N, self.row0, ptr(Q), None, None, 1.0, 0.0, 0.0,
                 yout, None, a["fptrs"], self.rank, w, a["epoch"], ptr(a["ticket"]), ptr(a["status"]), ptr(self._sp_ws),
                 stream_ptr())
            self.n_matvec += 1
            return a["ybufs"][out]
        ws = self.ops._mv_workspace(self.P_rows.shape[0], self.N)
        call("v3s_matvec_combine", ptr(self.P_rows), self.P_rows.shape[1], self.P_rows.shape[0], self.N, self.row0, ptr(Xf),
             None, None, 1.0, 0.0, 0.0, yout, None, a["fptrs"], self.rank, w, a["epoch"], ptr(a["ticket"]), ptr(a["status"]),
             ptr(ws), stream_ptr())
        self.n_matvec += 1
        return a["ybufs"][out]

    def cheb(self, Xf0, Q, degree, c, e):
        a = self.a
        res = C.c_int(0)
        if self.sparse:
            sp = self.P_rows
            call("v3s_sparse_cheb_filter", C.byref(sp.desc), sp.n_rows, self.N, self.row0, ptr(Q), self._index_of(Q), int(degree),
                 float(c), float(e), a["yptrs"], a["fptrs"], self.rank, self.world, a["epoch"], ptr(a["ticket"]),
                 ptr(a["status"]), ptr(self._sp_ws), C.byref(res), stream_ptr())
            if self.world > 1:
                a["epoch"] += int(degree)
            self.n_matvec += int(degree)
            return a["ybufs"][res.value]
        ws = self.ops._mv_workspace(self.P_rows.shape[0], self.N)
        call("v3s_cheb_filter", ptr(self.P_rows), self.P_rows.shape[1], self.P_rows.shape[0], self.N, self.row0, ptr(Xf0),
             ptr(Q), self._index_of(Q), int(degree), float(c), float(e), a["yptrs"], a["xptrs"], a["fptrs"], self.rank,
             self.world, a["epoch"], ptr(a["ticket"]), ptr(a["status"]), ptr(ws), C.byref(res), stream_ptr())
        if self.world > 1:
            a["epoch"] += int(degree)
        self.n_matvec += int(degree)
        return a["ybufs"][res.value]

    def eig_topk(self, nev, tol=1e-8, max_dim=320, degree=6, probe_steps=10, seed=0):
        """The whole eigen-solver through ONE C-ABI call (v3s_eig_topk: block Lanczos + Chebyshev filter + Ritz
        extraction driven from the library, no Python / LAPACK in the loop).
        -> lambda [nev] numpy (by decreasing magnitude), Y [N,nev] f64 tensor, residual norms [nev], info dict."""
        a = self.a
        pb = _lib.EigProblem()
        if self.sparse:
            pb.P_rows, pb.ldp, pb.sparse = None, 0, C.cast(C.pointer(self.P_rows.desc), C.c_void_p)
            pb.n_rows = self.P_rows.n_rows
            ws_mv = self._sp_ws
        else:
            pb.P_rows, pb.ldp, pb.sparse = self.P_rows.data_ptr(), self.P_rows.shape[1], None
            pb.n_rows = self.P_rows.shape[0]
            ws_mv = self.ops._mv_workspace(self.P_rows.shape[0], self.N)
        pb.N, pb.row0, pb.rank, pb.world = self.N, self.row0, self.rank, self.world
        pb.ybuf_ptrs = C.cast(a["yptrs"], C.c_void_p)
        pb.xf_ptrs = C.cast(a["xptrs"], C.c_void_p)
        pb.flag_ptrs = C.cast(a["fptrs"], C.c_void_p)
        pb.ticket_dev, pb.status_dev = a["ticket"].data_ptr(), a["status"].data_ptr()
        pb.mv_workspace = ws_mv.data_ptr()
        pb.epoch = a["epoch"]
        md = max(16, (min(int(max_dim), self.N) // 8) * 8)
        ws = self.ops.empty((int(_lib.lib.v3s_eig_topk_workspace_doubles(self.N, md)),), torch.float64)
        Y = self.ops.empty((self.N, nev), torch.float64)
        lam = (C.c_double * nev)()
        res = (C.c_double * nev)()
        info = (C.c_int * 8)()
        call("v3s_eig_topk", C.byref(pb), int(nev), float(tol), md, int(degree), int(probe_steps), int(seed), ptr(ws), lam, ptr(Y),
             res, info, stream_ptr())
        a["epoch"] = int(pb.epoch)
        self.n_matvec += int(info[1])
        import numpy as _np
        return (_np.array(lam[:nev]), Y, _np.array(res[:nev]),
                {"converged": bool(info[0]), "matvecs": int(info[1]), "krylov_dim": int(info[2]),
                 "driver": "cabi-chebyshev" if info[3] else "cabi-plain", "ritz_checks": int(info[4]),
                 "saturated_cluster": bool(info[5])})

    def check(self):
        """Raise if a peer barrier timed out (one small D2H; called once after every solve).  The status word
        lives in the cached arena: it is cleared once reported, so one failed solve does not poison the next."""
        if self.world > 1:
            st = int(self.ops.read_small(self.a["status"])[0])
            if st != 0:
                self.a["status"].zero_()
                raise Exception("v3s: a peer rank never reached the mat-vec barrier (status %d): the result blocks of "
                                "this solve are incomplete" % st)
