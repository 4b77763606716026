"""Device-resident pipeline: patterns -> kNN round distances -> diffusion coordinates -> orientations.

Host logic only; all arithmetic is done by the backend `ops` (CudaOps in the product).  The same
functions drive one GPU or a row-sharded job (one process per GPU, torch.distributed):
every N x N object (Gram panel, P_ep) is split by contiguous row blocks, rank g owning rows
[lo_g, hi_g).  Exchanges:
    all-reduce  per-pixel sums (mean image)                    P doubles
    all-gather  centred unit rows (BF16 split done locally)    N x P x 4 B
    peer stores symmetric Gram tiles into the owners' blocks   N x N x 4 B / world per rank (NVLink, in-kernel)
    all-gather  kNN lists                                      N x k x 12 B
    all-gather  block mat-vec result, once per Lanczos step    N x b x 8 B
    all-reduce  Sigma_T partial sum                            1 double
Small replicated work (Lanczos basis algebra, the 9x9 fit) is recomputed on every rank from
identical inputs, so all ranks hold identical results without further communication.
"""
import math

import numpy as np
import torch

from .eig import block_lanczos_device, block_lanczos_topk, chebyshev_lanczos_device, topk_with_locking

NEV = 10   # diffusion.py:42 (k = 10 eigenpairs)


class ShardPlan:
    """Contiguous near-equal row blocks over the ranks of `group` (None = single process)."""

    def __init__(self, N, rank=0, world=1, group=None):
        self.N, self.rank, self.world, self.group = int(N), int(rank), int(world), group
        base, rem = divmod(self.N, self.world)
        self.counts = [base + (1 if r < rem else 0) for r in range(self.world)]
        self.offsets = [sum(self.counts[:r]) for r in range(self.world)]
        self.lo = self.offsets[self.rank]
        self.hi = self.lo + self.counts[self.rank]

    @staticmethod
    def from_env(N):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return ShardPlan(N, dist.get_rank(), dist.get_world_size(), dist.group.WORLD)
        return ShardPlan(N)

    _side_groups = {}

    def side_group(self):
        """A second communicator over the same ranks: a long transfer issued on it (the all-gather of the FP32 rows,
        which runs beside the Gram kernels) does not hold up the short collectives of the main group -- operations of
        one NCCL communicator execute in issue order."""
        import torch.distributed as dist
        key = (id(self.group), self.world)
        g = ShardPlan._side_groups.get(key)
        if g is None:
            ranks = dist.get_process_group_ranks(self.group) if self.group is not None else list(range(self.world))
            g = ShardPlan._side_groups[key] = dist.new_group(ranks=ranks)
        return g

    # ---- collectives (no-ops for a single process) ---------------------------------------
    def all_reduce_sum(self, t):
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_gather_rows(self, t_local, out=None):
        """Concatenate per-rank row blocks [count_r, ...] into [N, ...] on every rank.
        `out` (optional, [N, ...]) avoids an allocation per call when the blocks are equal-sized."""
        if self.world == 1:
            return t_local
        import torch.distributed as dist
        cmax = max(self.counts)
        tail = tuple(t_local.shape[1:])
        if all(c == cmax for c in self.counts):
            if out is None:
                out = torch.empty((self.N,) + tail, dtype=t_local.dtype, device=t_local.device)
            dist.all_gather_into_tensor(out, t_local.contiguous(), group=self.group)
            return out
        pad = torch.zeros((cmax,) + tail, dtype=t_local.dtype, device=t_local.device)
        pad[: t_local.shape[0]] = t_local
        buf = torch.empty((self.world * cmax,) + tail, dtype=t_local.dtype, device=t_local.device)
        dist.all_gather_into_tensor(buf, pad, group=self.group)
        return torch.cat([buf[r * cmax: r * cmax + self.counts[r]] for r in range(self.world)], dim=0)


def knn_stage(ops, plan, img_local, k):
    """distance.py:17-22 (DistanceMatrix.knn) on the rank's pattern rows.
    -> S2_all [N,k] f64, Nbr_all [N,k] i32 (replicated), a32_all."""
    N = plan.N
    sums = ops.column_sums(img_local)
    plan.all_reduce_sum(sums)
    mean = sums / N
    fused = getattr(ops, "gram_impl", None) == "fused"       # the Gram operand (FP16) is derived from a32 inside knn_rows
    a16_all = None
    if plan.world == 1 and fused and hasattr(ops, "center_normalize_f16"):
        a32_all, a16_all, zf_all = ops.center_normalize_f16(img_local, mean)    # centred rows + Gram operand in one pass
        hi_all = lo_all = None
    elif plan.world == 1:
        a32_all, hi_all, lo_all, zf_all = ops.center_normalize(img_local, mean, want_split=not fused)
    else:
        fused_sharded = fused and hasattr(ops, "knn_rows_fused") and len(set(plan.counts)) == 1
        arena = ops.rows_arena(plan, int(img_local.shape[1])) if (fused_sharded and hasattr(ops, "rows_arena")) else None
        nr = plan.hi - plan.lo
        a16 = None
        if fused_sharded and hasattr(ops, "center_normalize_f16"):
            a32, a16, zf = ops.center_normalize_f16(img_local, mean,
                                                    **({"out_a32": arena["a32"][:nr], "out_a16": arena["a16"][:nr]} if arena is not None else {}))
        else:
            a32, _, _, zf = ops.center_normalize(img_local, mean, want_split=False,
                                                 **({"out_a32": arena["a32"][:nr]} if arena is not None else {}))
        zf_all = plan.all_gather_rows(zf)
        if fused_sharded:
            # the Gram kernel needs the FP16 operand of every pattern (2 B/element), the exact-distance refinement the
            # FP32 rows (4 B/element): both are gathered BESIDE the Gram kernels
            import torch.distributed as dist
            if a16 is None:
                a16 = ops.split_f16(a32, **({"out": arena["a16"][:nr]} if arena is not None else {}))
            a32_all = torch.empty((plan.N,) + tuple(a32.shape[1:]), dtype=a32.dtype, device=a32.device)
            a16_all = torch.empty((plan.N,) + tuple(a16.shape[1:]), dtype=a16.dtype, device=a16.device)
            sym = bool(getattr(ops, "wants_symmetric_gram", None) and ops.wants_symmetric_gram(plan.N, nr, k, plan))
            sample = None
            if sym:
                # the sampling pass of the symmetric form needs a SAMPLE of the patterns only (any subset gives valid
                # limits): gather every stride-th local row first (1/stride of the volume)
                smp = a16[::ops.sym_stride()].contiguous()           # (equal row counts: checked above)
                sample = torch.empty((plan.world * smp.shape[0],) + tuple(smp.shape[1:]), dtype=smp.dtype, device=smp.device)
                dist.all_gather_into_tensor(sample, smp, group=plan.group)
            if arena is not None:
                # copy engines pull the blocks from the peers' symmetric memory: no SMs taken from the Gram kernels
                arena["hdl"].barrier()                   # (device-side, stream-ordered) every rank's rows are complete
                w16 = ops.gather_rows_by_copy_engines(arena, plan, "a16", a16_all)
                w32 = ops.gather_rows_by_copy_engines(arena, plan, "a32", a32_all)
            else:
                # NCCL all-gathers on a second communicator (operations of one communicator execute in issue order:
                # the short collectives of the main group must not queue behind 6.5 GB)
                side = plan.side_group()
                w16 = dist.all_gather_into_tensor(a16_all, a16, group=side, async_op=True)
                w32 = dist.all_gather_into_tensor(a32_all, a32.contiguous(), group=side, async_op=True)
            if sym:
                S2_loc, Nbr_loc = ops.knn_rows_fused(a16_all, a32_all, zf_all, plan.lo, plan.hi, k, plan=plan, a32_ready=w32,
                                                     a16_ready=w16, a16_rows=a16, a16_sample=sample)
            else:
                w16.wait()
                S2_loc, Nbr_loc = ops.knn_rows_fused(a16_all, a32_all, zf_all, plan.lo, plan.hi, k, plan=plan, a32_ready=w32)
            return plan.all_gather_rows(S2_loc), plan.all_gather_rows(Nbr_loc)
        # ship the FP32 unit rows only (4 B/element) and split them locally: half the all-gather volume
        a32_all = plan.all_gather_rows(a32)
        hi_all, lo_all = (None, None) if fused else ops.split_bf16(a32_all)
    S2_loc, Nbr_loc = ops.knn_rows(hi_all, lo_all, a32_all, zf_all, plan.lo, plan.hi, k, plan=plan,
                                   **({"a16_all": a16_all} if a16_all is not None else {}))
    S2_all = plan.all_gather_rows(S2_loc)
    Nbr_all = plan.all_gather_rows(Nbr_loc)
    return S2_all, Nbr_all


def check_scalars(scalars_host):
    """Raise like diffusion.py:302 when the kernel scale is degenerate."""
    status = int(scalars_host[3])
    if status == 1:
        raise Exception("S2_Max = 0")
    if status == 2:
        raise Exception("s2_to_w_matrix: row index out of range (S2 has fewer rows than the reference indexes)")


def markov_stage(ops, plan, S2_all, Nbr_all, eps, med_row=4, idx_off=1, validate=True):
    """diffusion.py:291-325, 89-105, 244-260 fused.  Mutates S2_all in place like the reference.
    -> P_rows, dinv_sqrt [N] f64, scalars [5].  P_rows is the rank's row block of P_ep: dense
    [hi-lo, ldp] f32, or -- backend option `operator` = "sparse", or "auto" beyond
    DENSE_OPERATOR_BYTES per rank -- the FP64 list form (`SparseMarkov`), which never forms N x N."""
    scalars = ops.s2_scalars(S2_all, eps, med_row, idx_off)
    if validate:
        check_scalars(ops.read_small(scalars) if hasattr(ops, "read_small") else scalars.cpu().numpy())
    ops.s2_threshold_(S2_all, scalars)
    if hasattr(ops, "wants_sparse_operator") and ops.wants_sparse_operator(plan.hi - plan.lo, plan.N, int(S2_all.shape[1])):
        P_rows, dis = ops.build_markov_sparse(S2_all, Nbr_all, scalars, plan.lo, plan.hi)
    else:
        P_rows, dis = ops.build_markov(S2_all, Nbr_all, scalars, plan.lo, plan.hi)
    return P_rows, dis, scalars


def eigen_stage(ops, plan, P_rows, dis, nev=NEV, validate=True, tol=None, stats=None, seed=0, dense_below=600):
    """diffusion.py:56-81: top-`nev` largest-magnitude eigenpairs, sort descending, scale by
    lambda and D^-1/2.  -> Ps [N,nev] f64 tensor, Lambda [nev] f64 tensor."""
    N = plan.N
    if nev >= N - 1:
        raise Exception("No spare-matrix for eigs-calculation")   # diffusion.py:59-60
    dev = P_rows.device
    sparse = not isinstance(P_rows, torch.Tensor)            # ops_cuda.SparseMarkov

    def matvec(X):
        X = X.contiguous()
        Y = ops.sparse_block_matvec(P_rows, X) if sparse else ops.block_matvec(P_rows, N, X)
        return plan.all_gather_rows(Y)

    if N <= dense_below:
        # tiny problems (unit-test sizes): the Krylov space is the whole space; apply the operator to I
        eye = torch.eye(N, dtype=torch.float64, device=dev)
        cols = [matvec(eye[:, j:j + 8].contiguous()) for j in range(0, N, 8)]
        Pd = torch.cat(cols, dim=1)
        Pd = 0.5 * (Pd + Pd.T)
        # full eigendecomposition of the small matrix by the library's own Householder + QL solver (host; the
        # same routine that extracts Ritz pairs inside v3s_eig_topk) -- no LAPACK / cuSOLVER on the path
        if hasattr(ops, "sym_eig_host"):
            w_np, V_np = ops.sym_eig_host(Pd.cpu().numpy())
        else:                                                 # CPU stand-in of the tests
            w_np, V_np = np.linalg.eigh(Pd.cpu().numpy())
        idx = np.argsort(-np.abs(w_np), kind="stable")[:nev]
        lam = w_np[idx]
        V = torch.as_tensor(np.ascontiguousarray(V_np[:, idx]), dtype=torch.float64, device=dev)
        if stats is not None:
            stats.update({"matvecs": len(cols), "dense": True})
    elif hasattr(ops, "lanczos_step") and N >= 64:
        y_loc = ops.empty((plan.hi - plan.lo, 8), torch.float64)       # reused every step: no allocation in the loop
        w_all = ops.empty((N, 8), torch.float64) if plan.world > 1 else None

        def matvec_f32(Xf):
            if sparse:
                raise Exception("v3s: the sparse Markov operator needs the fused mat-vec path (symmetric memory unavailable?)")
            return plan.all_gather_rows(ops.block_matvec_f32(P_rows, N, Xf, out=y_loc), out=w_all)
        tol_ = 1e-8 if tol is None else tol
        # reduction + recurrence + cross-rank exchange inside the mat-vec's own kernels (no NCCL per step)
        fused = ops.fused_operator(P_rows, N, plan) if (hasattr(ops, "fused_operator") and plan.world <= 8
                                                        and min(plan.counts) >= 1
                                                        and getattr(ops, "use_fused_matvec", True)) else None

        def run_cabi():
            # the whole solver behind one C-ABI call (v3s_eig_topk: block Lanczos + Chebyshev filter + Ritz extraction
            # driven by the library, no Python / LAPACK in the loop).  -> result, or None when it did not converge
            # within its Krylov cap (the host-driven solver below then takes over: it restarts and deflates)
            lam_, Y_, res_, info = fused.eig_topk(nev, tol=tol_, seed=seed)
            if stats is not None:
                stats["matvecs"] = stats.get("matvecs", 0) + info["matvecs"]
            if not info["converged"]:
                if stats is not None:
                    stats["cabi_not_converged"] = True
                return None
            if stats is not None:
                n_mv = stats["matvecs"]
                stats.update(info)
                stats["matvecs"] = n_mv
                stats["residual_max"] = float(np.max(res_))
            run.residual = max(run.residual, float(np.max(res_)))
            return lam_, Y_, res_

        def run(lock):
            # (a saturated eigenvalue cluster -- disconnected graph -- needs deflation rounds: host-driven solver)
            if lock is None and fused is not None and getattr(ops, "eig_driver", "python") == "cabi":
                r_ = run_cabi()
                if r_ is not None:
                    return r_
            st_ = {}
            plain_op = fused.plain if fused is not None else ((lambda Xf, Q: matvec(Q)) if sparse else (lambda Xf, Q: matvec_f32(Xf)))
            pre = fused.sync if fused is not None else None
            if N >= 2048 and nev <= 16 and (fused is not None or not sparse):
                r = chebyshev_lanczos_device(ops, matvec_f32, matvec, N, nev, tol=tol_, stats=st_, seed=seed, lock=lock,
                                             fused=fused)
            else:
                r = block_lanczos_device(ops, plain_op, N, nev, tol=tol_, stats=st_, seed=seed, lock=lock, pre_sync=pre)
            # Krylov cap reached without convergence (ARPACK would restart implicitly): restart from the best
            # Ritz block, at most three times; what is still unconverged then is reported / raised below
            restarts, n_mv = 0, st_.get("matvecs", 0)
            while not st_.get("converged", True) and restarts < 3:
                start = r[1][:, :8].contiguous()
                st_ = {}
                r = block_lanczos_device(ops, plain_op, N, nev, tol=tol_, stats=st_, seed=seed, lock=lock, pre_sync=pre,
                                         start_block=start)
                n_mv += st_.get("matvecs", 0)
                restarts += 1
            st_["matvecs"] = n_mv
            st_["restarts"] = restarts
            if stats is not None:
                prev_mv = stats.get("matvecs", 0)
                prev_conv = stats.get("converged", True)
                stats.update(st_)
                stats["matvecs"] = prev_mv + st_.get("matvecs", 0)
                stats["converged"] = bool(prev_conv and st_.get("converged", True))
            run.converged = run.converged and bool(st_.get("converged", True))
            run.residual = max(run.residual, float(st_.get("true_residual_max", st_.get("residual_max", 0.0))))
            return r

        run.converged, run.residual = True, 0.0
        lam, V, _ = topk_with_locking(run, nev, stats=stats)
        if fused is not None:
            fused.sync()
            fused.check()                                  # a peer that never reached a barrier is an error, validated or not
            if stats is not None:
                stats["fused_matvec"] = True
        if not run.converged:
            msg = ("v3s eigen-solver: no convergence to %g within the Krylov cap after 3 restarts (largest residual %.3e); "
                   "scipy's eigs would raise ArpackNoConvergence here" % (tol_, run.residual))
            if validate:
                raise Exception(msg)
            import warnings
            warnings.warn(msg)
        if stats is not None:
            stats["operator"] = "sparse-fp64" if sparse else "dense-fp32"
    else:
        lam, V, _ = block_lanczos_topk(matvec, N, nev, dev, tol=(1e-8 if tol is None else tol), stats=stats, seed=seed)
    V = -V                                                   # diffusion.py:62 (sign is arbitrary upstream)
    if validate:                                             # compute.py:70-72 guards, then clip
        for w_ in lam:
            if not (0 < w_):
                raise Exception("negative eigen_value")
            if not (w_ < 1 + 1e-12 + 1e-6):
                raise Exception("too large an eigen_value " + str(w_))
        lam = np.clip(lam, 0.0, 1.0)
    order = np.flip(np.argsort(lam)).astype(int)             # diffusion.py:74-76
    lam = lam[order]
    lam_t = torch.as_tensor(lam.copy(), dtype=torch.float64, device=dev)
    V = V[:, torch.as_tensor(order.copy(), device=dev)]
    Ps = (V * lam_t[None, :]) * dis[:, None]                 # diffusion.py:78-81
    return Ps.contiguous(), lam_t


def fit_stage(ops, plan, Ps, Y, polar_mode=0, check=True):
    """diffusion.py:109-183 (a-priori branch).  Ps [N,10], Y = Rotation_R.T [N,9] replicated.
    -> dict(c, recon_pre, rproj, quat_est, quat_true, sigma_T)."""
    N = plan.N
    g = ops.fit_gram(Ps, Y)
    c = ops.fit_solve(g)
    recon, rproj, qe, qt = ops.fit_project(Ps, Y, c, polar_mode)
    part = ops.sigma_pairs(qe, qt, plan.lo, plan.hi)
    plan.all_reduce_sum(part)
    sigma_t = (180.0 / math.pi) * 2.0 * part / (N * (N - 1))
    out = {"c": c, "recon_pre": recon, "rproj": rproj, "quat_est": qe, "quat_true": qt, "sigma_T": sigma_t}
    if check:
        s = float(ops.read_small(sigma_t.reshape(1))[0]) if hasattr(ops, "read_small") else float(sigma_t.item())
        out["sigma_T_host"] = s
        if s > 60:
            raise Exception("Sigma_T = " + str(s))           # diffusion.py:175-176
    return out


def run_pipeline(ops, plan, ed, rot9_local, rot_y_all, n_p, k, eps, polar_mode=0, check=True, stats=None,
                 validate=True, on_images=None, photons_per_pattern=None, noise_seed=0, on_lists=None):
    """Full path on device.  ed [n,n,n] f32; rot9_local [hi-lo,9] f64 row-major Rm of the rank's
    patterns; rot_y_all [N,9] f64 = Rotation_R.T (the reference's column-major unfold, transposed)."""
    # stage boundaries on the stream (bench.py: stage time minus the stage's kernels = exchanges + host gaps)
    tic = getattr(ops, "_tic", lambda name: None)
    toc = getattr(ops, "_toc", lambda h: None)
    h = tic("stage1_patterns")
    img = ops.synth_patterns(ed, rot9_local, n_p)
    if photons_per_pattern is not None:                      # BASELINE config 5: Poisson photon noise
        ops.poisson_noise_(img, photons_per_pattern, noise_seed, plan.lo)
    toc(h)
    if on_images is not None:
        on_images(img)                                       # e.g. start the host transfer of the patterns now
    h = tic("stage2_knn")
    S2, Nbr = knn_stage(ops, plan, img, k)
    toc(h)
    S2_out = S2.clone()                                      # pre-mutation copy for callers that want it
    h = tic("stage3_markov_eigen")
    P_rows, dis, scalars = markov_stage(ops, plan, S2, Nbr, eps, validate=validate)
    knn_overflow = getattr(ops, "knn_overflow", None)
    if validate and knn_overflow is not None:
        # rows whose near-tie band did not fit the candidate buffers (dense ties: noisy or duplicate-heavy data):
        # for those the exact-top-k containment proof does not hold (markov_stage has just synchronised anyway)
        n_over = int(ops.read_small(knn_overflow)[0]) if hasattr(ops, "read_small") else int(knn_overflow.item())
        if n_over:
            import warnings
            warnings.warn("v3s kNN: %d candidate lists saturated (more near-ties around the k-th neighbour than the buffers "
                          "hold); their neighbours are the best of an approximate ranking" % n_over)
    if on_lists is not None:
        on_lists(S2, Nbr)                                    # final (thresholded) kNN lists: their host transfer can start now
    Ps, lam = eigen_stage(ops, plan, P_rows, dis, validate=validate, stats=stats)
    toc(h)
    h = tic("stage4_fit")
    fit = fit_stage(ops, plan, Ps, rot_y_all, polar_mode=polar_mode, check=check)
    toc(h)
    out = {"Images_local": img, "S2": S2, "S2_unmutated": S2_out, "N": Nbr, "Ps": Ps, "Lambda": lam, "dinv_sqrt": dis,
           "scalars": scalars, "knn_overflow": knn_overflow, "P_rows": P_rows}
    out.update(fit)
    return out
