"""Block-Lanczos eigen-solver for the largest-magnitude eigenpairs of the symmetric Markov
matrix P_ep -- the replacement for `scipy.sparse.linalg.eigs(P_ep, k)` (ARPACK, which='LM') at
diffusion.py:61.

Host-side driver.  The only O(N^2) work per step is the caller's block mat-vec (the HBM-bound
CUDA kernel, `CudaOps.block_matvec`, FP32 matrix x FP64 block); the Lanczos basis, the two-pass
block Gram-Schmidt re-orthogonalisation, the Cholesky-QR of the new block and the Ritz extraction
are FP64 (an all-FP32 Lanczos misses the 1e-6 eigenvalue target).  The projected
matrix is block tridiagonal; its few extreme eigenpairs come from LAPACK's banded solver on the
host every `check_every` steps (a (b+1) x m band, a few kB).

P_ep is exactly symmetric, so a symmetric Lanczos targeting largest |lambda| returns the same
eigenpairs as ARPACK's non-symmetric driver, up to the sign / basis freedom the reference itself
has (it starts ARPACK from a random vector).  Here the start block is seeded,
so results are reproducible run to run.
"""
import contextlib

import numpy as np
import scipy.linalg as sla
import torch

import time

try:                                     # small LAPACK calls are faster and far more predictable on one thread
    from threadpoolctl import ThreadpoolController as _TPC
except Exception:                        # pragma: no cover
    _TPC = None
_tp_controller = None
HOST_TIMERS = {"ritz_s": 0.0, "wait_s": 0.0}     # accumulated host time (diagnostics, bench.py)


def _one_thread():
    """Context limiting BLAS/LAPACK to one thread.  The controller (a scan of the loaded shared
    libraries) is built once; entering the context afterwards only calls the libraries' setters."""
    global _tp_controller
    if _TPC is None:
        return contextlib.nullcontext()
    if _tp_controller is None:
        _tp_controller = _TPC()
    return _tp_controller.limit(limits=1)


def _chol_qr(W):
    """Thin QR of a tall block by two rounds of Cholesky-QR (FP64)."""
    R_tot = None
    for _ in range(2):
        G = W.T @ W
        L, info = torch.linalg.cholesky_ex(G)
        # W <- W L^-T
        W = torch.linalg.solve_triangular(L, W.T, upper=False).T.contiguous()
        R = L.T
        R_tot = R if R_tot is None else R @ R_tot
    return W, R_tot


def _ritz_from_band(band, m, b, nev, both_ends=True):
    t0 = time.perf_counter()
    with _one_thread():
        r = _ritz_from_band_impl(band, m, b, nev, both_ends)
    HOST_TIMERS["ritz_s"] += time.perf_counter() - t0
    return r


def _ritz_from_band_impl(band, m, b, nev, both_ends=True):
    """Largest-|lambda| eigenpairs of the m x m block-tridiagonal matrix held in lower band
    storage band[i, j] = T[j + i, j].

    Eigenvalues come from LAPACK's banded solver without vectors (O(m^2 b)); the few wanted
    vectors from two rounds of banded inverse iteration (O(m b^2) each) followed by a
    Rayleigh-Ritz step inside their span, which also resolves (near-)degenerate clusters.
    Accumulating the full orthogonal reduction (dsbevx with vectors) would cost O(m^3) per check.
    """
    Tb = band[:, :m].copy()
    for i in range(1, b + 1):
        Tb[i, max(m - i, 0):] = 0.0
    want = min(nev, m)
    if m <= 320:
        # small projected matrices: dense LAPACK on the m x m matrix (dsyevr on the wanted end only when
        # the caller targets the algebraically largest) beats the banded driver with vectors
        T = np.zeros((m, m))
        for i in range(b + 1):
            ii = np.arange(m - i)
            T[ii + i, ii] = Tb[i, : m - i]
        if both_ends or m <= want:
            w, S = sla.eigh(T, lower=True, check_finite=False, overwrite_a=True)
            _ritz_from_band_impl.last_min = float(w[0])
        else:
            w, S = sla.eigh(T, lower=True, check_finite=False, overwrite_a=True, subset_by_index=[m - want, m - 1])
        idx = np.argsort(-np.abs(w), kind="stable")[:want]
        return w[idx], S[:, idx]
    w = sla.eig_banded(Tb, lower=True, eigvals_only=True, select="i", select_range=(m - want, m - 1))
    if both_ends:      # largest |lambda| may sit at the negative end ('LM' semantics of eigs)
        w = np.concatenate([sla.eig_banded(Tb, lower=True, eigvals_only=True, select="i", select_range=(0, want - 1)), w])
    theta = w[np.argsort(-np.abs(w), kind="stable")[:want]]
    # general band storage for solve_banded: ab[b + i - j, j] = T[i, j]
    ab = np.zeros((2 * b + 1, m))
    for i in range(0, b + 1):
        ab[b + i, : m - i] = Tb[i, : m - i]          # sub-diagonal i
        if i:
            ab[b - i, i:] = Tb[i, : m - i]           # super-diagonal i (symmetric)
    scale = max(np.abs(theta).max(), 1e-300)
    rng = np.random.default_rng(12345)
    X = rng.standard_normal((m, want))
    for j in range(want):
        sh = ab.copy()
        sh[b] -= theta[j] + 4e-13 * scale * (1 + j)   # tiny offset keeps the shifted matrix non-singular
        x = X[:, j]
        for _ in range(2):
            x = sla.solve_banded((b, b), sh, x, check_finite=False)
            x /= np.linalg.norm(x)
        X[:, j] = x
    Q, _ = np.linalg.qr(X)
    # T @ Q through the band
    TQ = ab[b][:, None] * Q
    for i in range(1, b + 1):
        TQ[i:] += Tb[i, : m - i][:, None] * Q[: m - i]
        TQ[: m - i] += Tb[i, : m - i][:, None] * Q[i:]
    Hs = Q.T @ TQ
    wz, Z = np.linalg.eigh(0.5 * (Hs + Hs.T))
    idx = np.argsort(-np.abs(wz), kind="stable")
    return wz[idx], Q @ Z[:, idx]


_FILL_IDX = {}


def _fill_band(band, b, m_of_block, blk):
    """Columns [m_of_block-b, m_of_block) of the lower band storage from the step's [A_j; B_j] stack
    (A_j symmetric b x b, B_j upper triangular): band[i, col_c] = stack[c + i, c]."""
    idx = _FILL_IDX.get(b)
    if idx is None:
        c = np.arange(b)[None, :]
        i = np.arange(b + 1)[:, None]
        r = c + i
        idx = (r, np.broadcast_to(c, r.shape), (r < b) | ((r - b) <= c))
        _FILL_IDX[b] = idx
    r, c, keep = idx
    band[:, m_of_block - b:m_of_block] = np.where(keep, blk[r, c], 0.0)


def _next_check(steps, res_max, target, prev, max_gap):
    """Schedule the next Ritz extraction from the observed convergence rate (it is host work)."""
    gap = 4
    if prev is not None and res_max < prev[1]:
        rate = np.log(prev[1] / res_max) / (steps - prev[0])
        gap = int(np.ceil(0.85 * np.log(res_max / target) / rate))
    return steps + max(1, min(max_gap, gap))


def block_lanczos_device(ops, op, N, nev, tol=1e-8, max_dim=None, first_check=6, seed=0, max_gap=16,
                         stats=None, start_block=None, both_ends=True, lock=None, want_theta_min=False, pre_sync=None):
    """Same algorithm as `block_lanczos_topk`, with every per-step operation enqueued on the GPU by
    the C-ABI (`v3s_lanczos_start` / `v3s_lanczos_step`, csrc/lanczos.cu): the host only launches,
    and reads back the small (A_j, B_j) blocks when it extracts Ritz pairs.

    op(Xf, Q): the operator applied to the current block -- Xf is the backend's FP32 copy of the
    block (opaque layout, what `ops.block_matvec_f32` consumes), Q the same block in FP64 [N,8];
    returns [N,8] float64 (all-gathered when sharded), which the step overwrites with the next block.
    lock: optional [L,N] orthonormal rows (L a multiple of 8, zero rows allowed) that every block is
    kept orthogonal to -- converged eigenvectors of earlier runs (see `topk_with_locking`).
    pre_sync: optional callable enqueuing a cross-rank barrier before the first step (operators that
    store into peers' buffers, `FusedOperator.sync`).
    """
    b = 8
    if pre_sync is not None:
        pre_sync()
    if max_dim is None:
        max_dim = min(N, 1024)
    max_dim = max(b, (min(max_dim, N) // b) * b)
    max_steps = max_dim // b
    L = 0 if lock is None else int(lock.shape[0])
    st = ops.lanczos_alloc(N, max_dim + L, max_steps)
    if L:
        st["V"][:L] = lock
    if start_block is None:
        gen = torch.Generator(device="cpu")
        gen.manual_seed(seed)
        W0 = torch.randn((N, b), generator=gen, dtype=torch.float64).to(ops.device)
    else:
        W0 = start_block.clone()
    if L:
        W0 -= lock.T @ (lock @ W0)
        W0 -= lock.T @ (lock @ W0)
    ops.lanczos_start(st, W0, L)
    Q = W0                               # FP64 copy of the current block
    m, steps, copied = b, 0, 0           # m counts the Krylov basis only (rows L .. L+m of V)
    band = np.zeros((b + 1, max_dim))
    next_check, prev, n_checks = min(first_check, max_steps), None, 0
    w = S = res = None
    def do_step():
        nonlocal Q, steps
        W = op(st["Xf"], Q)
        ops.lanczos_step(st, L + m, W, steps)
        Q = W
        steps += 1

    n_spec = 0
    _ritz_from_band_impl.last_min = None
    history = []                         # (step, max residual estimate) of every check (diagnostics)
    while True:
        do_step()
        last = (m + b > max_dim)
        if steps >= next_check or last:
            # snapshot the new (A_j, B_j) blocks asynchronously, keep the GPU busy with one speculative
            # step while the host extracts Ritz pairs, then wait for the snapshot only.  (More than one
            # speculative step hides more host time at intermediate checks but is wasted device time in
            # front of the next stage at the last one: measured net loss at cfg 2.)
            s_chk, m_chk = steps, m
            snap = ops.lanczos_snapshot(st, copied, s_chk)
            if not last and m + 2 * b <= max_dim:
                m += b
                do_step()
                n_spec += 1
            t0 = time.perf_counter()
            blks, status = ops.lanczos_snapshot_wait(snap)
            HOST_TIMERS["wait_s"] += time.perf_counter() - t0
            blks = blks.reshape(-1, 2 * b, b)
            if status != 0:
                raise Exception("block Lanczos broke down (Cholesky-QR of a rank-deficient block)")
            for i_, blk in enumerate(blks):
                _fill_band(band, b, (copied + i_ + 1) * b, blk)
            copied = s_chk
            w, S = _ritz_from_band(band, m_chk, b, nev, both_ends)
            res = np.linalg.norm(blks[-1][b:] @ S[m_chk - b:m_chk, :], axis=0)
            if not np.all(np.isfinite(res)):
                raise Exception("block Lanczos broke down (non-finite residuals)")
            target = tol * max(np.abs(w).max(), 1e-300)
            n_checks += 1
            history.append((s_chk, float(res.max() / max(np.abs(w).max(), 1e-300))))
            if res.max() <= target or last or m + b > max_dim:
                m = m_chk
                break
            next_check = max(steps + 1, _next_check(s_chk, res.max(), target, prev, max_gap))
            prev = (s_chk, res.max())
        m += b
    St = torch.from_numpy(np.ascontiguousarray(S)).to(ops.device)
    Y = st["V"][L:L + m].T @ St
    if stats is not None and want_theta_min:
        if both_ends and m <= 320 and getattr(_ritz_from_band_impl, "last_min", None) is not None:
            stats["theta_min"] = _ritz_from_band_impl.last_min     # the last check diagonalised this very matrix
        else:
            Tb = band[:, :m].copy()
            for i in range(1, b + 1):
                Tb[i, max(m - i, 0):] = 0.0
            with _one_thread():
                stats["theta_min"] = float(sla.eig_banded(Tb, lower=True, eigvals_only=True, select="i", select_range=(0, 0))[0])
    if stats is not None:
        stats.update({"matvecs": steps, "block": b, "krylov_dim": m, "ritz_checks": n_checks, "speculative_steps": n_spec,
                      "residual_max": float(res.max()), "check_history": history,
                      "converged": bool(res.max() <= tol * max(np.abs(w).max(), 1e-300)), "driver": "device"})
    return w, Y, res


def chebyshev_lanczos_device(ops, matvec_f32, matvec_f64, N, nev, tol=1e-8, degree=6, probe_steps=10, seed=0,
                             stats=None, lock=None, safe_lower=False, fused=None):
    """Largest eigenpairs of the symmetric operator P by block Lanczos on a Chebyshev polynomial
    of P (device driver).  Clustered top eigenvalues (gaps ~1e-4 at N = 10^4) make plain Lanczos
    need a Krylov space of ~600 vectors, whose Ritz extraction is O(m^2 b) host work per check;
    Lanczos on T_d((P - c)/e), with [c-e, c+e] = [-1, cut] the unwanted part of the spectrum,
    converges in ~4x fewer steps for ~1.5x more (HBM-bound, host-free) mat-vecs (degree 6 and the
    probe-estimated lower end minimise mat-vecs x 80 us + steps x 100 us at cfg 2).

      1. probe: `probe_steps` plain block-Lanczos steps.  Ritz values are lower bounds of the
         eigenvalues (Cauchy interlacing), so cut = theta_nev(probe) <= lambda_nev: the wanted
         eigenvalues are never inside the damped interval.
      2. Lanczos on p(P), started from the probe's best Ritz block.
      3. Rayleigh-Ritz of P itself on the converged subspace (eigenvalues of P, not of p(P)).

    matvec_f32: Xf [N,8] f32 -> P Xf [N,8] f64;  matvec_f64: X [N,b<=8] f64 -> P X.
    fused: optional `FusedOperator` -- the mat-vecs of phases 1 and 2 then run through it (reduction,
    recurrence and the cross-rank exchange inside the kernels, one host call per filter application).
    Falls back to plain largest-magnitude Lanczos when the probe shows a negative eigenvalue whose
    magnitude competes with the wanted ones (the filter targets the algebraically largest).
    """
    b = 8
    st1 = {}
    plain_op = (lambda Xf, Q: matvec_f32(Xf)) if fused is None else fused.plain
    pre_sync = None if fused is None else fused.sync
    w0, Y0, res0 = block_lanczos_device(ops, plain_op, N, nev, tol=tol, max_dim=probe_steps * b, seed=seed,
                                        first_check=probe_steps, stats=st1, both_ends=True, lock=lock, want_theta_min=True,
                                        pre_sync=pre_sync)
    n_mv = st1["matvecs"]
    info = {"probe_steps": st1["matvecs"], "degree": degree}
    if st1["converged"]:
        lam, Y, res = w0, Y0, res0
        info.update({"cheb_steps": 0, "krylov_dim": st1["krylov_dim"]})
    elif np.any(w0 < 0) or len(w0) < nev:
        st2 = {}
        lam, Y, res = block_lanczos_device(ops, plain_op, N, nev, tol=tol, seed=seed, stats=st2, lock=lock, pre_sync=pre_sync)
        if stats is not None:
            stats.update(st2)
            stats["matvecs"] += n_mv
            stats["driver"] = "device-plain(LM)"
        return lam, Y, res
    else:
        cut = float(np.min(w0))
        # lower end of the damped interval: the spectrum of a normalised Markov matrix is inside [-1, 1];
        # the probe's smallest Ritz value (extremal values converge first) minus a 15 % safety margin is
        # much tighter and amplifies the wanted end far more per degree.  Guarded below: if an eigenvalue
        # under `lo` existed it would be amplified too and show up in the Rayleigh-Ritz step -> safe rerun.
        lo = -1.0
        if safe_lower is False and "theta_min" in st1:
            lo = max(-1.0, st1["theta_min"] - 0.15 * (float(np.max(w0)) - st1["theta_min"]))
        c, e = 0.5 * (cut + lo), 0.5 * (cut - lo)
        count = [0]
        if fused is None:
            bufs = [ops.empty((N, b), torch.float64) for _ in range(3)]
            Xf = ops.alloc_xf(N)

        def cheb_fused(Xf0, Q):
            count[0] += degree
            return fused.cheb(Xf0, Q, degree, c, e)

        def cheb_op(Xf0, Q):
            # Y0 = Q, Y1 = (P Y0 - c Y0)/e, Y_{k+1} = 2 (P Y_k - c Y_k)/e - Y_{k-1}
            Yk, spare = [t for t in bufs if t.data_ptr() != Q.data_ptr()][:2]
            Ykm1 = Q
            ops.cheb_combine(matvec_f32(Xf0), Ykm1, None, 1.0 / e, -c / e, 0.0, Yk, Xf)
            count[0] += 1
            for i in range(degree - 1):
                ops.cheb_combine(matvec_f32(Xf), Yk, Ykm1, 2.0 / e, -2.0 * c / e, -1.0, spare, Xf)
                Ykm1, Yk, spare = Yk, spare, (Ykm1 if i > 0 else [t for t in bufs if t.data_ptr() not in (Yk.data_ptr(), spare.data_ptr())][0])
                count[0] += 1
            return Yk          # scratch: the Lanczos step overwrites it with Q_{j+1}

        order = np.argsort(-w0, kind="stable")[:b]
        start = Y0[:, torch.as_tensor(order.copy(), device=Y0.device)].contiguous()
        st2 = {}
        _, Y, res = block_lanczos_device(ops, cheb_op if fused is None else cheb_fused, N, nev, tol=tol, seed=seed,
                                         first_check=4, max_gap=4, stats=st2, start_block=start, both_ends=False, lock=lock,
                                         pre_sync=pre_sync)
        n_mv += count[0]
        info.update({"cheb_steps": st2["matvecs"], "krylov_dim": st2["krylov_dim"], "cut": cut,
                     "cheb_check_history": st2.get("check_history"),
                     "ritz_checks": st2["ritz_checks"], "converged": st2["converged"]})
        # Rayleigh-Ritz with P itself
        PY = torch.cat([matvec_f64(Y[:, j:j + b].contiguous()) for j in range(0, Y.shape[1], b)], dim=1)
        n_mv += (Y.shape[1] + b - 1) // b
        H = Y.T @ PY
        # nev x nev eigenproblem on the host (a device syevd is ~10 latency-bound launches plus its own sync)
        Hh = H.cpu().numpy()
        lam_h, Zh = np.linalg.eigh(0.5 * (Hh + Hh.T))
        Z = torch.from_numpy(Zh).to(Y.device)
        lam_t = torch.from_numpy(lam_h).to(Y.device)
        Y = Y @ Z
        R = PY @ Z - Y * lam_t[None, :]
        lam, res = lam_h, torch.linalg.norm(R, dim=0).cpu().numpy()
        idx = np.argsort(-np.abs(lam), kind="stable")
        lam, res = lam[idx], res[idx]
        Y = Y[:, torch.as_tensor(idx.copy(), device=Y.device)]
        info["true_residual_max"] = float(res.max())
        info["damped_interval"] = [lo, cut]
        if lo > -1.0 and float(np.min(lam)) < cut - 1e-6 * max(1.0, abs(cut)):
            # something from the supposedly damped region was amplified: the lower bound was not a bound
            return chebyshev_lanczos_device(ops, matvec_f32, matvec_f64, N, nev, tol=tol, degree=degree,
                                            probe_steps=probe_steps, seed=seed, stats=stats, lock=lock, safe_lower=True,
                                            fused=fused)
    if stats is not None:
        stats.update(info)
        stats.update({"matvecs": n_mv, "block": b, "residual_max": float(np.max(res)), "driver": "device-chebyshev",
                      "converged": bool(info.get("converged", st1["converged"]))})
    return lam, Y, res


def topk_with_locking(run, nev, block=8, rel_tol=1e-6, max_rounds=8, stats=None):
    """A block Krylov space holds at most `block` independent vectors of one eigenspace, so an
    eigenvalue of multiplicity >= block (disconnected kNN graphs: lambda = 1 once per component) can
    hide further copies.  When the converged set contains such a saturated cluster, rerun with the
    found eigenvectors locked (deflated) until a run contributes nothing new to the top `nev`.

    run(lock) -> (lam [nev] numpy, Y [N,nev] tensor, res); lock is None or [L,N] rows."""
    lam, Y, res = run(None)
    rounds = 1
    found_lam, found_Y = lam.copy(), Y
    while rounds < max_rounds:
        a = np.abs(lam)
        saturated = any(int(np.sum(np.abs(a - x) <= rel_tol * max(x, 1e-300))) >= block for x in a)
        if not saturated:
            break
        Lrows = found_Y.T.contiguous()
        pad = (-Lrows.shape[0]) % 8
        if pad:
            Lrows = torch.cat([Lrows, torch.zeros((pad, Lrows.shape[1]), dtype=Lrows.dtype, device=Lrows.device)], dim=0)
        lam2, Y2, res2 = run(Lrows)
        rounds += 1
        floor = np.min(np.abs(lam))
        new = np.abs(lam2) > floor * (1 + rel_tol) + 1e-300
        found_lam = np.concatenate([found_lam, lam2])
        found_Y = torch.cat([found_Y, Y2], dim=1)
        all_lam = np.concatenate([lam, lam2])
        all_res = np.concatenate([res, res2])
        all_Y = torch.cat([Y, Y2], dim=1)
        # prefer the already-held vectors on ties (stable sort on -|lambda|)
        idx = np.argsort(-np.abs(all_lam), kind="stable")[:nev]
        contributed = bool(np.any(idx >= len(lam)) and np.any(new))
        lam, res = all_lam[idx], all_res[idx]
        Y = all_Y[:, torch.as_tensor(idx.copy(), device=all_Y.device)]
        if not contributed:
            break
    if stats is not None:
        stats["locking_rounds"] = rounds
    return lam, Y, res


def block_lanczos_topk(matvec, N, nev, device, block=8, tol=1e-8, max_dim=None, first_check=6, seed=0,
                       max_gap=16, stats=None):
    """Largest-magnitude `nev` eigenpairs of the symmetric operator `matvec`.

    matvec: callable X [N,b] float64 tensor -> P @ X [N,b] float64 tensor (same device).
    Returns (lambda [nev] numpy float64 sorted by decreasing |lambda|, V [N,nev] float64 tensor,
    residual estimates [nev]).  Converged when every Ritz residual estimate
    ||B_j S_last|| <= tol * max|theta|.
    """
    b = int(min(block, max(1, N // 4)))
    if max_dim is None:
        max_dim = min(N, 1024)
    max_dim = max(b, (min(max_dim, N) // b) * b)
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    Q0 = torch.randn((N, b), generator=gen, dtype=torch.float64).to(device)
    Q, _ = _chol_qr(Q0)
    V = torch.zeros((max_dim, N), dtype=torch.float64, device=device)     # basis vectors as rows
    V[:b] = Q.T
    m = b
    band = np.zeros((b + 1, max_dim))
    steps = 0
    n_matvec = 0
    w = S = res = None
    pending = []                         # device [2b,b] blocks not yet copied into `band`
    next_check, prev, n_checks = min(first_check, max(1, max_dim // b)), None, 0
    while True:
        W = matvec(Q)
        n_matvec += 1
        Vm = V[:m]
        H = Vm @ W
        W = W - Vm.T @ H
        H2 = Vm @ W                      # second Gram-Schmidt pass (full re-orthogonalisation)
        W = W - Vm.T @ H2
        H = H + H2
        A = H[m - b:m]
        A = 0.5 * (A + A.T)
        Qn, B = _chol_qr(W)
        pending.append((m, torch.cat([A, B], dim=0)))      # [2b, b]: A_j over B_j
        steps += 1
        last = (m + b > max_dim)
        if steps >= next_check or last:
            blks = torch.stack([t for _, t in pending]).cpu().numpy()   # one small D2H per check
            for (mm, _), blk in zip(pending, blks):
                _fill_band(band, b, mm, blk)
            pending = []
            w, S = _ritz_from_band(band, m, b, nev)
            Bn = blk[b:]
            res = np.linalg.norm(Bn @ S[m - b:m, :], axis=0)
            if not np.all(np.isfinite(res)):
                raise Exception("block Lanczos broke down (non-finite residuals)")
            target = tol * max(np.abs(w).max(), 1e-300)
            n_checks += 1
            if res.max() <= target or last:
                break
            next_check = _next_check(steps, res.max(), target, prev, max_gap)
            prev = (steps, res.max())
        V[m:m + b] = Qn.T
        m += b
        Q = Qn
    St = torch.from_numpy(np.ascontiguousarray(S)).to(device)
    Y = V[:m].T @ St
    if stats is not None:
        stats.update({"matvecs": n_matvec, "block": b, "krylov_dim": m, "ritz_checks": n_checks, "residual_max": float(res.max()),
                      "converged": bool(res.max() <= tol * max(np.abs(w).max(), 1e-300))})
    return w, Y, res
