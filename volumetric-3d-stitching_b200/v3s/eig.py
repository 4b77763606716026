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
import numpy as np
import scipy.linalg as sla
import torch


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


def _ritz_from_band(band, m, b, nev):
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
    if m <= 96:
        w, S = sla.eig_banded(Tb, lower=True)
        idx = np.argsort(-np.abs(w), kind="stable")[:want]
        return w[idx], S[:, idx]
    w_hi = sla.eig_banded(Tb, lower=True, eigvals_only=True, select="i", select_range=(m - want, m - 1))
    w_lo = sla.eig_banded(Tb, lower=True, eigvals_only=True, select="i", select_range=(0, want - 1))
    w = np.concatenate([w_lo, w_hi])
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


def _fill_band(band, b, m_of_block, blk):
    for c in range(b):
        col = m_of_block - b + c
        for i in range(0, b + 1):
            r = c + i                           # row offset inside the [A; B] stack
            band[i, col] = blk[r, c] if (r < b or (r - b) <= c) else 0.0


def _next_check(steps, res_max, target, prev, max_gap):
    """Schedule the next Ritz extraction from the observed convergence rate (it is host work)."""
    gap = 4
    if prev is not None and res_max < prev[1]:
        rate = np.log(prev[1] / res_max) / (steps - prev[0])
        gap = int(np.ceil(0.85 * np.log(res_max / target) / rate))
    return steps + max(1, min(max_gap, gap))


def block_lanczos_device(ops, matvec_f32, N, nev, tol=1e-8, max_dim=None, first_check=6, seed=0, max_gap=16,
                         stats=None):
    """Same algorithm as `block_lanczos_topk`, with every per-step operation enqueued on the GPU by
    the C-ABI (`v3s_lanczos_start` / `v3s_lanczos_step`, csrc/lanczos.cu): the host only launches,
    and reads back the small (A_j, B_j) blocks when it extracts Ritz pairs.

    matvec_f32: Xf [N,8] float32 tensor -> P @ Xf as [N,8] float64 tensor (all-gathered when sharded).
    """
    b = 8
    if max_dim is None:
        max_dim = min(N, 1024)
    max_dim = max(b, (min(max_dim, N) // b) * b)
    max_steps = max_dim // b
    st = ops.lanczos_alloc(N, max_dim, max_steps)
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    W0 = torch.randn((N, b), generator=gen, dtype=torch.float64).to(ops.device)
    ops.lanczos_start(st, W0)
    m, steps, copied = b, 0, 0
    band = np.zeros((b + 1, max_dim))
    next_check, prev, n_checks = min(first_check, max_steps), None, 0
    w = S = res = None
    while True:
        W = matvec_f32(st["Xf"])
        ops.lanczos_step(st, m, W, steps)
        steps += 1
        last = (m + b > max_dim)
        if steps >= next_check or last:
            blks = st["T"][copied:steps].cpu().numpy().reshape(-1, 2 * b, b)     # one small D2H per check
            if int(st["status"].item()) != 0:
                raise Exception("block Lanczos broke down (Cholesky-QR of a rank-deficient block)")
            for i_, blk in enumerate(blks):
                _fill_band(band, b, (copied + i_ + 1) * b, blk)
            copied = steps
            w, S = _ritz_from_band(band, m, b, nev)
            res = np.linalg.norm(blks[-1][b:] @ S[m - b:m, :], axis=0)
            if not np.all(np.isfinite(res)):
                raise Exception("block Lanczos broke down (non-finite residuals)")
            target = tol * max(np.abs(w).max(), 1e-300)
            n_checks += 1
            if res.max() <= target or last:
                break
            next_check = _next_check(steps, res.max(), target, prev, max_gap)
            prev = (steps, res.max())
        m += b
    St = torch.from_numpy(np.ascontiguousarray(S)).to(ops.device)
    Y = st["V"][:m].T @ St
    if stats is not None:
        stats.update({"matvecs": steps, "block": b, "krylov_dim": m, "ritz_checks": n_checks, "residual_max": float(res.max()),
                      "converged": bool(res.max() <= tol * max(np.abs(w).max(), 1e-300)), "driver": "device"})
    return w, Y, res


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
