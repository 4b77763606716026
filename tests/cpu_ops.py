"""CPU stand-in for `v3s.ops_cuda.CudaOps`, built on the oracle (test infrastructure only).

It lets the host logic of the product (pipeline.py sharding + collectives, eig.py block Lanczos,
the stage adapters) run on CPU tensors -- single process or world_size-2 gloo -- so that logic is
covered by the `-m "not gpu"` suite.  It is never importable from the product package.
"""
import numpy as np
import torch

import v3s_oracle as orc


def _np(t):
    return t.detach().cpu().numpy()


class _SparseRows:
    """Stand-in for ops_cuda.SparseMarkov: the rank's rows of P_ep as a scipy CSR matrix (FP64)."""

    def __init__(self, P_rows_csr, r0, r1, N):
        self.P, self.r0, self.r1, self.N = P_rows_csr, r0, r1, N
        self.device = torch.device("cpu")


class OracleOps:
    name = "oracle-cpu"
    device = torch.device("cpu")
    operator = "dense"

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype)

    def zeros(self, shape, dtype):
        return torch.zeros(shape, dtype=dtype)

    def to_device(self, x, dtype=None):
        t = torch.as_tensor(x)
        return (t.to(dtype) if dtype is not None else t).contiguous()

    def synth_patterns(self, ed, rot9, n_p, geom=None):
        n = ed.shape[0]
        R9 = _np(rot9).reshape(-1, 3, 3).transpose(0, 2, 1).reshape(-1, 9).T   # back to the reference unfold
        imgs = orc.generate_from_rotations(R9, n, n_p) if R9.shape[1] else np.zeros((0, n_p * n_p))
        return torch.from_numpy(imgs.astype(np.float32))

    def column_sums(self, img):
        return img.to(torch.float64).sum(0)

    def center_normalize(self, img, mean, want_split=True):
        x = img.to(torch.float64) - mean[None, :]
        nrm = torch.sqrt((x * x).sum(1))
        zf = (nrm == 0).to(torch.int32)
        a = torch.where(nrm[:, None] > 0, x / nrm[:, None], torch.zeros_like(x)).to(torch.float32)
        hi = a.to(torch.bfloat16) if want_split else None
        lo = (a - hi.to(torch.float32)).to(torch.bfloat16) if want_split else None
        return a, hi, lo, zf

    def knn_rows(self, hi_all, lo_all, a32_all, zero_flag, r0, r1, k, plan=None):
        A = _np(a32_all).astype(np.float64)
        if k > A.shape[0]:
            raise Exception("Number of nearest neighbors to preserve is too high.")
        nn = np.sum(A * A, axis=1)
        sq = np.maximum(nn[r0:r1, None] + nn[None, :] - 2 * (A[r0:r1] @ A.T), 0)
        B = 2 * np.arcsin(np.minimum(0.5 * np.sqrt(sq), 1.0))      # == arccos(<a,b>) for unit rows
        zf = _np(zero_flag).astype(bool)
        B[zf[r0:r1], :] = 0
        B[:, zf] = 0
        idx = np.argsort(B, axis=1, kind="stable")[:, :k]
        return torch.from_numpy(np.take_along_axis(B, idx, 1)), torch.from_numpy(idx.astype(np.int32))

    def round_distance_dense(self, a32, zero_flag):
        A = _np(a32).astype(np.float64)
        B = np.arccos(np.clip(A @ A.T, -1, 1))
        return torch.from_numpy(B)

    def knn_from_dense(self, B, k):
        S2, N = orc.knn_from_round_distance(_np(B), k)
        return torch.from_numpy(S2), torch.from_numpy(N.astype(np.int32))

    def s2_scalars(self, S2, eps, med_row=4, idx_off=1):
        try:
            s2max, thr = orc.s2_scalars(_np(S2), med_row, idx_off)
            return torch.tensor([s2max, thr, eps * s2max, 0.0, 0.0], dtype=torch.float64)
        except Exception:
            return torch.tensor([0.0, 0.0, 0.0, 1.0, 0.0], dtype=torch.float64)

    def s2_threshold_(self, S2, scalars):
        S2[S2 > scalars[1]] = float("inf")
        return S2

    def _dense_p(self, S2t, Nbr, scalars):
        n, d = S2t.shape
        W = np.zeros((n, n))
        np.add.at(W, (np.repeat(np.arange(n), d), _np(Nbr).astype(np.int64).reshape(-1)),
                  np.exp(-_np(S2t).reshape(-1) / float(scalars[2])))
        P, di = orc.dm_cov(orc.anisotropic_norm(W), dense_D=False)
        return P, di

    def build_markov(self, S2t, Nbr, scalars, r0, r1):
        P, di = self._dense_p(S2t, Nbr, scalars)
        N = P.shape[0]
        ldp = (N + 3) // 4 * 4
        out = torch.zeros((r1 - r0, ldp), dtype=torch.float32)
        out[:, :N] = torch.from_numpy(P[r0:r1].astype(np.float32))
        return out, torch.from_numpy(di)

    def wants_sparse_operator(self, n_rows, N, k=150):
        return self.operator == "sparse"

    def build_markov_sparse(self, S2t, Nbr, scalars, r0, r1):
        import scipy.sparse as sps
        P, di = self._dense_p(S2t, Nbr, scalars)
        return _SparseRows(sps.csr_matrix(P[r0:r1]), r0, r1, P.shape[0]), torch.from_numpy(di)

    def sparse_block_matvec(self, sp, X, out=None):
        return torch.from_numpy(sp.P @ _np(X))

    def w_dense(self, S2t, Nbr, eps_eff):
        n, d = S2t.shape
        W = np.zeros((n, n))
        np.add.at(W, (np.repeat(np.arange(n), d), _np(Nbr).astype(np.int64).reshape(-1)), np.exp(-_np(S2t).reshape(-1) / eps_eff))
        return torch.from_numpy(W)

    def anisotropic_norm_dense(self, W):
        return torch.from_numpy(orc.anisotropic_norm(_np(W).copy()))

    def dm_cov_dense(self, W):
        P, di = orc.dm_cov(_np(W).copy(), dense_D=False)
        return torch.from_numpy(P), torch.from_numpy(di)

    def block_matvec(self, P_rows, N, X):
        return (P_rows[:, :N].to(torch.float32) @ X.to(torch.float32)).to(torch.float64)

    def fit_gram(self, Ps, Y):
        X = Ps[:, 1:10]
        return torch.cat([(X.T @ X).reshape(-1), (X.T @ Y).reshape(-1)])

    def fit_solve(self, g162):
        g = _np(g162)
        return torch.from_numpy(np.linalg.pinv(g[:81].reshape(9, 9)) @ g[81:].reshape(9, 9))

    def fit_project(self, Ps, Y, c, polar_mode):
        recon = _np(Ps[:, 1:10] @ c)
        n = recon.shape[0]
        U, _, Vh = np.linalg.svd(recon.reshape(n, 3, 3))
        if polar_mode == 0:
            Rp = U @ Vh.transpose(0, 2, 1)
        elif polar_mode == 1:
            Rp = U @ Vh
        else:
            Rp = orc.project_so3(recon.reshape(n, 3, 3))
        qe = orc.rot_mat_to_quat_batch(Rp).T.copy()
        qt = orc.rot_mat_to_quat_batch(_np(Y).reshape(n, 3, 3)).T.copy()
        return torch.from_numpy(recon), torch.from_numpy(Rp.reshape(n, 9)), torch.from_numpy(qe), torch.from_numpy(qt)

    def sigma_pairs(self, qe, qt, r0, r1):
        a = np.arccos(np.clip(_np(qe)[r0:r1] @ _np(qe).T, 0, 1))
        b = np.arccos(np.clip(_np(qt)[r0:r1] @ _np(qt).T, 0, 1))
        return torch.tensor([np.abs(a - b).sum()], dtype=torch.float64)

    # device-driver Lanczos interface (mirrors CudaOps.lanczos_*), plain torch on CPU
    def alloc_xf(self, N):
        return torch.zeros((N, 8), dtype=torch.float32)

    def block_matvec_f32(self, P_rows, N, Xf, out=None):
        Y = (P_rows[:, :N].to(torch.float32) @ Xf).to(torch.float64)
        if out is not None:
            out.copy_(Y)
            return out
        return Y

    def lanczos_alloc(self, N, max_dim, max_steps):
        return {"N": N, "max_dim": max_dim, "V": torch.zeros((max_dim, N), dtype=torch.float64),
                "Xf": torch.zeros((N, 8), dtype=torch.float32), "T": torch.zeros((max_steps, 128), dtype=torch.float64),
                "status": torch.zeros((1,), dtype=torch.int32)}

    def lanczos_snapshot(self, st, s0, s1):
        return (st["T"][s0:s1].reshape(-1).clone().numpy(), int(st["status"].item()))

    def lanczos_snapshot_wait(self, snap):
        return snap

    def lanczos_start(self, st, W0, row0=0):
        Q, _ = torch.linalg.qr(W0)
        W0.copy_(Q)
        st["V"][row0:row0 + 8] = Q.T
        st["Xf"].copy_(Q.to(torch.float32))

    def lanczos_step(self, st, m, W, step):
        Vm = st["V"][:m]
        H = Vm @ W
        W -= Vm.T @ H
        H2 = Vm @ W
        W -= Vm.T @ H2
        H += H2
        A = 0.5 * (H[m - 8:m] + H[m - 8:m].T)
        Q, R = torch.linalg.qr(W)
        sg = torch.sign(torch.diagonal(R))
        Q, R = Q * sg[None, :], R * sg[:, None]          # positive diagonal, like a Cholesky factor
        W.copy_(Q)
        if m + 8 <= st["max_dim"]:
            st["V"][m:m + 8] = Q.T
        st["Xf"].copy_(Q.to(torch.float32))
        st["T"][step] = torch.cat([A, R], dim=0).reshape(-1)

    def cheb_combine(self, PY, Yk, Ykm1, alpha, beta, gamma, Y_out, Xf_out):
        v = alpha * PY + beta * Yk
        if Ykm1 is not None:
            v = v + gamma * Ykm1
        Y_out.copy_(v)
        if Xf_out is not None:
            Xf_out.copy_(v.to(torch.float32))

    def poisson_noise_(self, img, photons_per_pattern, seed=0, row0_global=0):
        # stand-in: numpy Poisson keyed by (seed, global row) -- sharding-independent like the kernel
        a = img.numpy()
        for r in range(a.shape[0]):
            rng = np.random.default_rng([int(seed), int(row0_global) + r])
            s = a[r].sum()
            if s > 0:
                a[r] = rng.poisson(a[r] * (photons_per_pattern / s))
        return img

    def split_bf16(self, a32):
        hi = a32.to(torch.bfloat16)
        return hi, (a32 - hi.to(torch.float32)).to(torch.bfloat16)
