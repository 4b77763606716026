"""DiffusionMapSolver -- diffusion coordinates and orientation fit (mirror of the reference's
diffusion.py).  `diffusion_coordinate` is the fused path: kNN lists -> FP32 Markov row block
(markov.cu) -> block-Lanczos on the HBM-bound block mat-vec (matvec.cu, eig.py).  The dense
per-function forms (`s2_to_w_matrix`, `anisotropic_norm`, `dm_cov`) are kept for drop-in use.
"""
import numpy as np
import torch

from . import _util
from .pipeline import NEV, ShardPlan, check_scalars, eigen_stage, fit_stage, markov_stage


class DiffusionMapSolver:

    @staticmethod
    def solve(S2, N, R, parameters):
        """diffusion.py:23-34 -> [Ps, Lambda, c0, c, Sigma_T]."""
        print('DiffusionMapSolver.solve started.')
        [Ps, Lambda] = DiffusionMapSolver.diffusion_coordinate(S2, N, parameters['Epsilon'])
        [c, c0, Sigma_T] = DiffusionMapSolver.dm_fit_all(Ps, R, parameters['aPrioriFlag'])
        print('DiffusionMapSolver.solve ended.')
        return [Ps, Lambda, c0, c, Sigma_T]

    @staticmethod
    def _lists_to_dev(S2, N):
        """kNN lists on the device: S2 as FP64 (a CUDA FP64 tensor is used as is, so the thresholding
        mutates the caller's tensor like the reference mutates its array), N as int32."""
        return _util.to_dev(S2, torch.float64), _util.to_dev(N, torch.int32)

    @staticmethod
    def diffusion_coordinate(S2, N, Epsilon, k=10, validate=True, stats=None):
        """diffusion.py:38-85 -> [Ps (N,k), Lambda (k,)].  Mutates S2 in place (entries beyond the
        distance threshold become inf) exactly like the reference (diffusion.py:311)."""
        print('DiffusionMapSolver.diffusion_coordinate started.')
        ops = _util.get_ops()
        S2d, Nd = DiffusionMapSolver._lists_to_dev(S2, N)
        plan = ShardPlan(S2d.shape[0])
        P_rows, dis, _ = markov_stage(ops, plan, S2d, Nd, float(Epsilon))
        if not _util.is_torch(S2):
            np.copyto(S2, S2d.cpu().numpy())       # reference-compat in-place mutation of the caller's array
        Ps, lam = eigen_stage(ops, plan, P_rows, dis, nev=int(k), validate=validate, stats=stats)
        print('DiffusionMapSolver.diffusion_coordinate ended.')
        return [_util.back(Ps, S2), _util.back(lam, S2)]

    @staticmethod
    def s2_to_w_matrix(S2, N, Epsilon, s2_median_index=4, index_offset=1):
        """diffusion.py:291-325 -> dense W (N,N).  Mutates S2 in place."""
        ops = _util.get_ops()
        S2d, Nd = DiffusionMapSolver._lists_to_dev(S2, N)
        sc = ops.s2_scalars(S2d, float(Epsilon), int(s2_median_index), int(index_offset))
        sch = sc.cpu().numpy()
        check_scalars(sch)
        ops.s2_threshold_(S2d, sc)
        if not _util.is_torch(S2):
            np.copyto(S2, S2d.cpu().numpy())
        return _util.back(ops.w_dense(S2d, Nd, float(sch[2])), S2)

    @staticmethod
    def anisotropic_norm(W):
        """diffusion.py:89-105."""
        ops = _util.get_ops()
        return _util.back(ops.anisotropic_norm_dense(_util.to_dev(W, torch.float64)), W)

    @staticmethod
    def dm_cov(W):
        """diffusion.py:244-260 -> [P_ep, D] with D the dense diagonal matrix, as in the reference."""
        ops = _util.get_ops()
        P, dis = ops.dm_cov_dense(_util.to_dev(W, torch.float64))
        return [_util.back(P, W), _util.back(torch.diag(dis), W)]

    @staticmethod
    def dm_fit_all(Ps, Rotation_R, aPrioriFlag, polar_mode=0, return_all=False):
        """diffusion.py:109-183 -> [c, c0, Sigma_T] (c0 is c on the a-priori branch)."""
        print('DiffusionMapSolver.dm_fit_all started.')
        if not aPrioriFlag:
            # the reference's other branch cannot run (diffusion.py:231 raises TypeError)
            raise Exception('aPrioriFlag=False is not runnable in the reference either (diffusion.py:231)')
        ops = _util.get_ops()
        Psd = _util.to_dev(Ps, torch.float64)
        if Psd.shape[1] < 10:
            raise Exception('Ps needs the trivial + 9 non-trivial eigenvectors (10 columns)')
        Y = _util.to_dev(Rotation_R, torch.float64).T.contiguous()
        out = fit_stage(ops, ShardPlan(Psd.shape[0]), Psd, Y, polar_mode=polar_mode, check=True)
        Sigma_T = out['sigma_T_host']
        print('Measure of error in Relative Orientations of All Pairs')
        print('Sigma_All_Pairs: ' + str(Sigma_T) + ' degrees')
        print('DiffusionMapSolver.dm_fit_all ended.')
        c = _util.back(out['c'], Ps)
        if return_all:
            return [c, c, Sigma_T, out]
        return [c, c, Sigma_T]
