"""Rows (f) of the scope table that sit next to the hot path:

* stage hand-off files -- the Octave / MATLAB implementations exchange every intermediate through
  files (`src/octave/DiffusionMap3dOrientations.m:10-27`: Images, S2, N, Axis, ...); here one `.npz`
  with the result map of `OrientationRecoverySolver.solve`, so a run can be resumed or cross-checked;
* a *correct* orientation metric next to the bug-compatible Sigma_T: per-pattern geodesic error of
  the recovered rotations after the best global alignment (the diffusion map fixes orientations only
  up to one global rotation / reflection of the frame).
"""
import numpy as np
import torch

from . import _util

RESULT_KEYS = ('Ps', 'Lambda', 'c0', 'c', 'S2', 'N', 'Images', 'R', 'Sigma_T', 'Axis', 'Quat')


def save_stage_outputs(path, result):
    """Write the result map of `solve` (solution.py:49-61 keys) to a compressed .npz."""
    missing = [k for k in RESULT_KEYS if k not in result]
    if missing:
        raise Exception('result map lacks ' + str(missing))
    np.savez_compressed(path, **{k: np.asarray(result[k]) for k in RESULT_KEYS})


def load_stage_outputs(path):
    with np.load(path) as f:
        return {k: f[k] for k in RESULT_KEYS}


def orientation_errors_deg(Recon, R_true):
    """Per-pattern geodesic error (degrees) between the fitted rotations and the truth.

    Recon: (N,9) rows `Ps[:,1:] @ c` (any near-rotation 3x3, row-major as in diffusion.py:155);
    R_true: (9,N) in the reference's column-major unfold.  Each fitted matrix is projected to the
    nearest rotation (SVD, det fixed), one global rotation A minimising sum ||R_est_i - R_true_i A||_F
    is removed (orthogonal Procrustes), then angle_i = acos((tr(R_est_i^T R_true_i A) - 1)/2).
    Device arithmetic (batched 3x3 in FP64); returns numpy (N,)."""
    ops = _util.get_ops()
    M = _util.to_dev(Recon, torch.float64).reshape(-1, 3, 3)
    Rt = _util.to_dev(R_true, torch.float64).T.reshape(-1, 3, 3)      # same (transposed) convention as the fit target
    U, _, Vh = torch.linalg.svd(M)
    det = torch.linalg.det(U @ Vh)
    U = U.clone()
    U[:, :, 2] *= det[:, None]
    Re = U @ Vh
    # global alignment: A = argmin sum ||Re_i - Rt_i A||  ->  polar factor of sum Rt_i^T Re_i
    C = (Rt.transpose(1, 2) @ Re).sum(0)
    Uc, _, Vc = torch.linalg.svd(C)
    d = torch.linalg.det(Uc @ Vc)
    Uc = Uc.clone()
    Uc[:, 2] *= d
    A = Uc @ Vc
    tr = torch.einsum('nij,nij->n', Re, Rt @ A)
    ang = torch.rad2deg(torch.arccos(torch.clamp((tr - 1) / 2, -1, 1)))
    return ang.cpu().numpy()
