"""Rows (f) of the scope table that sit next to the hot path:

* stage hand-off files -- the Octave / MATLAB implementations exchange every intermediate through
  files (`src/octave/DiffusionMap3dOrientations.m:10-27`: Images, S2, N, Axis, ...); here the result
  map of `OrientationRecoverySolver.solve` as one `.npz`, one MATLAB `.mat`, or a folder of Octave
  `save -text` files named like the Octave pipeline's artifacts, so a run can be resumed or
  cross-checked against Octave / MATLAB users' files;
* a *correct* orientation metric next to the bug-compatible Sigma_T: per-pattern geodesic error of
  the recovered rotations after the best global alignment (the diffusion map fixes orientations only
  up to one global rotation / reflection of the frame);
* the orientation fit WITHOUT known orientations (`or_func`, `fit_without_prior`): dead code in the
  Python reference (diffusion.py:124-141, 228-240, 265-287 raise before doing anything), live in
  Octave (src/octave/CalculateDiffusion.m:29-38, 136-157).  `DiffusionMapSolver.dm_fit_all(..., False)`
  keeps raising like the reference; these are separate entry points.
"""
import numpy as np
import torch

from . import _util

RESULT_KEYS = ('Ps', 'Lambda', 'c0', 'c', 'S2', 'N', 'Images', 'R', 'Sigma_T', 'Axis', 'Quat')


OCTAVE_ARTIFACTS = ('Images', 'Axis', 'S2', 'N', 'Ps', 'Lambda', 'c0', 'c')   # files of src/octave/DiffusionMap3dOrientations.m:10-27


def save_stage_outputs(path, result):
    """Write the result map of `solve` (solution.py:49-61 keys).

    path ending in .npz -> one compressed numpy archive; .mat -> one MATLAB v5 file (scipy.io.savemat;
    Octave and MATLAB `load` both read it); a directory (existing, or a path ending in '/') -> one
    Octave `save -text` file per variable, `<dir>/<name>.mat`, the layout of the Octave pipeline's
    `artifacts/` folder.  In the two MATLAB / Octave forms the neighbour indices `N` are written
    1-based, as DistanceMatrix.m produces them; `load_stage_outputs` undoes that."""
    import os
    missing = [k for k in RESULT_KEYS if k not in result]
    if missing:
        raise Exception('result map lacks ' + str(missing))
    arrays = {k: np.asarray(result[k], dtype=np.float64) for k in RESULT_KEYS}
    if path.endswith(os.sep) or os.path.isdir(path):
        os.makedirs(path, exist_ok=True)
        for k in RESULT_KEYS:
            a = arrays[k] + 1.0 if k == 'N' else arrays[k]
            _write_octave_text(os.path.join(path, k + '.mat'), k, a)
    elif path.endswith('.mat'):
        import scipy.io
        arrays['N'] = arrays['N'] + 1.0
        scipy.io.savemat(path, arrays, do_compression=True)
    else:
        np.savez_compressed(path, **arrays)


def load_stage_outputs(path):
    """Inverse of `save_stage_outputs` for all three forms (a directory may also hold files written by the
    Octave pipeline itself: whatever of RESULT_KEYS is present is returned)."""
    import os
    if os.path.isdir(path):
        out = {}
        for k in RESULT_KEYS:
            f = os.path.join(path, k + '.mat')
            if os.path.exists(f):
                out[k] = _read_octave_text(f)
                if k == 'N':
                    out[k] = out[k] - 1.0
        return out
    if path.endswith('.mat'):
        import scipy.io
        m = scipy.io.loadmat(path)
        out = {k: np.asarray(m[k], dtype=np.float64) for k in RESULT_KEYS if k in m}
        if 'N' in out:
            out['N'] = out['N'] - 1.0
        for k in ('Lambda', 'Sigma_T'):                  # loadmat returns 1-D arrays as (1, n)
            if k in out and out[k].ndim == 2 and out[k].shape[0] == 1:
                out[k] = out[k][0]
        return out
    with np.load(path) as f:
        return {k: f[k] for k in RESULT_KEYS}


def _write_octave_text(fname, name, a):
    """One variable in Octave's `save -text` format (matrix / scalar), full double precision."""
    a = np.asarray(a, dtype=np.float64)
    with open(fname, 'w') as f:
        f.write('# Created by v3s\n# name: %s\n' % name)
        if a.ndim == 0:
            f.write('# type: scalar\n%.17g\n' % float(a))
        else:
            a2 = a.reshape(-1, 1) if a.ndim == 1 else a.reshape(a.shape[0], -1)
            f.write('# type: matrix\n# rows: %d\n# columns: %d\n' % a2.shape)
            np.savetxt(f, a2, fmt=' %.17g', delimiter='')
        f.write('\n\n')


def _read_octave_text(fname):
    rows = cols = None
    kind = None
    data = []
    with open(fname) as f:
        for line in f:
            t = line.strip()
            if t.startswith('#'):
                key, _, val = t[1:].partition(':')
                key, val = key.strip(), val.strip()
                if key == 'type':
                    kind = val
                elif key == 'rows':
                    rows = int(val)
                elif key == 'columns':
                    cols = int(val)
            elif t:
                data.extend(float(x) for x in t.split())
    if kind == 'scalar':
        return np.float64(data[0])
    if kind != 'matrix' or rows is None or cols is None or len(data) != rows * cols:
        raise Exception('unsupported or truncated Octave text file: ' + fname)
    a = np.array(data, dtype=np.float64).reshape(rows, cols)
    return a[:, 0] if cols == 1 else a


def orientation_errors_deg(Recon, R_true):
    """Per-pattern geodesic error (degrees) between the fitted rotations and the truth.

    Recon: (N,9) rows `Ps[:,1:] @ c` (any near-rotation 3x3, row-major as in diffusion.py:155);
    R_true: (9,N) in the reference's column-major unfold.  Each fitted matrix is projected to the
    nearest rotation (SVD, det fixed), one global rotation A minimising sum ||R_est_i - R_true_i A||_F
    is removed (orthogonal Procrustes), then angle_i = acos((tr(R_est_i^T R_true_i A) - 1)/2).
    Device arithmetic (batched 3x3 in FP64); returns numpy (N,)."""
    ops = _util.get_ops()
    Rd = _util.to_dev(Recon, torch.float64).reshape(-1, 9)
    Y = _util.to_dev(R_true, torch.float64).T.contiguous()            # same (transposed) convention as the fit target
    Rt = Y.reshape(-1, 3, 3)
    # nearest rotation of every fitted matrix by the library's own projection kernel (fit.cu, polar mode 2: SVD through
    # a 3x3 Jacobi eigen-solver, det fixed): Recon = [0 | Recon] @ I in the kernel's terms
    Psf = torch.cat([torch.zeros((Rd.shape[0], 1), dtype=torch.float64, device=Rd.device), Rd], dim=1).contiguous()
    eye = torch.eye(9, dtype=torch.float64, device=Rd.device)
    _, rproj, _, _ = ops.fit_project(Psf, Y, eye, 2)
    Re = rproj.reshape(-1, 3, 3)
    # global alignment: A = argmin sum ||Re_i - Rt_i A||  ->  polar factor of sum Rt_i^T Re_i (one 3x3, on the host)
    C = torch.einsum('nji,njk->ik', Rt, Re).cpu().numpy()
    Uc, _, Vc = np.linalg.svd(C)
    if np.linalg.det(Uc @ Vc) < 0:
        Uc = Uc.copy()
        Uc[:, 2] *= -1.0
    A = torch.as_tensor(Uc @ Vc, dtype=torch.float64, device=Rd.device)
    tr = torch.einsum('nij,nij->n', Re, Rt @ A)
    ang = torch.rad2deg(torch.arccos(torch.clamp((tr - 1) / 2, -1, 1)))
    return ang.cpu().numpy()


def or_func(c, PsMod, form="python"):
    """DiffusionMapSolver.or_func (diffusion.py:265-287): residuals G (r,) of the rotation-matrix
    constraints for C = c.reshape(9,9) and PsMod = Ps[:, 1:10].T (9, r).  form "octave" evaluates
    src/octave/CalculateDiffusion.m:153 instead of diffusion.py:283 (they differ in a parenthesis)."""
    ops = _util.get_ops()
    X = _util.to_dev(PsMod, torch.float64).T
    Ps = torch.cat([torch.zeros((X.shape[0], 1), dtype=torch.float64, device=X.device), X], dim=1).contiguous()
    cd = _util.to_dev(np.asarray(c, dtype=np.float64).reshape(81) if not _util.is_torch(c) else c.reshape(81), torch.float64)
    G, _ = ops.or_func(cd, Ps, Ps.shape[0], 0 if form == "python" else 1, want_jac=False)
    return _util.back(G, PsMod)


def fit_without_prior(Ps, x0=None, c_scale=5.0, seed=0, max_iter=250, form="python", ftol=1e-10, xtol=1e-10, info=None):
    """The `aPrioriFlag=False` branch of dm_fit_all as written upstream (diffusion.py:124-141;
    CalculateDiffusion.m:29-38): minimise sum or_func(C)^2 over the 81 entries of C from
    x0 = c_scale * rand(81), then c = pinv(C), c0 = pinv(x0 / c_scale).  The residuals and their
    forward-difference Jacobian are one CUDA kernel (csrc/fit.cu `or_func_kernel`); the 81 x 81
    Levenberg-Marquardt system is solved on the host (upstream: scipy 'trf' / lsqnonlin -- another
    trust-region least-squares method, so iterates differ while the minimised functional is the same).
    -> [c (9,9), c0 (9,9), cost]; `info` (dict) receives iterations / final x."""
    ops = _util.get_ops()
    Psd = _util.to_dev(Ps, torch.float64)
    if Psd.shape[1] < 10:
        raise Exception('Ps needs the trivial + 9 non-trivial eigenvectors (10 columns)')
    n = Psd.shape[0]
    if x0 is None:
        x0 = c_scale * np.random.default_rng(seed).random(81)
    x0 = np.asarray(x0, dtype=np.float64).reshape(81).copy()
    fm = 0 if form == "python" else 1
    x = x0.copy()

    def evaluate(xv, jac):
        G, J = ops.or_func(ops.to_device(xv, torch.float64), Psd, n, fm, want_jac=jac)
        if not jac:
            return float((G @ G).item()) * 0.5, None, None
        pack = torch.cat([(J.T @ J).reshape(-1), J.T @ G, (G @ G).reshape(1)]).cpu().numpy()      # one D2H
        return 0.5 * pack[-1], pack[:81 * 81].reshape(81, 81), pack[81 * 81:81 * 81 + 81]

    cost, A, g = evaluate(x, True)
    mu = 1e-3 * max(np.max(np.diag(A)), 1e-300)
    it, nfev = 0, 1
    while it < max_iter:
        it += 1
        D = np.maximum(np.diag(A), 1e-12 * np.max(np.diag(A)))
        accepted = False
        for _ in range(12):
            try:
                dx = np.linalg.solve(A + mu * np.diag(D / np.max(D)), -g)
            except np.linalg.LinAlgError:
                mu *= 4.0
                continue
            new_cost, _, _ = evaluate(x + dx, False)
            nfev += 1
            pred = -(g @ dx) - 0.5 * dx @ (A @ dx)
            rho = (cost - new_cost) / max(pred, 1e-300)
            if new_cost < cost and np.isfinite(new_cost):
                accepted = True
                mu *= max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3)
                break
            mu *= 2.5
        if not accepted:
            break
        small_step = np.linalg.norm(dx) <= xtol * (xtol + np.linalg.norm(x))
        small_gain = (cost - new_cost) <= ftol * cost
        x = x + dx
        cost, A, g = evaluate(x, True)
        nfev += 1
        if small_step or small_gain:
            break
    if info is not None:
        info.update({"iterations": it, "nfev": nfev, "x": x, "x0": x0, "cost": cost})
    c = np.linalg.pinv(x.reshape(9, 9))
    c0 = np.linalg.pinv((x0 / c_scale).reshape(9, 9))
    return [c, c0, cost]
