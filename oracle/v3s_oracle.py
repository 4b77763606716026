"""CPU oracle: a numpy restatement of the reference's hot path.

TEST INFRASTRUCTURE ONLY.  This module restates, in vectorised float64 numpy, the
algorithm of nurlicht/Volumetric-3D-Stitching's Python pipeline
(`/root/reference/src/python/*.py`).  It is the checker for the CUDA path.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import it; the product package (`v3s`) never does.

Pinning: every function below is checked against the reference's own golden
vectors (`tests/test_oracle_golden.py`, vectors restated from
`distance_test.py`, `diffusion_test.py`, `compute_test.py`, `rotation_test.py`)
and against outputs of the unmodified reference run in the build container
(`tests/golden/make_golden.py` -> `tests/golden/*.npz`).

Each function cites the reference file:line it follows.  Bugs of the Python
port that change numbers are reproduced on purpose and
marked `bug-compat`.

Third-party arithmetic the reference delegates (numpy 2.3.5 / scipy 1.18.1 in
this image, unpinned upstream): BLAS matmul, pocketfft `fftn`, LAPACK
`svd`/`pinv`, `argsort`, scipy `RegularGridInterpolator` (restated here as
`interp3`), scipy `csr_array(...).toarray()` (restated as a scatter-add) and
ARPACK `eigs` (restated as a dense symmetric `eigh` + largest-magnitude
selection: P_ep is exactly symmetric, so the two agree up to the sign / basis
ambiguity the reference itself has).
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------- #
# compute.py
# --------------------------------------------------------------------------- #


def meshgrid_3d(x, y, z):
    """compute.py:26-42 -- X[i,j,k]=x[i], Y[i,j,k]=y[j], Z[i,j,k]=z[k] ('ij' order)."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    z = np.asarray(z, dtype=float)
    X, Y, Z = np.meshgrid(x, y, z, indexing="ij")
    return [np.ascontiguousarray(X), np.ascontiguousarray(Y), np.ascontiguousarray(Z)]


def interp3(x1, y1, z1, F1, x2, y2, z2):
    """compute.py:13-22 -- scipy RegularGridInterpolator((x1,y1,z1), F1, 'linear',
    bounds_error=False, fill_value=0) evaluated at (x2,y2,z2).

    Restated algorithm (scipy/interpolate/_rgi.py `_find_indices` +
    `_evaluate_linear`): per axis i = clip(searchsorted(grid, x) - 1, 0, n-2),
    t = (x - grid[i]) / (grid[i+1] - grid[i]); out of bounds iff x < grid[0] or
    x > grid[-1]; value = sum over the 8 corners of prod(weights).
    """
    grids = [np.asarray(g, dtype=float) for g in (x1, y1, z1)]
    q = [np.asarray(a, dtype=float) for a in (x2, y2, z2)]
    shape = q[0].shape
    F1 = np.asarray(F1, dtype=float)
    idx, t, oob = [], [], np.zeros(q[0].size, dtype=bool)
    for g, x in zip(grids, q):
        x = x.reshape(-1)
        n = g.size
        if n == 1:
            # degenerate axis (reference test at n_voxel_1d=1): scipy refuses < 2 points
            raise ValueError("interp3 needs at least 2 grid points per axis")
        i = np.searchsorted(g, x) - 1
        i = np.clip(i, 0, n - 2)
        denom = g[i + 1] - g[i]
        t.append((x - g[i]) / denom)
        idx.append(i)
        oob |= (x < g[0]) | (x > g[-1])
    i0, i1, i2 = idx
    t0, t1, t2 = t
    out = np.zeros(q[0].size)
    for a in (0, 1):
        wa = t0 if a else (1 - t0)
        for b in (0, 1):
            wb = t1 if b else (1 - t1)
            for c in (0, 1):
                wc = t2 if c else (1 - t2)
                out += F1[i0 + a, i1 + b, i2 + c] * wa * wb * wc
    out[oob] = 0.0
    # NaN queries propagate NaN in scipy; the reference path never produces them.
    return out.reshape(shape)


def polar_decompose(R, compat=True):
    """compute.py:51-55.  bug-compat: returns U @ Vh.T (numpy's svd returns Vh, so this
    is NOT the polar factor).  compat=False gives the intended
    U @ Vh (src/octave/PolarDecompose.m:2-3)."""
    U, _, Vh = np.linalg.svd(np.asarray(R, dtype=float))
    return U @ (Vh.T if compat else Vh)


def lsq_linear(A, b):
    """compute.py:103-112 -- M = ((b^T A) pinv(A^T A))^T."""
    A = np.asarray(A, dtype=float)
    b = np.asarray(b, dtype=float)
    return np.transpose(np.matmul(np.matmul(b.T, A), np.linalg.pinv(np.matmul(A.T, A))))


def validate_eigen_values_vectors(A, eigen_values, eigen_vectors, tolerance=1e-14):
    """compute.py:59-79 (guards only)."""
    for w in eigen_values:
        if not (0 < w):
            raise Exception("negative eigen_value")
        if not (w < 1 + 1e-12):
            raise Exception("too large an eigen_value " + str(w))
    for i in range(eigen_values.size):
        v = eigen_vectors[:, i]
        err = abs(np.max(A @ v - eigen_values[i] * v))
        if not (err < tolerance):
            raise Exception("residual of eigenvalue " + str(i) + " = " + str(err))


# --------------------------------------------------------------------------- #
# utils.py comparators (reused by the parity harness)
# --------------------------------------------------------------------------- #


def almost_equal_hybrid(x, y, p):
    """utils.py:76-80, vectorised: (|x|<p and |y|<p) or (y!=0 and |(x-y)/y|<p)."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    both_small = (np.abs(x) < p) & (np.abs(y) < p)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs((x - y) / y) < p
    return both_small | ((y != 0) & rel)


# --------------------------------------------------------------------------- #
# rotation.py
# --------------------------------------------------------------------------- #


def unit_mag_pos(q):
    """rotation.py:60-72 -- normalise columns, flip so q0 >= 0 (sign(0) = 0 kept)."""
    q = np.array(q, dtype=float)
    nrm = np.sqrt(q[0] ** 2 + q[1] ** 2 + q[2] ** 2 + q[3] ** 2)
    return q * (np.sign(q[0]) / nrm)[None, :]


def uniform_so3_hopf(n_orient):
    """rotation.py:28-56."""
    n1 = round(math.pow(n_orient, 1 / 3))
    if n_orient != math.pow(n1, 3):
        raise Exception("Input argument should be the cube of a positive integer.")
    Psi_ = (2 * np.pi) * np.linspace(0, 1, n1 + 1)[0:n1]
    Theta_ = np.arccos(np.linspace(1, -1, n1)[0:n1])
    Phi_ = (2 * np.pi) * np.linspace(0, 1, n1 + 1)[0:n1]
    Psi_ = Psi_ - np.mean(Psi_)
    Theta_ = Theta_ - np.mean(Theta_) + (np.pi / 2)
    Phi_ = Phi_ - np.mean(Phi_)
    Phi, Psi, Theta = meshgrid_3d(Phi_, Psi_, Theta_)
    Psi = Psi.reshape((1, n_orient))
    Theta = Theta.reshape((1, n_orient))
    Phi = Phi.reshape((1, n_orient))
    Q = np.vstack([
        np.cos(Theta / 2) * np.cos(Psi / 2),
        np.sin(Theta / 2) * np.sin(Phi + Psi / 2),
        np.sin(Theta / 2) * np.cos(Phi + Psi / 2),
        np.cos(Theta / 2) * np.sin(Psi / 2),
    ])
    return unit_mag_pos(Q)


def quat_to_axis(Q):
    """rotation.py:118-139 -- angle = 2 acos(clip |q0|), axis = angle * q[1:4]/|q[1:4]|."""
    Q = np.asarray(Q, dtype=float)
    angle = 2 * np.arccos(np.clip(np.abs(Q[0]), 0.0, 1.0))
    nrm = np.sqrt(Q[1] ** 2 + Q[2] ** 2 + Q[3] ** 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        axis = angle[None, :] * (Q[1:4] / nrm[None, :])
    axis[:, angle == 0] = 0.0
    return axis


def axis_to_rotmat(axis, clip_bug=True):
    """rotation.py:75-98.  bug-compat: the unit axis is clipped to [0,1]
    (rotation.py:81).  clip_bug=False is src/octave/Axis2RotMat.m."""
    axis = np.asarray(axis, dtype=float).reshape(3)
    theta = np.linalg.norm(axis)
    if theta == 0:
        return np.eye(3)
    u = axis / theta
    if clip_bug:
        u = np.clip(u, 0.0, 1.0)
    a = math.cos(theta)
    la = 1 - a
    b = math.sin(theta)
    m, n, p = u
    return np.array([
        [a + m ** 2 * la, m * n * la - p * b, m * p * la + n * b],
        [n * m * la + p * b, a + n ** 2 * la, n * p * la - m * b],
        [p * m * la - n * b, p * n * la + m * b, a + p ** 2 * la],
    ])


def axis_to_rotmat_batch(Axis, clip_bug=True):
    """rotation.py:102-114 -- column c holds Rm.T.flatten() (column-major unfold)."""
    N = Axis.shape[1]
    R = np.zeros((9, N))
    for c in range(N):
        R[:, c] = axis_to_rotmat(Axis[:, c], clip_bug).T.flatten()
    return R


def all_rot_variables(N, clip_bug=True):
    """rotation.py:18-25 -> [R (9,N), Axis (3,N), Quat (4,N)]."""
    Q = uniform_so3_hopf(N)
    Axis = quat_to_axis(Q)
    R = axis_to_rotmat_batch(Axis, clip_bug)
    return [R, Axis, Q]


def random_rotations(N, seed=0):
    """Proper i.i.d. uniform SO(3) rotations (normalised Gaussian quaternions, the
    reference's own 'random' mode, src/octave/UniformSO3PDF.m:1-7).  Not part of the Python
    reference; shared input generator for tests and benches.  Returns R (9,N) in the
    reference's column-major unfold, Quat (4,N)."""
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((4, N))
    q = unit_mag_pos(q)
    w, x, y, z = q
    Rm = np.empty((N, 3, 3))
    Rm[:, 0, 0] = 1 - 2 * (y * y + z * z)
    Rm[:, 0, 1] = 2 * (x * y - z * w)
    Rm[:, 0, 2] = 2 * (x * z + y * w)
    Rm[:, 1, 0] = 2 * (x * y + z * w)
    Rm[:, 1, 1] = 1 - 2 * (x * x + z * z)
    Rm[:, 1, 2] = 2 * (y * z - x * w)
    Rm[:, 2, 0] = 2 * (x * z - y * w)
    Rm[:, 2, 1] = 2 * (y * z + x * w)
    Rm[:, 2, 2] = 1 - 2 * (x * x + y * y)
    R = np.ascontiguousarray(Rm.transpose(0, 2, 1).reshape(N, 9).T)  # column c = Rm.T.flatten()
    return R, q


def rotate_structure_index(F, R, _cache={}):
    """rotation.py:143-162 -- ED_rot[i,j,k] = Tri(F)(R @ (x,y,z)) with z=U[i], x=U[j], y=U[k]."""
    N = max(F.shape)
    key = N
    if key not in _cache:
        U = (np.linspace(1, N, N) - (N + 1) / 2) / (N / 2)
        z, x, y = meshgrid_3d(U, U, U)
        _cache[key] = (U, np.vstack([x.reshape(1, -1), y.reshape(1, -1), z.reshape(1, -1)]), x.shape)
    U, P, shape = _cache[key]
    Q = np.matmul(np.asarray(R, dtype=float), P)
    return interp3(U, U, U, F, Q[0].reshape(shape), Q[1].reshape(shape), Q[2].reshape(shape))


def rot_mat_to_axis(R):
    """rotation.py:165-182."""
    x = R[2, 1] - R[1, 2]
    y = R[0, 2] - R[2, 0]
    z = R[1, 0] - R[0, 1]
    r_2sin = np.linalg.norm(np.array([x, y, z]))
    if r_2sin != 0:
        theta = np.arctan2(r_2sin, np.trace(R) - 1)
        return (theta / r_2sin) * np.array([x, y, z]).reshape((3, 1))
    return np.zeros((3, 1))


def axis_to_quat(Axis):
    """rotation.py:185-202."""
    N = Axis.shape[1]
    Q = np.zeros((4, N))
    for c in range(N):
        a = Axis[:, c]
        theta = np.linalg.norm(a)
        if theta == 0:
            Q[:, c] = [1, 0, 0, 0]
        else:
            a = (math.sin(theta / 2) / theta) * a
            Q[:, c] = [math.cos(theta / 2), a[0], a[1], a[2]]
    return Q


def rot_mat_to_quat(R):
    """rotation.py:205-210."""
    return axis_to_quat(rot_mat_to_axis(R))


def rot_mat_to_quat_batch(Rb):
    """Vectorised rotation.py:165-210 over Rb (N,3,3) -> (4,N)."""
    x = Rb[:, 2, 1] - Rb[:, 1, 2]
    y = Rb[:, 0, 2] - Rb[:, 2, 0]
    z = Rb[:, 1, 0] - Rb[:, 0, 1]
    v = np.stack([x, y, z])
    s = np.sqrt(x * x + y * y + z * z)
    tr = Rb[:, 0, 0] + Rb[:, 1, 1] + Rb[:, 2, 2]
    theta = np.arctan2(s, tr - 1)
    with np.errstate(divide="ignore", invalid="ignore"):
        axis = (theta / s)[None, :] * v
    axis[:, s == 0] = 0.0
    th = np.sqrt(np.sum(axis ** 2, axis=0))
    Q = np.zeros((4, Rb.shape[0]))
    Q[0] = np.cos(th / 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        Q[1:4] = (np.sin(th / 2) / th)[None, :] * axis
    Q[:, th == 0] = np.array([1.0, 0, 0, 0])[:, None]
    return Q


def geodesic_distance_matrix(Q):
    """rotation.py:214-220 -- acos(clip(Q^T Q, 0, 1)) (bug-compat: clip to [0,1], not abs)."""
    if len(Q.shape) != 2 or Q.shape[0] != 4:
        raise Exception("Quaternion shape was " + str(Q.shape))
    return np.arccos(np.clip(np.matmul(Q.T, Q), 0, 1.0))


def mean_geodesic_distance(Q1, Q2, block=2048):
    """rotation.py:228-244, evaluated in row blocks (same sum, bounded memory)."""
    N = Q1.shape[1]
    tot = 0.0
    for s in range(0, N, block):
        a = np.arccos(np.clip(Q1[:, s:s + block].T @ Q1, 0, 1.0))
        b = np.arccos(np.clip(Q2[:, s:s + block].T @ Q2, 0, 1.0))
        tot += np.sum(np.abs(a - b))
    return 2 * tot / (N * (N - 1))


def mean_geodesic_distance_degrees(Q1, Q2):
    """rotation.py:223-225."""
    return (180 / math.pi) * mean_geodesic_distance(Q1, Q2)


# --------------------------------------------------------------------------- #
# projection.py
# --------------------------------------------------------------------------- #


def synthesize_3d(n_voxel_1d):
    """projection.py:75-91 -- analytic object; array axes are (z, x, y)."""
    N = n_voxel_1d
    U = (np.linspace(1, N, N) - (N + 1) / 2) / (N / 2)
    z, x, y = meshgrid_3d(U, U, U)
    A = x / 0.47
    B = y / 0.37
    C = z / 0.29
    F = 1 - 0.4 * ((A - 0.15) ** 2 + (B + 0.2) ** 2 + (C - 0.1) * (A - 0.15) * (B + 0.2))
    F[np.cos(20 * np.pi * (x - z - 0.2) * abs(y + z + 0.1) * abs(z - 0.3)) < 0.2] = 0
    F[A ** 2 + B ** 2 + C ** 2 > 1] = 0
    F[F < 0] = 0
    return {"ED": F, "Grid_3D": {"x": 2e-7 * x, "y": 2e-7 * y, "z": 2e-7 * z}}


def experiment_parameters(n_voxel_1d, N_p=None):
    """projection.py:95-117.  N_p defaults to the reference's 1 + 2 n; a free detector size
    is the builder's extension and keeps Width = 75e-6 * 1024."""
    e = {"n_voxel_1d": n_voxel_1d, "Pixel": 75e-6, "zD": 0.5, "Lambda": 2 * 1e-9}
    e["N_p"] = (1 + 2 * n_voxel_1d) if N_p is None else int(N_p)
    e["SuperPixel"] = e["Pixel"] * (1024 / e["N_p"])
    e["Width"] = e["SuperPixel"] * e["N_p"]
    return {"N_P_NoBin": 1024, "Experiment": e}


def camera_x_y(Experiment):
    """projection.py:183-190."""
    Width = Experiment["Width"]
    N = Experiment["N_p"]
    U = (Width / (N - 1)) * (np.linspace(1, N, N) - (N + 1) / 2)
    return np.meshgrid(U, U, indexing="xy")


def q_and_mask(Experiment):
    """projection.py:164-179."""
    Lambda, zD, Width = Experiment["Lambda"], Experiment["zD"], Experiment["Width"]
    Cx, Cy = camera_x_y(Experiment)
    T = Lambda * np.sqrt(Cx ** 2 + Cy ** 2 + zD ** 2)
    Q = {"x": Cx / T, "y": Cy / T, "z": (zD / T - 1 / Lambda)}
    mask = (Cx ** 2 + Cy ** 2) > (Width / 2) ** 2
    return [Q, mask]


def fourier_scaled_axis(Number, Length):
    """projection.py:121-126."""
    d = Length / (Number - 1)
    Nyquist = 0.5 / d
    N1 = (Number - 1) / 2
    return [Nyquist, np.arange(-N1, N1 + 1) * (2 * Nyquist / Number)]


def fourier_axes_1d(Grid_3D):
    """projection.py:144-160 + 129-141, the `[2]` element only (1-D k axes)."""
    out = {}
    shape = Grid_3D["x"].shape
    for ax, num in zip(("x", "y", "z"), shape):
        t = Grid_3D[ax].flatten()
        out[ax] = fourier_scaled_axis(num, max(t) - min(t))[1]
    return out


def generate_single_image(ED, R, k, Q, Circle_Index):
    """projection.py:55-71 -> [Image, ED_rot, ED_rot_f, Camera_I]."""
    ED_rot = rotate_structure_index(ED, R)
    ED_rot_f = np.fft.fftshift(abs(np.fft.fftn(ED_rot)))
    Camera_I = interp3(k["x"], k["y"], k["z"], ED_rot_f, Q["x"], Q["y"], Q["z"])
    Camera_I[Circle_Index] = 0
    return [Camera_I.reshape((Camera_I.size,)), ED_rot, ED_rot_f, Camera_I]


def generate_from_rotations(R9, n_voxel_1d, N_p=None):
    """projection.py:17-51 with an explicit rotation array R9 (9,N) (column-major unfold)."""
    Experiment = experiment_parameters(n_voxel_1d, N_p)["Experiment"]
    Protein = synthesize_3d(n_voxel_1d)
    Q, mask = q_and_mask(Experiment)
    k = fourier_axes_1d(Protein["Grid_3D"])
    N = R9.shape[1]
    Images = np.empty((N, Experiment["N_p"] ** 2))
    for c in range(N):
        Images[c] = generate_single_image(Protein["ED"], R9[:, c].reshape((3, 3)).T, k, Q, mask)[0]
    if not Images[Images != 0].size > 0:
        raise Exception("Images is an all-zero array.")
    return Images


def generate_direct(R9, n_voxel_1d, N_p=None):
    """Direct-sum form of the patterns (verified there against the reference:
    corr 0.998): I = | sum_ijk F[i,j,k] exp(-2 pi i (M kappa).(a_i,a_j,a_k)) |, M = R Pm,
    Pm:(c0,c1,c2)->(c1,c2,c0), a = 2e-7 U.  Not a reference function: the checker of the optional
    `direct` synthesis mode.  Proper rotations only."""
    n = n_voxel_1d
    Experiment = experiment_parameters(n, N_p)["Experiment"]
    F = synthesize_3d(n)["ED"]
    Q, mask = q_and_mask(Experiment)
    U = (np.linspace(1, n, n) - (n + 1) / 2) / (n / 2)
    a = 2e-7 * U
    kap = np.stack([Q["x"].ravel(), Q["y"].ravel(), Q["z"].ravel()])          # (3, P)
    out = np.empty((R9.shape[1], kap.shape[1]))
    for c in range(R9.shape[1]):
        Rm = R9[:, c].reshape(3, 3).T
        M = Rm[:, [2, 0, 1]]                                                   # (M x) = R (x1, x2, x0)
        kp = M @ kap
        e0 = np.exp(-2j * np.pi * np.outer(kp[0], a))                          # (P, n)
        e1 = np.exp(-2j * np.pi * np.outer(kp[1], a))
        e2 = np.exp(-2j * np.pi * np.outer(kp[2], a))
        t = np.einsum("ijk,pk->pij", F, e2)
        t = np.einsum("pij,pj->pi", t, e1)
        out[c] = np.abs(np.einsum("pi,pi->p", t, e0))
    out[:, mask.ravel()] = 0.0
    return out


def generate(nRot1D, n_voxel_1d, clip_bug=True):
    """projection.py:17-51 -> [Images (N,P), R (9,N), Axis (3,N), Quat (4,N)]."""
    R, Axis, Quat = all_rot_variables(nRot1D ** 3, clip_bug)
    return [generate_from_rotations(R, n_voxel_1d), R, Axis, Quat]


# --------------------------------------------------------------------------- #
# distance.py
# --------------------------------------------------------------------------- #


def remove_offset(A):
    """distance.py:70-71."""
    A = np.asarray(A, dtype=float)
    rep = np.empty(A.shape)          # C-ordered operand, like matlib.repmat: keeps the result's
    rep[:] = np.mean(A, axis=0)      # memory layout (and so BLAS's rounding) equal to the reference's
    return A - rep


def set_norm_to_one(A):
    """distance.py:66-67 (0/0 -> NaN for an all-zero row, as in the reference)."""
    A = np.asarray(A, dtype=float)
    rep = np.empty(A.shape[::-1])
    rep[:] = np.sqrt(np.sum(A ** 2, axis=1))
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.divide(A, rep.T)


def convert_to_round_distance(A):
    """distance.py:51-63 -- arccos(clip(A A^T)) on centred unit rows, NaN->0, symmetrised."""
    A = set_norm_to_one(remove_offset(A))
    with np.errstate(invalid="ignore"):
        B = np.arccos(np.clip(np.matmul(A, A.T), -1, 1))
    B[B != B] = 0
    return (B + B.T) / 2


def knn_from_round_distance(B, k):
    """distance.py:26-48 -- per-row argsort, first k; indices returned as float64."""
    B = np.asarray(B, dtype=float)
    if len(B.shape) != 2 or B.shape[0] != B.shape[1]:
        raise Exception("Non-square round-distance matrix was encountered.")
    n = B.shape[0]
    if k > n:
        raise Exception("Number of nearest neighbors to preserve is too high.")
    S2 = np.empty((n, k))
    N = np.empty((n, k))
    for r in range(n):
        II = np.argsort(B[r, :].reshape(1, n), axis=1).astype(int)[0, range(k)]
        N[r, :] = II
        S2[r, :] = B[r, II]
    return [S2, N]


def knn(A, k):
    """distance.py:17-22."""
    return knn_from_round_distance(convert_to_round_distance(A), k)


# --------------------------------------------------------------------------- #
# diffusion.py
# --------------------------------------------------------------------------- #


def s2_scalars(S2, s2_median_index=4, index_offset=1):
    """diffusion.py:297-310 -- (S2_Max, Dist_Thr).  bug-compat: indexes ROWS of S2 (N,k)
    where Octave indexes columns."""
    n, d = S2.shape
    S2_Max = np.median(S2[min(s2_median_index, d - 1), :])
    if S2_Max == 0:
        S2_Max = np.max(S2)
    if not S2_Max != 0:
        raise Exception("S2_Max = 0")
    Index = 0
    Dist_Max = np.min(S2[d - 1, :])
    while np.max(S2[Index, :]) < Dist_Max:
        Index = Index + 1
    Dist_Thr = np.max(S2[Index - index_offset, :])
    return S2_Max, Dist_Thr


def s2_to_w_matrix(S2, N, Epsilon, s2_median_index=4, index_offset=1):
    """diffusion.py:291-325.  bug-compat: mutates S2 in place (entries > Dist_Thr -> inf);
    duplicate (row, col) pairs sum, as csr_array(...).toarray() does."""
    n, d = S2.shape
    S2_Max, Dist_Thr = s2_scalars(S2, s2_median_index, index_offset)
    S2[S2 > Dist_Thr] = np.inf
    eps = Epsilon * S2_Max
    W = np.zeros((n, n))
    rows = np.repeat(np.arange(n), d)
    cols = np.asarray(N).astype(np.int64).reshape(-1)
    np.add.at(W, (rows, cols), np.exp(-S2.astype(float).reshape(-1) / eps))
    return W


def anisotropic_norm(W):
    """diffusion.py:89-105 -- q = max(rowsum, 1/N); W <- diag(1/q) W diag(1/q); symmetrise.
    The reference multiplies by dense diagonal matrices; scaling rows/columns gives the same
    floats (the extra terms are exact zeros)."""
    n = W.shape[0]
    q = np.sum(W, axis=1)
    q[q < 1 / n] = 1 / n
    qi = np.divide(1, q)
    W = (qi[:, None] * W) * qi[None, :]
    return (W + W.T) / 2


def dm_cov(W, dense_D=True):
    """diffusion.py:244-260 -- d = max(colsum, 1/N); D = diag(1/sqrt(d)); P = D W D; symmetrise.
    Returns [P_ep, D]; D dense (N,N) like the reference, or the diagonal when dense_D=False."""
    n = W.shape[0]
    d = np.sum(W, axis=0)
    d[d < 1 / n] = 1 / n
    di = np.divide(1, np.sqrt(d))
    P = (di[:, None] * W) * di[None, :]
    P = (P + P.T) / 2
    return [P, np.diag(di) if dense_D else di]


def eig_top_lm(P_ep, k):
    """Stand-in for scipy.sparse.linalg.eigs(P_ep, k) (diffusion.py:61; ARPACK, 'LM').
    P_ep is exactly symmetric, so a dense symmetric eigendecomposition gives the same
    eigenpairs up to sign / basis inside degenerate clusters."""
    w, V = np.linalg.eigh(P_ep)
    sel = np.argsort(-np.abs(w), kind="stable")[:k]
    return w[sel], V[:, sel]


def diffusion_coordinate(S2, N, Epsilon, k=10, validate=True, return_all=False):
    """diffusion.py:38-85 -> [Ps (N,k), Lambda (k,)] (+ W, P_ep, d^-1/2 when return_all)."""
    W1 = s2_to_w_matrix(S2, N, Epsilon)
    W = anisotropic_norm(W1)
    P_ep, di = dm_cov(W, dense_D=False)
    if k >= P_ep.shape[0] - 1:
        raise Exception("No spare-matrix for eigs-calculation")
    Lambda, Ps = eig_top_lm(P_ep, k)
    Ps = -Ps
    if validate:
        for w in Lambda:
            if not (0 < w):
                raise Exception("negative eigen_value")
            if not (w < 1 + 1e-12):
                raise Exception("too large an eigen_value " + str(w))
        Lambda = np.clip(Lambda, 0.0, 1.0)
    Index = np.flip(np.argsort(Lambda)).astype(int)
    Lambda = Lambda[Index]
    Ps = Ps[:, Index] * Lambda[None, :]
    Ps = di[:, None] * Ps
    if return_all:
        return [Ps, Lambda, W1, W, P_ep, di]
    return [Ps, Lambda]


def dm_fit_all(Ps, Rotation_R, aPrioriFlag=True, compat_polar=True, check=True, return_all=False):
    """diffusion.py:109-183 (a-priori branch; the other branch is dead upstream, SURVEY 0.2
    item 7) -> [c, c0, Sigma_T]."""
    if not aPrioriFlag:
        raise Exception("aPrioriFlag=False is not runnable in the reference (diffusion.py:231)")
    X = np.asarray(Ps, dtype=float)[:, 1::]
    Y = np.asarray(Rotation_R, dtype=float).T
    c0 = lsq_linear(X, Y)
    c = c0
    Recon_pre = np.matmul(X, c)
    n = X.shape[0]
    U, _, Vh = np.linalg.svd(Recon_pre.reshape(n, 3, 3))
    Vuse = Vh.transpose(0, 2, 1) if compat_polar else Vh
    Recon = np.matmul(U, Vuse)
    Q_1 = rot_mat_to_quat_batch(Recon)
    Q_2 = rot_mat_to_quat_batch(Y.reshape(n, 3, 3))
    Sigma_T = mean_geodesic_distance_degrees(Q_1, Q_2)
    if check and Sigma_T > 60:
        raise Exception("Sigma_T = " + str(Sigma_T))
    if return_all:
        return [c, c0, Sigma_T, Recon_pre, Recon.reshape(n, 9), Q_1, Q_2]
    return [c, c0, Sigma_T]


def project_so3(M):
    """The *correct* nearest-rotation projection (U diag(1,1,det) Vh) used by the parity
    harness on both sides; M is (N,3,3)."""
    U, _, Vh = np.linalg.svd(M)
    det = np.linalg.det(np.matmul(U, Vh))
    U = U.copy()
    U[:, :, 2] *= det[:, None]
    return np.matmul(U, Vh)


def geodesic_angle_deg(Ra, Rb):
    """Per-pattern geodesic angle (degrees) between rotation batches (N,3,3)."""
    tr = np.einsum("nij,nij->n", Ra, Rb)
    return np.degrees(np.arccos(np.clip((tr - 1) / 2, -1, 1)))


def principal_angles(A, B):
    """Principal angles (rad) between the column spaces of A and B."""
    Qa, _ = np.linalg.qr(A)
    Qb, _ = np.linalg.qr(B)
    s = np.linalg.svd(Qa.T @ Qb, compute_uv=False)
    # stable for small angles: sin via the complement
    C = Qb - Qa @ (Qa.T @ Qb)
    sn = np.linalg.svd(C, compute_uv=False)
    sn = np.sort(np.clip(sn, 0, 1))[::-1]
    return np.arcsin(sn) if s.min() > 0.5 else np.arccos(np.clip(np.sort(s), 0, 1))


# --------------------------------------------------------------------------- #
# solution.py
# --------------------------------------------------------------------------- #


def default_parameters():
    """solution.py:79-87."""
    return {"n_voxel_1d": 7, "nRot1D": 8, "k": 24, "Epsilon": 0.7, "plot": True, "aPrioriFlag": True}


def validate_parameters(p):
    """solution.py:65-75 (bug-compat: exact key ORDER required)."""
    return (
        isinstance(p, dict)
        and list(p.keys()) == ["n_voxel_1d", "nRot1D", "k", "Epsilon", "plot", "aPrioriFlag"]
        and isinstance(p["n_voxel_1d"], int) and p["n_voxel_1d"] > 0
        and isinstance(p["k"], int) and p["k"] > 8
        and isinstance(p["nRot1D"], int) and p["nRot1D"] ** 3 - 1 > p["k"]
        and isinstance(p["Epsilon"], float) and p["Epsilon"] > 0
        and isinstance(p["plot"], bool)
        and isinstance(p["aPrioriFlag"], bool)
    )


def solve(parameters, check=True):
    """solution.py:34-61."""
    if not validate_parameters(parameters):
        parameters = default_parameters()
    Images, R, Axis, Quat = generate(parameters["nRot1D"], parameters["n_voxel_1d"])
    S2, N = knn(Images, parameters["k"])
    Ps, Lambda = diffusion_coordinate(S2, N, parameters["Epsilon"])
    c, c0, Sigma_T = dm_fit_all(Ps, R, parameters["aPrioriFlag"], check=check)
    return {"Ps": Ps, "Lambda": Lambda, "c0": c0, "c": c, "S2": S2, "N": N, "Images": Images,
            "R": R, "Sigma_T": np.array([Sigma_T]), "Axis": Axis.T, "Quat": Quat.T}


# ------------------------------------------------------------------------------------------
# fit WITHOUT known orientations.  Parity unpinned for this block: the Python
# reference's branch cannot run (diffusion.py:231 `np.random.rand((81, 1))` and :283 `math.abs`
# raise), Octave is not available here; the functions restate the code as written, minus the two
# defects, and `form="octave"` gives src/octave/CalculateDiffusion.m:136-157.
# ------------------------------------------------------------------------------------------

def or_func(c, PsMod, form="python"):
    """diffusion.py:265-287: rotation-matrix constraint residuals G (r,) for C = c.reshape(9,9) and
    PsMod = Ps[:, 1:10].T (9, r)."""
    r = PsMod.shape[1]
    C = np.asarray(c, dtype=np.float64).reshape(9, 9)
    RB = (C @ PsMod).T.reshape(-1, 3, 3)                      # R = R_Big[:, l].reshape((3, 3))
    T = np.einsum('nki,nkj->nij', RB, RB) - np.eye(3)         # R.T @ R - I
    s2 = np.sum(T * T, axis=(1, 2))
    d = np.abs(np.linalg.det(RB) - 1.0)
    G = np.sqrt(s2 + d) if form == "python" else np.sqrt(s2) + d     # diffusion.py:283 / CalculateDiffusion.m:153
    return np.sqrt(G) / math.sqrt(r)                          # diffusion.py:285


def fit_without_prior(Ps, x0, form="python", max_nfev=250):
    """diffusion.py:124-141: scipy least_squares (trust-region reflective) on or_func from x0 (81,);
    -> c = pinv(c_1D.reshape(9,9)), c0 = pinv((x0/5).reshape(9,9)), final cost."""
    from scipy.optimize import least_squares
    X = np.asarray(Ps, dtype=np.float64)[:, 1:10]
    res = least_squares(or_func, np.asarray(x0, dtype=np.float64).ravel(), args=(X.T, form), method='trf', max_nfev=max_nfev)
    return np.linalg.pinv(res.x.reshape(9, 9)), np.linalg.pinv((np.asarray(x0).ravel() / 5.0).reshape(9, 9)), float(res.cost), res.x
