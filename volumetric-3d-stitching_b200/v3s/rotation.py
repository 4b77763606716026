"""Rotation -- SO(3) sampling and conversions (mirror of the reference's rotation.py).

O(N) host-side set-up arithmetic, vectorised numpy float64; the O(N^2) all-pairs geodesic error
runs on the GPU (fit.cu).  Bug-compatible with the Python reference where that changes numbers
(SURVEY 0.2): the unit axis is clipped to [0,1] in Axis2RotMat (rotation.py:81).
"""
import math

import numpy as np

from . import _util


class Rotation:

    @staticmethod
    def all_rot_variables(N):
        """rotation.py:18-25 -> [R (9,N), Axis (3,N), Quat (4,N)]."""
        Q = Rotation.uniform_so3_hopf(N)
        Axis = Rotation.Quat2Axis(Q)
        R = Rotation.Axis2RotMatBatch(Axis)
        return [R, Axis, Q]

    @staticmethod
    def uniform_so3_hopf(n_orient):
        """rotation.py:28-56."""
        n1 = round(math.pow(n_orient, (1 / 3)))
        if n_orient != math.pow(n1, 3):
            raise Exception('Input argument should be the cube of a positive integer.')
        Psi_ = (2 * np.pi) * np.linspace(0, 1, n1 + 1)[0:n1]
        Theta_ = np.arccos(np.linspace(1, -1, n1)[0:n1])
        Phi_ = (2 * np.pi) * np.linspace(0, 1, n1 + 1)[0:n1]
        Psi_ = Psi_ - np.mean(Psi_)
        Theta_ = Theta_ - np.mean(Theta_) + (np.pi / 2)
        Phi_ = Phi_ - np.mean(Phi_)
        Phi, Psi, Theta = np.meshgrid(Phi_, Psi_, Theta_, indexing='ij')    # compute.py:26-42
        Psi = Psi.reshape((1, n_orient))
        Theta = Theta.reshape((1, n_orient))
        Phi = Phi.reshape((1, n_orient))
        Q = np.vstack([np.cos(Theta / 2) * np.cos(Psi / 2),
                       np.sin(Theta / 2) * np.sin(Phi + Psi / 2),
                       np.sin(Theta / 2) * np.cos(Phi + Psi / 2),
                       np.cos(Theta / 2) * np.sin(Psi / 2)])
        return Rotation.unit_mag_pos(Q)

    @staticmethod
    def unit_mag_pos(q):
        """rotation.py:60-72."""
        q = np.array(q, dtype=float)
        nrm = np.sqrt(q[0] ** 2 + q[1] ** 2 + q[2] ** 2 + q[3] ** 2)
        return q * (np.sign(q[0]) / nrm)[None, :]

    @staticmethod
    def Quat2AxisSingle(Q):
        """rotation.py:130-139."""
        return Rotation.Quat2Axis(np.asarray(Q, dtype=float).reshape(4, 1))[:, 0]

    @staticmethod
    def Quat2Axis(Q):
        """rotation.py:118-139."""
        Q = np.asarray(Q, dtype=float)
        angle = 2 * np.arccos(np.clip(np.abs(Q[0]), 0.0, 1.0))
        nrm = np.sqrt(Q[1] ** 2 + Q[2] ** 2 + Q[3] ** 2)
        with np.errstate(divide='ignore', invalid='ignore'):
            axis = angle[None, :] * (Q[1:4] / nrm[None, :])
        axis[:, angle == 0] = 0.0
        return axis

    @staticmethod
    def Axis2RotMat(Axis, clip_bug=True):
        """rotation.py:75-98 (bug-compat clip of the unit axis to [0,1] at :81)."""
        return Rotation._axis_to_rotmat_batch(np.asarray(Axis, dtype=float).reshape(3, 1), clip_bug)[0]

    @staticmethod
    def _axis_to_rotmat_batch(Axis, clip_bug=True):
        theta = np.sqrt(np.sum(Axis ** 2, axis=0))
        with np.errstate(divide='ignore', invalid='ignore'):
            u = Axis / theta[None, :]
        if clip_bug:
            u = np.clip(u, 0.0, 1.0)
        a = np.cos(theta)
        la = 1 - a
        b = np.sin(theta)
        m, n, p = u
        Rm = np.empty((Axis.shape[1], 3, 3))
        Rm[:, 0, 0] = a + m ** 2 * la; Rm[:, 0, 1] = m * n * la - p * b; Rm[:, 0, 2] = m * p * la + n * b
        Rm[:, 1, 0] = n * m * la + p * b; Rm[:, 1, 1] = a + n ** 2 * la; Rm[:, 1, 2] = n * p * la - m * b
        Rm[:, 2, 0] = p * m * la - n * b; Rm[:, 2, 1] = p * n * la + m * b; Rm[:, 2, 2] = a + p ** 2 * la
        Rm[theta == 0] = np.eye(3)
        return Rm

    @staticmethod
    def Axis2RotMatBatch(Axis, clip_bug=True):
        """rotation.py:102-114 -- column c holds Rm.T.flatten()."""
        Rm = Rotation._axis_to_rotmat_batch(np.asarray(Axis, dtype=float), clip_bug)
        return np.ascontiguousarray(Rm.transpose(0, 2, 1).reshape(-1, 9).T)

    @staticmethod
    def random_rotations(N, seed=0):
        """Proper i.i.d. uniform rotations from normalised Gaussian quaternions (the reference's
        'random' mode, src/octave/UniformSO3PDF.m:1-7).  -> R (9,N) reference unfold, Quat (4,N)."""
        rng = np.random.default_rng(seed)
        q = Rotation.unit_mag_pos(rng.standard_normal((4, N)))
        w, x, y, z = q
        Rm = np.empty((N, 3, 3))
        Rm[:, 0, 0] = 1 - 2 * (y * y + z * z); Rm[:, 0, 1] = 2 * (x * y - z * w); Rm[:, 0, 2] = 2 * (x * z + y * w)
        Rm[:, 1, 0] = 2 * (x * y + z * w); Rm[:, 1, 1] = 1 - 2 * (x * x + z * z); Rm[:, 1, 2] = 2 * (y * z - x * w)
        Rm[:, 2, 0] = 2 * (x * z - y * w); Rm[:, 2, 1] = 2 * (y * z + x * w); Rm[:, 2, 2] = 1 - 2 * (x * x + y * y)
        return np.ascontiguousarray(Rm.transpose(0, 2, 1).reshape(N, 9).T), q

    @staticmethod
    def rot_mat_to_axis(R):
        """rotation.py:165-182."""
        R = np.asarray(R, dtype=float)
        x = R[2, 1] - R[1, 2]
        y = R[0, 2] - R[2, 0]
        z = R[1, 0] - R[0, 1]
        r_2sin = np.linalg.norm(np.array([x, y, z]))
        if r_2sin != 0:
            theta = np.arctan2(r_2sin, np.trace(R) - 1)
            return (theta / r_2sin) * np.array([x, y, z]).reshape((3, 1))
        return np.zeros((3, 1))

    @staticmethod
    def axis_to_quat(Axis):
        """rotation.py:185-202."""
        Axis = np.asarray(Axis, dtype=float)
        if Axis.shape[0] != 3:
            print('Error in size of Axis.')
            return
        th = np.sqrt(np.sum(Axis ** 2, axis=0))
        Q = np.zeros((4, Axis.shape[1]))
        Q[0] = np.cos(th / 2)
        with np.errstate(divide='ignore', invalid='ignore'):
            Q[1:4] = (np.sin(th / 2) / th)[None, :] * Axis
        Q[:, th == 0] = np.array([1.0, 0, 0, 0])[:, None]
        return Q

    @staticmethod
    def rot_mat_to_quat(R):
        """rotation.py:205-210."""
        return Rotation.axis_to_quat(Rotation.rot_mat_to_axis(R))

    @staticmethod
    def geodesic_distance_matrix(Q):
        """rotation.py:214-220 (dense N x N; small-N helper, evaluated with torch on the device)."""
        if len(Q.shape) != 2 or Q.shape[0] != 4:
            raise Exception('Quaternion shape was ' + str(Q.shape))
        import torch
        q = _util.to_dev(Q, torch.float64)
        return _util.back(torch.arccos(torch.clip(q.T @ q, 0, 1.0)), Q)

    @staticmethod
    def mean_geodesic_distance_degrees(Q1, Q2):
        """rotation.py:223-225."""
        return (180 / math.pi) * Rotation.mean_geodesic_distance(Q1, Q2)

    @staticmethod
    def mean_geodesic_distance(Q1, Q2):
        """rotation.py:228-244 -- all-pairs sum on the GPU (fit.cu: sigma_pairs_kernel)."""
        for q, name in ((Q1, 'Q1 shape'), (Q2, 'Q2 shape')):
            if len(q.shape) != 2:
                raise Exception(name)
        if Q1.shape != Q2.shape:
            raise Exception('Q1 shape vs. Q2 shape')
        if Q2.shape[0] != 4:
            raise Exception('Q.shape[0]')
        import torch
        ops = _util.get_ops()
        N = Q1.shape[1]
        q1 = _util.to_dev(Q1, torch.float64).T.contiguous()
        q2 = _util.to_dev(Q2, torch.float64).T.contiguous()
        s = ops.sigma_pairs(q1, q2, 0, N)
        return 2 * float(s.item()) / (N * (N - 1))
