"""ComputeUtils -- numeric helpers on the path (mirror of the reference's compute.py)."""
import numpy as np
import torch

from . import _util


class ComputeUtils:

    @staticmethod
    def polar_decompose(R, mode=0):
        """compute.py:51-55 on the GPU (fit.cu).  mode 0 = the reference's U @ Vh.T form (depends on
        the SVD sign convention; ours is Jacobi's), 1 = polar factor U @ Vh,
        2 = nearest rotation."""
        ops = _util.get_ops()
        M = _util.to_dev(R, torch.float64).reshape(1, 9)
        Ps = torch.cat([torch.zeros((1, 1), dtype=torch.float64, device=M.device), M], dim=1).contiguous()
        eye = torch.eye(9, dtype=torch.float64, device=M.device)
        Y = torch.eye(3, dtype=torch.float64, device=M.device).reshape(1, 9).contiguous()
        _, rproj, _, _ = ops.fit_project(Ps, Y, eye, mode)
        return _util.back(rproj.reshape(3, 3), R)

    @staticmethod
    def lsq_linear(A, b):
        """compute.py:103-112: M = ((b^T A) pinv(A^T A))^T.  9-column inputs use the fit kernels;
        other shapes (the reference's 4x2 test) go through FP64 device matmuls."""
        ops = _util.get_ops()
        At = _util.to_dev(A, torch.float64)
        bt = _util.to_dev(b, torch.float64)
        if At.shape[1] == 9 and bt.shape[1] == 9:
            Ps = torch.cat([torch.zeros((At.shape[0], 1), dtype=torch.float64, device=At.device), At], dim=1).contiguous()
            return _util.back(ops.fit_solve(ops.fit_gram(Ps, bt)), A)
        M = ((bt.T @ At) @ torch.linalg.pinv(At.T @ At, hermitian=True)).T
        return _util.back(M.contiguous(), A)

    @staticmethod
    def validate_eigen_values_vectors(A, eigen_values, eigen_vectors, tolerance=1e-14):
        """compute.py:59-79 (host checks on small inputs)."""
        A = np.asarray(A); w = np.asarray(eigen_values); V = np.asarray(eigen_vectors)
        for i in range(w.size):
            if not (0 < w[i]):
                raise Exception('negative eigen_value')
            if not (w[i] < 1 + 1e-12):
                raise Exception('too large an eigen_value ' + str(w[i]))
        for i in range(w.size):
            err = abs(np.max(A @ V[:, i] - w[i] * V[:, i]))
            if not (err < tolerance):
                raise Exception('residual of eigenvalue ' + str(i) + ' = ' + str(err))

    @staticmethod
    def validate_lsq_linear(A, b, M, tolerance=1e-14):
        """compute.py:82-99."""
        A = np.asarray(A); b = np.asarray(b); M = np.asarray(M)
        for i in range(A.shape[0]):
            err = abs(np.max((A[i, :] @ M - b[i, :]).flatten()))
            if not (err < tolerance):
                raise Exception('residual of lsq element ' + str(i) + ' = ' + str(err))
