"""Projector -- (Fourier) projections of the object at given orientations (mirror of the
reference's projection.py).  The per-pattern hot loop (projection.py:35-43, 55-71 and
rotation.py:143-162) is one CUDA kernel (csrc/synth.cu); object / geometry set-up is host numpy.
"""
import numpy as np
import torch

from . import _util
from .rotation import Rotation


class Projector:

    @staticmethod
    def generate(nRot1D, n_voxel_1d):
        """projection.py:17-51 -> [Images (N,P), R (9,N), Axis (3,N), Quat (4,N)] (numpy float64)."""
        print('Projector.generate started.')
        N_loop = nRot1D ** 3
        [R, Axis, Quat] = Rotation.all_rot_variables(N_loop)
        Images = Projector.generate_from_rotations(R, n_voxel_1d)
        print('Projector.generate ended.')
        return [Images, R, Axis, Quat]

    @staticmethod
    def generate_from_rotations(R, n_voxel_1d, N_p=None, as_tensor=False, mode="exact"):
        """The hot loop of `generate` for an explicit rotation array R (9,N) in the reference's
        column-major unfold (any N, free detector size N_p).  Images (N, N_p^2).
        mode="direct" evaluates the analytical amplitude on the rotated q-grid instead (no
        interpolation; proper rotations only; the same image up to ~7 % rel-L2)."""
        ops = _util.get_ops()
        Experiment = Projector.experiment_parameters(n_voxel_1d)['Experiment']
        n_p = Experiment['N_p'] if N_p is None else int(N_p)
        ED = Projector.synsthesize_3d(n_voxel_1d)['ED']
        ed = ops.to_device(ED.astype(np.float32), torch.float32)
        rot9 = Projector.rot9_rowmajor(R, ops)
        img = ops.synth_patterns(ed, rot9, n_p, mode=mode)
        if as_tensor:
            return img
        Images = img.cpu().numpy().astype(np.float64)
        if not Images[Images != 0].size > 0:
            raise Exception('Images is an all-zero array.')           # projection.py:47
        return Images

    @staticmethod
    def add_poisson_noise(Images, photons_per_pattern, seed=0, row0=0):
        """Poisson photon noise (not in the reference; BASELINE config 5).  Images (N,P); returns
        the photon counts in the caller's kind (numpy float64 or device tensor)."""
        ops = _util.get_ops()
        img = _util.to_dev(Images, torch.float32)
        if _util.is_torch(Images) and img.data_ptr() == Images.data_ptr():
            img = img.clone()
        ops.poisson_noise_(img, photons_per_pattern, seed, row0)
        return _util.back(img, Images)

    _ed_cache = {}

    @staticmethod
    def object_on_device(n_voxel_1d, ops):
        """The analytic object (projection.py:75-91) as an FP32 device tensor; it depends on n only, so
        it is built once per size and device (the reference rebuilds it on every call: ~1 ms at n = 31)."""
        key = (int(n_voxel_1d), str(ops.device))
        ed = Projector._ed_cache.get(key)
        if ed is None:
            ed = ops.to_device(Projector.synsthesize_3d(n_voxel_1d)['ED'].astype(np.float32), torch.float32)
            Projector._ed_cache[key] = ed
        return ed

    @staticmethod
    def rot9_rowmajor(R, ops):
        """R (9,N) with column c = Rm.T.flatten()  ->  [N,9] row-major Rm (= R[:,c].reshape(3,3).T,
        projection.py:38) as an FP64 device tensor."""
        if _util.is_torch(R):
            Rt = R.to(device=ops.device, dtype=torch.float64)
            return Rt.T.reshape(-1, 3, 3).transpose(1, 2).reshape(-1, 9).contiguous()
        Rn = np.asarray(R, dtype=np.float64)
        return ops.to_device(np.ascontiguousarray(Rn.T.reshape(-1, 3, 3).transpose(0, 2, 1).reshape(-1, 9)), torch.float64)

    @staticmethod
    def generate_single_image(ED, R, k, Q, Circle_Index):
        """projection.py:55-71 for one rotation matrix R (3,3) -> [Image, None, None, Camera_I].
        The k / Q / Circle_Index arguments are accepted for signature compatibility; the kernel
        derives the same geometry from the object size and the detector size of Q."""
        ops = _util.get_ops()
        ED = np.asarray(ED)
        n_p = np.asarray(Q['x']).shape[0]
        ed = ops.to_device(ED.astype(np.float32), torch.float32)
        rot9 = ops.to_device(np.asarray(R, dtype=np.float64).reshape(1, 9), torch.float64)
        Image = ops.synth_patterns(ed, rot9, n_p)[0].cpu().numpy().astype(np.float64)
        return [Image, None, None, Image.reshape(n_p, n_p)]

    @staticmethod
    def synsthesize_3d(n_voxel_1d):
        """projection.py:75-91 (array axes are (z, x, y))."""
        N = n_voxel_1d
        U = (np.linspace(1, N, N) - (N + 1) / 2) / (N / 2)
        z, x, y = np.meshgrid(U, U, U, indexing='ij')
        A = x / 0.47
        B = y / 0.37
        C = z / 0.29
        F = 1 - 0.4 * ((A - 0.15) ** 2 + (B + 0.2) ** 2 + (C - 0.1) * (A - 0.15) * (B + 0.2))
        F[np.cos(20 * np.pi * (x - z - 0.2) * abs(y + z + 0.1) * abs(z - 0.3)) < 0.2] = 0
        F[A ** 2 + B ** 2 + C ** 2 > 1] = 0
        F[F < 0] = 0
        return {'ED': F, 'Grid_3D': {'x': 2e-7 * x, 'y': 2e-7 * y, 'z': 2e-7 * z}}

    @staticmethod
    def experiment_parameters(n_voxel_1d):
        """projection.py:95-117."""
        parameters = {
            'N_P_NoBin': 1024,
            'Experiment': {'n_voxel_1d': n_voxel_1d, 'Pixel': 75e-6, 'zD': 0.5, 'Lambda': 2 * 1e-9,
                           'N_p': np.nan, 'SuperPixel': np.nan, 'Width': np.nan}
        }
        e = parameters['Experiment']
        e['N_p'] = 1 + 2 * e['n_voxel_1d']
        e['SuperPixel'] = e['Pixel'] * (parameters['N_P_NoBin'] / e['N_p'])
        e['Width'] = e['SuperPixel'] * e['N_p']
        return parameters

    @staticmethod
    def fourier_scaled_axis(Number, Length):
        """projection.py:121-126."""
        d = Length / (Number - 1)
        Nyquist = 0.5 / d
        N1 = (Number - 1) / 2
        k = np.arange(-N1, N1 + 1) * (2 * Nyquist / Number)
        return [Nyquist, k]

    @staticmethod
    def fourier_scaled_axes(Number, Length):
        """projection.py:129-141."""
        [Nyquist_x, k_x_1D] = Projector.fourier_scaled_axis(Number[0], Length['x'])
        [Nyquist_y, k_y_1D] = Projector.fourier_scaled_axis(Number[1], Length['y'])
        [Nyquist_z, k_z_1D] = Projector.fourier_scaled_axis(Number[2], Length['z'])
        k_z, k_x, k_y = np.meshgrid(k_z_1D, k_x_1D, k_y_1D, indexing='ij')
        return [{'x': Nyquist_x, 'y': Nyquist_y, 'z': Nyquist_z},
                {'x': k_x, 'y': k_y, 'z': k_z},
                {'x': k_x_1D, 'y': k_y_1D, 'z': k_z_1D}]

    @staticmethod
    def fourier_scaled_axes_from_grid(Grid_3D):
        """projection.py:144-147."""
        [Length, Number] = Projector.extract_coordinates(Grid_3D)
        return Projector.fourier_scaled_axes(Number, Length)

    @staticmethod
    def extract_coordinates(Grid_3D):
        """projection.py:150-160."""
        L = {}
        for ax in ('x', 'y', 'z'):
            t = Grid_3D[ax].flatten()
            L[ax] = max(t) - min(t)
        return [L, Grid_3D['x'].shape]

    @staticmethod
    def q_and_mask(Experiment):
        """projection.py:164-179."""
        Lambda, zD, Width = Experiment['Lambda'], Experiment['zD'], Experiment['Width']
        [Camera_x, Camera_y] = Projector.camera_x_y(Experiment)
        Temp = Lambda * np.sqrt(Camera_x ** 2 + Camera_y ** 2 + zD ** 2)
        Q = {'x': Camera_x / Temp, 'y': Camera_y / Temp, 'z': (zD / Temp - 1 / Lambda)}
        Circle_Index = (Camera_x ** 2 + Camera_y ** 2) > (Width / 2) ** 2
        return [Q, Circle_Index]

    @staticmethod
    def camera_x_y(Experiment):
        """projection.py:183-190."""
        Width = Experiment['Width']
        N = Experiment['N_p']
        U = (Width / (N - 1)) * (np.linspace(1, N, N) - (N + 1) / 2)
        [Camera_x, Camera_y] = np.meshgrid(U, U, indexing='xy')
        return [Camera_x, Camera_y]
