"""OrientationRecoverySolver -- orchestration (mirror of the reference's solution.py)."""
import numpy as np
import torch

from . import _util
from .pipeline import ShardPlan, run_pipeline
from .projection import Projector
from .rotation import Rotation


# results up to this size are handed out as private arrays (copied by worker threads while the GPU is busy); larger
# ones stay views of the reusable pinned transfer buffers -- in a sharded job every rank of the box would otherwise
# copy the replicated kNN lists at the same time, on the cores that have to keep eight GPUs fed with launches
PRIVATE_COPY_BYTES = 64 << 20


class OrientationRecoverySolver:

    @staticmethod
    def solve_and_plot(parameters):
        """solution.py:26-30.  Plotting (plot.py) is outside the hot path; the result map is returned."""
        return OrientationRecoverySolver.solve(parameters)

    @staticmethod
    def solve(parameters, rotations=None, N_p=None, polar_mode=0, stats=None, check_sigma=True,
              photons_per_pattern=None, noise_seed=0, copy_results=None, validate=True, images_dtype=np.float64):
        """solution.py:34-61 -> dict of numpy float64 arrays with the reference's 11 keys.
        Extensions that never change reference-mode results: `rotations` = explicit
        R (9,N) instead of the Hopf grid, `N_p` = free detector size, `check_sigma=False` = report
        Sigma_T > 60 instead of raising (diffusion.py:175-176), `photons_per_pattern` = Poisson photon
        noise on the patterns (BASELINE config 5; the reference has no noise model).
        The large outputs ('Images', 'S2', 'N') are numpy views of reusable pinned host buffers (filled by
        transfers that overlap stages 2-4).  `copy_results` None (default): results up to 64 MB are
        copied into private arrays like the reference's, larger ones stay views that are valid until the next
        call of solve; True / False force either.  `images_dtype=np.float32` returns the pattern matrix in the precision
        it is computed in (half the bytes over PCIe; the float64 default only widens the same values).  `validate=False` skips the eigenvalue / convergence guards (compute.py:59-79).
        Under torch.distributed (one process per GPU, world > 1) the job is row-sharded: every rank calls
        solve with the same arguments and gets the same replicated results, except 'Images', which holds
        the rank's own pattern rows [rows[0], rows[1]) (key 'rows'; the full matrix never leaves the GPUs)."""
        print('OrientationRecoverySolver.solve started.')
        import time
        t_mark = [time.perf_counter()]

        def lap(name):                                               # host-side diagnostics (bench.py's e2e.host_ms)
            now = time.perf_counter()
            _util.TIMERS[name] = _util.TIMERS.get(name, 0.0) + now - t_mark[0]
            t_mark[0] = now
        if not OrientationRecoverySolver.validate_parameters(parameters):
            parameters = OrientationRecoverySolver.default_parameters()
        if not OrientationRecoverySolver.validate_parameters(parameters):
            raise Exception('Default parameters are invalid.')
        if not parameters['aPrioriFlag']:
            raise Exception('aPrioriFlag=False is not runnable in the reference either (diffusion.py:231)')
        ops = _util.get_ops()
        n = parameters['n_voxel_1d']
        if rotations is None:
            [R, Axis, Quat] = Rotation.all_rot_variables(parameters['nRot1D'] ** 3)
        else:
            R = np.asarray(rotations, dtype=np.float64)
            Axis = np.full((3, R.shape[1]), np.nan)
            Quat = np.full((4, R.shape[1]), np.nan)
        Nn = R.shape[1]
        n_p = (1 + 2 * n) if N_p is None else int(N_p)
        ed = Projector.object_on_device(n, ops)
        rot9 = Projector.rot9_rowmajor(R, ops)
        rot_y = ops.to_device(np.ascontiguousarray(R.T), torch.float64)
        xfer = {}
        # the pattern matrix is by far the largest output: its host transfer starts as soon as stage 1
        # is done and overlaps stages 2-4 (side stream + worker thread)
        on_gpu = ops.device.type == 'cuda'
        def private(t, itemsize=8):          # results the caller gets as private arrays (see `copy_results`)
            return bool(copy_results or (copy_results is None and t.numel() * itemsize <= PRIVATE_COPY_BYTES))
        start_images = (lambda img: xfer.setdefault('Images', _util.AsyncFetch(
            img, out_dtype=images_dtype, private=private(img, np.dtype(images_dtype).itemsize)))) if on_gpu else None
        # so are the kNN lists (final once the Markov matrix is built): they travel during the eigen-solver

        def start_lists(S2_dev, N_dev):
            # (private copies are made by worker threads while the eigen-solver runs)
            xfer['S2'] = _util.AsyncFetch(S2_dev, slot='S2', private=private(S2_dev))
            xfer['N'] = _util.AsyncFetch(N_dev, slot='N', private=private(N_dev))   # int32 on the wire, float64 (distance.py:37) on the host
        plan = ShardPlan.from_env(Nn)
        if plan.world > 1:
            rot9 = rot9[plan.lo:plan.hi].contiguous()
        lap('setup_s')
        out = run_pipeline(ops, plan, ed, rot9, rot_y, n_p, parameters['k'], parameters['Epsilon'],
                           polar_mode=polar_mode, check=check_sigma, stats=stats, on_images=start_images,
                           photons_per_pattern=photons_per_pattern, noise_seed=noise_seed,
                           on_lists=start_lists if on_gpu else None, validate=validate)
        lap('pipeline_host_s')
        # the reference returns float64 everywhere: widen on the device, one staged transfer for all arrays
        f64 = lambda t: t.to(torch.float64)
        small = {'Ps': f64(out['Ps']), 'Lambda': f64(out['Lambda']), 'c': f64(out['c'])}
        if 'S2' not in xfer:
            small.update({'S2': f64(out['S2']), 'N': f64(out['N'])})
        host = _util.fetch_many({
            **small,
            'nonzero': plan.all_reduce_sum(torch.any(out['Images_local'] != 0).reshape(1).to(torch.int32)).to(torch.uint8),
            'Sigma_T': out['sigma_T'].reshape(1).to(torch.float64)})
        lap('small_results_s')
        host['Images'] = xfer['Images'].result() if 'Images' in xfer else \
            _util.fetch_many({'Images': f64(out['Images_local'])})['Images']
        if 'S2' in xfer:
            host['S2'], host['N'] = xfer['S2'].result(), xfer['N'].result()
        lap('large_results_wait_s')
        for key in ('Images', 'S2', 'N'):
            if key not in xfer and (copy_results or (copy_results is None and host[key].nbytes <= PRIVATE_COPY_BYTES)):
                host[key] = np.array(host[key], copy=True)
        lap('private_copies_s')
        if not host['nonzero'][0]:
            raise Exception('Images is an all-zero array.')          # projection.py:47
        print('OrientationRecoverySolver.solve ended.')
        extra = {'rows': np.array([plan.lo, plan.hi])} if plan.world > 1 else {}
        return {
            **extra,
            'Ps': host['Ps'], 'Lambda': host['Lambda'], 'c0': host['c'].copy(), 'c': host['c'],
            'S2': host['S2'], 'N': host['N'], 'Images': host['Images'], 'R': R,
            'Sigma_T': np.array([out.get('sigma_T_host', float(host['Sigma_T'][0]))]), 'Axis': Axis.T, 'Quat': Quat.T,
        }

    @staticmethod
    def validate_parameters(parameters):
        """solution.py:65-75 (bug-compat: the exact key ORDER is required)."""
        return (
            isinstance(parameters, dict) and
            list(parameters.keys()) == ['n_voxel_1d', 'nRot1D', 'k', 'Epsilon', 'plot', 'aPrioriFlag'] and
            isinstance(parameters['n_voxel_1d'], int) and parameters['n_voxel_1d'] > 0 and
            isinstance(parameters['k'], int) and parameters['k'] > 8 and
            isinstance(parameters['nRot1D'], int) and parameters['nRot1D'] ** 3 - 1 > parameters['k'] and
            isinstance(parameters['Epsilon'], float) and parameters['Epsilon'] > 0 and
            isinstance(parameters['plot'], bool) and
            isinstance(parameters['aPrioriFlag'], bool)
        )

    @staticmethod
    def default_parameters():
        """solution.py:79-87."""
        return {'n_voxel_1d': 7, 'nRot1D': 8, 'k': 24, 'Epsilon': 0.7, 'plot': True, 'aPrioriFlag': True}
