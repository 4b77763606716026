"""v3s -- B200-native hot path of nurlicht/Volumetric-3D-Stitching behind the reference's stage API.

    from v3s import Projector, DistanceMatrix, DiffusionMapSolver, OrientationRecoverySolver

Same class / method names, argument meaning and error behaviour as the reference's
src/python/{projection,distance,diffusion,solution,rotation,compute}.py; the arithmetic runs in
hand-written sm_100a CUDA kernels (csrc/) behind the C-ABI of include/v3s.h.  There is no CPU
fallback: the kernels need a B200.
"""
from . import _lib                      # fails loudly when libv3s.so is not built
from .compute import ComputeUtils
from .diffusion import DiffusionMapSolver
from .distance import DistanceMatrix
from .projection import Projector
from .rotation import Rotation
from .solution import OrientationRecoverySolver
from .extras import fit_without_prior, load_stage_outputs, or_func, orientation_errors_deg, save_stage_outputs

__all__ = ["Projector", "DistanceMatrix", "DiffusionMapSolver", "OrientationRecoverySolver", "Rotation",
           "ComputeUtils", "save_stage_outputs", "load_stage_outputs", "orientation_errors_deg",
           "or_func", "fit_without_prior"]
