"""DistanceMatrix -- round distances of the patterns and their k nearest neighbours (mirror of
the reference's distance.py).  `knn` is the fused fast path (centre/normalise -> tcgen05 Gram row
panels -> radix select -> exact refinement); the individual functions are kept for drop-in use.
"""
import numpy as np
import torch

from . import _util
from .pipeline import ShardPlan, knn_stage


class DistanceMatrix:

    @staticmethod
    def knn(A, k):
        """distance.py:17-22 -> [S2 (N,k) ascending round distances, N (N,k) neighbour indices]."""
        print('DistanceMatrix.knn started.')
        ops = _util.get_ops()
        img = _util.to_dev(A, torch.float32)
        if img.dim() != 2:
            raise Exception('Pattern matrix must be 2-D (N, P).')
        S2, Nbr = knn_stage(ops, ShardPlan(img.shape[0]), img, int(k))
        print('DistanceMatrix.knn ended.')
        return [_util.back(S2, A), _util.back(Nbr, A)]

    @staticmethod
    def knn_from_round_distance(B, k):
        """distance.py:26-48."""
        if len(B.shape) != 2 or B.shape[0] != B.shape[1]:
            raise Exception('Non-square round-distance matrix was encountered.')
        if k > B.shape[0]:
            raise Exception('Number of nearest neighbors to preserve is too high.')
        ops = _util.get_ops()
        S2, Nbr = ops.knn_from_dense(_util.to_dev(B, torch.float64), int(k))
        return [_util.back(S2, B), _util.back(Nbr, B)]

    @staticmethod
    def convert_to_round_distance(A):
        """distance.py:51-63 -> dense (N,N) float64.  d = 2 asin(|a-b|/2) on the centred unit rows,
        which equals arccos(<a,b>) and keeps small distances accurate in FP32."""
        ops = _util.get_ops()
        img = _util.to_dev(A, torch.float32)
        mean = ops.column_sums(img) / img.shape[0]
        a32, _, _, zf = ops.center_normalize(img, mean, want_split=False)
        return _util.back(ops.round_distance_dense(a32, zf), A)

    @staticmethod
    def set_norm_to_one(A):
        """distance.py:66-67 (an all-zero row gives NaN, as in the reference)."""
        ops = _util.get_ops()
        img = _util.to_dev(A, torch.float32)
        zero_mean = ops.zeros((img.shape[1],), torch.float64)
        a32, _, _, zf = ops.center_normalize(img, zero_mean, want_split=False)
        a32 = torch.where(zf[:, None] != 0, torch.full_like(a32, float('nan')), a32)
        return _util.back(a32, A)

    @staticmethod
    def remove_offset(A):
        """distance.py:70-71 (column sums on the device kernel, FP64 mean)."""
        ops = _util.get_ops()
        img = _util.to_dev(A, torch.float32)
        mean = ops.column_sums(img) / img.shape[0]
        return _util.back(img.to(torch.float64) - mean[None, :], A)
