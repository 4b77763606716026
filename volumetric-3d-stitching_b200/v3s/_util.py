"""Conversions at the API edge: numpy / torch in, the caller's kind out."""
import numpy as np
import torch

_ops = None


def get_ops():
    """The process-wide CUDA backend (created on first use; raises without a B200)."""
    global _ops
    if _ops is None:
        from .ops_cuda import CudaOps
        _ops = CudaOps()
    return _ops


def set_ops(ops):
    """Install a backend (tests inject a CPU stand-in for the host logic)."""
    global _ops
    _ops = ops


def is_torch(x):
    return isinstance(x, torch.Tensor)


def to_dev(x, dtype):
    ops = get_ops()
    if is_torch(x):
        return x.to(device=ops.device, dtype=dtype).contiguous()
    return ops.to_device(np.ascontiguousarray(np.asarray(x)), dtype)


def back(t, like, dtype=np.float64):
    """Return `t` in the kind of `like`: numpy float64 for numpy callers (the reference's dtype)."""
    if is_torch(like):
        return t
    return t.detach().cpu().numpy().astype(dtype, copy=False)
