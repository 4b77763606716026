"""One run of the direct analytical synthesis kernel (north-star kernel 1) at BASELINE config 2 size, for ncu:
    ncu --set full -k regex:synth_direct python tests/gpu_synth_direct_run.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import v3s_oracle as orc
from v3s import _util
from v3s.projection import Projector

ops = _util.get_ops()
N, n, n_p = 10000, 31, 64
R9, _ = orc.random_rotations(N, seed=0)
ed = ops.to_device(orc.synthesize_3d(n)["ED"].astype(np.float32), torch.float32)
rot9 = Projector.rot9_rowmajor(R9, ops)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    img = ops.synth_patterns(ed, rot9, n_p, mode="direct")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
P = n_p * n_p
flop = N * (P * (8.0 * n ** 3 + 16.0 * n ** 2 + 16.0 * n))       # real x complex MACs of the factorised triple sum, 2 flop each
print("synth_direct: N=%d %dx%d n=%d: %.3f ms, %.1f TFLOP/s FP32 (algorithmic)" % (N, n_p, n_p, n, ms, flop / ms / 1e9))
