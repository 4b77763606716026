"""Stand-alone GPU diagnostic (not a pytest file): mat-vec launch time against the size of the row
block (library-side CUDA events around block_matvec_kernel alone, back-to-back launches): the fit
t = t0 + bytes / BW separates the per-launch fixed cost from the streaming bandwidth."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
from v3s import _util

ops = _util.get_ops()
pts = []
for (rows, N) in ((1250, 10000), (2500, 10000), (5000, 10000), (10000, 10000), (6250, 50000), (12500, 50000), (12500, 100000)):
    P = torch.rand((rows, (N + 3) // 4 * 4), dtype=torch.float32, device="cuda")
    Xf = ops.alloc_xf(N)
    Xf.normal_()
    Y = ops.empty((rows, 8), torch.float64)
    for _ in range(6):
        ops.block_matvec_f32(P, N, Xf, out=Y)
    ops.profile = True
    ops.pop_timings()
    for _ in range(40):
        ops.block_matvec_f32(P, N, Xf, out=Y)
    t = ops.pop_timings()["matvec"]
    ops.profile = False
    us = 1e3 * sum(t) / len(t)
    mb = 4.0 * rows * N / 1e6
    pts.append((mb, us))
    print("rows %6d x N %6d (%7.0f MB): %7.1f us  %6.0f GB/s" % (rows, N, mb, us, mb / us * 1e3), flush=True)
    del P
A = np.array([[1.0, m] for m, _ in pts[3:]])
t0, slope = np.linalg.lstsq(A, np.array([u for _, u in pts[3:]]), rcond=None)[0]
print("fit over the blocks >= 400 MB: fixed %.1f us per launch, streaming %.0f GB/s" % (t0, 1e3 / slope))
