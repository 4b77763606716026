"""Stand-alone host micro-benchmark (not a pytest file): cost of the Ritz extraction variants on this
machine's CPU -- the only host arithmetic inside the eigen-solver loop."""
import os
import sys
import time

import numpy as np
import scipy.linalg as sla
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
from v3s import eig


def t(f, n=10):
    f()
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    b = 8
    rng = np.random.default_rng(0)
    print("cpus", os.cpu_count(), "torch threads", torch.get_num_threads())
    def ctx():
        with eig._one_thread():
            pass
    print("one_thread ctx enter/exit: %.3f ms" % t(ctx, 50))
    for m in (32, 64, 80, 96, 128, 160):
        band = np.zeros((b + 1, 1024))
        band[:, :m] = rng.standard_normal((b + 1, m))
        Tb = band[:, :m].copy()
        for i in range(1, b + 1):
            Tb[i, m - i:] = 0
        T = np.zeros((m, m))
        for i in range(b + 1):
            ii = np.arange(m - i)
            T[ii + i, ii] = Tb[i, :m - i]
        Tt = torch.from_numpy(T + np.tril(T, -1).T)
        with eig._one_thread():
            r = {
                "ritz(top)": t(lambda: eig._ritz_from_band_impl(band, m, b, 10, False)),
                "ritz(both)": t(lambda: eig._ritz_from_band_impl(band, m, b, 10, True)),
                "eig_banded full": t(lambda: sla.eig_banded(Tb, lower=True)),
                "scipy eigh subset": t(lambda: sla.eigh(T, lower=True, check_finite=False, subset_by_index=[m - 10, m - 1])),
                "scipy eigh full": t(lambda: sla.eigh(T, lower=True, check_finite=False)),
                "numpy eigh": t(lambda: np.linalg.eigh(T, UPLO="L")),
                "torch eigh (1 thr ctx)": t(lambda: torch.linalg.eigh(Tt)),
            }
        r["ritz via wrapper"] = t(lambda: eig._ritz_from_band(band, m, b, 10, False))
        r["torch eigh (default threads)"] = t(lambda: torch.linalg.eigh(Tt))
        print("m=%d: " % m + ", ".join("%s %.3f" % kv for kv in r.items()))


if __name__ == "__main__":
    main()
