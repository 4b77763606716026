"""Stand-alone GPU diagnostic (not a pytest file): exercises each kernel once, prints error
statistics and timings.  Used during bring-up: `python tests/gpu_diag.py > gpurun_out/diag.log`."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import v3s_oracle as orc
from v3s import _util


def timed(fn, reps=3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps


def section(name):
    print("\n==== " + name, flush=True)


def main():
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    ops = _util.get_ops()
    rng = np.random.default_rng(0)

    section("gram: tcgen05 vs simt vs fp64")
    try:
        for (N, P) in ((256, 64), (300, 200), (1024, 1024), (4096, 4096)):
            A = (rng.random((N, P)).astype(np.float32)) ** 3
            img = ops.to_device(A, torch.float32)
            mean = ops.column_sums(img) / N
            a32, hi, lo, zf = ops.center_normalize(img, mean)
            ref = a32.double() @ a32.double().T
            ops.gram_impl = "simt"
            Gs = ops.gram_rows(hi, lo, a32, 0, N)[:, :N].double()
            ops.gram_impl = "tcgen05"
            Gt = ops.gram_rows(hi, lo, a32, 0, N)[:, :N].double()
            torch.cuda.synchronize()
            es, et = (Gs - ref).abs().max().item(), (Gt - ref).abs().max().item()
            print(f"N={N} P={P}: simt err {es:.2e}  tcgen05 err {et:.2e}  diag tc {Gt.diag()[:3].tolist()}")
            if et > 1e-4:
                d = (Gt - ref).abs()
                bad = (d > 1e-4).nonzero()
                print("  bad entries:", bad.shape[0], "first:", bad[:8].tolist())
                print("  Gt[0,:8]", Gt[0, :8].tolist(), "\n  ref[0,:8]", ref[0, :8].tolist())
                hh = hi.float().double()
                print("  hi.hi^T[0,:8]", (hh @ hh.T)[0, :8].tolist())
        from v3s import _lib
        for (N, P) in ((10000, 4096), (30000, 16384)):
            A = (rng.random((N, P)).astype(np.float32)) ** 3
            img = ops.to_device(A, torch.float32)
            a32, hi, lo, zf = ops.center_normalize(img, ops.column_sums(img) / N)
            del img
            G = ops.empty((N, (N + 3) // 4 * 4), torch.float32)
            for bk in (64, 32):
                _lib.call("v3s_set_gram_kblock", bk)
                _, ms = timed(lambda: ops.gram_rows(hi, lo, a32, 0, N, out=G))
                _, ms2 = timed(lambda: ops.gram_symmetric(hi, lo, N, out=G))
                print(f"gram N={N} P={P} bk={bk}: full {ms:.3f} ms ({2*N*N*P/ms/1e9:.0f} TF/s alg, issued {6*N*N*P/ms/1e9:.0f}) | symmetric {ms2:.3f} ms ({2*N*N*P/ms2/1e9:.0f} TF/s alg)")
            del a32, hi, lo, G
        _lib.call("v3s_set_gram_kblock", 32)
        N, P = 10000, 4096
        A = (rng.random((N, P)).astype(np.float32)) ** 3
        img = ops.to_device(A, torch.float32)
        a32, hi, lo, zf = ops.center_normalize(img, ops.column_sums(img) / N)
        G = ops.empty((N, N), torch.float32)
        _, ms = timed(lambda: ops.gram_rows(hi, lo, a32, 0, N, out=G))
        print(f"gram tcgen05 N={N} P={P}: {ms:.3f} ms  -> algorithmic {2*N*N*P/ms/1e9:.1f} TFLOP/s, issued x3 = {6*N*N*P/ms/1e9:.1f}")
        ref_rows = a32[:64].double() @ a32.double().T
        print("  err on 64 rows:", (G[:64].double() - ref_rows).abs().max().item())
    except Exception:
        traceback.print_exc()

    section("synth vs oracle")
    try:
        for n, n_p in ((7, 15), (15, 31), (31, 63), (31, 64)):
            R9, _ = orc.random_rotations(4, seed=n)
            ref = orc.generate_from_rotations(R9, n, n_p)
            ed = ops.to_device(orc.synthesize_3d(n)["ED"].astype(np.float32), torch.float32)
            rot = ops.to_device(np.ascontiguousarray(R9.T.reshape(-1, 3, 3).transpose(0, 2, 1).reshape(-1, 9)), torch.float64)
            got = ops.synth_patterns(ed, rot, n_p).cpu().numpy()
            print(f"n={n} n_p={n_p}: max err/max {np.max(np.abs(got-ref))/ref.max():.2e}  nonzero frac {np.mean(got!=0):.3f} vs {np.mean(ref!=0):.3f}")
        R9, _ = orc.random_rotations(10000, seed=0)
        rot = ops.to_device(np.ascontiguousarray(R9.T.reshape(-1, 3, 3).transpose(0, 2, 1).reshape(-1, 9)), torch.float64)
        ed = ops.to_device(orc.synthesize_3d(31)["ED"].astype(np.float32), torch.float32)
        _, ms = timed(lambda: ops.synth_patterns(ed, rot, 64))
        print(f"synth 10k x 64x64 (n=31): {ms:.3f} ms")
    except Exception:
        traceback.print_exc()

    section("knn + markov + matvec + eig + fit at N=10k")
    try:
        from v3s.pipeline import ShardPlan, knn_stage, markov_stage, eigen_stage, fit_stage
        img = ops.synth_patterns(ed, rot, 64)
        plan = ShardPlan(10000)
        (S2, Nbr), ms = timed(lambda: knn_stage(ops, plan, img, 150), reps=2)
        print(f"knn_stage: {ms:.3f} ms")
        S2c = S2.clone()
        (P_rows, dis, sc), ms = timed(lambda: markov_stage(ops, plan, S2c.copy_(S2), Nbr, 0.7), reps=2)
        print(f"markov_stage: {ms:.3f} ms  scalars {sc.tolist()}")
        X = torch.randn((10000, 8), dtype=torch.float64, device="cuda")
        Y, ms = timed(lambda: ops.block_matvec(P_rows, 10000, X), reps=10)
        Yref = P_rows[:, :10000].double() @ X
        print(f"block_matvec: {ms*1e3:.1f} us -> {4*1e8/ms/1e6:.1f} GB/s; rel err {((Y-Yref).abs().max()/Yref.abs().max()).item():.2e}")
        stats = {}
        t0 = time.perf_counter()
        Ps, lam = eigen_stage(ops, plan, P_rows, dis, stats=stats, validate=False)
        torch.cuda.synchronize()
        print(f"eigen_stage: {(time.perf_counter()-t0)*1e3:.1f} ms  {stats}  lambda {lam.tolist()}")
        rot_y = ops.to_device(np.ascontiguousarray(R9.T), torch.float64)
        t0 = time.perf_counter()
        fit = fit_stage(ops, plan, Ps, rot_y, polar_mode=0, check=False)
        torch.cuda.synchronize()
        print(f"fit_stage: {(time.perf_counter()-t0)*1e3:.1f} ms  sigma_T {fit['sigma_T'].item():.3f}")
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
