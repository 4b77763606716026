"""Stand-alone check of the fused Gram / top-k kernel (gram_topk.cu) against FP64, for both cta_group shapes.

    python tests/gpu_fused_check.py [cg ...]        (run under `timeout`: a barrier bug would hang, not fail)

For every shape: the union of a row's candidate lists must contain the exact (FP64) k nearest, every listed
value must be within the documented error bound of the FP64 Gram entry, and the final S2 / N must equal the
materialised-Gram path's bit for bit.
"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def check(ops, N, P, k, r0, r1, seed, clustered=False):
    from v3s import _lib
    from v3s._lib import call, ptr, stream_ptr
    from v3s.ops_cuda import fused_gram_error_bound, round_up
    rng = np.random.default_rng(seed)
    if clustered:
        base = rng.standard_normal((60, P)).astype(np.float32)
        A = (base[rng.integers(0, 60, N)] + 0.35 * rng.standard_normal((N, P))).astype(np.float32)
    else:
        A = rng.random((N, P)).astype(np.float32) ** 3
    img = ops.to_device(A, torch.float32)
    a32, _, _, zf = ops.center_normalize(img, ops.column_sums(img) / N, want_split=False)
    a16 = ops.split_f16(a32)
    ref = a32.cpu().numpy().astype(np.float64)
    G64 = ref[r0:r1] @ ref.T
    nr = r1 - r0
    pl = (C.c_int * 4)()
    call("v3s_gram_topk_plan", nr, N, k, pl)
    S, Cc, rows_pad = int(pl[0]), int(pl[1]), int(pl[2])
    lists = torch.zeros((rows_pad * S * Cc,), dtype=torch.int64, device=ops.device)
    counts = torch.zeros((rows_pad * S,), dtype=torch.int32, device=ops.device)
    ovf = torch.zeros((1,), dtype=torch.int32, device=ops.device)
    tws = torch.empty((int(pl[3]),), dtype=torch.int32, device=ops.device)
    eps = fused_gram_error_bound(P)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    call("v3s_gram_topk", ptr(a16), N, ptr(a16[r0:r1]), nr, a16.shape[1], k, eps, ptr(lists), ptr(counts), ptr(ovf), ptr(tws), stream_ptr())
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    L = lists.cpu().numpy().reshape(rows_pad, S, Cc)
    cn = counts.cpu().numpy().reshape(rows_pad, S)
    vals = (L >> 32).astype(np.uint32).view(np.float32)
    cols = (L & 0xFFFFFFFF).astype(np.int64)
    top = np.argsort(-G64, axis=1, kind="stable")[:, :k]
    miss, max_err, max_union = 0, 0.0, 0
    for r in range(nr):
        cs, vs = [], []
        for s in range(S):
            c = cn[r, s]
            assert 0 <= c <= Cc, (r, s, c)
            cs.append(cols[r, s, :c])
            vs.append(vals[r, s, :c])
        cs, vs = np.concatenate(cs), np.concatenate(vs)
        assert cs.size == np.unique(cs).size, "duplicate columns in the lists of row %d" % r
        assert np.all((cs >= 0) & (cs < N))
        max_err = max(max_err, float(np.max(np.abs(vs - G64[r, cs]))))
        miss += len(set(top[r]) - set(cs.tolist()))
        max_union = max(max_union, cs.size)
    print("  N=%d P=%d k=%d rows[%d,%d) S=%d C=%d: %.2f ms, misses %d, max|g~-g| %.2e (bound %.2e), max union %d, overflow %d"
          % (N, P, k, r0, r1, S, Cc, dt * 1e3, miss, max_err, eps, max_union, int(ovf.item())))
    assert miss == 0 and max_err <= eps and int(ovf.item()) == 0
    # end to end against the materialised path (same untiled refinement kernel on both sides: bit for bit)
    tiled = ops.tiled_refine
    try:
        ops.tiled_refine = False
        ops.gram_impl = "fused"
        S2f, Nf = ops.knn_rows(None, None, a32, zf, r0, r1, k)
        ops.gram_impl = "tcgen05"
        S2m, Nm = ops.knn_rows(None, None, a32, zf, r0, r1, k)
        ops.gram_impl = "fused"
        assert torch.equal(S2f, S2m) and torch.equal(Nf, Nm), "fused and materialised kNN differ"
        # the tiled refinement sums the pixel axis in another order: same neighbours, distances to FP32 rounding
        ops.tiled_refine = True
        S2t, Nt = ops.knn_rows(None, None, a32, zf, r0, r1, k)
        same = float((Nt == Nf).float().mean().item())
        rel = float(((S2t - S2f).abs() / torch.clamp(S2f, min=1e-3)).max().item())
        print("    tiled refinement vs untiled: same neighbours %.6f, max rel distance difference %.2e" % (same, rel))
        assert same > 0.9995 and rel < 3e-6
        assert bool(torch.all(S2t[:, 1:] >= S2t[:, :-1]))
    finally:
        ops.tiled_refine = tiled
        ops.gram_impl = "fused"
    return dt


def check_sym(ops, N, P, k, seed, clustered=False):
    """Symmetric form (pass 1: limits from every 8th pattern; pass 2: tiles on / above the diagonal, both directions)
    against the rectangular fused kernel: identical S2 / N (same refinement kernel on both sides)."""
    rng = np.random.default_rng(seed)
    if clustered:
        base = rng.standard_normal((200, P)).astype(np.float32)
        A = (base[rng.integers(0, 200, N)] + 0.5 * rng.standard_normal((N, P))).astype(np.float32)
    else:
        A = rng.random((N, P)).astype(np.float32) ** 3
    img = ops.to_device(A, torch.float32)
    a32, _, _, zf = ops.center_normalize(img, ops.column_sums(img) / N, want_split=False)
    tiled, sym = ops.tiled_refine, ops.symmetric_gram
    try:
        ops.tiled_refine = False
        ops.symmetric_gram = False
        S2r, Nr = ops.knn_rows(None, None, a32, zf, 0, N, k)
        ovr = int(ops.knn_overflow.item())
        ops.symmetric_gram = True
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        S2s, Ns = ops.knn_rows(None, None, a32, zf, 0, N, k)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        ovs = int(ops.knn_overflow.item())
        same = float((Ns == Nr).float().mean().item())
        print("  symmetric N=%d P=%d k=%d: %.1f ms (whole kNN), identical S2 %s, same neighbours %.6f, overflow %d / %d"
              % (N, P, k, dt * 1e3, bool(torch.equal(S2s, S2r)), same, ovs, ovr))
        assert torch.equal(S2s, S2r) and torch.equal(Ns, Nr) and ovs == 0 and ovr == 0
    finally:
        ops.tiled_refine, ops.symmetric_gram = tiled, sym


def perf(ops, N, P, k, nr, cg):
    """Throughput of the fused kernel alone on random unit rows (device-generated)."""
    from v3s import _lib
    from v3s._lib import call, ptr, stream_ptr
    from v3s.ops_cuda import fused_gram_error_bound
    _lib.call("v3s_set_gram_topk_cta_group", cg)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    a32 = torch.empty((N, P), dtype=torch.float32, device="cuda")
    for r in range(0, N, 8192):
        blk = torch.randn((min(8192, N - r), P), generator=g, device="cuda")
        a32[r:r + blk.shape[0]] = blk / blk.norm(dim=1, keepdim=True)
    a16 = ops.split_f16(a32)
    del a32
    pl = (C.c_int * 4)()
    call("v3s_gram_topk_plan", nr, N, k, pl)
    S, Cc, rows_pad = int(pl[0]), int(pl[1]), int(pl[2])
    lists = torch.empty((rows_pad * S * Cc,), dtype=torch.int64, device="cuda")
    counts = torch.empty((rows_pad * S,), dtype=torch.int32, device="cuda")
    ovf = torch.zeros((1,), dtype=torch.int32, device="cuda")
    tws = torch.empty((int(pl[3]),), dtype=torch.int32, device="cuda")
    eps = fused_gram_error_bound(P)
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        call("v3s_gram_topk", ptr(a16), N, ptr(a16), nr, a16.shape[1], k, eps, ptr(lists), ptr(counts), ptr(ovf), ptr(tws), stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:])
    print("  perf cta_group::%d N=%d P=%d rows=%d S=%d C=%d: %.2f ms  %.0f TFLOP/s algorithmic (2 rows N P), mean appended/row %.0f, overflow %d"
          % (cg, N, P, nr, S, Cc, t, 2.0 * nr * N * P / t / 1e9, float(counts[: nr * S].sum().item()) / nr, int(ovf.item())))


def main():
    from v3s import _lib, _util
    ops = _util.get_ops()
    if len(sys.argv) > 1 and sys.argv[1] == "perf1":        # perf1 cg N P rows   (one configuration, e.g. under ncu)
        cg, N, P, nr = (int(a) for a in sys.argv[2:6])
        perf(ops, N, P, 150, nr, cg)
        _lib.call("v3s_set_gram_topk_cta_group", 0)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sym":
        check_sym(ops, 20000, 1024, 150, 31)
        check_sym(ops, 17001, 640, 100, 32, clustered=True)
        print("symmetric check ok")
        return
    if len(sys.argv) > 1 and sys.argv[1] == "perf":
        for cg in (1, 2):
            perf(ops, 10000, 4096, 150, 10000, cg)
            perf(ops, 50000, 16384, 150, 50000, cg)
        _lib.call("v3s_set_gram_topk_cta_group", 0)
        return
    cgs = [int(a) for a in sys.argv[1:]] or [1, 2]
    for cg in cgs:
        _lib.call("v3s_set_gram_topk_cta_group", cg)
        print("cta_group::%d" % cg)
        check(ops, 301, 225, 50, 0, 301, 1)
        check(ops, 77, 64, 19, 0, 77, 2)
        check(ops, 1000, 961, 150, 0, 1000, 3)
        check(ops, 3000, 640, 150, 1000, 2177, 4, clustered=True)
        check(ops, 5000, 192, 150, 0, 5000, 5)
        check(ops, 6000, 4096, 150, 0, 6000, 6)
        check(ops, 20000, 1024, 200, 2500, 5000, 7)
    _lib.call("v3s_set_gram_topk_cta_group", 0)
    print("fused check ok")


if __name__ == "__main__":
    main()
