"""Diagnostics of the tiled refinement on a BASELINE-config-3/4-like problem: tile statistics (rows, pairs, distinct
candidates per tile = the reuse the tiling achieves) and the time of each of its kernels (CUDA events).
    python tests/gpu_refine_diag.py [N] [n_p]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import v3s_oracle as orc


def main():
    from v3s import _lib, _util
    from v3s._lib import call, ptr, stream_ptr
    from v3s.ops_cuda import fused_gram_error_bound, KNN_MARGIN, KNN_EXTRA_CAP
    from v3s.projection import Projector
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    n_p = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    k, n = 150, 31
    ops = _util.get_ops()
    R9, _ = orc.random_rotations(N, seed=0)
    ed = ops.to_device(orc.synthesize_3d(n)["ED"].astype(np.float32), torch.float32)
    img = ops.synth_patterns(ed, Projector.rot9_rowmajor(R9, ops), n_p)
    a32, _, _, zf = ops.center_normalize(img, ops.column_sums(img) / N, want_split=False)
    del img
    a16 = ops.split_f16(a32)
    P = a32.shape[1]
    pl = (C.c_int * 4)()
    call("v3s_gram_topk_plan", N, N, k, pl)
    S, Cc, rows_pad = int(pl[0]), int(pl[1]), int(pl[2])
    lists = ops.empty((rows_pad * S * Cc,), torch.int64)
    counts = ops.empty((rows_pad * S,), torch.int32)
    tws = ops.empty((int(pl[3]),), torch.int32)
    ovf = ops.zeros((1,), torch.int32)
    eps = fused_gram_error_bound(P)

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        print("  %-28s %8.3f ms" % (name, e0.elapsed_time(e1)))

    timed("gram_topk", lambda: call("v3s_gram_topk", ptr(a16), N, ptr(a16), N, a16.shape[1], k, eps, ptr(lists), ptr(counts), ptr(ovf), ptr(tws), stream_ptr()))
    print("  appended per row (all lists): %.0f" % (float(counts[: N * S].sum().item()) / N))
    cap = min(k + KNN_MARGIN + KNN_EXTRA_CAP, N)
    qt = int(_lib.lib.v3s_knn_refine_tile_rows(cap))
    cand, cnt = ops.empty((N, cap), torch.int32), ops.empty((N,), torch.int32)
    zws = ops.empty((cap + 1,), torch.int32)
    n_bins = (N + qt - 1) // qt * 2
    order_ws = ops.empty((2 * N + n_bins,), torch.int32)
    max_tiles = n_bins + N // qt + 2
    tiles_ws = ops.empty((1 + 2 * max_tiles,), torch.int32)
    timed("merge + order + tiles", lambda: call("v3s_knn_merge_lists", ptr(lists), ptr(counts), S, Cc, N, N, 0, k, eps, cap, ptr(zf), ptr(zws), ptr(cand),
                                                ptr(cnt), ptr(ovf), qt, ptr(order_ws), qt, max_tiles, ptr(tiles_ws), stream_ptr()))
    print("  candidates per row: mean %.1f max %d; rows without a listed anchor: %.2f %%"
          % (float(cnt.float().mean().item()), int(cnt.max().item()),
             100.0 * float((order_ws[N:2 * N] >= (N + qt - 1) // qt).float().mean().item())))
    S2, Nbr = ops.empty((N, k), torch.float64), ops.empty((N, k), torch.int32)
    dws, pref = ops.empty((N, cap), torch.float64), ops.empty((N, cap), torch.int32)
    keys, npw = ops.empty((max_tiles * 16384,), torch.int64), ops.zeros((max_tiles,), torch.int32)
    for dedup in (1, 0):
        timed("refine tiled (dedup=%d)" % dedup, lambda: call("v3s_knn_refine_tiled", ptr(a32), ptr(a32), P, 0, N, N, k, cap, ptr(zf), ptr(cand), ptr(cnt),
                                                           ptr(S2), ptr(Nbr), ptr(dws), ptr(order_ws), ptr(tiles_ws), max_tiles, ptr(pref), ptr(keys), ptr(npw), dedup, stream_ptr()))
        nt = int(tiles_ws[0].item())
        tl = tiles_ws[1:1 + 2 * nt].view(nt, 2).cpu().numpy()
        npn = npw[:nt].cpu().numpy()
        kk = keys.view(max_tiles, 16384)[:nt].cpu().numpy()
        du = [np.unique(kk[t, :npn[t]] >> 32).size for t in range(0, nt, max(1, nt // 200))]
        sel = list(range(0, nt, max(1, nt // 200)))
        print("    tiles %d (max %d), rows/tile mean %.1f, pairs/tile mean %.0f, distinct candidates/tile mean %.0f -> pairs per candidate row read %.2f"
              % (nt, max_tiles, tl[:, 1].mean(), npn.mean(), np.mean(du), npn[sel].sum() / max(1, np.sum(du))))
    timed("refine untiled (dedup)", lambda: call("v3s_knn_refine", ptr(a32), ptr(a32), P, 0, N, N, k, cap, ptr(zf), ptr(cand), ptr(cnt), ptr(S2), ptr(Nbr),
                                                 ptr(dws), None, stream_ptr()))


if __name__ == "__main__":
    main()
