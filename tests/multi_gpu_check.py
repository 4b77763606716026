"""Multi-GPU parity check, run under torchrun on a multi-GPU box (tests/test_multi_gpu.py launches it from
the `-m gpu` suite when the box has >= 2 GPUs).  Compares the row-sharded pipeline with the same job on one GPU
on CONNECTED configurations (one graph component, lambda_10 > lambda_11: the eigen-subspace is unique, so
every comparison below is asserted, none is conditional).

    python -m torch.distributed.run --nproc-per-node 2 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import v3s_oracle as orc

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from v3s import _util
from v3s.pipeline import ShardPlan, knn_stage, run_pipeline
from v3s.projection import Projector

ops = _util.get_ops()
ok = True
only = os.environ.get("V3S_CHECK_ONLY", "")          # "sparse": only the sparse-operator section (cheap re-check)
# k chosen so that the thresholded graph has ONE component (oracle, scipy connected_components: 3001 -> 14 components at
# k = 100, 1 at k = 250; 5000 -> 3 at k = 120, 1 at k = 200)
for (N, n, n_p, k) in ((3001, 15, 31, 250), (5000, 15, 48, 200)):
    R9, _ = orc.random_rotations(N, seed=3)
    ed = ops.to_device(orc.synthesize_3d(n)["ED"].astype(np.float32), torch.float32)
    rot9_all = Projector.rot9_rowmajor(R9, ops)
    rot_y = ops.to_device(np.ascontiguousarray(R9.T), torch.float64)
    plan = ShardPlan.from_env(N)
    one = ShardPlan(N)
    # --- sparse (FP64 kNN-list) Markov operator: sharded fused mat-vec (peer stores + flag barrier) vs one GPU
    ops.operator = "sparse"
    st_s, st_1 = {}, {}
    sp_s = run_pipeline(ops, plan, ed, rot9_all[plan.lo:plan.hi].contiguous(), rot_y, n_p, k, 0.7, check=False, stats=st_s)
    sp_1 = run_pipeline(ops, one, ed, rot9_all, rot_y, n_p, k, 0.7, check=False, stats=st_1)
    # the same sharded job with the whole eigen-solver behind one C-ABI call (v3s_eig_topk, peer stores + barriers inside)
    ops.eig_driver = "cabi"
    st_c = {}
    sp_c = run_pipeline(ops, plan, ed, rot9_all[plan.lo:plan.hi].contiguous(), rot_y, n_p, k, 0.7, check=False, stats=st_c)
    ops.eig_driver = "python"
    lam_c = float((sp_c["Lambda"] - sp_1["Lambda"]).abs().max().item())
    ang_c = float(np.max(orc.principal_angles(sp_c["Ps"].cpu().numpy(), sp_1["Ps"].cpu().numpy())))
    print(f"rank {rank} N={N}: v3s_eig_topk sharded vs Python driver on one GPU: |dLambda| {lam_c:.2e}, subspace angle {ang_c:.2e}, "
          f"driver {st_c.get('driver')} matvecs {st_c.get('matvecs')}", flush=True)
    ok &= str(st_c.get("driver", "")).startswith("cabi") and bool(st_c.get("converged")) and lam_c < 1e-8 and ang_c < 1e-6
    ops.operator = "auto"
    lam_sp = float((sp_s["Lambda"] - sp_1["Lambda"]).abs().max().item())
    ang_sp = float(np.max(orc.principal_angles(sp_s["Ps"].cpu().numpy(), sp_1["Ps"].cpu().numpy())))
    conn_sp = bool(np.all(sp_1["Lambda"].cpu().numpy()[1:] < 1 - 1e-6))
    print(f"rank {rank} N={N}: sparse operator sharded vs one GPU: |dLambda| {lam_sp:.2e}, subspace angle {ang_sp:.2e}, "
          f"operator {st_s.get('operator')} fused={st_s.get('fused_matvec')} matvecs {st_s.get('matvecs')} vs {st_1.get('matvecs')}",
          flush=True)
    ok &= st_s.get("operator") == "sparse-fp64" and bool(st_s.get("fused_matvec")) and lam_sp < 1e-9 and conn_sp and ang_sp < 1e-6
    if only == "sparse":
        continue
    img_all = ops.synth_patterns(ed, rot9_all, n_p)
    # --- Gram row block: fused symmetric + peer stores vs rectangular panel
    sums = ops.column_sums(img_all)
    a32, hi, lo, zf = ops.center_normalize(img_all, sums / N)
    Gs = ops.gram_symmetric_sharded(hi, lo, plan)
    if Gs is None:
        print(rank, "symmetric memory unavailable:", ops.symm_error, flush=True)
        ok = False
    else:
        Gr = ops.gram_rows(hi, lo, a32, plan.lo, plan.hi)
        d = (Gs[:, :N] - Gr[:, :N]).abs().max().item()
        print(f"rank {rank} N={N}: |G_sharded_symmetric - G_panel| max = {d:.2e}", flush=True)
        ok &= d < 3e-6
    # --- whole pipeline: sharded (peer Gram) vs single GPU
    out_s = run_pipeline(ops, plan, ed, rot9_all[plan.lo:plan.hi].contiguous(), rot_y, n_p, k, 0.7, check=False)
    out_1 = run_pipeline(ops, one, ed, rot9_all, rot_y, n_p, k, 0.7, check=False)
    S2s, S21 = out_s["S2_unmutated"].cpu().numpy(), out_1["S2_unmutated"].cpu().numpy()
    okS = (orc.almost_equal_hybrid(S2s, S21, 1e-6) | (np.abs(S2s - S21) < 1e-9)).all()
    same_idx = float((out_s["N"] == out_1["N"]).float().mean().item())
    lam_err = float((out_s["Lambda"] - out_1["Lambda"]).abs().max().item())
    ang = float(np.max(orc.principal_angles(out_s["Ps"].cpu().numpy(), out_1["Ps"].cpu().numpy())))
    print(f"rank {rank} N={N}: S2 equal {bool(okS)}, same neighbours {same_idx:.6f}, |dLambda| {lam_err:.2e}, "
          f"subspace angle {ang:.2e}, sigma_T {out_s['sigma_T'].item():.4f} vs {out_1['sigma_T'].item():.4f}", flush=True)
    # --- the eigen-solver's fused mat-vec (peer stores + flag barrier) vs the NCCL all-gather form
    ops.use_fused_matvec = False
    st_n = {}
    out_n = run_pipeline(ops, plan, ed, rot9_all[plan.lo:plan.hi].contiguous(), rot_y, n_p, k, 0.7, check=False, stats=st_n)
    ops.use_fused_matvec = True
    st_f = {}
    out_f = run_pipeline(ops, plan, ed, rot9_all[plan.lo:plan.hi].contiguous(), rot_y, n_p, k, 0.7, check=False, stats=st_f)
    lam_fn = float((out_f["Lambda"] - out_n["Lambda"]).abs().max().item())
    ang_fn = float(np.max(orc.principal_angles(out_f["Ps"].cpu().numpy(), out_n["Ps"].cpu().numpy())))
    print(f"rank {rank} N={N}: fused vs NCCL mat-vec exchange: |dLambda| {lam_fn:.2e}, subspace angle {ang_fn:.2e}, "
          f"fused={st_f.get('fused_matvec')} matvecs {st_f.get('matvecs')} vs {st_n.get('matvecs')}", flush=True)
    ok &= bool(st_f.get("fused_matvec")) and not st_n.get("fused_matvec") and lam_fn < 1e-9
    lam1 = out_1["Lambda"].cpu().numpy()
    connected = bool(np.all(lam1[1:] < 1 - 1e-6))            # a degenerate lambda = 1 cluster has no unique basis
    ok &= connected and bool(okS) and same_idx > 0.9999 and lam_err < 1e-6 and ang < 1e-4
    ok &= ang_fn < 1e-6
    if not ok:
        print(f"rank {rank} N={N}: CHECK FAILED (connected={connected})", flush=True)
# --- sharded SYMMETRIC fused Gram / top-k (N >= 16384: each rank computes its share of the upper-triangle tiles and
#     appends candidates into the owners' lists over NVLink) against the same kNN on one GPU
N, n, n_p, k = 20000, 7, 40, 100
R9, _ = orc.random_rotations(N, seed=5)
ed = ops.to_device(orc.synthesize_3d(n)["ED"].astype(np.float32), torch.float32)
img_all = ops.synth_patterns(ed, Projector.rot9_rowmajor(R9, ops), n_p)
plan, one = ShardPlan.from_env(N), ShardPlan(N)
S2s, Ns_ = knn_stage(ops, plan, img_all[plan.lo:plan.hi].contiguous(), k)
ov_s = int(ops.knn_overflow.item())
S21, N1_ = knn_stage(ops, one, img_all, k)
same = float((Ns_ == N1_).float().mean().item())
dmax = float((S2s - S21).abs().max().item())
print(f"rank {rank} N={N}: sharded symmetric kNN vs one GPU: max |dS2| {dmax:.2e}, same neighbours {same:.6f}, overflow {ov_s}", flush=True)
ok &= dmax == 0.0 and same == 1.0 and ov_s == 0
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
