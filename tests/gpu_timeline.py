"""Stand-alone GPU diagnostic (not a pytest file): kernel timeline of one device-resident pass of
the bench workload, taken with torch.profiler (CUPTI).  Prints the GPU-busy time, the idle gaps
and which kernel follows each large gap; writes the raw kernel list to gpurun_out/timeline.csv.

    python tests/gpu_timeline.py [workload] > gpurun_out/timeline.log
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
sys.path.insert(0, ROOT)


def main():
    import bench
    from torch.profiler import ProfilerActivity, profile
    from v3s import _util
    from v3s.pipeline import ShardPlan, run_pipeline
    from v3s.projection import Projector
    from v3s.rotation import Rotation
    wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    ops = _util.get_ops()
    N, n, n_p, k, eps = wl["N"], wl["n"], wl["n_p"], wl["k"], wl["eps"]
    plan = ShardPlan.from_env(N)
    R9, _ = Rotation.random_rotations(N, seed=0)
    ed = ops.to_device(Projector.synsthesize_3d(n)["ED"].astype(np.float32), torch.float32)
    rot9 = Projector.rot9_rowmajor(R9, ops)[plan.lo:plan.hi].contiguous()
    rot_y = ops.to_device(np.ascontiguousarray(R9.T), torch.float64)

    def one_pass():
        return run_pipeline(ops, plan, ed, rot9, rot_y, n_p, k, eps, polar_mode=0, check=False, validate=False)

    for _ in range(3):
        one_pass()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        one_pass()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rank != 0:
        dist.destroy_process_group()
        return
    ev = []
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            ev.append((e.time_range.start, e.time_range.end, e.name))
    ev.sort()
    t0, t1 = ev[0][0], max(e[1] for e in ev)
    busy, cur_end, gaps = 0.0, t0, []
    prev_name = None
    for s, e_, name in ev:
        if s > cur_end:
            gaps.append((s - cur_end, prev_name, name, cur_end - t0))
            busy += e_ - s
            cur_end = e_
        else:
            busy += max(0.0, e_ - cur_end)
            cur_end = max(cur_end, e_)
        prev_name = name
    print("pass span %.3f ms, GPU busy %.3f ms, idle %.3f ms, %d device activities" % ((t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3, len(ev)))
    gaps.sort(reverse=True)
    print("largest gaps (us, at ms, after -> before):")
    for g, a, b, at in gaps[:40]:
        print("  %8.1f us  @%8.3f ms  %s  ->  %s" % (g, at / 1e3, (a or "")[:60], b[:60]))
    hist = {}
    for g, a, b, at in gaps:
        hist.setdefault(b[:50], [0, 0.0])
        hist[b[:50]][0] += 1
        hist[b[:50]][1] += g
    print("idle time by the kernel that follows the gap:")
    for name, (c, tot) in sorted(hist.items(), key=lambda x: -x[1][1])[:25]:
        print("  %8.1f us  n=%4d  %s" % (tot, c, name))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "timeline_%dgpu.csv" % world), "w") as f:
        for s, e_, name in ev:
            f.write("%.3f,%.3f,%s\n" % ((s - t0), (e_ - s), name.replace(",", ";")))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
