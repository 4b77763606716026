"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

    python bench.py --gpus N --steps K --warmup W            (N>1: under torch.distributed.run)
    python bench.py --impl reference --steps K --warmup W     (the reference's CPU path, oracle port)

metric: end-to-end seconds, N patterns -> orientations (synthesis -> kNN round distances ->
diffusion-map eigenvectors -> orientation fit).  Default workload at every N: BASELINE config 4,
N = 100 000 patterns of 128 x 128 from the reference's analytic object (n = 31), k = 150,
Epsilon = 0.7, proper random rotations (seed 0) -- the configuration the north-star target is stated on
(it fits one B200); `--workload cfg2|cfg3|cfg5|small` selects the others.  A "step" is one pass of the
whole path.
`value`  = seconds per pass, inputs resident in HBM (strong scaling: the same job on N GPUs).
`e2e`    = the same pass through the reference-facing call OrientationRecoverySolver.solve with
           host (numpy) inputs and host outputs (float64 like the reference's), copies inside the timed region.
`roofline` is for the kernel with the largest share of the pass: from cfg 3 up the fused tcgen05 Gram /
top-k kernel (algorithmic 2 rows N P flop per launch / its mean CUDA-event time, against the measured
dense BF16 = FP16 tensor peak), at cfg 2 the mat-vec of the eigen-solver; the other one is reported under
`roofline_other`.  `alt_operator` = the same pass with the other form of the Markov operator (dense FP32
rows <-> FP64 kNN lists; the dense form is the HBM-bound kernel the north star names),
`host_wall_ms_of_each_pass` = host clock per timed pass.
The reference arm times the oracle port on the box's host cores: the whole workload when one pass fits the
time box (cfg 2, small), otherwise a row subsample with the extrapolation model printed and
`extrapolated: true` (cfg 3-5: the unmodified reference needs >= 80 GB and hours there).
"""
import argparse
import contextlib
import gc
import io
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "volumetric-3d-stitching_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    "cfg2": dict(N=10000, n=31, n_p=64, k=150, eps=0.7, desc="BASELINE config 2: N=10k patterns 64x64, full pipeline, 1 B200"),
    "cfg3": dict(N=50000, n=31, n_p=128, k=150, eps=0.7, desc="BASELINE config 3: N=50k patterns 128x128, row-sharded"),
    "cfg4": dict(N=100000, n=31, n_p=128, k=150, eps=0.7, desc="BASELINE config 4: N=100k patterns 128x128, 8 B200, row-sharded ~40 GB Markov matrix"),
    "cfg5": dict(N=200000, n=31, n_p=256, k=150, eps=0.7, photons=1e8, desc="BASELINE config 5: N=200k patterns 256x256 with Poisson noise (1e8 photons per pattern), Gram tiles streamed (never materialised), 8 B200"),
    "small": dict(N=2000, n=15, n_p=31, k=150, eps=0.7, desc="debug size"),
}
METRIC = "end_to_end_seconds_patterns_to_orientations"
EXTRA_WARMUP = 4

# stdout carries exactly one JSON line: everything else any library prints (NCCL banners, progress
# prints of the stage API, ...) is routed to stderr at the file-descriptor level
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md's clocks line).
    In-process NVML polling from a thread (a few tens of microseconds per sample, every 10 ms); an
    `nvidia-smi -lms` child costs tens of milliseconds of driver time per poll, which a 100 ms timed region
    cannot absorb.  Falls back to one nvidia-smi query after the region if NVML cannot be loaded."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"),
               (0x80, "hw_power_brake_slowdown"))

    def __init__(self, index=0):
        self.index = index
        self.samples = []          # (sm_mhz, reasons bit mask)
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        self._nvml = None
        self._handle = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self._nvml, self._handle = pynvml, handle
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self._sample()         # first call pays NVML's lazy initialisation, outside the timed region
            self.samples.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    def _sample(self):
        nv, h = self._nvml, self._handle
        sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        self.samples.append((sm, int(get(h))))

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception:
                break
            self._stop.wait(0.01)

    def reset(self):
        self.samples.clear()

    def stop(self):
        if self._nvml is None:
            return self._fallback()
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        sm = sorted(s for s, _ in self.samples)
        mask = 0
        for _, m in self.samples:
            mask |= m
        reasons = [name for bit, name in self.REASONS if mask & bit]
        return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(sm), "source": "nvml, 10 ms polling inside the timed region"}

    def _fallback(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=20)
            f = [x.strip() for x in r.stdout.strip().splitlines()[0].split(",")]
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6])
                       if v.lower().startswith("active")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": reasons, "samples": 1,
                    "source": "nvidia-smi, one query right after the timed region (NVML python binding unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference's NumPy/SciPy path on a bounded sample
# ------------------------------------------------------------------------------------------

def _synth_chunk(args):
    import v3s_oracle as orc
    R9, n, n_p = args
    return orc.generate_from_rotations(R9, n, n_p)


def cpu_reference_pass(wl, sample_n, cores):
    """One pass of the oracle port on `sample_n` patterns of the workload; every stage is timed
    and extrapolated to the full N with the stage's own complexity (stated in `sample`)."""
    import numpy as np
    import scipy.sparse.linalg as spla
    import v3s_oracle as orc
    from multiprocessing import get_context
    N, n, n_p, k, eps = wl["N"], wl["n"], wl["n_p"], wl["k"], wl["eps"]
    Ns = min(sample_n, N)
    R9, _ = orc.random_rotations(N, seed=0)
    R9s = R9[:, :Ns]
    t = {}
    t0 = time.perf_counter()
    chunks = [np.ascontiguousarray(c) for c in np.array_split(R9s, cores * 4, axis=1) if c.shape[1]]
    with get_context("fork").Pool(cores) as pool:
        imgs = np.concatenate(pool.map(_synth_chunk, [(c, n, n_p) for c in chunks]), axis=0)
    t["synth"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    S2, Nn = orc.knn(imgs, min(k, Ns - 2))
    t["knn"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    W = orc.anisotropic_norm(orc.s2_to_w_matrix(S2, Nn, eps))
    P_ep, di = orc.dm_cov(W, dense_D=False)
    lam, V = spla.eigs(P_ep, 10)                      # the reference's own eigen-solver call (diffusion.py:61)
    lam, V = np.real(lam), np.real(V)
    order = np.argsort(-lam)
    Ps = di[:, None] * (V[:, order] * lam[order][None, :])
    t["diffusion"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.dm_fit_all(Ps, R9s, True, check=False)
    t["fit"] = time.perf_counter() - t0
    f = N / Ns
    full = t["synth"] * f + (t["knn"] + t["diffusion"] + t["fit"]) * f * f
    return full, t, Ns


def _use_all_host_threads():
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core (numpy is
    # imported lazily, after this)
    n = str(os.cpu_count() or 1)
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = n


FULL_N_LIMIT = 12000            # the oracle port runs the whole workload up to this N (cfg 2: ~15-25 s per pass)


def reference_sample_size(wl, asked):
    """Patterns the CPU arm really processes: all of them when a pass fits the time box, else a subsample."""
    return wl["N"] if wl["N"] <= FULL_N_LIMIT else min(asked, wl["N"])


def cpu_baseline_block(wl, sample_n):
    _use_all_host_threads()
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    full, t, Ns = cpu_reference_pass(wl, reference_sample_size(wl, sample_n), cores)
    measured = time.perf_counter() - t0
    extrap = Ns < wl["N"]
    return {"value": round(full, 3), "unit": "s", "cores": cores, "kind": "port", "extrapolated": extrap,
            "measured_seconds": round(measured, 3), "patterns_run": Ns,
            "sample": ("oracle (vectorised numpy/scipy port of src/python) on %s %d of %d patterns: synth %.2fs%s, knn %.2fs, "
                       "W+normalise+eigs %.2fs, fit %.2fs%s; the as-is reference also spends ~3 us x N^2 in per-element Python "
                       "checks, not included"
                       % ("the first" if extrap else "all", Ns, wl["N"], t["synth"], " (x N/Ns)" if extrap else "", t["knn"],
                          t["diffusion"], t["fit"], " (each x (N/Ns)^2)" if extrap else ""))}


def workload_config(wl, world):
    """The `config` object of the JSON line -- the same for both arms (it names the workload, not the implementation)."""
    N, P = wl["N"], wl["n_p"] ** 2
    cfg = {"workload": wl["desc"], "N": N, "pixels": P, "object_voxels_1d": wl["n"], "k": wl["k"], "epsilon": wl["eps"],
           "rotations": "uniform random SO(3), seed 0", "parallelism": "rows%d" % world,
           "l2": "inputs larger than L2 (pattern matrix %.0f MB)" % (N * P * 4 / 1e6)}
    if wl.get("photons"):
        cfg["photons_per_pattern"] = wl["photons"]
    return cfg


def run_reference_arm(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    _use_all_host_threads()
    cores = os.cpu_count() or 1
    budget = float(os.environ.get("V3S_REFERENCE_BUDGET_S", "150"))
    sample = reference_sample_size(wl, args.cpu_sample)
    vals, walls, t_start, done = [], [], time.perf_counter(), 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        full, t, Ns = cpu_reference_pass(wl, sample, cores)
        walls.append(time.perf_counter() - t0)
        done += 1
        vals.append(full)
        # keep the whole arm within a few minutes whatever K and W the caller asks for: stop before the next
        # pass would leave the time box (a CPU pass has no warm-up effect worth discarding once the pool is forked)
        if time.perf_counter() - t_start + walls[-1] > budget:
            break
    timed = vals[min(args.warmup, len(vals) - 1):] if len(vals) > 1 else vals
    v = sum(timed) / len(timed)
    extrap = Ns < wl["N"]
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 3), "unit": "s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(v * 1e3, 1), "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(wl, int(os.environ.get("WORLD_SIZE", "1"))),
            "steps_run": done, "extrapolated": extrap,
            "wall_seconds_of_each_pass_run": [round(w, 2) for w in walls],
            "cpu_baseline": {"value": round(v, 3), "unit": "s", "cores": cores, "kind": "port", "extrapolated": extrap,
                             "patterns_run": Ns,
                             "sample": ("oracle port on %s %d of %d patterns, %d of %d passes run inside the %d s time box, %.1f s of "
                                        "CPU work per pass%s"
                                        % ("the first" if extrap else "all", Ns, wl["N"], done, args.warmup + args.steps, int(budget),
                                           sum(walls) / len(walls),
                                           "; value = stage times extrapolated to N (synth x N/Ns, kNN / diffusion / fit x (N/Ns)^2): a model, "
                                           "not a timing of the whole workload" if extrap else "; value = the measured time, no extrapolation"))},
            "e2e": {"value": round(v, 3), "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="v3s", choices=["v3s", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample", type=int, default=8000,
                    help="patterns of the CPU arm when the workload is too large to run whole (N > %d)" % FULL_N_LIMIT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-only", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--operator", default="auto", choices=["auto", "dense", "sparse"],
                    help="Markov operator of the eigen-solver (auto: the form a cost model predicts to be faster -- "
                         "dense FP32 rows for small N, the FP64 kNN-list form otherwise)")
    ap.add_argument("--no-alt-operator", action="store_true", help="skip the extra passes with the other operator form")
    ap.add_argument("--eig-driver", default=None, choices=["python", "cabi"],
                    help="eigen-solver driver: the Python host loop or the whole solver behind v3s_eig_topk (default: the backend's)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        return run_reference_arm(args, wl)
    if args.cpu_baseline_only:                       # child process of the v3s arm (see below)
        return emit(cpu_baseline_block(wl, args.cpu_sample))

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    cpu_base = None

    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the v3s path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # the all-gather of the FP32 rows runs beside the persistent Gram kernel on the SMs it leaves free
        os.environ.setdefault("NCCL_MAX_CTAS", "16")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from v3s import _lib, _util
    from v3s.pipeline import ShardPlan, run_pipeline
    from v3s.projection import Projector
    from v3s.rotation import Rotation
    from v3s.solution import OrientationRecoverySolver
    ops = _util.get_ops()
    ops.operator = args.operator
    if args.eig_driver:
        ops.eig_driver = args.eig_driver
    N, n, n_p, k, eps = wl["N"], wl["n"], wl["n_p"], wl["k"], wl["eps"]
    plan = ShardPlan.from_env(N)
    R9, _ = Rotation.random_rotations(N, seed=0)
    ed = ops.to_device(Projector.synsthesize_3d(n)["ED"].astype(np.float32), torch.float32)
    rot9_all = Projector.rot9_rowmajor(R9, ops)
    rot9 = rot9_all[plan.lo:plan.hi].contiguous()
    rot_y = ops.to_device(np.ascontiguousarray(R9.T), torch.float64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass(stats=None):
        if stats is not None:
            stats.clear()
        return run_pipeline(ops, plan, ed, rot9, rot_y, n_p, k, eps, polar_mode=0, check=False, stats=stats, validate=False,
                            photons_per_pattern=wl.get("photons"))

    trace = None
    if os.environ.get("V3S_BENCH_TRACE"):
        # diagnostic: host timestamps around every C-ABI call (no synchronisation added) -> where a stalled pass lost its time
        trace = []
        _orig_call = _lib.call

        def _traced_call(name, *a):
            t0_ = time.perf_counter()
            r_ = _orig_call(name, *a)
            trace.append((name, t0_, time.perf_counter()))
            return r_
        import v3s.ops_cuda as _oc
        _lib.call = _oc.call = _traced_call

    # the clock sampler starts before the warm-up: NVML's initialisation must not fall into the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("V3S_BENCH_NO_SAMPLER"):
        sampler.start()
    # W warm-up passes as asked, plus EXTRA_WARMUP untimed ones: on every box tried the 6th pass of a process is
    # slower (0.6 ms to >100 ms, one-off driver-side work after ~3000 launches); it belongs to start-up, not to a step
    extra_warmup = EXTRA_WARMUP if N <= 20000 else 0       # (a start-up effect of ~20 ms passes; irrelevant for 0.5 s ones)
    for _ in range(args.warmup + extra_warmup):
        out = one_pass()
    barrier()
    if rank == 0:
        sampler.reset()                   # keep the samples of the timed region only
    ops.profile = not os.environ.get("V3S_BENCH_NO_PROFILE")
    ops.pop_timings()
    from v3s import eig as _eig
    _eig.HOST_TIMERS.update({"ritz_s": 0.0, "wait_s": 0.0})
    launches0 = int(_lib.lib.v3s_kernel_launches())
    stats = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # a generation-2 collection of the cyclic GC (torch + scipy imported: ~10^5 tracked objects) is a 50-100 ms
    # pause at a random point of a ~115 ms timed region: collect now, keep the collector off while timing
    gc.collect()
    gc.disable()
    barrier()
    e0.record()
    pass_wall = [time.perf_counter()]
    for _ in range(args.steps):
        if trace is not None:
            trace.append(("PASS", time.perf_counter(), time.perf_counter()))
        out = one_pass(stats)
        pass_wall.append(time.perf_counter())          # host clock only (no synchronisation): where a stall fell
    if trace is not None:
        trace.append(("END", time.perf_counter(), time.perf_counter()))
        ev = [i for i, r_ in enumerate(trace) if r_[0] in ("PASS", "END")]
        for a_, b_ in zip(ev[-args.steps - 1:-1], ev[-args.steps:]):
            seg = trace[a_:b_ + 1]
            total = (seg[-1][1] - seg[0][1]) * 1e3
            slow = []
            for x_, y_ in zip(seg[:-1], seg[1:]):
                if (x_[2] - x_[1]) > 2e-3:
                    slow.append("in %s %.1f ms" % (x_[0], (x_[2] - x_[1]) * 1e3))
                if (y_[1] - x_[2]) > 2e-3:
                    slow.append("%s -> %s %.1f ms" % (x_[0], y_[0], (y_[1] - x_[2]) * 1e3))
            sys.stderr.write("TRACE pass %.1f ms: %s\n" % (total, "; ".join(slow)))
    e1.record()
    barrier()
    gc.enable()
    ms = e0.elapsed_time(e1) / args.steps
    host_timers = dict(_eig.HOST_TIMERS)                 # of the timed region only (the e2e calls below add to them)
    launches = (int(_lib.lib.v3s_kernel_launches()) - launches0) // args.steps      # kernels of libv3s.so only
    kt = ops.pop_timings()
    ops.profile = False
    clocks = sampler.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    sigma_T = float(out["sigma_T"].item())
    lam = out["Lambda"].cpu().numpy()
    if not stats.get("converged", False):
        raise SystemExit("bench.py: the eigen-solver did not converge (%r): no valid number" % (stats,))
    from v3s import orientation_errors_deg
    ang = orientation_errors_deg(out["recon_pre"], R9)     # per-pattern geodesic error after one global alignment (correct metric)
    med_err, p90_err = float(np.median(ang)), float(np.percentile(ang, 90))
    knn_over = int(out["knn_overflow"].item()) if out.get("knn_overflow") is not None else None
    del ang
    sparse_run = stats.get("operator") == "sparse-fp64"
    nnz_local = None
    if sparse_run:      # stored entries of the rank's rows: its own lists + the transposed copies of the non-zero entries
        nnz_local = (plan.hi - plan.lo) * k + int((torch.isfinite(out["S2"]) & (out["N"] >= plan.lo) & (out["N"] < plan.hi)).sum().item())
    del out

    # ---- the same pass with the other form of the Markov operator (reported beside the headline) ----
    alt = None
    if not args.no_alt_operator and world == 1:          # (sharded: tests/multi_gpu_check.py covers the sparse form)
        try:
            ops.operator = "dense" if sparse_run else "sparse"
            alt_stats = {}
            alt_steps = args.steps if N <= 20000 else min(args.steps, 3)
            ops.profile = True
            ops.pop_timings()
            for _ in range(2 if N <= 20000 else 1):
                o2 = one_pass(alt_stats)
            ops.pop_timings()
            gc.collect()
            gc.disable()
            barrier()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(alt_steps):
                o2 = one_pass(alt_stats)
            a1.record()
            barrier()
            gc.enable()
            kt_alt = ops.pop_timings()
            ops.profile = False
            t_alt = torch.tensor([a0.elapsed_time(a1) / alt_steps], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(t_alt, op=dist.ReduceOp.MAX)
            lam2 = o2["Lambda"].cpu().numpy()
            alt = {"operator": alt_stats.get("operator"), "value": round(float(t_alt.item()) / 1e3, 5), "unit": "s",
                   "matvecs": alt_stats.get("matvecs"), "converged": alt_stats.get("converged"), "steps": alt_steps,
                   "matvec_ms_per_launch": (round(sum(kt_alt["matvec"]) / len(kt_alt["matvec"]), 4) if kt_alt.get("matvec") else None),
                   "max_abs_lambda_difference_to_headline": float(np.max(np.abs(lam2 - lam))),
                   "sigma_T_deg": round(float(o2["sigma_T"].item()), 4),
                   "note": "same device-resident pass, other operator form (dense FP32 row block <-> FP64 kNN-list form); not the headline"}
            del o2
        except Exception as e_alt:                      # never lose the headline line over the side measurement
            gc.enable()
            alt = {"error": repr(e_alt)}
        ops.operator = args.operator

    # ---- e2e: reference-facing call, host inputs -> host outputs, on every rank of the job ----
    e2e = None
    if not args.no_e2e:
        params = {"n_voxel_1d": n, "nRot1D": 22, "k": k, "Epsilon": eps, "plot": False, "aPrioriFlag": True}
        def call():
            with contextlib.redirect_stdout(io.StringIO()):
                return OrientationRecoverySolver.solve(params, rotations=R9, N_p=n_p, check_sigma=False,
                                                       photons_per_pattern=wl.get("photons"))
        for _ in range(2):
            res = call()
        barrier()
        for k_ in _util.TIMERS:
            _util.TIMERS[k_] = 0.0
        del res
        gc.collect()
        gc.disable()
        t0 = time.perf_counter()
        reps = max(2, min(args.steps, 10 if N <= 20000 else 5))
        for _ in range(reps):
            res = call()
        barrier()
        gc.enable()
        host_e2e = {k_: round(v_ / reps * 1e3, 3) for k_, v_ in _util.TIMERS.items()}
        t_e2e = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e_s = float(t_e2e.item())
        # bytes actually copied: rotations (two layouts) + object in; every output travels as the reference's
        # float64 (widened on the GPU, copied into reusable pinned buffers while stages 2-4 run)
        h2d = R9.nbytes * 2 + ed.numel() * 4
        d2h = sum(v.nbytes for kk_, v in res.items() if kk_ not in ("R", "Axis", "Quat", "rows", "c0"))
        e2e = {"value": round(e2e_s, 5), "unit": "s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "host_ms": host_e2e,
               "note": "OrientationRecoverySolver.solve with numpy inputs/outputs on each rank; max over ranks; bytes per rank; "
                       "every output float64 like the reference's (the pattern matrix is %.1f GB of it: the transfer overlaps "
                       "stages 2-4 and is what bounds this number when it is slower than they are)" % (res["Images"].nbytes / 1e9)}
        # the same call returning the pattern matrix in the FP32 it is computed in (half the PCIe bytes)
        del res
        import numpy as _np

        def call32():
            with contextlib.redirect_stdout(io.StringIO()):
                return OrientationRecoverySolver.solve(params, rotations=R9, N_p=n_p, check_sigma=False,
                                                       photons_per_pattern=wl.get("photons"), images_dtype=_np.float32)
        res = call32()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            res = call32()
        barrier()
        t32 = torch.tensor([(time.perf_counter() - t0) / 2], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t32, op=dist.ReduceOp.MAX)
        e2e["with_float32_images"] = {"value": round(float(t32.item()), 5), "unit": "s",
                                      "d2h_bytes_per_step": int(sum(v.nbytes for kk_, v in res.items() if kk_ not in ("R", "Axis", "Quat", "rows", "c0")))}
        del res

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # CPU baseline (rank 0, N=1 only) in a child process, after the GPU measurements: its BLAS thread pools,
    # forked workers and multi-GB temporaries stay out of the process (and of the time window) being measured
    if world == 1 and not args.no_cpu_baseline:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-only", "--workload", args.workload,
                                "--cpu-sample", str(args.cpu_sample)], capture_output=True, text=True, timeout=900)
            cpu_base = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:
            cpu_base = {"value": None, "unit": "s", "cores": os.cpu_count(), "kind": "port", "sample": "failed: %r" % (e,)}
    peaks, peak_src = load_peaks()
    P = n_p * n_p
    gram_ms = kt.get("gram", [])
    gram_per_pass = sum(gram_ms) / args.steps if gram_ms else float("nan")
    rows_local = plan.hi - plan.lo
    gram_flop = 2.0 * rows_local * N * P                       # algorithmic: no split-precision multiplier
    achieved = gram_flop / (gram_per_pass * 1e-3) / 1e12 if gram_ms else None
    mv = kt.get("matvec", [])
    mv_bytes = (12.0 * nnz_local + 64.0 * N + 64.0 * rows_local) if sparse_run else (4.0 * rows_local * N + 8.0 * N * 8 * 2)
    mv_gbs = mv_bytes / (sum(mv) / len(mv) * 1e-3) / 1e9 if mv else None
    gaps = sorted(kt.pop("matvec_gap", []))
    stage_ms = {name: round(sum(v) / args.steps, 3) for name, v in kt.items()}
    gap_stats = None
    if gaps:
        inner = [g for g in gaps if g < 5.0]                 # (the gap between two passes is the rest of the pipeline)
        gap_stats = {"median_us": round(1e3 * gaps[len(gaps) // 2], 1), "p10_us": round(1e3 * gaps[len(gaps) // 10], 1),
                     "p90_us": round(1e3 * gaps[(9 * len(gaps)) // 10], 1), "sum_ms_per_pass": round(sum(inner) / args.steps, 3),
                     "largest_ms": [round(g, 2) for g in gaps[-4 * args.steps:]],
                     "note": "device time between the end of one mat-vec kernel and the start of the next: the fused reduction / "
                             "recurrence / exchange kernel with its cross-rank flag barrier, Lanczos basis kernels, host gaps"}
    line = {
        "metric": METRIC, "value": round(ms / 1e3, 5), "unit": "s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms, 3), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, world),
        "untimed_passes_before_the_timed_region": args.warmup + extra_warmup,
        "arithmetic": "FP16 tcgen05 Gram (one product, FP32 accumulate; ranks kNN candidates only), FP32 exact "
                      "distances / synthesis, FP64 Markov operator, reductions and eigen-solver",
        "clocks": clocks,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": None, "roofline_other": None,
        "kernel_ms_per_pass": stage_ms,
        "between_matvecs": gap_stats,
        "eigen": {k_: stats.get(k_) for k_ in ("driver", "matvecs", "probe_steps", "cheb_steps", "degree", "krylov_dim",
                                               "ritz_checks", "locking_rounds", "residual_max", "converged", "fused_matvec", "cheb_check_history")},
        "host_ms_per_pass": {"ritz_extraction": round(host_timers["ritz_s"] / args.steps * 1e3, 3),
                             "waiting_for_gpu_in_eigensolver": round(host_timers["wait_s"] / args.steps * 1e3, 3)},
        "result": {"sigma_T_deg": round(sigma_T, 4), "lambda_1_9": [round(float(x), 8) for x in lam[:10]],
                   "median_orientation_error_deg": round(med_err, 4), "p90_orientation_error_deg": round(p90_err, 4),
                   "orientation_error_note": "per-pattern geodesic error of the recovered rotations against the truth after one global "
                                             "alignment (v3s.orientation_errors_deg, proper SO(3) projection); sigma_T is the reference's "
                                             "bug-compatible all-pairs statistic",
                   "converged": bool(stats.get("converged")), "knn_lists_saturated": knn_over},
        "cpu_baseline": cpu_base,
        "host_wall_ms_of_each_pass": [round((b_ - a_) * 1e3, 2) for a_, b_ in zip(pass_wall[:-1], pass_wall[1:])],
        "operator": stats.get("operator"),
        "alt_operator": alt,
    }
    sym_run = "gram_limits_pass" in kt
    # MMAs really issued: rectangular = every tile of the row block; symmetric = 1/stride sample pass + tiles on/above the diagonal
    from v3s import ops_cuda as _oc
    gram_issued = gram_flop * ((1.0 / _oc.SYM_SAMPLE_STRIDE + 0.5 * (1.0 + 256.0 / N)) if sym_run else 1.0)
    long_kernel = gram_per_pass > 20.0                    # timed inside a long step: the sustained (power-capped) peak applies
    gram_peak = peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]) if long_kernel else peaks["bf16_tflops"]
    gram_roof = {"kernel": "gram_topk_kernel (tcgen05 cta_group::2 / TMEM / TMA, top-k selection in the epilogue)", "bound": "tensor",
                 "achieved": (round(achieved, 2) if achieved else None),
                 "peak": gram_peak, "unit": "TFLOP/s", "frac": (round(achieved / gram_peak, 4) if achieved else None),
                 "frac_of_burst_peak": (round(achieved / peaks["bf16_tflops"], 4) if achieved else None),
                 "traffic": None, "ms_per_pass": round(gram_per_pass, 3),
                 "peak_source": peak_src + (" bf16_tflops_sustained (the kernel runs %.0f ms inside a long step)" % gram_per_pass if long_kernel
                                            else " bf16_tflops (burst: the kernel runs %.1f ms)" % gram_per_pass),
                 "note": "algorithmic 2*rows*N*P flop per pass (SURVEY 8(d): no symmetry discount) over the time of the Gram kernels of the "
                         "pass; one FP16 product per element; G is never written (the epilogue keeps per-pattern candidate lists). "
                         "On one GPU from N = 16384 the kernels are the symmetric pair: a 1/8-sample pass fixing every pattern's "
                         "lower limit + the upper-triangle tiles serving rows and columns (0.625 of the rectangular MMA work), so "
                         "the algorithmic rate can exceed the pipe's peak; issued_mma_frac is what the tensor pipe really does",
                 "issued_mma_frac": (round((gram_issued / (gram_per_pass * 1e-3) / 1e12) / gram_peak, 4) if gram_ms else None),
                 "symmetric": bool(sym_run)}
    mv_roof = {"kernel": "block_matvec_kernel (TMA-fed, warp-shuffle-reduced)", "bound": "hbm", "achieved": (round(mv_gbs, 1) if mv_gbs else None),
               "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": (round(mv_gbs / peaks["hbm_gbs"], 4) if mv_gbs else None),
               "traffic": None, "ms_per_pass": round(sum(mv) / args.steps, 3) if mv else None, "launches_per_pass": len(mv) // args.steps,
               "peak_source": peak_src + " hbm_gbs", "note": "algorithmic 4*rows*N + 8*N*8*2 bytes per launch"}
    if sparse_run:
        mv_roof["kernel"] = "sparse_block_matvec_kernel (FP64 kNN-list operator, warp per row, shuffle-broadcast entries)"
        mv_roof["note"] = ("algorithmic HBM bytes per launch: 12 B per stored entry + the FP64 block read once and the rank's rows "
                           "written once; the kernel itself is bound by L2 gathers of the block (64 B per entry), not by HBM")
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(args.workload, {})
    except Exception:
        traffic = {}
    try:
        pipes = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json"))).get(args.workload + "_ncu_pipes")
    except Exception:
        pipes = None
    if world == 1 and pipes:
        line["ncu_pipe_utilisation"] = pipes          # committed ncu capture of this workload (not measured in this run)
    try:     # north-star kernel 1 (direct analytical synthesis, an optional mode): its committed FP32-pipe capture (cfg-2 size)
        line["ncu_direct_synthesis_fp32_pipe"] = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["cfg2_ncu_pipes"]["synth_direct_kernel"]
    except Exception:
        pass
    if world == 1:
        mv_roof["traffic"] = traffic.get("sparse_block_matvec_kernel" if sparse_run else "block_matvec_kernel")
        gram_roof["traffic"] = traffic.get("gram_topk_kernel")
        mv_roof["traffic_source"] = gram_roof["traffic_source"] = "profiles/r02_traffic.json (ncu --set full, bytes per launch)"
    dominant_is_mv = bool(mv) and (not gram_ms or sum(mv) >= sum(gram_ms))
    line["roofline"], line["roofline_other"] = (mv_roof, gram_roof) if dominant_is_mv else (gram_roof, mv_roof)
    line.pop("matvec", None)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
