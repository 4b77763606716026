"""Conversions at the API edge: numpy / torch in, the caller's kind out."""
import numpy as np
import torch

_ops = None
TIMERS = {"widen_s": 0.0, "result_wait_s": 0.0, "fetch_many_s": 0.0}     # host-side diagnostics (bench.py)


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


# ---- bulk device -> host transfer of a result set -------------------------------------------
_pinned = {"buf": None}
_pool = None


def fetch_many(tensors):
    """{name: CUDA tensor} -> {name: fresh numpy array}.  All copies go through one cached pinned
    staging buffer (async D2H on the current stream, one synchronisation), then are copied out to
    freshly allocated numpy arrays by a few threads (numpy releases the GIL while copying), so the
    caller owns the results and the next call may reuse the staging buffer."""
    global _pool
    import concurrent.futures as cf
    import time
    t_start = time.perf_counter()
    items = [(k, t.contiguous()) for k, t in tensors.items()]
    offs, total = [], 0
    for _, t in items:
        offs.append(total)
        total += (t.numel() * t.element_size() + 255) // 256 * 256
    if _pinned["buf"] is None or _pinned["buf"].numel() < total:
        _pinned["buf"] = torch.empty((max(total, 1),), dtype=torch.uint8, pin_memory=True)
    stage = _pinned["buf"]
    views = []
    for (k, t), o in zip(items, offs):
        nbytes = t.numel() * t.element_size()
        v = stage[o:o + nbytes].view(t.dtype).view(t.shape)
        v.copy_(t, non_blocking=True)
        views.append(v)
    torch.cuda.current_stream().synchronize()
    if _pool is None:
        _pool = cf.ThreadPoolExecutor(max_workers=8)
    out, jobs = {}, []
    for (k, t), v in zip(items, views):
        src = v.numpy()
        dst = np.empty(src.shape, dtype=src.dtype)
        out[k] = dst
        flat_s, flat_d = src.reshape(-1), dst.reshape(-1)
        n = flat_s.size
        step = max(1 << 20, (n + 7) // 8)
        for a in range(0, n, step):
            jobs.append(_pool.submit(np.copyto, flat_d[a:a + step], flat_s[a:a + step]))
    for j in jobs:
        j.result()
    TIMERS["fetch_many_s"] += time.perf_counter() - t_start
    return out


class AsyncFetch:
    """Device -> host transfer of one large tensor that overlaps with later GPU work: the copy runs
    on a side stream into a cached pinned buffer as soon as the producer kernel has finished, and a
    worker thread widens / copies it into a fresh numpy array (`out_dtype`) while the main stream
    keeps computing.  `result()` returns the array."""
    _stream = None
    _pins = {}                      # one cached pinned staging buffer per slot (transfers of different slots overlap)

    def __init__(self, t, out_dtype=np.float64, slot="images"):
        global _pool
        import concurrent.futures as cf
        if _pool is None:
            _pool = cf.ThreadPoolExecutor(max_workers=8)
        cls = AsyncFetch
        if cls._stream is None:
            cls._stream = torch.cuda.Stream()
        nbytes = t.numel() * t.element_size()
        pin = cls._pins.get(slot)
        if pin is None or pin.numel() < nbytes:
            pin = torch.empty((max(nbytes, 1),), dtype=torch.uint8, pin_memory=True)
            cls._pins[slot] = pin
        self._keep = t
        ready = torch.cuda.Event()
        ready.record()
        cls._stream.wait_event(ready)
        view = pin[:nbytes].view(t.dtype).view(t.shape)
        with torch.cuda.stream(cls._stream):
            view.copy_(t, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        src = view.numpy()
        self._out = np.empty(src.shape, dtype=out_dtype)

        def work():
            import time
            done.synchronize()
            t0 = time.perf_counter()
            fs, fd = src.reshape(-1), self._out.reshape(-1)
            n = fs.size
            step = max(1 << 20, (n + 7) // 8)
            jobs = [_pool.submit(np.copyto, fd[a:a + step], fs[a:a + step]) for a in range(0, n, step)]
            for j in jobs:
                j.result()
            TIMERS["widen_s"] += time.perf_counter() - t0
            return self._out

        import threading
        self._res = None
        self._thr = threading.Thread(target=lambda: setattr(self, "_res", work()), daemon=True)
        self._thr.start()

    def result(self):
        import time
        t0 = time.perf_counter()
        self._thr.join()
        TIMERS["result_wait_s"] += time.perf_counter() - t0
        self._keep = None
        return self._res
