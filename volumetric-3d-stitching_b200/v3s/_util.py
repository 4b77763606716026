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
        _pool = cf.ThreadPoolExecutor(max_workers=_host_threads())
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


def _host_threads():
    """Worker threads for host-side copies: the cores of the box shared by the ranks of the job."""
    import os
    world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")) or 1)
    return max(1, min(8, (os.cpu_count() or 1) // max(world, 1)))


class AsyncFetch:
    """Device -> host transfer of one large tensor that overlaps with later GPU work.  The tensor is
    widened to `out_dtype` ON THE GPU (side stream, in chunks through one cached device buffer) and
    copied straight into a cached pinned host buffer; `result()` waits for the copy and returns a
    numpy view of that buffer -- no host-side pass over the data, no allocation per call.

    The buffer of a slot is reused: an array returned by `result()` is valid until the second next
    transfer of the same slot starts (two alternating buffers per slot; one for arrays above
    `SINGLE_BUFFER_BYTES`, which are then valid until the next transfer).  Callers that keep results
    across calls copy them (`OrientationRecoverySolver.solve(..., copy_results=True)`)."""
    _stream = None
    _pins = {}                      # slot -> [buffers], [next index]
    _chunk = {}                     # slot -> cached device staging buffer (uint8)
    SINGLE_BUFFER_BYTES = 1 << 30
    import os as _os
    CHUNK_BYTES = int(_os.environ.get("V3S_FETCH_CHUNK_MB", "256")) << 20

    def __init__(self, t, out_dtype=np.float64, slot="images", private=False):
        """private=True: `result()` returns a freshly allocated array the caller owns (copied out of the pinned
        buffer by a worker thread as soon as the transfer is complete, i.e. while the GPU is still busy with the
        later stages) instead of a view of the reusable buffer."""
        cls = AsyncFetch
        if cls._stream is None:
            cls._stream = torch.cuda.Stream()
        tdt = torch.from_numpy(np.empty(0, dtype=out_dtype)).dtype
        nbytes = t.numel() * tdt.itemsize
        nbuf = 1 if nbytes > cls.SINGLE_BUFFER_BYTES else 2
        ent = cls._pins.get(slot)
        if ent is None or ent[0][0].numel() < nbytes or len(ent[0]) != nbuf:
            ent = [[torch.empty((max(nbytes, 1),), dtype=torch.uint8, pin_memory=True) for _ in range(nbuf)], 0]
            cls._pins[slot] = ent
        pin = ent[0][ent[1] % nbuf]
        ent[1] += 1
        self._keep = t
        ready = torch.cuda.Event()
        ready.record()
        cls._stream.wait_event(ready)
        view = pin[:nbytes].view(tdt).view(t.shape)
        with torch.cuda.stream(cls._stream):
            if t.dtype == tdt:
                view.copy_(t, non_blocking=True)
            else:
                rows = t.shape[0]
                row_bytes = max(1, (t.numel() // max(rows, 1)) * tdt.itemsize)
                step = max(1, cls.CHUNK_BYTES // row_bytes)
                need = min(rows, step) * row_bytes
                dev = cls._chunk.get(slot)
                if dev is None or dev.numel() < need:
                    dev = torch.empty((max(need, 1),), dtype=torch.uint8, device=t.device)
                    cls._chunk[slot] = dev
                for r0 in range(0, rows, step):
                    r1 = min(rows, r0 + step)
                    tmp = dev[: (r1 - r0) * row_bytes].view(tdt).view((r1 - r0,) + tuple(t.shape[1:]))
                    tmp.copy_(t[r0:r1])                       # widening cast on the GPU
                    view[r0:r1].copy_(tmp, non_blocking=True)
            self._done = torch.cuda.Event()
            self._done.record()
        self._out = view.numpy()
        self._future = None
        if private:
            global _pool
            if _pool is None:
                import concurrent.futures as cf
                _pool = cf.ThreadPoolExecutor(max_workers=_host_threads())
            self._future = _pool.submit(self._copy_out)

    def _copy_out(self):
        self._done.synchronize()
        dst = np.empty(self._out.shape, dtype=self._out.dtype)
        np.copyto(dst, self._out)                      # (numpy releases the GIL for the copy)
        return dst

    def result(self):
        import time
        t0 = time.perf_counter()
        if self._future is not None:
            out = self._future.result()
        else:
            self._done.synchronize()
            out = self._out
        TIMERS["result_wait_s"] += time.perf_counter() - t0
        self._keep = None
        return out
