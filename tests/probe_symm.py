"""2-rank probe (not a pytest file): does torch symmetric memory give peer-mapped pointers our
kernels can store through?  torchrun --nproc-per-node 2 tests/probe_symm.py"""
import os, sys, ctypes as C
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "volumetric-3d-stitching_b200"))
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from v3s import _lib
import torch.distributed._symmetric_memory as symm
t = symm.empty((1024,), dtype=torch.float32, device=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
t.fill_(-1.0)
hdl = symm.rendezvous(t, dist.group.WORLD)
ptrs = list(hdl.buffer_ptrs)
print(rank, "buffer_ptrs", [hex(p) for p in ptrs], "own", hex(t.data_ptr()), flush=True)
dist.barrier()
peer = (rank + 1) % world
_lib.call("v3s_peer_fill", C.c_void_p(ptrs[peer] + 4 * 100 * rank), C.c_float(10.0 + rank), 100, _lib.stream_ptr())
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
src = (rank - 1) % world
got = t[100 * src:100 * src + 100]
print(rank, "received from", src, got[:3].tolist(), "ok" if bool((got == 10.0 + src).all()) else "MISMATCH", flush=True)
dist.destroy_process_group()
