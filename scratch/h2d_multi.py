"""Concurrent host-to-device bandwidth of N ranks (torchrun): where does the end-to-end path stop scaling?
H2D_MODE = plain (torch pin_memory) | bind (CPU affinity to the GPU's NUMA node first) | wc (write-combined pinned)
           | numa (bind + explicit first touch by the bound thread)
Prints per-rank GB/s and the aggregate; rank 0 also prints the topology once."""
import ctypes as C, os, subprocess, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
mode = os.environ.get("H2D_MODE", "plain")
torch.cuda.set_device(local)
ncores = 0
if mode in ("bind", "numa", "wc"):
    from dspsr_b200 import sharding
    ncores = sharding.bind_cpu_affinity(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 512 << 20
if mode == "wc":
    rt = C.CDLL("libcudart.so.12")
    p = C.c_void_p()
    assert rt.cudaHostAlloc(C.byref(p), C.c_size_t(n), C.c_uint(0x04)) == 0     # cudaHostAllocWriteCombined
    C.memset(p, 1, n)
    class V:
        __array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (p.value, False), "version": 3}
    import numpy as np
    h = torch.from_numpy(np.asarray(V()))
else:
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
with torch.cuda.stream(s):
    for _ in range(2): d.copy_(h, non_blocking=True)
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(s):
    e0.record(s)
    for _ in range(20): d.copy_(h, non_blocking=True)
    e1.record(s)
barrier()
gbs = 20 * n / e0.elapsed_time(e1) / 1e6
t = torch.tensor([gbs], device="cuda")
if world > 1:
    all_ = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(all_, t)
    vals = [float(x.item()) for x in all_]
else:
    vals = [gbs]
if rank == 0:
    print("mode=%s N=%d bound_cores=%d per-rank GB/s %s aggregate %.1f" % (mode, world, ncores, [round(v, 1) for v in vals], sum(vals)))
    if os.environ.get("H2D_TOPO"):
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
        print(subprocess.run(["lscpu"], capture_output=True, text=True).stdout[:1800])
if world > 1: dist.destroy_process_group()
