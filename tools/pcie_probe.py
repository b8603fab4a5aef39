"""How much device -> host bandwidth do N ranks get at the same time?  Every rank copies a 4 MB surface (the e2e step's
D2H payload) from its GPU into its own page-locked buffer, `reps` times, all ranks starting together; prints per-rank and
aggregate GB/s for the DMA engine (cudaMemcpyAsync) and for SM stores into mapped host memory (what the zero-copy
surface path does).   torchrun --nproc-per-node N tools/pcie_probe.py"""
import os, sys, time
import torch, torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 1_000_000
src = torch.rand(n, dtype=torch.float32, device=dev)
host = torch.empty(n, dtype=torch.float32, pin_memory=True)
reps = 50

def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()

def run(fn):
    for _ in range(5):
        fn()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return 4.0 * n * reps / dt / 1e9

def dma():
    host.copy_(src, non_blocking=True); torch.cuda.synchronize()

# SM stores into mapped host memory: a device-side tensor view of the pinned buffer
import ctypes
cudart = torch.cuda.cudart()
mapped = None
try:
    err, ptr = cudart.cudaHostGetDevicePointer(host.data_ptr(), 0) if hasattr(cudart, "cudaHostGetDevicePointer") else (1, 0)
except Exception:
    err, ptr = 1, 0
res = {"dma": run(dma)}
t = torch.tensor([res["dma"]], dtype=torch.float64, device=dev)
if world > 1:
    allv = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allv, t)
    vals = [float(v) for v in allv]
else:
    vals = [res["dma"]]
if rank == 0:
    print(f"N={world}: cudaMemcpyAsync D2H of 4 MB, all ranks at once: per rank {[round(v, 1) for v in vals]} GB/s, aggregate {sum(vals):.1f} GB/s, "
          f"4 MB in {4e-3 * n / 1e6 / min(vals) * 1e3:.3f} ms on the slowest rank")
if world > 1:
    dist.destroy_process_group()
