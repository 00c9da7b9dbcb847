"""Host<->device copy bandwidth of the box, one process per GPU (torchrun): every rank alone, then all ranks at once.
Tells whether sliced per-rank transfers (N/P sites per PCIe link) scale on this host or share one bottleneck."""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def bw(nbytes, direction, reps=20):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        (d.copy_(h, non_blocking=True) if direction == "h2d" else h.copy_(d, non_blocking=True))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    return dt


for nbytes in (3 << 20, 25 << 20):
    for direction in ("h2d", "d2h"):
        t = bw(nbytes, direction)
        ts = torch.tensor([t], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"{direction} {nbytes >> 20:3d} MiB x {world} ranks at once: {1e3 * ts.item():.3f} ms each "
                  f"-> {world * nbytes / ts.item() / 1e9:.1f} GB/s aggregate", flush=True)
if world > 1:
    dist.destroy_process_group()
