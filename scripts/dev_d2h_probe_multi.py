"""aggregate D2H rate of N ranks copying concurrently into (a) their own cudaHostAlloc buffers, (b) one shared /dev/shm
segment page-locked by every rank -- the platform ceiling of the multi-GPU e2e leg.  torchrun --nproc-per-node N."""
import mmap
import os
import time

import numpy as np
import torch
import torch.distributed as dist

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
N = 128 << 20
src = torch.empty(N, dtype=torch.uint8, device="cuda")
src.fill_(rank)
own = torch.empty(N, dtype=torch.uint8).pin_memory()
box = [None]
if rank == 0:
    path = f"/dev/shm/probe_multi_{os.getpid()}"
    with open(path, "wb") as fh:
        fh.truncate(N * world)
    box = [path]
dist.broadcast_object_list(box, src=0)
fd = os.open(box[0], os.O_RDWR)
mm = mmap.mmap(fd, N * world)
os.close(fd)
shm = np.frombuffer(mm, dtype=np.uint8)
shm[rank * N:(rank + 1) * N:4096] = 0
dist.barrier()
cudart = torch.cuda.cudart()
assert int(cudart.cudaHostRegister(shm.ctypes.data, N * world, 0)) == 0
dist.barrier()
if rank == 0:
    os.unlink(box[0])
shared = torch.from_numpy(shm)[rank * N:(rank + 1) * N]
for name, dst in (("own cudaHostAlloc buffer per rank", own), ("one shared /dev/shm segment, cudaHostRegister", shared)):
    for active in sorted({1, 2, 4, world}):
        if active > world:
            continue
        ts = []
        for _ in range(5):
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if rank < active:
                for _ in range(4):
                    dst.copy_(src, non_blocking=True)
                torch.cuda.synchronize()
            t = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(t.item())
        if rank == 0:
            print(f"{name}: {active} ranks copying 4 x {N >> 20} MB each: aggregate {active * 4 * N / min(ts) / 1e9:.1f} GB/s "
                  f"({4 * N / min(ts) / 1e9:.1f} GB/s per rank)", flush=True)
dist.destroy_process_group()
