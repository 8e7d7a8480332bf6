"""D2H rate probe: torch pinned (cudaHostAlloc) vs a /dev/shm segment page-locked with cudaHostRegister, and an
anonymous MADV_HUGEPAGE mapping page-locked the same way.  One GPU; prints GB/s for 16 / 92 / 256 MB copies."""
import mmap
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libmat_b200.rpd import Context

ctx = Context()
dev = torch.device("cuda", 0)
N = 256 << 20
src = torch.empty(N, dtype=torch.uint8, device=dev)
src.fill_(3)


def rate(dst_ptr, nbytes, reps=10):
    out = []
    dst = np.ctypeslib.as_array((__import__("ctypes").c_ubyte * nbytes).from_address(dst_ptr))
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.copy_to_host(dst, src.data_ptr(), nbytes)
        torch.cuda.synchronize()
        out.append(time.perf_counter() - t0)
    return nbytes / min(out) / 1e9, nbytes / float(np.median(out)) / 1e9


pinned = torch.empty(N, dtype=torch.uint8).pin_memory()
path = f"/dev/shm/probe_{os.getpid()}"
with open(path, "wb") as fh:
    fh.truncate(N)
fd = os.open(path, os.O_RDWR)
mm = mmap.mmap(fd, N)
os.close(fd)
os.unlink(path)
shm = np.frombuffer(mm, dtype=np.uint8)
shm[::4096] = 0
r = ctx.host_register(shm.ctypes.data, N)
print("register shm:", r)
an = mmap.mmap(-1, N + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
an.madvise(mmap.MADV_HUGEPAGE)
anon = np.frombuffer(an, dtype=np.uint8)
base = (anon.ctypes.data + (2 << 20) - 1) // (2 << 20) * (2 << 20)
off = base - anon.ctypes.data
anon[off:off + N:4096] = 0
r = ctx.host_register(base, N)
print("register anon THP:", r, [l for l in open("/proc/meminfo") if "AnonHuge" in l])
for nb in (16 << 20, 92 << 20, 256 << 20):
    print(f"{nb >> 20:4d} MB  pinned(best, median) {rate(pinned.data_ptr(), nb)}  shm-registered {rate(shm.ctypes.data, nb)}  "
          f"anon-THP-registered {rate(base, nb)}")
