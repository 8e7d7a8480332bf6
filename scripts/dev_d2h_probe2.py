"""D2H rate with the copy split over several streams (does a second copy engine / more outstanding DMA help?)"""
import time
import numpy as np
import torch
dev = torch.device("cuda", 0)
N = 96 << 20
src = torch.empty(N, dtype=torch.uint8, device=dev); src.fill_(3)
dst = torch.empty(N, dtype=torch.uint8).pin_memory()
for ns in (1, 2, 3, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    part = N // ns
    ts = []
    for _ in range(8):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                dst[i * part:(i + 1) * part].copy_(src[i * part:(i + 1) * part], non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"D2H {N >> 20} MB over {ns} streams: best {N / min(ts) / 1e9:.1f} GB/s, median {N / np.median(ts) / 1e9:.1f} GB/s")
for chunk in (1 << 20, 4 << 20, 16 << 20, 32 << 20):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for o in range(0, N, chunk):
            dst[o:o + chunk].copy_(src[o:o + chunk], non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"D2H {N >> 20} MB in {chunk >> 20} MB chunks on one stream: best {N / min(ts) / 1e9:.1f} GB/s")
h = torch.empty(N, dtype=torch.uint8).pin_memory()
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); src.copy_(h, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
print(f"H2D {N >> 20} MB: best {N / min(ts) / 1e9:.1f} GB/s")
