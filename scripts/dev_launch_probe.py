"""kernel launch throughput with and without a bulk D2H copy in flight on another stream, plain launches vs a CUDA graph"""
import time
import torch
dev = torch.device("cuda", 0)
big = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
host = torch.empty(512 << 20, dtype=torch.uint8).pin_memory()
x = torch.zeros(1024, device=dev)
cs = torch.cuda.Stream()
N = 200


def burst():
    for _ in range(N):
        x.add_(1.0)


g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    burst()
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        burst()
torch.cuda.synchronize()
for name, fn in (("plain launches", burst), ("one graph launch", g.replay)):
    for copy in (False, True):
        ts = []
        for _ in range(5):
            torch.cuda.synchronize()
            if copy:
                with torch.cuda.stream(cs):
                    host.copy_(big, non_blocking=True)
                time.sleep(0.001)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
            torch.cuda.synchronize()
        print(f"{name:18s} D2H in flight={copy}: {N} tiny kernels in {min(ts):.3f} ms (best), {sorted(ts)[2]:.3f} ms (median) -> {1e3 * min(ts) / N:.1f} us per kernel")
