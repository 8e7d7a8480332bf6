"""Developer diagnostic (GPU box): wall-clock breakdown of the e2e path + PCIe bandwidth."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from libmat_b200 import synth
from libmat_b200.rpd import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
t = time.time(); mesh = synth.make_ball_mesh(n); sites = synth.make_spheres(ns); print("synth", time.time() - t, mesh.n_tet, mesh.n_vert)
ctx = Context(0)
def T(f, reps=5):
    torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best * 1e3, r
ms, _ = T(lambda: ctx.set_mesh(mesh)); print("set_mesh ms", ms)
ms, _ = T(lambda: ctx.upload_sites(sites.site_soa, sites.weights, sites.flags)); print("upload_sites ms", ms)
res = None
def run():
    global res
    if res: res.free()
    res = ctx.run(); return res
ms, _ = T(run); print("run ms (wall)", ms, res.kernel_ms, "cells", res.n_cells, "pairs", res.n_pairs, "bytes", res.compact_bytes, "ovf", res.n_cand_overflow, "big", res.n_big_pass_tets, "exact", res.n_exact, "clips", res.n_clips)
tb = torch.empty(res.compact_bytes // 4 + 16, dtype=torch.int32).pin_memory(); to = torch.empty(res.n_cells + 16, dtype=torch.int64).pin_memory()
ms, _ = T(lambda: ctx._check(ctx.lib.mb_rpd_fetch_compact(res._h, tb.numpy().ctypes.data, to.numpy().ctypes.data))); print("fetch_compact pinned ms", ms, "GB/s", res.compact_bytes / ms / 1e6)
ms, _ = T(lambda: res.records(), reps=2); print("records() (D2H + expansion to 3456-B records) ms", ms)
# raw PCIe
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); h = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
ms, _ = T(lambda: h.copy_(d)); print("D2H 256MB pinned GB/s", 268.4 / ms)
ms, _ = T(lambda: d.copy_(h)); print("H2D 256MB pinned GB/s", 268.4 / ms)
print("status hist", res.status_histogram.tolist())
