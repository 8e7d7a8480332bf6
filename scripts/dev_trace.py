"""host-side timeline of the streamed run (MB_TRACE=2 prints per call): where the host thread waits vs launches
usage: dev_trace.py [n] [n_sites] [chunks,...] [calls]"""
import os, sys, time
os.environ.setdefault("MB_TRACE", "2")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libmat_b200 import synth
from libmat_b200.rpd import Context

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
chunk_list = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 4, 8]
N = int(sys.argv[4]) if len(sys.argv) > 4 else 4
mesh = synth.make_ball_mesh(n); sites = synth.make_spheres(ns)
ctx = Context(0)
ctx.set_mesh(mesh)
ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, None, 0)
ctx.run().free()
for chunks in chunk_list:
    for _ in range(3):
        ctx.run_to_host(n_chunks=chunks).free()
    sys.stderr.write(f"---- chunks {chunks}\n"); sys.stderr.flush()
    t0 = time.perf_counter()
    for _ in range(N):
        r = ctx.run_to_host(n_chunks=chunks); ms = r.kernel_ms; r.free()
    dt = (time.perf_counter() - t0) / N
    print(f"chunks {chunks}: run_to_host {dt*1e3:.3f} ms wall; device stages {ms}", flush=True)
ctx.close()
