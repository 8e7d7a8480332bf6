import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libmat_b200 import synth
from libmat_b200.rpd import Context
from oracle import oracle as O
d = synth.make_dist2mat(20000)
# one sample per (sample, prim) entry
smp = np.repeat(np.arange(len(d.samples)), d.count)
samples = d.samples[smp]
n = len(smp)
d1 = synth.Dist2MatInput(d.spheres, samples, np.arange(n, dtype=np.uint32), np.ones(n, np.uint32), d.prims)
ctx = Context(0)
r, c, tie = ctx.compute_closest_dist2mat(d1.spheres, d1.samples, d1.offset, d1.count, d1.prims)
ro, co, to = O.dist2mat(d1, "ref")
kind = np.where(d.prims[:, 1] == -1, 0, np.where(d.prims[:, 0] == -1, 1, 2))
for k, name in enumerate(["sphere", "cone", "slab"]):
    m = kind == k
    eq = r[m].view(np.uint32) == ro[m].view(np.uint32)
    rel = np.abs(r[m] - ro[m]) / np.maximum(np.abs(ro[m]), 1e-6)
    print(name, m.sum(), "bit-equal", eq.mean(), "maxrel", np.nanmax(rel), "nan gpu", np.isnan(r[m]).sum(), "nan ref", np.isnan(ro[m]).sum())
    bad = np.where(m)[0][~eq][:5]
    for b in bad:
        print("   ", b, d.prims[b], samples[b], r[b], ro[b], [d.spheres[j].tolist() for j in d.prims[b] if j >= 0])
