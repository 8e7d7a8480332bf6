"""dev: where does the reference's dist2mat DEVICE build differ from its host build / this library?"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libmat_b200 import synth
from libmat_b200.rpd import Context
from oracle import oracle as O

d = synth.make_dist2mat(200000)
rr, rc, _ = O.ref_d2m_gpu(d)
ctx = Context(0)
r, cid, tie = ctx.compute_closest_dist2mat(d.spheres, d.samples, d.offset, d.count, d.prims)
rh, ch, _ = O.dist2mat(d, "ref")
rel = np.abs(r - rr) / np.maximum(np.abs(rr), 1e-3)
bad = np.flatnonzero(rel > 1e-6)
print("samples", len(r), "ours!=refgpu", len(bad), "ours!=refhost", int((np.abs(r - rh) / np.maximum(np.abs(rh), 1e-3) > 1e-6).sum()))
print("id mismatch ours/refgpu", int((cid != rc).sum()), " ours/refhost", int((cid != ch).sum()))
# per-primitive: evaluate every primitive of the first bad samples on device and on host
l = O.ref("d2m")
for s in bad[:6]:
    off, cnt = int(d.offset[s]), int(d.count[s])
    pr = np.ascontiguousarray(d.prims[off:off + cnt])
    pos = np.ascontiguousarray(np.repeat(d.samples[s][None, :], cnt, axis=0).astype(np.float32))
    out = np.zeros(cnt, np.float32)
    rc_ = l.ref_d2m_eval_prims_gpu(O._p(np.ascontiguousarray(d.spheres, dtype=np.float32)), C.c_int(len(d.spheres)), O._p(pos), O._p(pr), C.c_int(cnt), O._p(out))
    host = np.zeros(cnt, np.float32)
    for i in range(cnt):
        p = pr[i]
        sp = d.spheres.astype(np.float32)
        if p[0] == -1 and p[1] == -1:
            host[i] = l.ref_d2m_sphere(O._p(pos[i]), O._p(np.ascontiguousarray(sp[p[2]])))
        elif p[0] == -1:
            host[i] = l.ref_d2m_cone(O._p(pos[i]), O._p(np.ascontiguousarray(sp[p[1]])), O._p(np.ascontiguousarray(sp[p[2]])))
        else:
            host[i] = l.ref_d2m_slab(O._p(pos[i]), O._p(np.ascontiguousarray(sp[p[0]])), O._p(np.ascontiguousarray(sp[p[1]])), O._p(np.ascontiguousarray(sp[p[2]])))
    dd = np.flatnonzero(np.abs(out - host) > 1e-6 * np.maximum(np.abs(host), 1e-3))
    print("sample", s, "ours", r[s], cid[s], "refgpu", rr[s], rc[s], "refhost", rh[s], ch[s], "prims differing dev/host:", [(int(i), pr[i].tolist(), float(out[i]), float(host[i])) for i in dd[:5]])
