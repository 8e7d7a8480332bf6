"""TEST INFRASTRUCTURE ONLY.  Generates the committed golden vectors under tests/golden/ by
running the REFERENCE's own code (oracle/_ref, built in place from /root/reference by
oracle/Makefile).  Run in the build container (the reference tree is absent on the GPU box):

    make -C oracle all && python oracle/gen_golden.py

KAT-1    single tet + 2 sites through the reference's clipped_voro_cell_test_GPU_param_tet
                    (src/rpd3d/convex_cell.cu:1166-1337, host build)            -> kat1_rpd.npz
KAT-2/2b            the reference's dist2mat distance functions
                    (src/dist2mat/dist2mat.cu:5-193, host build)                -> kat2_dist2mat.json
KAT-3               src/predicate_generator/main.cpp output                     -> kat3_predicates.json
MINI                seeded synthetic mini mesh (n=4 -> 384 tets, 40 sites, all-to-all neighbours):
                    reference records, reload_active, euler, vertex coordinates -> mini_rpd.npz
                    seeded dist2mat (2 000 samples): reference results          -> mini_dist2mat.npz
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from libmat_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def kat1_inputs():
    verts = np.array([[0, 0, 0], [1000, 0, 0], [0, 1000, 0], [0, 0, 1000]], np.float32)
    idx = np.array([[0, 1, 2, 3]], np.int32)
    mesh = synth.TetMesh(verts, idx, np.ones(4, np.int32), np.ones((1, 6), np.int32),
                         np.ones((1, 4), np.int32), np.arange(4, dtype=np.int32).reshape(1, 4), 4)
    c = np.array([[100, 100, 100], [600, 200, 150]], np.float32)
    r = np.array([50, 80], np.float32)
    sites = synth.Sites(np.ascontiguousarray(c.T).ravel(), r * r, np.ones(2, np.uint32), r)
    knn = np.array([[1, 0], [-1, -1]], np.int32)  # (site_k+1) x n_site, site_k = 1
    return mesh, sites, knn, 1


def mini_inputs():
    mesh = synth.make_ball_mesh(4)
    sites = synth.make_spheres(40)
    n = sites.n_site
    knn, k = synth.site_lists_from_sets([[m for m in range(n) if m != s] for s in range(n)], n)
    return mesh, sites, knn, k


def pack_records(recs):
    """only the defined fields, so that the fixture stays small"""
    return {f: recs[f] for f in ("status", "voro_id", "tet_id", "weight", "nb_v", "nb_p", "nb_e", "ver",
                                 "clip", "id2", "edge")}


def main():
    assert O.ref("rpd") is not None and O.ref("d2m") is not None and O.ref("host") is not None, \
        "build oracle/_ref first (make -C oracle all)"
    os.makedirs(OUT, exist_ok=True)

    # ---- KAT-1
    mesh, sites, knn, k = kat1_inputs()
    pt = np.array([0, 0], np.int32)
    ps = np.array([0, 1], np.int32)
    recs, stat, _, vol, _ = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="ref", want_vol=True)
    recs = O.zero_undefined(recs)
    np.savez_compressed(os.path.join(OUT, "kat1_rpd.npz"), stat=stat, site_vol=vol, **pack_records(recs))

    # ---- KAT-2 / 2b
    d = O.ref("d2m")
    f = lambda *a: [np.asarray(x, np.float32) for x in a]

    def sph(p, a):
        p, a = f(p, a)
        return float(d.ref_d2m_sphere(O._p(p), O._p(a)))

    def cone(p, a, b):
        p, a, b = f(p, a, b)
        return float(d.ref_d2m_cone(O._p(p), O._p(a), O._p(b)))

    def slab(p, a, b, c):
        p, a, b, c = f(p, a, b, c)
        return float(d.ref_d2m_slab(O._p(p), O._p(a), O._p(b), O._p(c)))

    pos = [0, 0.7, -0.35]
    m2, m3, m4 = [0.5, 0, 0, 0.35], [-0.5, 0, 0, 0.5], [0, 0, -0.6, 0.2]
    q = [0.3, 0.4, 0]
    cases = []

    def add(kind, p, prims, val):
        cases.append({"kind": kind, "pos": p, "prims": prims, "value": np.float32(val).item(),
                      "bits": int(np.float32(val).view(np.uint32))})

    add("slab", pos, [m2, m3, m4], slab(pos, m2, m3, m4))
    add("cone", pos, [m2, m3], cone(pos, m2, m3))
    add("cone", pos, [m3, m2], cone(pos, m3, m2))
    add("cone", pos, [m2, m4], cone(pos, m2, m4))
    add("cone", pos, [m3, m4], cone(pos, m3, m4))
    add("sphere", pos, [m2], sph(pos, m2))
    add("cone", [0.5, 1, 0], [[0, 0, 0, 0.3], [1, 0, 0, 0.3]], cone([0.5, 1, 0], [0, 0, 0, 0.3], [1, 0, 0, 0.3]))
    # KAT-2b: degenerate / branch cases
    add("cone", q, [[0, 0, 0, 0.1], [0, 0, 0, 0.2]], cone(q, [0, 0, 0, 0.1], [0, 0, 0, 0.2]))
    add("cone", q, [[0, 0, 0, 0.1], [0.05, 0, 0, 0.5]], cone(q, [0, 0, 0, 0.1], [0.05, 0, 0, 0.5]))
    s3 = [[0, 0, 0, 0.2], [1, 0, 0, 0.2], [0, 1, 0, 0.2]]
    add("slab", [0.25, 0.25, 0.5], s3, slab([0.25, 0.25, 0.5], *s3))
    add("slab", [2, 2, 0.5], s3, slab([2, 2, 0.5], *s3))
    s4 = [[0, 0, 0, 0.3], [1, 0, 0, 0.2], [0, 1, 0, 0.1]]
    add("slab", [0.3, 0.3, 0.6], s4, slab([0.3, 0.3, 0.6], *s4))
    s5 = [[0, 0, 0, 0.3], [0, 1, 0, 0.1], [0, 1, 0, 0.1]]
    add("slab", [0.3, 0.3, 0.6], s5, slab([0.3, 0.3, 0.6], *s5))
    # one radius equal to the third (R1 == 0, R2 != 0 and the converse branch)
    s6 = [[0, 0, 0, 0.2], [1, 0, 0, 0.3], [0, 1, 0, 0.2]]
    add("slab", [0.3, 0.3, 0.6], s6, slab([0.3, 0.3, 0.6], *s6))
    s7 = [[0, 0, 0, 0.3], [1, 0, 0, 0.2], [0, 1, 0, 0.2]]
    add("slab", [0.3, 0.3, 0.6], s7, slab([0.3, 0.3, 0.6], *s7))
    with open(os.path.join(OUT, "kat2_dist2mat.json"), "w") as fh:
        json.dump(cases, fh, indent=1)

    # ---- KAT-3
    out = subprocess.check_output([os.path.join(ROOT, "oracle", "_ref", "predgen")], text=True)
    vals = [line.split(":")[1].strip() for line in out.strip().splitlines()]
    with open(os.path.join(OUT, "kat3_predicates.json"), "w") as fh:
        json.dump({"bound_double": vals[0], "bound_float": vals[1],
                   "source": "src/predicate_generator/main.cpp (reference), stdout"}, fh, indent=1)

    # ---- MINI RPD
    mesh, sites, knn, k = mini_inputs()
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    recs, stat, _, vol, bary = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="ref", n_threads=1, want_vol=True)
    recs = O.zero_undefined(recs)
    ok = recs["status"] == 4
    ap, ae, eu = O.reload_active(recs[ok], "ref")
    vc = O.vertex_coordinates(recs[ok], "ref")
    np.savez_compressed(os.path.join(OUT, "mini_rpd.npz"), pair_tet=pt, pair_site=ps, stat=stat,
                        site_vol=vol, site_bary=bary, active_planes=np.packbits(ap, axis=1),
                        active_edges=np.packbits(ae, axis=1), euler=eu, vertex_xyzw=vc,
                        **pack_records(recs))

    # ---- MINI dist2mat
    dd = synth.make_dist2mat(2000, nu=20, nv=40, n_slabs=2400, n_cones=1200)
    res, cid, _ = O.dist2mat(dd, "ref", n_threads=1)
    np.savez_compressed(os.path.join(OUT, "mini_dist2mat.npz"), result=res, closest_id=cid)
    print("golden vectors written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(f"  {fn}: {os.path.getsize(os.path.join(OUT, fn))} bytes")


if __name__ == "__main__":
    main()
