import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libmat_b200 import synth
from libmat_b200.rpd import Context
from oracle import oracle as O
mesh = synth.make_ball_mesh(6)
s = synth.make_spheres(200)
c = s.centers() * 0.15 + 200.0
r = s.radii * 0.15
sites = synth.Sites(np.ascontiguousarray(c.T.astype(np.float32)).ravel(), (r * r).astype(np.float32), np.ones(200, np.uint32), r.astype(np.float32))
knn, k, valid = synth.rt_site_lists(sites); sites.flags[:] = valid
pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
ra, sa, _ = O.run_pairs(mesh, sites, knn, k, pt, ps)
want = ra[ra["status"] == 4]
ctx = Context(0); ctx.set_mesh(mesh)
gv = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
print("given gpu cells", gv.n_cells, "oracle", len(want), O.defined_equal(want, gv.records()) if gv.n_cells == len(want) else "")
for gk in (0, 256):
    g = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0, grid_k=gk)
    got = g.records()
    print("grid k", gk, "cells", g.n_cells, "pairs", g.n_pairs, "ovf", g.n_cand_overflow, "big", g.n_big_pass_tets, "hist", g.status_histogram.tolist())
    ka = want["tet_id"].astype(np.int64) * 200 + want["voro_id"]; kb = got["tet_id"].astype(np.int64) * 200 + got["voro_id"]
    oa = np.setdiff1d(ka, kb); ob = np.setdiff1d(kb, ka)
    print("  only oracle", len(oa), "only grid", len(ob))
    va = O.cell_volumes(want[np.isin(ka, oa)]); print("  missing cell volumes", np.sort(va)[::-1][:10])
    for key in oa[:6]:
        t, sid = key // 200, key % 200
        ptt, pss, st = g.pairs()
        print("   missing (tet, site)", t, sid, "grid pairs of tet", pss[ptt == t].tolist(), st[ptt == t].tolist(), "oracle cells", want["voro_id"][want["tet_id"] == t].tolist())
