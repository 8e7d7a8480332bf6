"""Developer diagnostic (GPU box): grid-kNN mode vs given-neighbours mode (regular-triangulation
neighbour lists) at a chosen size: cell sets, canonical parity, volume conservation, stage times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libmat_b200 import synth
from libmat_b200.rpd import Context
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
mesh = synth.make_ball_mesh(n); sites = synth.make_spheres(ns)
t = time.time(); knn, k, valid = synth.rt_site_lists(sites); print("rt lists", time.time() - t, "site_k", k, "valid", valid.sum())
sites.flags[:] = valid.astype(np.uint32)
ctx = Context(0); ctx.set_mesh(mesh)
for rep in range(2):
    rg = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
    print("[given] cells", rg.n_cells, "pairs", rg.n_pairs, "clips", rg.n_clips, "culled", rg.n_culled, "hist", rg.status_histogram.tolist(), rg.kernel_ms)
    a = rg.records(); rg.free()
tv = mesh.tet_volumes()
cva = O.cell_volumes(a); pva = np.zeros(mesh.n_tet); np.add.at(pva, a["tet_id"], cva)
print("[given] volume: cells", cva.sum(), "mesh", tv.sum(), "max rel", (np.abs(pva - tv) / tv).max())
for gk in (0, 0, 256):
    r = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0, grid_k=gk)
    print(f"[grid k={gk}] cells", r.n_cells, "pairs", r.n_pairs, "clips", r.n_clips, "culled", r.n_culled, "cand_overflow", r.n_cand_overflow, "hist", r.status_histogram.tolist(), r.kernel_ms)
    b = r.records()
    ka = a["tet_id"].astype(np.int64) * ns + a["voro_id"]; kb = b["tet_id"].astype(np.int64) * ns + b["voro_id"]
    oa, ob, cm = np.setdiff1d(ka, kb), np.setdiff1d(kb, ka), np.intersect1d(ka, kb)
    print("   only given", len(oa), "only grid", len(ob), "common", len(cm))
    if len(oa): print("     vol of only-given cells: max", cva[np.isin(ka, oa)].max())
    cv = O.cell_volumes(b); pv = np.zeros(mesh.n_tet); np.add.at(pv, b["tet_id"], cv)
    if len(ob): print("     vol of only-grid cells: max", cv[np.isin(kb, ob)].max())
    rel = np.abs(pv - tv) / tv
    print("   volume: cells", cv.sum(), "mesh", tv.sum(), "max rel", rel.max(), "n>1e-2", int((rel > 1e-2).sum()))
    d = O.defined_equal(O.canonicalize(a[np.isin(ka, cm)]), O.canonicalize(b[np.isin(kb, cm)]))
    print("   canonical parity on common cells:", d)
    pt, ps, st = r.pairs()
    cnt = np.bincount(pt, minlength=mesh.n_tet)
    print("   pairs/tet: mean", cnt.mean(), "max", cnt.max(), "p99", np.percentile(cnt, 99))
    bad = np.where(rel > 1e-2)[0][:5]
    for tt in bad:
        print("   bad tet", tt, "rel", rel[tt], "grid cells", b["voro_id"][b["tet_id"] == tt].tolist(), "given cells", a["voro_id"][a["tet_id"] == tt].tolist(),
              "grid pairs", ps[pt == tt].tolist(), "status", st[pt == tt].tolist())
    r.free()
