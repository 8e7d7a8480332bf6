"""Developer check (GPU box): given-neighbours parity vs the CPU oracle / reference at config-1 size,
grid mode sanity, dist2mat parity.  Prints diagnostics; the real tests live in tests/."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libmat_b200 import synth
from libmat_b200.rpd import Context
from oracle import oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 15
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
mesh = synth.make_ball_mesh(n); sites = synth.make_spheres(ns)
knn, k = synth.knn_site_lists(sites, 80)
ctx = Context(0)
ctx.set_mesh(mesh)
for G in (8, 16, 32):
    t = time.time()
    res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k, lanes_per_cell=G)
    dt = time.time() - t
    print(f"[given G={G}] cells {res.n_cells} pairs {res.n_pairs} clips {res.n_clips} hist {res.status_histogram.tolist()} ms {res.kernel_ms} wall {dt*1e3:.1f} ms compact {res.compact_bytes}")
recs = res.records()
pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
print("oracle pairs", len(pt), "gpu pairs", res.n_pairs)
impl = "ref" if O.ref("rpd") is not None else "oracle"
ra, sa, ta = O.run_pairs(mesh, sites, knn, k, pt, ps, impl=impl)
ok = ra["status"] == 4
ra = ra[ok]
print("oracle valid", len(ra), "gpu valid", len(recs), "impl", impl)
if len(ra) == len(recs):
    ra2 = ra.copy(); ra2["thread_id"] = 0; 
    print("defined_equal:", O.defined_equal(ra2, recs))
    print("ids ok", np.array_equal(recs["id"], np.arange(len(recs))))
else:
    ka = set(zip(ra["tet_id"].tolist(), ra["voro_id"].tolist())); kb = set(zip(recs["tet_id"].tolist(), recs["voro_id"].tolist()))
    print("only oracle", len(ka - kb), "only gpu", len(kb - ka), list(ka - kb)[:5], list(kb - ka)[:5])
hist_o = np.bincount(sa + 1, minlength=10)
print("oracle stat hist", hist_o.tolist())
# emit
em = res.emit(mesh.n_surf_faces - 1)
ap, ae, eu = O.reload_active(recs, "oracle")
print("emit counts", {k_: len(v) for k_, v in em.items()}, "euler eq", np.array_equal(eu, em["cell_euler"]), "facets oracle", int(ap.sum()))
# grid mode
t = time.time()
rg = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, None, 0)
print(f"[grid] cells {rg.n_cells} pairs {rg.n_pairs} clips {rg.n_clips} hist {rg.status_histogram.tolist()} ms {rg.kernel_ms} wall {(time.time()-t)*1e3:.1f}")
rr = rg.records()
cv = O.cell_volumes(rr)
pv = np.zeros(mesh.n_tet); np.add.at(pv, rr["tet_id"], cv); tv = mesh.tet_volumes()
rel = np.abs(pv - tv) / tv
print("grid volume: sum cells", cv.sum(), "mesh", tv.sum(), "per-tet rel err max", rel.max(), "n>1e-3", int((rel > 1e-3).sum()), "n>1e-5", int((rel>1e-5).sum()))
# dist2mat
d = synth.make_dist2mat(200000)
t = time.time(); r, c, tie = ctx.compute_closest_dist2mat(d.spheres, d.samples, d.offset, d.count, d.prims); dt = time.time() - t
ro, co, to, s2 = O.dist2mat(d, "ref" if O.ref("d2m") is not None else "oracle", want_second=False) + (None,)
print(f"[d2m] wall {dt*1e3:.1f} ms cpu {to:.3f}s bit-equal {np.mean(r.view(np.uint32)==ro.view(np.uint32)):.6f} maxrel {np.max(np.abs(r-ro)/np.maximum(np.abs(ro),1e-6)):.3e} id-equal {np.mean(c==co):.6f} ties {int(tie.sum())} id-mismatch-not-tie {int(((c!=co)&(tie==0)).sum())}")
