import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from libmat_b200 import synth
from libmat_b200.rpd import Context
from oracle import oracle as O
mesh = synth.make_ball_mesh(15); sites = synth.make_spheres(1000)
knn, k, valid = synth.rt_site_lists(sites); sites.flags[:] = valid.astype(np.uint32)
ctx = Context(0); ctx.set_mesh(mesh)
res = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k)
recs = res.records(); em = res.emit(mesh.n_surf_faces - 1); tp = res.topology()
want = O.topology(em, recs["voro_id"], em["cell_euler"])
bad = np.nonzero(tp["cell_cc"] != want["cell_cc"])[0]
print("cells", len(recs), "mismatch", len(bad), bad[:10])
for c in bad[:5]:
    s = recs["voro_id"][c]
    print("cell", c, "site", s, "gpu", tp["cell_cc"][c], "want", want["cell_cc"][c])
    fs = np.nonzero(em["facet_cell"] == c)[0]
    print("  facets", [(int(em["facet_key"][f]), int(em["facet_is_tet"][f])) for f in fs])
    for f in fs:
        if em["facet_is_tet"][f]:
            same = np.nonzero((em["facet_key"] == em["facet_key"][f]) & (em["facet_is_tet"] == 1))[0]
            print("   tfid", int(em["facet_key"][f]), "cells", [(int(em["facet_cell"][g]), int(recs["voro_id"][em["facet_cell"][g]])) for g in same])
badf = np.nonzero(tp["facet_cc"] != want["facet_cc"])[0]
print("facet mismatch", len(badf), badf[:10])
