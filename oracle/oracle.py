"""TEST INFRASTRUCTURE ONLY: ctypes access to oracle/liboracle.so (plain-C restatement) and
oracle/_ref/*.so (the reference's own sources compiled in place).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package libmat_b200 never does.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# numpy view of ConvexCellTransfer (reference src/rpd3d/convex_cell.h:189-217), 3456 bytes
RECORD_DTYPE = np.dtype({
    "names": ["status", "thread_id", "voro_id", "tet_id", "weight", "is_active", "nb_v", "nb_p",
              "nb_e", "ver", "clip", "id2", "edge", "euler", "cell_vol", "id"],
    "formats": ["<i4", "<i4", "<i4", "<i4", "<f4", "u1", "u1", "u1", "u1", ("u1", (96, 4)),
                ("<f4", (64, 8)), ("<i4", (64, 2)), ("u1", (152, 3)), "<f4", "<f4", "<i4"],
    "offsets": [0, 4, 8, 12, 16, 20, 21, 22, 23, 24, 416, 2464, 2976, 3432, 3436, 3440],
    "itemsize": 3456,
})

STATUS = {"early_return": -1, "triangle_overflow": 0, "vertex_overflow": 1,
          "inconsistent_boundary": 2, "security_radius_not_reached": 3, "success": 4,
          "needs_exact_predicates": 5, "no_intersection": 6, "edge_overflow": 7, "needs_perturb": 8}


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _load(path):
    if not os.path.exists(path):
        return None
    return C.CDLL(path)


_oracle = None
_refs = {}


def lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)
        _oracle = C.CDLL(path)
        _oracle.orc_rpd_run_pairs.restype = C.c_double
        _oracle.orc_tet_sphere_relation.restype = C.c_long
        _oracle.orc_dist2mat.restype = C.c_double
        for f in ("orc_distance_to_sphere", "orc_distance_to_cone", "orc_distance_to_slab"):
            getattr(_oracle, f).restype = C.c_float
    return _oracle


def ref(name):
    """name in {'rpd','host','d2m','rpd_gpu'}; returns None if the prebuilt .so is absent."""
    if name not in _refs:
        l = _load(os.path.join(HERE, "_ref", f"libref_{name}.so"))
        if l is not None:
            if name in ("rpd", "rpd_filter"):
                l.ref_rpd_run_pairs.restype = C.c_double
            if name == "d2m":
                l.ref_d2m_run_host.restype = C.c_double
                l.ref_d2m_run_gpu.restype = C.c_double
                l.ref_d2m_kernel_ms.restype = C.c_double
                for f in ("ref_d2m_sphere", "ref_d2m_cone", "ref_d2m_slab"):
                    getattr(l, f).restype = C.c_float
            if name == "rpd_gpu":
                l.ref_rpd_gpu_run.restype = C.c_long
        _refs[name] = l
    return _refs[name]


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def set_security_radius(on: bool):
    """a9: the oracle port's security-radius option (the blocks the live reference comments out); tests only"""
    lib().orc_set_security_radius(C.c_int(int(on)))


def run_pairs(mesh, sites, site_knn, site_k, pair_tet, pair_site, impl="oracle", n_threads=0,
              want_vol=False):
    """Run the per-(tet,site) clipping of the reference kernel body on the CPU.
    impl='oracle' -> plain-C restatement, impl='ref' -> reference sources (oracle/_ref).
    Returns (records[RECORD_DTYPE], stat[int32], seconds[, site_vol, site_bary])."""
    n = int(len(pair_tet))
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    stat = np.zeros(n, dtype=np.int32)
    verts = _c(mesh.vertices, np.float32)
    idx = _c(mesh.indices, np.int32)
    v_adjs = _c(mesh.v_adjs, np.int32)
    e6 = _c(mesh.e_adj6, np.int32)
    fa = _c(mesh.f_adjs, np.int32)
    fi = _c(mesh.f_ids, np.int32)
    ss = _c(sites.site_soa, np.float32)
    sw = _c(sites.weights, np.float32)
    sf = _c(sites.flags, np.uint32)
    knn = _c(site_knn, np.int32)
    pt = _c(pair_tet, np.int32)
    ps = _c(pair_site, np.int32)
    vol = np.zeros(sites.n_site, dtype=np.float32)
    bary = np.zeros(3 * sites.n_site, dtype=np.float32)
    if impl == "oracle":
        fn = lib().orc_rpd_run_pairs
    else:
        r = ref("rpd_filter" if impl == "ref_filter" else "rpd")
        if r is None:
            raise RuntimeError("oracle/_ref/libref_rpd%s.so not built" % ("_filter" if impl == "ref_filter" else ""))
        fn = r.ref_rpd_run_pairs
    sec = fn(_p(verts), _p(idx), C.c_int(mesh.n_tet), _p(v_adjs), _p(e6), _p(fa), _p(fi), _p(ss),
             _p(sw), _p(sf), C.c_int(sites.n_site), _p(knn), C.c_int(site_k), _p(pt), _p(ps),
             C.c_long(n), _p(recs), _p(stat), _p(vol), _p(bary), C.c_int(n_threads))
    if want_vol:
        return recs, stat, sec, vol, bary
    return recs, stat, sec


def flagged_pairs(mesh, sites, site_knn, site_k, pair_tet, pair_site, impl="oracle"):
    """The flagged class per candidate pair (uint8): some conflict test of the pair's clipping had
    |det| < 1.2466136531027298e-13 * maxx*maxy*maxz*max^2 (predicate_generator/main.cpp:52-78).
    impl='ref': the reference's own USE_ARITHMETIC_FILTER build (status needs_exact_predicates, convex_cell.cu:479-497);
    impl='oracle': the plain-C restatement."""
    if impl == "ref":
        _, stat, _ = run_pairs(mesh, sites, site_knn, site_k, pair_tet, pair_site, impl="ref_filter")
        return (stat == STATUS["needs_exact_predicates"]).astype(np.uint8)
    n = int(len(pair_tet))
    out = np.zeros(n, np.uint8)
    lib().orc_rpd_flagged_pairs(
        _p(_c(mesh.vertices, np.float32)), _p(_c(mesh.indices, np.int32)), C.c_int(mesh.n_tet), _p(_c(mesh.v_adjs, np.int32)),
        _p(_c(mesh.e_adj6, np.int32)), _p(_c(mesh.f_adjs, np.int32)), _p(_c(mesh.f_ids, np.int32)),
        _p(_c(sites.site_soa, np.float32)), _p(_c(sites.weights, np.float32)), _p(_c(sites.flags, np.uint32)),
        C.c_int(sites.n_site), _p(_c(site_knn, np.int32)), C.c_int(site_k), _p(_c(pair_tet, np.int32)),
        _p(_c(pair_site, np.int32)), C.c_long(n), _p(out))
    return out


def tet_sphere_relation(mesh, sites, site_knn, site_k, cap=None):
    """a3+a4 restated: candidate (tet, site) pairs sorted by (tet, site)."""
    cap = cap or 16 * mesh.n_tet
    while True:
        pt = np.empty(cap, dtype=np.int32)
        ps = np.empty(cap, dtype=np.int32)
        n = lib().orc_tet_sphere_relation(
            _p(_c(mesh.vertices, np.float32)), _p(_c(mesh.indices, np.int32)), C.c_int(mesh.n_tet),
            _p(_c(sites.site_soa, np.float32)), _p(_c(sites.weights, np.float32)),
            _p(_c(sites.flags, np.uint32)), C.c_int(sites.n_site), _p(_c(site_knn, np.int32)),
            C.c_int(site_k), _p(pt), _p(ps), C.c_long(cap))
        if n >= 0:
            return pt[:n].copy(), ps[:n].copy()
        cap = -n


def ref_rpd_gpu(mesh, sites, site_knn, site_k):
    """The reference's own CUDA build (oracle/_ref/libref_rpd_gpu.so: voronoi.cu + convex_cell.cu + knncuda.cu compiled
    in place with the reference's --use_fast_math) through its real entry point compute_clipped_voro_diagram_GPU
    (voronoi.cu:455-795).  Needs a GPU and the dense e_adjs table; memory-feasible up to config 2.  Returns
    (records of the valid cells sorted by (tet, site), {"call_ms", "kernel_ms", "d2h_ms"}) or None without the build /
    a device.  The reference prints to stdout and appends record.csv in the CWD."""
    l = ref("rpd_gpu")
    if l is None:
        return None
    e = _c(mesh.dense_e_adjs(), np.int32)
    ms = C.c_double(0)
    n = l.ref_rpd_gpu_run(
        _p(_c(mesh.vertices, np.float32)), C.c_int(mesh.n_vert), _p(_c(mesh.indices, np.int32)), C.c_int(mesh.n_tet),
        _p(_c(mesh.v_adjs, np.int32)), _p(e), C.c_long(e.size), _p(_c(mesh.f_adjs, np.int32)), _p(_c(mesh.f_ids, np.int32)),
        _p(_c(sites.site_soa, np.float32)), _p(_c(sites.weights, np.float32)), _p(_c(sites.flags, np.uint32)),
        C.c_int(sites.n_site), _p(_c(site_knn, np.int32)), C.c_int(site_k), C.byref(ms))
    if n < 0:
        return None
    recs = np.zeros(n, dtype=RECORD_DTYPE)
    l.ref_rpd_gpu_fetch(_p(recs))
    k_ms, d_ms = C.c_double(-1), C.c_double(-1)
    l.ref_rpd_gpu_last_ms(C.byref(k_ms), C.byref(d_ms))
    return recs, {"call_ms": ms.value, "kernel_ms": k_ms.value, "d2h_ms": d_ms.value}


def ref_d2m_gpu(inp, kernel_only=False, warmup=1, reps=3):
    """The reference's dist2mat on the GPU: its whole entry point compute_closest_dist2mat (7 H2D + kernel + 2 D2H,
    dist2mat.cu:280-315) or, kernel_only, its kernel ClosestDistanceToLocalMat alone on resident buffers.
    Returns (result, closest_id, milliseconds) or None without the build / a device."""
    l = ref("d2m")
    if l is None:
        return None
    n = len(inp.samples)
    res = np.zeros(n, np.float32)
    cid = np.zeros(n, np.int32)
    sph = _c(inp.spheres, np.float32)
    pr = _c(inp.prims, np.int32)
    args = (_p(sph), C.c_int(len(sph)), _p(_c(inp.samples, np.float32)), C.c_int(n), _p(_c(inp.offset, np.uint32)),
            _p(_c(inp.count, np.uint32)), _p(pr), C.c_long(len(pr)), _p(res), _p(cid))
    ms = l.ref_d2m_kernel_ms(*args, C.c_int(warmup), C.c_int(reps)) if kernel_only else l.ref_d2m_run_gpu(*args)
    return None if ms < 0 else (res, cid, ms)


def reload_active(recs, impl="oracle"):
    n = len(recs)
    recs = np.ascontiguousarray(recs)
    ap = np.zeros((n, 64), dtype=np.uint8)
    ae = np.zeros((n, 152), dtype=np.uint8)
    eu = np.zeros(n, dtype=np.float32)
    if impl == "oracle":
        lib().orc_reload_active(_p(recs), C.c_long(n), _p(ap), _p(ae), _p(eu))
    else:
        ref("host").ref_host_reload_active(_p(recs), C.c_long(n), _p(ap), _p(ae), _p(eu))
    return ap, ae, eu


def canonicalize(recs):
    recs = np.ascontiguousarray(recs)
    out = np.zeros(len(recs), dtype=RECORD_DTYPE)
    lib().orc_canonicalize(_p(recs), C.c_long(len(recs)), _p(out))
    return out


def vertex_coordinates(recs, impl="oracle"):
    recs = np.ascontiguousarray(recs)
    out = np.zeros((len(recs), 96, 4), dtype=np.float32)
    if impl == "oracle":
        lib().orc_vertex_coordinates(_p(recs), C.c_long(len(recs)), _p(out))
    else:
        ref("host").ref_host_vertex_coordinates(_p(recs), C.c_long(len(recs)), _p(out))
    return out


def cell_volumes(recs):
    recs = np.ascontiguousarray(recs)
    out = np.zeros(len(recs), dtype=np.float64)
    lib().orc_cell_volumes(_p(recs), C.c_long(len(recs)), _p(out))
    return out


def defined_equal(a, b):
    """Compare two record arrays on the DEFINED entries only (entries < nb_v/nb_p/nb_e; the
    reference leaves the rest uninitialised).  Returns a dict of mismatch counts."""
    assert len(a) == len(b)
    out = {}
    for f in ("status", "voro_id", "tet_id"):
        out[f] = int((a[f] != b[f]).sum())
    ok = (a["status"] == 4) & (b["status"] == 4)
    for f in ("nb_v", "nb_p", "nb_e", "weight"):
        out[f] = int((a[f][ok] != b[f][ok]).sum())
    iv = np.arange(96)[None, :] < a["nb_v"][:, None]
    ip = np.arange(64)[None, :] < a["nb_p"][:, None]
    ie = np.arange(152)[None, :] < a["nb_e"][:, None]
    same_n = ok & (a["nb_v"] == b["nb_v"]) & (a["nb_p"] == b["nb_p"]) & (a["nb_e"] == b["nb_e"])
    iv &= same_n[:, None]
    ip &= same_n[:, None]
    ie &= same_n[:, None]
    out["ver"] = int(((a["ver"] != b["ver"]).any(axis=2) & iv).sum())
    # planes: bitwise compare of the 5 floats
    ca = a["clip"][:, :, :5].view(np.uint32)
    cb = b["clip"][:, :, :5].view(np.uint32)
    out["clip"] = int(((ca != cb).any(axis=2) & ip).sum())
    out["id2"] = int(((a["id2"] != b["id2"]).any(axis=2) & ip).sum())
    out["edge"] = int(((a["edge"] != b["edge"]).any(axis=2) & ie).sum())
    out["cells_compared"] = int(same_n.sum())
    return out


def dist2mat(inp, impl="oracle", n_threads=0, n=None, want_second=False):
    n = len(inp.samples) if n is None else n
    res = np.full(n, 1e28, dtype=np.float32)
    cid = np.full(n, -1, dtype=np.int32)
    sec2 = np.zeros(n, dtype=np.float32)
    sph = _c(inp.spheres, np.float32)
    smp = _c(inp.samples[:n], np.float32)
    off = _c(inp.offset[:n], np.uint32)
    cnt = _c(inp.count[:n], np.uint32)
    pr = _c(inp.prims, np.int32)
    if impl == "oracle":
        t = lib().orc_dist2mat(_p(sph), _p(smp), C.c_long(n), _p(off), _p(cnt), _p(pr), _p(res),
                               _p(cid), _p(sec2), C.c_int(n_threads))
    else:
        t = ref("d2m").ref_d2m_run_host(_p(sph), _p(smp), C.c_long(n), _p(off), _p(cnt), _p(pr),
                                        _p(res), _p(cid), C.c_int(n_threads))
    if want_second:
        return res, cid, t, sec2
    return res, cid, t


def zero_undefined(recs):
    """Zero every entry the reference leaves uninitialised (entries >= nb_v/nb_p/nb_e, padding,
    whole records that are not `success`), so that records can be stored and compared bytewise."""
    out = np.zeros(len(recs), dtype=RECORD_DTYPE)
    for f in ("status", "voro_id", "tet_id"):
        out[f] = recs[f]
    ok = recs["status"] == 4
    for f in ("weight", "nb_v", "nb_p", "nb_e", "is_active"):
        out[f][ok] = recs[f][ok]
    iv = (np.arange(96)[None, :] < recs["nb_v"][:, None]) & ok[:, None]
    ip = (np.arange(64)[None, :] < recs["nb_p"][:, None]) & ok[:, None]
    ie = (np.arange(152)[None, :] < recs["nb_e"][:, None]) & ok[:, None]
    out["ver"][iv] = recs["ver"][iv]
    clip = np.zeros_like(recs["clip"])
    clip[..., :5] = recs["clip"][..., :5]
    out["clip"][ip] = clip[ip]
    out["id2"][ip] = recs["id2"][ip]
    out["edge"][ie] = recs["edge"][ie]
    return out


def emit(recs, max_surf_fid):
    """get_all_voro_info's per-cell emission (restated, rpd_update.cxx:112-301) for success records."""
    recs = np.ascontiguousarray(recs)
    n = len(recs)
    cap = max(16, 40 * n)
    out = {
        "facet_cell": np.zeros(cap, np.int32), "facet_key": np.zeros(cap, np.int32),
        "facet_is_tet": np.zeros(cap, np.uint8), "facet_centroid": np.zeros((cap, 3), np.float32),
        "vert_cell": np.zeros(cap, np.int32), "vert_lvid": np.zeros(cap, np.int32),
        "vert_key": np.zeros((cap, 3), np.int32), "vert_pos": np.zeros((cap, 3), np.float32),
        "vert_surf_fid": np.zeros(cap, np.int32),
        "edge_cell": np.zeros(cap, np.int32), "edge_key": np.zeros((cap, 2), np.int32),
        "edge_lvid": np.zeros((cap, 2), np.int32),
    }
    cnt = np.zeros(3, np.int64)
    order = ["facet_cell", "facet_key", "facet_is_tet", "facet_centroid", "vert_cell", "vert_lvid", "vert_key",
             "vert_pos", "vert_surf_fid", "edge_cell", "edge_key", "edge_lvid"]
    rc = lib().orc_emit(_p(recs), C.c_long(n), C.c_int(int(max_surf_fid)), C.c_long(cap), _p(cnt),
                        *[_p(out[k]) for k in order])
    assert rc == 0
    for k in order:
        m = cnt[0] if k.startswith("facet") else (cnt[1] if k.startswith("vert") else cnt[2])
        out[k] = out[k][:m].copy()
    return out


def feature_edges(recs, fe_map):
    """TEST INFRASTRUCTURE.  Restates the feature-edge part of get_all_voro_info (rpd_update.cxx:209-259) per cell:
    an ACTIVE edge whose two planes are tet faces and whose (tet, lf_min, lf_max) is in TetMesh::tet_es2fe_map is a
    covered sharp (SE = 1) or concave (CE = 2) edge; its end vertices are the two vertices lying on both planes
    (convert_e_lfids_to_lvids :44-71); every end vertex that also lies on a half-plane records the sharp line's end
    position under (cell, lvid, neighbour of its FIRST half-plane, fe_line_id) (:237-252).
    Returns rows (site, kind, cell, lv1, lv2, fe_line_id, fe_id), end_rows (site, cell, lvid, neigh, line), end_pos."""
    fe = {(int(r[0]), int(r[1]), int(r[2])): (int(r[3]), int(r[4]), int(r[5])) for r in np.asarray(fe_map).reshape(-1, 6)}
    _, ae, _ = reload_active(recs, "oracle")
    pos = vertex_coordinates(recs)
    rows, end = set(), {}
    tets_with_fe = {k[0] for k in fe}
    for c in np.flatnonzero(np.isin(recs["tet_id"], list(tets_with_fe))):
        r = recs[c]
        site, tet = int(r["voro_id"]), int(r["tet_id"])
        for e in range(int(r["nb_e"])):
            if not ae[c, e]:
                continue
            a, b = int(r["edge"][e][0]), int(r["edge"][e][1])
            if r["id2"][a][1] != -1 or r["id2"][b][1] != -1:
                continue
            lo, hi = min(a, b), max(a, b)
            if (tet, lo, hi) not in fe:
                continue
            kind, fe_id, line = fe[(tet, lo, hi)]
            lvs = sorted(v for v in range(int(r["nb_v"])) if lo in r["ver"][v][:3] and hi in r["ver"][v][:3])
            assert len(lvs) == 2
            rows.add((site, kind, int(c), lvs[0], lvs[1], line, fe_id))
            for lv in lvs:
                for i in range(3):
                    hp = r["id2"][int(r["ver"][lv][i])]
                    if hp[1] == -1:
                        continue
                    neigh = int(hp[1]) if int(hp[0]) == site else int(hp[0])
                    end[(site, int(c), lv, neigh, line)] = pos[c, lv, :3].copy()
                    break
    keys = sorted(end)
    return {"rows": np.array(sorted(rows), np.int32).reshape(-1, 7), "end_rows": np.array(keys, np.int32).reshape(-1, 5),
            "end_pos": np.array([end[k] for k in keys], np.float32).reshape(-1, 3)}


def topology(em, cell_site, cell_euler):
    """TEST INFRASTRUCTURE.  Restates the remainder of update_power_cells per power cell, literally with
    dict / set / BFS like the reference:
      * tfid_to_cells -> cell_neighbors: the first and last cell of a tet-face id's set become neighbours
        (rpd_update.cxx:303-316);
      * get_CC_given_neighbors (include/common_cxx.h:447-489): BFS over cell_neighbors restricted to the
        visited set -- on all cells of the power cell (update_pc_cc_info, rpd_update.cxx:497-503) and on the
        cells of every half-plane facet_neigh_to_cells[neigh] (update_pc_facet_cc_info, :439-470);
      * msphere.euler_sum = double sum of ConvexCellHost::euler over cell_ids in ascending order
        (is_to_fix_voro_euler, fix_topo.cxx:117-131).
    `em` is the facet emission of emit() (facet_cell, facet_key, facet_is_tet), cell_site the voro_id of every
    cell, cell_euler the float per-cell Euler values.  Labels: smallest member of the component."""
    from collections import defaultdict, deque

    f_cell, f_key, f_tet = em["facet_cell"], em["facet_key"], em["facet_is_tet"]
    n_cells = len(cell_site)
    tfid_to_cells = defaultdict(list)          # (site, tfid) -> cells (ascending: facets come cell by cell)
    facet_neigh_to_cells = defaultdict(dict)   # (site, neigh) -> {cell: facet index}
    for f in range(len(f_cell)):
        c = int(f_cell[f])
        s = int(cell_site[c])
        if f_tet[f]:
            tfid_to_cells[(s, int(f_key[f]))].append(c)
        else:
            facet_neigh_to_cells[(s, int(f_key[f]))].setdefault(c, f)
    cell_neighbors = defaultdict(set)
    for cells in tfid_to_cells.values():
        if len(cells) == 1:
            continue
        c1, c2 = cells[0], cells[-1]
        if c1 != c2:
            cell_neighbors[c1].add(c2)
            cell_neighbors[c2].add(c1)

    def cc_given_neighbors(to_visit):
        unvisited = set(to_visit)
        out = []
        for start in sorted(to_visit):
            if start not in unvisited:
                continue
            comp, q = set(), deque([start])
            while q:
                c = q.popleft()
                if c in comp:
                    continue
                comp.add(c)
                unvisited.discard(c)
                for nb in cell_neighbors.get(c, ()):
                    if nb in to_visit and nb not in comp:
                        q.append(nb)
            out.append(comp)
        return out

    cells_of_site = defaultdict(list)
    for c in range(n_cells):
        cells_of_site[int(cell_site[c])].append(c)
    cell_cc = np.full(n_cells, -1, np.int64)
    site_stats = {}
    for s, cells in cells_of_site.items():
        comps = cc_given_neighbors(set(cells))
        for comp in comps:
            m = min(comp)
            for c in comp:
                cell_cc[c] = m
        esum = 0.0
        for c in cells:  # std::set<int> order
            esum += float(np.float32(cell_euler[c]))
        site_stats[s] = (len(cells), len(comps), esum)
    facet_cc = np.full(len(f_cell), -1, np.int64)
    pairs = {}
    for (s, n), c2f in facet_neigh_to_cells.items():
        comps = cc_given_neighbors(set(c2f.keys()))
        pairs[(s, n)] = len(comps)
        for comp in comps:
            m = min(c2f[c] for c in comp)
            for c in comp:
                facet_cc[c2f[c]] = m
    # e_to_cells -> edge_cc_cells (update_pc_edge_cc_info, rpd_update.cxx:507-521): per (site, neigh_min, neigh_max)
    edge_cc = None
    if "edge_cell" in em:
        e_cell, e_key = em["edge_cell"], em["edge_key"]
        e_to_cells = defaultdict(dict)
        for e in range(len(e_cell)):
            c = int(e_cell[e])
            e_to_cells[(int(cell_site[c]), int(e_key[e][0]), int(e_key[e][1]))].setdefault(c, e)
        edge_cc = np.full(len(e_cell), -1, np.int64)
        for key, c2e in e_to_cells.items():
            for comp in cc_given_neighbors(set(c2e.keys())):
                m = min(c2e[c] for c in comp)
                for c in comp:
                    edge_cc[c2e[c]] = m
    return {"cell_cc": cell_cc, "facet_cc": facet_cc, "edge_cc": edge_cc, "site_stats": site_stats, "pairs": pairs}


_UPDATE_WIDTH = {"facets": (3, 0), "tfids": (3, 0), "surf": (3, 3), "vertices": (7, 3), "edges": (6, 0), "e2cells": (4, 0),
                 "neighbours": (3, 0), "cc": (3, 0), "facet_cc": (4, 0), "edge_cc": (5, 0), "fe": (7, 0), "fe_end": (5, 3)}


def ref_update(recs, n_site, max_surf_fid, fe_map=None):
    """TEST INFRASTRUCTURE.  The reference's own update_power_cells (src/rpd3d_base/rpd_update.cxx:568-639 ->
    get_all_voro_info :80-340, update_pc_cc_info, update_pc_facet_cc_info, update_pc_edge_cc_info) compiled in place
    (oracle/_ref/libref_update.so) on success records with id = index.  fe_map: int rows (tet, lf_min, lf_max, fe_type,
    fe_id, fe_line_id) = TetMesh::tet_es2fe_map.  Returns a dict of the flattened PowerCell containers (row layouts in
    oracle/ref_shim_update.cpp) + "cell_euler"; integer rows as int arrays [n, width], positions as float64 [n, 3]."""
    l = ref("update")
    assert l is not None, "oracle/_ref/libref_update.so not built"
    l.ref_update_get.restype = C.c_long
    recs = np.ascontiguousarray(recs)
    fe = np.zeros((0, 6), np.int32) if fe_map is None else np.ascontiguousarray(fe_map, dtype=np.int32).reshape(-1, 6)
    rc = l.ref_update_run(_p(recs), C.c_long(len(recs)), C.c_int(int(n_site)), C.c_int(int(max_surf_fid)), _p(fe), C.c_long(len(fe)))
    assert rc == 0
    out = {}
    for what, (w, vw) in _UPDATE_WIDTH.items():
        n = l.ref_update_get(what.encode(), None, None, C.c_long(0))
        rows = np.zeros((n, w), np.int32)
        vals = np.zeros((n, 3), np.float64) if vw else None
        if n:
            l.ref_update_get(what.encode(), _p(rows), _p(vals) if vw else None, C.c_long(n))
        out[what] = rows
        if vw:
            out[what + "_pos"] = vals
    eu = np.zeros(len(recs), np.float32)
    l.ref_update_cell_euler(_p(eu))
    out["cell_euler"] = eu
    return out


def ref_bgeo(recs, max_sf_fid, is_boundary_only, work_dir, name="t"):
    """TEST INFRASTRUCTURE.  The reference's own save_convex_cells_houdini (io_cuda.cxx:152-187, compiled in place in
    oracle/_ref/libref_bgeo.so) on ConvexCellTransfer records; returns the bytes of the .bgeo it wrote under
    <work_dir>/../out/<name>/rpd/."""
    import glob
    l = ref("bgeo")
    assert l is not None, "oracle/_ref/libref_bgeo.so not built"
    recs = np.ascontiguousarray(recs)
    os.makedirs(work_dir, exist_ok=True)
    out_dir = os.path.normpath(os.path.join(work_dir, "..", "out", name, "rpd"))
    for old in glob.glob(os.path.join(out_dir, "*.bgeo")):
        os.remove(old)
    rc = l.ref_bgeo_write(_p(recs), C.c_long(len(recs)), C.c_int(int(max_sf_fid)), C.c_int(int(is_boundary_only)),
                          work_dir.encode(), name.encode())
    assert rc == 0, rc
    files = glob.glob(os.path.join(out_dir, "*.bgeo"))
    assert len(files) == 1, files
    with open(files[0], "rb") as fh:
        return fh.read()
