"""MATTopo-style iteration driver for the RPD path (BASELINE.json config 5): the tet mesh stays resident in
HBM, every iteration inserts / updates a few medial spheres (what fix_topo / fix_geo do between two
RPD3D_GPU::calculate_partial calls, reference src/rpd3d_api/rpd_api.cxx:147-313) and recomputes the
restricted power diagram.  The reference recomputes only the tets of the affected spheres because its
full recompute is slow; here a full recompute in grid-kNN mode costs a few milliseconds, so the
default is the full, exact recompute (no CGAL ring selection needed); callers that do know the
affected tets (load_partial_tet_given_spheres, rpd_api.cxx:482-535) can pass them as `tet_subset`.
"""
from __future__ import annotations

import time

import numpy as np

from . import synth


def evolve_sites(sites: synth.Sites, iteration: int, frac_insert=0.005, frac_update=0.005, R=500.0,
                 seed: int = synth.RAN_SEED):
    """Deterministic edit of the sphere set: insert frac_insert new spheres (same
    distribution), perturb frac_update existing ones (centre +-1 % R, radius +-5 %).
    Returns (new Sites, changed ids)."""
    n = sites.n_site
    n_ins = max(1, int(round(n * frac_insert)))
    n_upd = max(1, int(round(n * frac_update)))
    c = sites.centers().astype(np.float64)
    r = sites.radii.astype(np.float64)
    u = synth.uniform01(seed, 5 * n_upd, stream=1000 + iteration).reshape(n_upd, 5)
    idx = np.unique(np.minimum((u[:, 0] * n).astype(np.int64), n - 1))
    k = idx.size
    c[idx] += (2.0 * u[:k, 1:4] - 1.0) * (0.01 * R)
    r[idx] *= 1.0 + 0.05 * (2.0 * u[:k, 4] - 1.0)
    fresh = synth.make_spheres(n_ins, seed=seed, stream=2000 + iteration, R=R)
    spacing = (4.0 / 3.0 * np.pi * (0.9 * R) ** 3 / (n + n_ins)) ** (1.0 / 3.0)
    fr = np.minimum(fresh.radii.astype(np.float64), spacing)  # radii of the new spheres at the current density
    c = np.concatenate([c, fresh.centers().astype(np.float64)])
    r = np.concatenate([r, fr])
    # keep every sphere inside the domain
    d = np.linalg.norm(c - 500.0, axis=1)
    r = np.minimum(r, np.maximum(R - d, 1e-3))
    c32 = c.astype(np.float32)
    r32 = r.astype(np.float32)
    out = synth.Sites(np.ascontiguousarray(c32.T).ravel(), (r32 * r32).astype(np.float32),
                      np.ones(c32.shape[0], dtype=np.uint32), r32)
    changed = np.concatenate([idx, np.arange(n, n + n_ins)])
    return out, changed


class RpdLoop:
    """resident-mesh iteration loop; `step` returns (RpdResult, seconds end to end)"""

    def __init__(self, ctx, mesh):
        self.ctx = ctx
        self.mesh = mesh
        ctx.set_mesh(mesh)
        self.last = None

    def step(self, sites: synth.Sites, tet_subset=None, fetch=None, to_host=False, **opts):
        """upload the sites (H2D), recompute (all tets, or `tet_subset`), and deliver the result to the host:
        to_host=True streams it (mb_rpd_run_to_host: the D2H of tet span c overlaps the kernels of span c+1; the
        compact result is in the library's pinned memory on return, RpdResult.host_compact()); otherwise the
        result stays on the device and `fetch` = (blob_u32, offsets_i64) optionally copies it afterwards"""
        ctx = self.ctx
        t0 = time.perf_counter()
        ctx.set_tet_subset(tet_subset)
        ctx.upload_sites(sites.site_soa, sites.weights, sites.flags)
        res = ctx.run_to_host(**opts) if to_host else ctx.run(**opts)
        if fetch is not None and not to_host:
            ctx._check(ctx.lib.mb_rpd_fetch_compact(res._h, fetch[0].ctypes.data, fetch[1].ctypes.data))
        dt = time.perf_counter() - t0
        if self.last is not None:
            self.last.free()
        self.last = res
        return res, dt

    def step_incremental(self, sites: synth.Sites, to_host=False, **opts):
        """upload the sites and clip only the AFFECTED tets (mb_rpd_run_incremental: the tets whose candidate list
        changed since the previous incremental step -- exact, no neighbour rings, no CGAL).  Returns
        (RpdResult of the affected tets, ascending affected tet ids, seconds end to end); merge the patch into the
        previous result with capi.merge_compact (the counterpart of merge_convex_cells, rpd_api.cxx:432-479)."""
        ctx = self.ctx
        t0 = time.perf_counter()
        ctx.upload_sites(sites.site_soa, sites.weights, sites.flags)
        res, tets = ctx.run_incremental(to_host=to_host, **opts)
        dt = time.perf_counter() - t0
        if self.last is not None:
            self.last.free()
        self.last = res
        return res, tets, dt


def site_rings(pair_site, pair_neigh, seeds, n_site: int, depth: int = 2):
    """1-ring / 2-ring of `seeds` in the restricted power diagram, from the half-plane pairs K6 emits
    (mb_rpd_fetch_topology: pair_site / pair_neigh = every (site, neighbour) sharing a bisector facet inside the mesh)
    -- the neighbour source that replaces get_RT_partial_spheres_and_neighbors (triangulation.cxx:442-549) for callers
    that want the reference's N + 1-ring + 2-ring site selection (rpd_api.cxx:54-63).  Returns a list of id arrays,
    ring 0 = the seeds."""
    adj = [[] for _ in range(n_site)]
    for s, n in zip(np.asarray(pair_site).tolist(), np.asarray(pair_neigh).tolist()):
        if 0 <= n < n_site:
            adj[s].append(n)
    seen = set(int(s) for s in seeds)
    rings = [np.array(sorted(seen), np.int64)]
    frontier = set(seen)
    for _ in range(depth):
        nxt = set()
        for s in frontier:
            for n in adj[s]:
                if n not in seen:
                    nxt.add(n)
        seen |= nxt
        rings.append(np.array(sorted(nxt), np.int64))
        frontier = nxt
    return rings
