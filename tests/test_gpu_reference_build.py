"""The reference's OWN CUDA build on the B200 (oracle/_ref/libref_rpd_gpu.so, libref_d2m.so: the reference sources
compiled in place with its own flags, --use_fast_math and FMA contraction included) as a parity target next to its host
build and this library -- what LibMAT users actually run (SURVEY 2.1 / 8c "host-shim == device build on config 1")."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_reference_cuda_build_vs_host_build_vs_library_cfg1(ctx, O, cfg1, cfg1_oracle, capfd):
    if O.ref("rpd_gpu") is None or O.ref("rpd") is None:
        pytest.skip("oracle/_ref not built")
    mesh, sites, knn, k = cfg1
    pt, ps, _, _ = cfg1_oracle
    out = O.ref_rpd_gpu(mesh, sites, knn, k)
    capfd.readouterr()  # the reference prints its launch shape
    assert out is not None, "the reference CUDA build did not run"
    dev, ms = out
    assert ms["kernel_ms"] > 0 and ms["d2h_ms"] > 0
    host, _, _ = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="ref")
    host = host[host["status"] == 4]
    ctx.set_mesh(mesh)
    mine = ctx.compute_clipped_voro_diagram(sites.site_soa, sites.weights, sites.flags, knn, k).records()
    # the device build (FMA-contracted, fast-math division) returns the same cells in the same order ...
    assert len(dev) == len(host) == len(mine)
    for f in ("voro_id", "tet_id"):
        assert np.array_equal(dev[f], host[f]) and np.array_equal(dev[f], mine[f]), f
    assert np.array_equal(dev["id"], mine["id"]) and np.array_equal(dev["id"], np.arange(len(dev)))  # voronoi.cu:766-768
    # ... with the same combinatorics: any difference would have to be a flagged cell, and config 1 has none
    fr = O.flagged_pairs(mesh, sites, knn, k, pt, ps, "oracle")
    assert fr.sum() == 0
    d = O.defined_equal(host, dev)
    for f in ("status", "nb_v", "nb_p", "nb_e", "ver", "id2", "edge"):
        assert d[f] == 0, d
    assert d["cells_compared"] == len(host)
    # plane equations: FMA contraction / fast-math division move the last bits (relative to the plane's own scale: a
    # cancelling normal component of a sliver tet can lose all its relative accuracy)
    ip = np.arange(64)[None, :] < host["nb_p"][:, None]
    a, b = host["clip"][..., :4][ip].astype(np.float64), dev["clip"][..., :4][ip].astype(np.float64)
    nscale = np.abs(a[:, :3]).max(axis=1)
    assert np.max(np.abs(a[:, :3] - b[:, :3]).max(axis=1) / nscale) < 1e-5
    assert np.max(np.abs(a[:, 3] - b[:, 3]) / np.maximum(np.abs(a[:, 3]), 1000.0 * nscale)) < 1e-5
    print(f"reference device vs host build: {100.0 * np.mean((a == b).all(axis=1)):.2f} % of the planes bit-identical")
    # the library is byte-identical to the HOST build (checked elsewhere) -- and therefore combinatorially to this one
    d2 = O.defined_equal(dev, mine)
    for f in ("status", "nb_v", "nb_p", "nb_e", "ver", "id2", "edge"):
        assert d2[f] == 0, d2
    print(f"reference CUDA build, config 1: kernel {ms['kernel_ms']:.2f} ms, D2H {ms['d2h_ms']:.2f} ms, call {ms['call_ms']:.0f} ms")


def test_reference_dist2mat_device_vs_host_per_primitive(O, synth):
    """The reference's distance functions, device build against host build, primitive by primitive on 60 000 samples
    (~1.3 M primitive evaluations).  Pins the one semantic difference between the two compiles of the same source:
    clamp(NaN, 0, 1) is a saturate (-> 0) on the device, 1 on the host (nested spheres, dist2mat.cu:61-64); with the
    host functions following the device (oracle/ref_shim_d2m.cu) everything else agrees to rounding."""
    import ctypes as C
    l = O.ref("d2m")
    if l is None:
        pytest.skip("oracle/_ref not built")
    d = synth.make_dist2mat(60000)
    n = len(d.samples)
    cnt = d.count.astype(np.int64)
    rows = np.repeat(np.arange(n), cnt)
    pos = np.ascontiguousarray(d.samples[rows], dtype=np.float32)
    pr = np.ascontiguousarray(d.prims[: int(cnt.sum())], dtype=np.int32)  # lists are laid out back to back
    assert int(d.offset[-1]) + int(cnt[-1]) == len(pr)
    sph = np.ascontiguousarray(d.spheres, dtype=np.float32)
    dev = np.zeros(len(pr), np.float32)
    assert l.ref_d2m_eval_prims_gpu(O._p(sph), C.c_int(len(sph)), O._p(pos), O._p(pr), C.c_int(len(pr)), O._p(dev)) == 0
    # host build: one single-primitive list per evaluation
    one = synth.Dist2MatInput(d.spheres, pos, np.arange(len(pr), dtype=np.uint32), np.ones(len(pr), np.uint32), pr, d.n_cones, d.n_slabs)
    host, _, _ = O.dist2mat(one, "ref")
    orc, _, _ = O.dist2mat(one, "oracle")
    assert np.array_equal(host.view(np.uint32), orc.view(np.uint32))  # plain-C port == host build, bit for bit
    rel = np.abs(dev - host) / np.maximum(np.abs(host), 0.1)  # relative to the primitive scale (test_gpu_dist2mat.rel_err)
    kind = np.where(pr[:, 0] != -1, 2, np.where(pr[:, 1] != -1, 1, 0))
    nested = np.zeros(len(pr), bool)
    a, b = sph[np.maximum(pr[:, 1], 0)], sph[pr[:, 2]]
    nested[kind == 1] = (((a[:, :3] - b[:, :3]) ** 2).sum(axis=1) < (a[:, 3] - b[:, 3]) ** 2)[kind == 1]
    assert nested.sum() > 1000  # the NaN branch is really exercised
    print(f"device vs host build, {len(pr)} primitive evaluations: bit-identical {100 * np.mean(dev.view(np.uint32) == host.view(np.uint32)):.2f} %, "
          f"max rel spheres {rel[kind == 0].max():.2e} cones {rel[kind == 1].max():.2e} (nested cones {rel[nested].max():.2e}) "
          f"slabs {rel[kind == 2].max():.2e}; slabs beyond 1e-6: {int((rel[kind == 2] > 1e-6).sum())}")
    assert rel[kind == 0].max() <= 1e-6 and rel[kind == 1].max() <= 1e-6
    # the slab solve cancels catastrophically (W1..W3, dist2mat.cu:150-160): FMA contraction moves roots and flips the
    # discriminant's sign on a few per cent of the slabs -- the reason K5 is built with the device build's arithmetic
    assert np.mean(rel[kind == 2] > 1e-6) < 0.05


def test_reference_dist2mat_kernel_vs_library(ctx, O, synth):
    """the reference's own CUDA kernel (its whole entry point and the kernel alone) against the library, 200 000
    samples: bit-identical on >= 99.9 %, argmin ids equal apart from flagged ties"""
    from test_gpu_dist2mat import rel_err
    if O.ref("d2m") is None:
        pytest.skip("oracle/_ref not built")
    d = synth.make_dist2mat(200000)
    out = O.ref_d2m_gpu(d)
    assert out is not None, "the reference dist2mat CUDA build did not run"
    rr, rc, _ = out
    ko = O.ref_d2m_gpu(d, kernel_only=True, warmup=1, reps=2)
    assert ko is not None and np.array_equal(ko[0].view(np.uint32), rr.view(np.uint32)) and np.array_equal(ko[1], rc)
    r, cid, tie = ctx.compute_closest_dist2mat(d.spheres, d.samples, d.offset, d.count, d.prims)
    off = rel_err(r, rr) > 1e-6
    bitwise = np.mean(r.view(np.uint32) == rr.view(np.uint32))
    print(f"library vs the reference's CUDA kernel, {len(r)} samples: bit-identical {100 * bitwise:.3f} %, "
          f"{int(off.sum())} beyond 1e-6, argmin ids differing without a flagged tie: {int(((cid != rc) & (tie == 0) & ~off).sum())}")
    assert bitwise >= 0.999 and off.mean() <= 2e-4
    assert not ((cid != rc) & (tie == 0) & ~off).any()
