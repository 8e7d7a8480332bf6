"""GPU parity tests of dist2mat through the C ABI."""
import numpy as np
import pytest

from conftest import golden

pytestmark = pytest.mark.gpu


def test_kat2_golden(ctx):
    """each KAT as a 1-primitive list"""
    cases = golden("kat2_dist2mat.json")
    for c in cases:
        sph = np.asarray(c["prims"], np.float32)
        n = len(sph)
        prim = {1: [-1, -1, 0], 2: [-1, 0, 1], 3: [0, 1, 2]}[n]
        r, cid, _ = ctx.compute_closest_dist2mat(sph, np.asarray([c["pos"]], np.float32), np.zeros(1, np.uint32),
                                                 np.ones(1, np.uint32), np.asarray([prim], np.int32))
        want = np.uint32(c["bits"]).view(np.float32)
        assert cid[0] == 0
        assert r[0] == want or abs(float(r[0]) - float(want)) <= 1e-6 * max(0.1, abs(float(want))), c


SCALE = 0.1  # primitive scale of the synthetic medial mesh (sphere diameters in the unit box)


def rel_err(a, b):
    """north_star tolerance metric: relative to the larger of the distance itself and the primitive scale -- a
    signed distance |p - c| - r cancels near the surface, where 1e-6 of the distance alone would be far below one
    ulp of the operands"""
    return np.abs(a - b) / np.maximum(np.abs(b), SCALE)


def check_against_builds(O, d, r, cid, tie, min_bitwise=0.999):
    """K5 evaluates the reference's expressions with the arithmetic of the reference's DEVICE build (nvcc's FMA
    contraction, saturating clamp): the pin is the reference's own CUDA kernel run here on the same input.
    The plain-C oracle / the reference's host compile use non-fused IEEE arithmetic; the slab solve
    (dist2mat.cu:150-170) cancels catastrophically, so on a few per cent of the samples the reference's own two builds
    disagree by far more than any tolerance.  Those samples are the flagged 'unstable' class; everywhere else all
    four (library, device build, host build, oracle) agree within 1e-6."""
    ro, co, _ = O.dist2mat(d, "oracle")
    dev = O.ref_d2m_gpu(d) if O.ref("d2m") is not None else None
    if dev is None:  # no reference build on this box: the oracle alone, unstable slabs bounded by their frequency
        assert np.mean(rel_err(r, ro) > 1e-6) < 0.05
        return
    rr, cr, _ = dev
    bitwise = np.mean(r.view(np.uint32) == rr.view(np.uint32))
    off = rel_err(r, rr) > 1e-6
    assert bitwise >= min_bitwise, bitwise
    assert off.mean() <= 2e-4, int(off.sum())
    assert not np.any((cid != cr) & (tie == 0) & ~off)
    stable = rel_err(ro, rr) <= 1e-6  # the reference's two arithmetics agree
    assert stable.mean() > 0.9
    assert not np.any(stable & ~off & (rel_err(r, ro) > 2e-6))
    assert not np.any(stable & ~off & (cid != co) & (tie == 0))
    return {"bitwise_vs_device_build": float(bitwise), "beyond_1e-6": int(off.sum()), "unstable": int((~stable).sum()), "n": len(r)}


def test_mini_golden(ctx, O, synth):
    """the committed fixture (generated from the reference's host compile by oracle/gen_golden.py) on its stable samples,
    and the reference's CUDA kernel on all of them"""
    g = golden("mini_dist2mat.npz")
    d = synth.make_dist2mat(2000, nu=20, nv=40, n_slabs=2400, n_cones=1200)
    r, cid, tie = ctx.compute_closest_dist2mat(d.spheres, d.samples, d.offset, d.count, d.prims)
    ro, co, _ = O.dist2mat(d, "oracle")
    assert rel_err(ro, g["result"]).max() <= 1e-6  # the oracle against the fixture (the reference's host compile)
    info = check_against_builds(O, d, r, cid, tie, min_bitwise=0.98)
    print("mini fixture:", info)


def test_vs_reference_cuda_kernel_and_oracle_200k(ctx, O, synth):
    d = synth.make_dist2mat(200000)
    r, cid, tie = ctx.compute_closest_dist2mat(d.spheres, d.samples, d.offset, d.count, d.prims)
    info = check_against_builds(O, d, r, cid, tie)
    print("200k samples:", info)


def test_empty_and_ragged_lists(ctx, synth):
    """empty list -> (1e16f, -1) (dist2mat.cu:224-231); lists longer than one warp stride"""
    sph = np.array([[0, 0, 0, 1], [3, 0, 0, 0.5], [0, 4, 0, 0.25]], np.float32)
    samples = np.array([[0, 0, 2], [3, 0, 1], [9, 9, 9]], np.float32)
    prims = np.array([[-1, -1, 0]] + [[-1, -1, 1]] * 40 + [[-1, -1, 2]], np.int32)
    offset = np.array([0, 1, 0], np.uint32)
    count = np.array([1, 41, 0], np.uint32)
    r, cid, tie = ctx.compute_closest_dist2mat(sph, samples, offset, count, prims)
    assert r[0] == 1.0 and cid[0] == 0
    # 40 identical spheres then a farther one: winner = highest lane among exact ties (dist2mat.cu:269-276)
    assert r[1] == 0.5 and cid[1] == 31 and tie[1] == 1
    assert r[2] == np.float32(1e16) and cid[2] == -1


def test_nan_clamp_follows_the_device_build(ctx):
    """coincident / nested spheres: t = NaN -> the reference's device build saturates it to 0 -> distance to the LARGER
    sphere (its host compile would give the smaller one: KAT-2b, oracle/ref_shim_d2m.cu)"""
    sph = np.array([[0, 0, 0, 0.1], [0, 0, 0, 0.2]], np.float32)
    r, cid, _ = ctx.compute_closest_dist2mat(sph, np.array([[0.3, 0.4, 0]], np.float32), np.zeros(1, np.uint32),
                                             np.ones(1, np.uint32), np.array([[-1, 0, 1]], np.int32))
    assert abs(r[0] - 0.3) < 1e-6


def test_queue_kernel_equals_warp_per_sample_kernel(synth, monkeypatch):
    """the queue-compacted kernel (default) and the warp-per-sample kernel (MB_D2M_VARIANT=1, the
    reference's lane <-> primitive layout literally) must agree bit for bit: distances, argmin ids under
    the block tie rule, tie flags -- including lists longer than a warp stride, longer than the
    shared-memory batch (768 slots -> direct path), empty lists and ragged offsets."""
    from libmat_b200.rpd import Context

    d = synth.make_dist2mat(30000)
    rng = np.random.default_rng(7)
    n_prim = d.prims.shape[0]
    # ragged / long lists appended: counts 0, 1, 33, 70, 700, 800, 2000 at random offsets
    extra_cnt = np.array([0, 1, 33, 70, 700, 800, 2000, 5, 0, 64, 32, 31] * 8, np.uint32)
    extra_off = rng.integers(0, n_prim - 2100, len(extra_cnt)).astype(np.uint32)
    pos = d.samples[rng.integers(0, len(d.samples), len(extra_cnt))]
    samples = np.concatenate([d.samples, pos]).astype(np.float32)
    offset = np.concatenate([d.offset, extra_off]).astype(np.uint32)
    count = np.concatenate([d.count, extra_cnt]).astype(np.uint32)
    perm = rng.permutation(len(samples))  # long lists scattered through the batches
    samples, offset, count = samples[perm], offset[perm], count[perm]
    out = []
    for variant in ("0", "1"):
        monkeypatch.setenv("MB_D2M_VARIANT", variant)
        c = Context(0)
        out.append(c.compute_closest_dist2mat(d.spheres, samples, offset, count, d.prims))
        c.close()
    (r0, c0, t0), (r1, c1, t1) = out
    assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32))
    assert np.array_equal(c0, c1)
    assert np.array_equal(t0, t1)
    assert (c0[count == 0] == -1).all() and (r0[count == 0] == np.float32(1e16)).all()
