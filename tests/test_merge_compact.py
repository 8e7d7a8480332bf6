"""mb_rpd_merge_compact (host code, no GPU): the counterpart of merge_convex_cells (reference
src/rpd3d_api/rpd_api.cxx:432-479) on compact records -- the records of the affected tets are replaced by the patch,
everything stays in (tet, site) order.  Checked against a plain numpy merge on fabricated records."""
import numpy as np

from libmat_b200 import capi


def fabricate(rng, tets, max_cells=4):
    """(blob uint32, offsets int64 in bytes, list of (tet, site, words)) with 1..max_cells records per listed tet"""
    recs, words, offs = [], [], [0]
    for t in tets:
        for s in sorted(rng.choice(1000, size=rng.integers(1, max_cells + 1), replace=False)):
            n = int(rng.integers(6, 40))
            w = rng.integers(0, 2**32, size=n, dtype=np.uint64).astype(np.uint32)
            w[0], w[1] = t, s
            recs.append((int(t), int(s), w))
            words.append(w)
            offs.append(offs[-1] + 4 * n)
    blob = np.concatenate(words) if words else np.zeros(0, np.uint32)
    return blob, np.array(offs, np.int64), recs


def numpy_merge(prev, patch, affected):
    aff = set(int(t) for t in affected)
    keep = [r for r in prev if r[0] not in aff] + list(patch)
    keep.sort(key=lambda r: (r[0], r[1]))
    blob = np.concatenate([r[2] for r in keep]) if keep else np.zeros(0, np.uint32)
    offs = np.concatenate([[0], np.cumsum([4 * len(r[2]) for r in keep])]).astype(np.int64)
    return blob, offs


def test_merge_replaces_affected_tets():
    rng = np.random.default_rng(7)
    all_tets = np.arange(0, 400)
    prev_blob, prev_offs, prev = fabricate(rng, all_tets)
    affected = np.sort(rng.choice(all_tets, size=60, replace=False)).astype(np.int32)
    # the patch holds records for most affected tets; a few affected tets lose all their cells
    patch_tets = affected[rng.random(len(affected)) > 0.15]
    patch_blob, patch_offs, patch = fabricate(rng, patch_tets, max_cells=6)
    got_blob, got_offs = capi.merge_compact(prev_blob, prev_offs, patch_blob, patch_offs, affected)
    want_blob, want_offs = numpy_merge(prev, patch, affected)
    assert np.array_equal(got_offs, want_offs)
    assert np.array_equal(got_blob, want_blob)


def test_merge_edge_cases():
    rng = np.random.default_rng(11)
    prev_blob, prev_offs, prev = fabricate(rng, np.arange(10, 30))
    empty_blob, empty_offs = np.zeros(0, np.uint32), np.zeros(1, np.int64)
    # nothing affected, empty patch: identity
    b, o = capi.merge_compact(prev_blob, prev_offs, empty_blob, empty_offs, np.zeros(0, np.int32))
    assert np.array_equal(b, prev_blob) and np.array_equal(o, prev_offs)
    # everything affected: the result is the patch
    patch_blob, patch_offs, patch = fabricate(rng, np.arange(10, 30))
    b, o = capi.merge_compact(prev_blob, prev_offs, patch_blob, patch_offs, np.arange(10, 30, dtype=np.int32))
    assert np.array_equal(b, patch_blob) and np.array_equal(o, patch_offs)
    # empty previous result (the first incremental run affects every tet)
    b, o = capi.merge_compact(empty_blob, empty_offs, patch_blob, patch_offs, np.arange(10, 30, dtype=np.int32))
    assert np.array_equal(b, patch_blob) and np.array_equal(o, patch_offs)
    # affected tets at both ends and new tet ids beyond the previous range
    affected = np.array([10, 29, 35], np.int32)
    pb, po, p = fabricate(rng, affected)
    b, o = capi.merge_compact(prev_blob, prev_offs, pb, po, affected)
    wb, wo = numpy_merge(prev, p, affected)
    assert np.array_equal(o, wo) and np.array_equal(b, wb)
