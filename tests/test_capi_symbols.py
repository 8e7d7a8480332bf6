"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares; the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_header_symbols_exported():
    from libmat_b200 import capi
    lib = capi.load()
    decl = declared_symbols()
    assert len(decl) >= 25
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert sorted(capi.SYMBOLS) == decl


def test_version_and_bounds():
    from libmat_b200 import capi
    lib = capi.load()
    assert b"sm_100a" in lib.mb_version()
    d, f = ctypes.c_double(), ctypes.c_float()
    lib.mb_predicate_bounds(ctypes.byref(d), ctypes.byref(f))
    assert d.value == capi.FILTER_BOUND_F64
    assert np.float32(f.value) == np.float32(capi.FILTER_BOUND_F32)


def test_record_layout_constants():
    from libmat_b200 import capi
    assert capi.RECORD_DTYPE.itemsize == 3456
    off = {n: capi.RECORD_DTYPE.fields[n][1] for n in capi.RECORD_DTYPE.names}
    # SURVEY 8a row a11 (probe of sizeof/offsetof on the reference's ConvexCellTransfer)
    assert off == {"status": 0, "thread_id": 4, "voro_id": 8, "tet_id": 12, "weight": 16, "is_active": 20,
                   "nb_v": 21, "nb_p": 22, "nb_e": 23, "ver": 24, "clip": 416, "id2": 2464, "edge": 2976,
                   "euler": 3432, "cell_vol": 3436, "id": 3440}


def test_record_layout_matches_reference_build(O):
    r = O.ref("rpd")
    if r is None:
        pytest.skip("oracle/_ref not built")
    from libmat_b200 import capi
    assert r.ref_rpd_record_bytes() == capi.RECORD_BYTES
    off = (ctypes.c_int * 17)()
    r.ref_rpd_record_layout(off)
    assert list(off)[:16] == [capi.RECORD_DTYPE.fields[n][1] for n in capi.RECORD_DTYPE.names]
    assert off[16] == 32


def test_no_cpu_fallback():
    """without a CUDA device the product refuses to run (it must never route through the oracle)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from libmat_b200.rpd import Context, LibMatError
    with pytest.raises(LibMatError):
        Context(0)


def test_product_never_imports_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/"""
    bad = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|libref_|#include\s*[\"<][^\n]*oracle)")
    for path in glob.glob(os.path.join(ROOT, "libmat_b200", "**", "*"), recursive=True):
        if path.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
            assert not bad.search(open(path).read()), f"{path} reaches into oracle/"
    for path in glob.glob(os.path.join(ROOT, "include", "*")):
        assert not bad.search(open(path).read()), f"{path} reaches into oracle/"


def test_merge_compact_host_helper():
    """mb_rpd_merge_compact (no GPU): previous records with the affected tets' runs replaced by the patch"""
    import numpy as np
    from libmat_b200 import capi
    rng = np.random.default_rng(1)

    def make(tc):
        blob, offs = [], [0]
        for t, c in tc:
            for k in range(c):
                n = int(rng.integers(5, 12))
                blob += [t, 100 * t + k] + list(rng.integers(0, 1000, n - 2))
                offs.append(offs[-1] + 4 * n)
        return np.array(blob, np.uint32), np.array(offs, np.int64)

    def split(b, o):
        return [tuple(b[o[i] // 4:o[i + 1] // 4].tolist()) for i in range(len(o) - 1)]

    pb, po = make([(t, int(rng.integers(1, 4))) for t in range(12)])
    aff = np.array([0, 2, 5, 9, 11], np.int32)
    qb, qo = make([(0, 2), (2, 1), (5, 3), (11, 1)])  # tet 9 loses all its cells
    ob, oo = capi.merge_compact(pb, po, qb, qo, aff)
    want = []
    for t in range(12):
        want += [r for r in (split(qb, qo) if t in aff else split(pb, po)) if r[0] == t]
    assert split(ob, oo) == want
    ob, oo = capi.merge_compact(pb, po, np.zeros(0, np.uint32), np.zeros(1, np.int64), np.zeros(0, np.int32))
    assert np.array_equal(ob, pb) and np.array_equal(oo, po)
