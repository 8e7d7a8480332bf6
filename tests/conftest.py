"""Shared fixtures.  `-m "not gpu"` runs on the CPU container (oracle vs golden vectors, host logic,
C-ABI load/export checks); `-m gpu` runs the parity tests proper on a B200 through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    from libmat_b200 import build as B
    import shutil
    lib = os.path.join(ROOT, "libmat_b200", "libmat_b200.so")
    if not os.path.exists(lib) or (B.needs_build() and shutil.which("nvcc")):
        B.build()
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()


@pytest.fixture(scope="session")
def O():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="session")
def synth():
    from libmat_b200 import synth as s
    return s


def golden(name):
    path = os.path.join(GOLDEN, name)
    if name.endswith(".npz"):
        return np.load(path)
    import json
    with open(path) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def cfg1(synth):
    """BASELINE.json configs[0]: ~20k tets (n=15 -> 20 250), 1 000 medial spheres, k = 80."""
    mesh = synth.make_ball_mesh(15)
    sites = synth.make_spheres(1000)
    knn, k = synth.knn_site_lists(sites, 80)
    return mesh, sites, knn, k


@pytest.fixture(scope="session")
def cfg1_oracle(cfg1, O):
    """oracle run of config 1: candidate pairs + records (CPU, a fraction of a second)."""
    mesh, sites, knn, k = cfg1
    pt, ps = O.tet_sphere_relation(mesh, sites, knn, k)
    recs, stat, _ = O.run_pairs(mesh, sites, knn, k, pt, ps, impl="oracle")
    return pt, ps, recs, stat


@pytest.fixture(scope="session")
def cfg1_rt(synth):
    """config 1 with regular-triangulation neighbour lists (what the reference's callers pass:
    rpd_api.cxx:35,67) and hidden sites unflagged: a SUFFICIENT list, so the cells tile the mesh."""
    mesh = synth.make_ball_mesh(15)
    sites = synth.make_spheres(1000)
    knn, k, valid = synth.rt_site_lists(sites)
    sites.flags[:] = valid.astype(np.uint32)
    return mesh, sites, knn, k


@pytest.fixture(scope="session")
def ctx():
    """One mb_ctx on cuda:0 (GPU tests only)."""
    from libmat_b200.rpd import Context
    c = Context(0)
    yield c
    c.close()
