"""Builds libmat_b200/libmat_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

-fmad=false: the combinatorial decisions and the stored plane equations must be bit-identical to
the reference's code built for the host (no FMA contraction); filters call fmaf() explicitly.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmat_b200.so")
SOURCES = ["capi.cu", "rpd_kernels.cu", "rpd_emit.cu", "rpd_topo.cu", "dist2mat_kernels.cu", "dist2mat_lists.cu", "peaks.cu", "adjacency.cu", "bgeo.cu"]
HEADERS = ["mb_internal.h", "rpd_device.cuh", "rpd_clip.cuh", "rpd_clip2.cuh", "rpd_grid.cuh",
           os.path.join("..", "..", "include", "libmat_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-fmad=false", "-Xcompiler", "-fPIC,-fopenmp,-O2,-ffp-contract=off", "-Xptxas", "-v"]


# dist2mat is pinned against the reference's DEVICE build (nvcc's default FMA contraction on the reference's own
# expressions): its slab solve is ill-conditioned, and only the same contraction reproduces the same roots
# (tests/test_gpu_reference_build.py).  The RPD sources stay non-fused (explicit __f*_rn; host-build parity).
FMAD_SOURCES = {"dist2mat_kernels.cu", "peaks.cu"}


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace(".cu", ".o"))
        objs.append(o)
        flags = list(NVCC_FLAGS)
        if s in FMAD_SOURCES:  # see FMAD_SOURCES
            flags[flags.index("-fmad=false")] = "-fmad=true"
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {s}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [nvcc, "-shared", "-o", OUT, *objs, "-lgomp"]
    subprocess.check_call(cmd)
    with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
