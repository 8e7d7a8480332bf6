#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the RPD3D hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--mode grid|given] [--workload cfg1|cfg2|cfg4|cfg5|d2m] [--records full|lean|slim]

A "step" = one full restricted power diagram of the workload: candidate search (K1+K2), clipping (K3),
ordering/compaction (K4 first half), all in libmat_b200.so through the C ABI.

  N = 1   workload = BASELINE.json configs[1]: synthetic Kuhn ball mesh n=32 (196 608 tets, 35 937
          vertices), 10 000 medial spheres, neighbour cap k=80 (given mode: regular-triangulation neighbour lists, the
          reference's semantics; grid mode: the library's own uniform-grid search).
  N > 1   tets sharded in contiguous blocks of equal estimated WORK (libmat_b200.dist.balanced_shards from the per-tet cell
          counts of one untimed run + three measured refinements; --equal-shards for equal sizes), spheres replicated
          on every rank.  N = 2 / 4: ~196 608 tets per GPU
          (n = 40 / 51, 20 000 / 40 000 spheres); N = 8 IS BASELINE.json configs[3], the north-star target:
          n = 70, 2 058 000 tets (257 250 per GPU), 100 000 spheres.  Every rank's streamed run
          (mb_rpd_run_to_sink) writes its ordered shard straight into rank 0's HBM over NVLink peer memory
          (CUDA IPC, copy-engine DMA overlapped with the next tet span); the shard directory (bytes, cells per rank)
          goes through a shared-memory mailbox the ranks spin on, which is also the completion barrier -- neither
          payload nor directory travels through NCCL (a 16-byte all-gather per step cost 130-250 us;
          `--gather nccl` keeps the all-gather + grouped send/recv path for comparison).
          `value` = cells of all ranks / max-over-ranks device time.

`value`  : valid cells per second with the inputs resident in HBM (device time, CUDA events on the
           stream the kernels run on, L2 flushed between timed steps).
`e2e`    : the same through the C ABI with HOST buffers, every step: mb_set_tetmesh + mb_rpd_upload_sites (H2D from
           pinned memory) + mb_rpd_run_to_host (kernels + streamed D2H of the compact records), wall clock.
`e2e_shim`: (N = 1) the drop-in C++ call itself, compute_clipped_voro_diagram_GPU with the reference's signature,
           std::vector<ConvexCellHost> construction + expansion included (tests/cxx/shim_driver bench).
`gpu_reference`: (N = 1) the reference's OWN CUDA build (oracle/_ref/libref_rpd_gpu.so, its flags) timed on the same
           GPU at config 1 (and config 2 when memory allows), next to this library on the same input.
`--impl reference` : the reference's own clipping code (oracle/_ref = /root/reference's convex_cell.cu
           compiled for the host; else the oracle port) on ALL host threads (OMP_NUM_THREADS is overridden: torchrun
           sets it to 1), on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rpd_tet_cells_clipped_per_sec"
UNIT = "cells/s"
N_FOR_GPUS = {1: 32, 2: 40, 4: 51, 8: 70}        # N = 8: BASELINE configs[3] (2 058 000 tets)
SITES_FOR_GPUS = {1: 10000, 2: 20000, 4: 40000, 8: 100000}
L2_NOTE = "GPU arm: L2 flushed (512 MB write) between timed steps"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_traffic(kernel: str, key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
    (profiles/traffic.json, written by scripts/ncu_traffic.py from an `ncu --set full` report); None when no
    capture exists for this kernel / workload."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(p))
        e = d.get(kernel, {}).get(key)
        return (int(e["dram_bytes"]), e.get("source")) if e else (None, None)
    except Exception:
        return None, None


def measured_pipes(kernel: str, key: str):
    """pipe utilisation of the same ncu capture (FP64 / FMA / LSU pipes, issue slots, resident warps): what explains
    the time of a kernel that is not bandwidth-bound"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return d.get(kernel, {}).get(key, {}).get("pipes")
    except Exception:
        return None


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBps"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def make_workload(workload: str, n_gpus: int):
    from libmat_b200 import synth

    if workload == "cfg4":
        n, ns = 70, 100000
    elif workload == "cfg1":
        n, ns = 15, 1000
    else:
        n = N_FOR_GPUS.get(n_gpus, int(round(32 * n_gpus ** (1 / 3))))
        ns = SITES_FOR_GPUS.get(n_gpus, 10000 * n_gpus)
    mesh = synth.make_ball_mesh(n)
    sites = synth.make_spheres(ns)
    return mesh, sites, n, ns


from libmat_b200.dist import as_u8_tensor, gather_varlen, shard  # noqa: E402


def tot_bytes_guess(rec_bytes, world):
    return rec_bytes * world * 1.25 + 4096


def algorithmic_bytes(mesh_n_tet, mesh_n_vert, n_site, n_pairs, n_listed, recs_bytes):
    """B_rpd = n_tet*72 + n_vert*16 + n_site*16 + C*4 + N*4 + sum(compact records)."""
    return mesh_n_tet * 72 + mesh_n_vert * 16 + n_site * 16 + n_pairs * 4 + n_listed * 4 + recs_bytes


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_reference_run(mesh, sites, knn, k, pt, ps, repeats=2):
    """the reference's clipping code on the host cores (oracle/_ref if built, else the oracle port)"""
    from oracle import oracle as O

    kind = "reference" if O.ref("rpd") is not None else "port"
    impl = "ref" if kind == "reference" else "oracle"
    best, cells = None, 0
    for _ in range(repeats):
        recs, stat, sec = O.run_pairs(mesh, sites, knn, k, pt, ps, impl=impl, n_threads=host_threads())
        cells = int((recs["status"] == 4).sum())
        best = sec if best is None else min(best, sec)
    return kind, cells, best


def host_threads() -> int:
    """the host cores this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arms override it)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def bench_gpu_reference(ctx, with_cfg2=True):
    """The reference's own CUDA build on this GPU, next to this library on the SAME input and the same (given-
    neighbours) semantics.  Reference times: CUDA events around its clip kernel (clipped_voro_cell_test_GPU_param_tet,
    voronoi.cu:672-706) and the wall clock of its whole entry point compute_clipped_voro_diagram_GPU."""
    import contextlib
    import io
    import torch

    from libmat_b200 import synth
    from oracle import oracle as O

    out = {}
    if O.ref("rpd_gpu") is None:
        return {"unavailable": "oracle/_ref/libref_rpd_gpu.so not built"}
    for name, n, ns in (("config1", 15, 1000), ("config2", 32, 10000)):
        if name == "config2":
            if not with_cfg2:
                continue
            free_gpu = torch.cuda.mem_get_info()[0]
            try:
                free_host = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
            except (ValueError, OSError):
                free_host = 0
            if free_gpu < 60e9 or free_host < 96e9:  # dense e_adjs 2.6 GB, relation 7.9 GB x2, records ~13 GB x2
                out[name] = {"skipped": f"needs ~35 GB device / ~60 GB host (free: {free_gpu / 1e9:.0f} / {free_host / 1e9:.0f} GB)"}
                continue
        mesh = synth.make_ball_mesh(n)
        sites = synth.make_spheres(ns)
        knn, k, valid = synth.rt_site_lists(sites)
        sites.flags[:] = valid.astype(np.uint32)
        try:
            t0 = time.perf_counter()
            got = O.ref_rpd_gpu(mesh, sites, knn, k)  # the reference prints to fd 1: main() has redirected it to stderr
            t_ref = time.perf_counter() - t0
            if got is None:
                out[name] = {"unavailable": "reference CUDA build did not run"}
                continue
            if name == "config1":  # a second call: the first one carries context creation
                got = O.ref_rpd_gpu(mesh, sites, knn, k)
            recs, ms = got
            n_ref = len(recs)
            del recs
        except Exception as exc:  # noqa: BLE001
            out[name] = {"error": str(exc)}
            continue
        # this library, same semantics (given-neighbours mode, the reference's relation predicate), same box
        ctx.set_mesh(mesh)
        ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, knn, k)
        for _ in range(2):
            ctx.run().free()
        r = ctx.run()
        ours = dict(r.kernel_ms)
        n_ours = r.n_cells
        r.free()
        rf = ctx.run(grid_candidates=True)
        ours_fast = dict(rf.kernel_ms)
        rf.free()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            ctx.set_mesh(mesh)
            ctx.upload_sites(sites.site_soa, sites.weights, sites.flags, knn, k)
            ctx.run_to_host(grid_candidates=True).free()
        t_ours_call = (time.perf_counter() - t0) / reps
        out[name] = {
            "workload": f"n={n}: {mesh.n_tet} tets, {ns} spheres, RT lists site_k={k}",
            "cells": {"reference": n_ref, "libmat_b200": n_ours, "equal": bool(n_ref == n_ours)},
            "reference": {"clip_kernel_ms": ms["kernel_ms"], "d2h_ms": ms["d2h_ms"], "call_ms": ms["call_ms"],
                          "cells_per_s_kernel": n_ref / (ms["kernel_ms"] * 1e-3) if ms["kernel_ms"] > 0 else None,
                          "cells_per_s_call": n_ref / (ms["call_ms"] * 1e-3),
                          "what": "unmodified voronoi.cu / convex_cell.cu / knncuda.cu, nvcc -O2 --use_fast_math -DNDEBUG, sm_100"},
            "libmat_b200": {"clip_kernel_ms": ours["clip"], "candidates_ms_reference_predicate": ours["candidates"],
                            "candidates_ms_grid": ours_fast["candidates"], "device_total_ms": ours_fast["total"],
                            "call_ms": 1e3 * t_ours_call,
                            "cells_per_s_kernel": n_ours / (ours["clip"] * 1e-3), "cells_per_s_call": n_ours / t_ours_call},
            "speedup": {"clip_kernel": ms["kernel_ms"] / ours["clip"] if ms["kernel_ms"] > 0 else None,
                        "call": ms["call_ms"] / (1e3 * t_ours_call)},
        }
    return out


def bench_e2e_shim(mesh, sites, knn, k, reps=3):
    """the drop-in C++ call with the reference's signature and types (tests/cxx/shim_driver bench): everything the
    e2e leg times PLUS the std::vector<ConvexCellHost> the caller receives (3 776 B per cell, non-POD)"""
    import tempfile

    drv = os.path.join(ROOT, "tests", "cxx", "_build", "shim_driver")
    if not os.path.exists(drv):
        return {"unavailable": "tests/cxx/_build/shim_driver not prebuilt"}
    with tempfile.TemporaryDirectory() as td:
        fin = os.path.join(td, "in.bin")
        with open(fin, "wb") as f:
            for a in (mesh.vertices.astype(np.float32), mesh.indices.astype(np.int32), mesh.v_adjs.astype(np.int32),
                      mesh.e_adj6.astype(np.int32), mesh.f_adjs.astype(np.int32), mesh.f_ids.astype(np.int32),
                      sites.site_soa, sites.weights, sites.flags, knn.astype(np.int32), np.array([sites.n_site, k], np.int32)):
                a = np.ascontiguousarray(a)
                f.write(np.int64(a.size).tobytes())
                f.write(a.tobytes())
        env = dict(os.environ)
        env["OMP_NUM_THREADS"] = str(host_threads())
        r = subprocess.run([drv, "bench", fin, str(reps)], capture_output=True, text=True, env=env, timeout=600)
    if r.returncode != 0:
        return {"error": f"shim_driver rc={r.returncode}: {r.stderr[-300:]}"}
    d = json.loads(r.stdout.strip().splitlines()[-1])
    d["value"] = d["n_cells"] / (d["call_ms"] * 1e-3)
    d["unit"] = UNIT
    d["path"] = ("compute_clipped_voro_diagram_GPU(...) -> std::vector<ConvexCellHost> (include/libmat_b200_shim.hpp): "
                 "mb_set_tetmesh (dense e_adjs accepted) + mb_rpd_upload_sites + mb_rpd_run_to_host + expansion, median of %d calls" % reps)
    return d


def bench_dist2mat(ctx, n_samples, steps, warmup, with_cpu=True):
    """config 3 shape: 20 000 spheres, 60 000 slabs, 30 000 cones, replicated per-sample int3 lists
    (fix_geo_error.cxx:300-366).  Returns the dist2mat sub-object of the bench line."""
    import torch

    from libmat_b200 import synth

    d = synth.make_dist2mat(n_samples)
    n_prims = int(d.prims.shape[0])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    tp = [pin(x) for x in (d.spheres, d.samples, d.offset, d.count, d.prims)]
    hp = [t.numpy() for t in tp]
    ctx.dist2mat_upload(*hp)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        ctx.dist2mat_run()
    ms = []
    for _ in range(steps):
        flush.zero_()
        torch.cuda.synchronize()
        ms.append(ctx.dist2mat_run())  # CUDA events on the library's stream around the kernel
    k_ms = float(np.mean(ms))
    b_alg = n_samples * (12 + 8 + 8) + n_prims * 12 + len(d.spheres) * 16
    peak, peak_src = hbm_peak()
    # secondary, compute roofline: algorithmic FP32 flops (~12 per sphere, ~60 per cone, ~300 per
    # slab evaluation incl. its share of boundary-cone fallbacks) against the FFMA peak measured on this device
    try:
        fp32_peak, fp64_peak = ctx.measure_peaks()
    except Exception:  # noqa: BLE001
        fp32_peak = fp64_peak = None
    n_sl_eval = int((d.prims[:, 0] != -1).sum())
    n_co_eval = int(((d.prims[:, 0] == -1) & (d.prims[:, 1] != -1)).sum())
    n_sp_eval = n_prims - n_sl_eval - n_co_eval
    flops_alg = 300.0 * n_sl_eval + 60.0 * n_co_eval + 12.0 * n_sp_eval
    # e2e: the reference-facing call with host buffers (7 H2D + kernel + 2 D2H, like dist2mat.cu:287-307)
    res_h = torch.empty(n_samples, dtype=torch.float32).pin_memory()
    cid_h = torch.empty(n_samples, dtype=torch.int32).pin_memory()
    e_steps = max(3, min(steps, 5))
    # (every call returns with the results in host memory, so each one is timed by itself; the median is reported and
    # all of them listed: on some boxes the first call after another leg's large pinned allocations takes several times
    # as long as the rest)
    torch.cuda.synchronize()
    t_calls = []
    for _ in range(e_steps + 1):
        t0 = time.perf_counter()
        ctx._check(ctx.lib.mb_dist2mat(ctx._ctx, hp[0].ctypes.data, len(d.spheres), hp[1].ctypes.data, n_samples,
                                       hp[2].ctypes.data, hp[3].ctypes.data, hp[4].ctypes.data, n_prims,
                                       res_h.numpy().ctypes.data, cid_h.numpy().ctypes.data, None))
        t_calls.append(time.perf_counter() - t0)
    t_e2e = float(np.median(t_calls[1:]))
    out = {"metric": "dist2mat_queries_per_sec", "value": n_samples / (k_ms * 1e-3), "unit": "queries/s",
           "ms_per_step": k_ms,
           "config": {"workload": f"config 3 shape: {n_samples} samples, {len(d.spheres)} spheres, {d.n_slabs} slabs, "
                                  f"{d.n_cones} cones, {n_prims / n_samples:.1f} prims/sample (replicated int3 lists)"},
           "roofline": {"kernel": "k_dist2mat", "bound": "hbm", "achieved": b_alg / (k_ms * 1e-3) / 1e9, "peak": peak,
                        "unit": "GB/s", "frac": b_alg / (k_ms * 1e-3) / 1e9 / peak,
                        "traffic": measured_traffic("k_dist2mat_q", f"d2m-{n_samples}")[0],
                        "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg,
                        "compute": None if not fp32_peak else {
                            "bound": "fp32 issue", "achieved": flops_alg / (k_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                            "frac": flops_alg / (k_ms * 1e-3) / 1e12 / fp32_peak, "fp64_peak_tflops": fp64_peak,
                            "algorithmic_flops_per_launch": flops_alg,
                            "note": "peaks measured on this device with FFMA / DFMA loops (mb_measure_peaks); flops per "
                                    "primitive evaluation: slab 300, cone 60, sphere 12"}},
           "e2e": {"value": n_samples / t_e2e, "unit": "queries/s",
                   "h2d_bytes_per_step": int(sum(x.nbytes for x in hp)), "d2h_bytes_per_step": int(8 * n_samples),
                   "ms_median": 1e3 * t_e2e, "ms_calls": [round(1e3 * t, 2) for t in t_calls[1:]]}}
    # ---- f3: the candidate lists built on the device (mb_dist2mat_by_face): the caller hands over what the reference
    # STARTS from -- medial mesh, (surface fid, site) incidence, samples + their surface-face id -- instead of one
    # private int3 list per sample.  Everything is inside the timed region: medial mesh + incidence upload, list
    # construction on the device, 16 B per sample H2D, kernel, 8 B per sample D2H.
    try:
        smp_h, fid_h = pin(d.samples), pin(d.sample_fid)
        sph_h, mf_h, me_h, fs_h = pin(d.spheres), pin(d.mm_faces), pin(d.mm_edges), pin(d.fid_sites)

        bf_stage = {}

        def by_face_call():
            ta = time.perf_counter()
            ctx.dist2mat_set_medial_mesh(sph_h.numpy(), mf_h.numpy(), me_h.numpy())
            tb = time.perf_counter()
            ctx.dist2mat_set_face_sites(fs_h.numpy(), d.n_fid)
            tc = time.perf_counter()
            ctx._check(ctx.lib.mb_dist2mat_by_face(ctx._ctx, smp_h.numpy().ctypes.data, fid_h.numpy().ctypes.data, n_samples,
                                                   res_h.numpy().ctypes.data, cid_h.numpy().ctypes.data, None, None))
            td = time.perf_counter()
            bf_stage.update(set_medial_mesh=1e3 * (tb - ta), set_face_sites=1e3 * (tc - tb), by_face=1e3 * (td - tc))

        by_face_call()
        off_l, _ = ctx.dist2mat_face_lists()
        ms3 = []
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            ms3.append(ctx.dist2mat_run())
        torch.cuda.synchronize()
        t_calls3 = []
        for _ in range(e_steps + 1):
            t0 = time.perf_counter()
            by_face_call()
            t_calls3.append(time.perf_counter() - t0)
        t_bf = float(np.median(t_calls3[1:]))
        k3 = float(np.mean(ms3))
        n_list = int(off_l[-1])
        per_sample = float(np.mean(off_l[d.sample_fid.astype(np.int64) + 1] - off_l[d.sample_fid.astype(np.int64)]))
        b_alg3 = n_samples * (12 + 4 + 8) + n_list * 12 + len(d.spheres) * 16
        h2d3 = int(smp_h.numpy().nbytes + fid_h.numpy().nbytes + sph_h.numpy().nbytes + mf_h.numpy().nbytes + me_h.numpy().nbytes + fs_h.numpy().nbytes)
        out["by_face"] = {
            "what": "mb_dist2mat_set_medial_mesh + mb_dist2mat_set_face_sites + mb_dist2mat_by_face: per-surface-face lists built on "
                    "the device in the reference's order (fix_geo_error.cxx:149-215), samples carry a face id",
            "value": n_samples / (k3 * 1e-3), "unit": "queries/s", "ms_per_step": k3, "prims_per_sample": per_sample,
            "list_entries_on_device": n_list, "surface_faces": int(d.n_fid),
            "roofline": {"bound": "hbm", "achieved": b_alg3 / (k3 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": b_alg3 / (k3 * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": b_alg3},
            "e2e": {"value": n_samples / t_bf, "unit": "queries/s", "h2d_bytes_per_step": h2d3, "d2h_bytes_per_step": int(8 * n_samples),
                    "ms": 1e3 * t_bf, "ms_calls": [round(1e3 * t, 2) for t in t_calls3[1:]],
                    "stage_ms_last_call": dict(bf_stage)}}
    except Exception as exc:  # noqa: BLE001
        out["by_face"] = {"error": str(exc)}
    # the same query with every distinct list stored once (samples of one surface face share their list:
    # fix_geo_error.cxx:149-215 builds one list per face and :300-366 replicates it per sample).  Offsets may
    # point anywhere, so this needs no new entry point -- only a caller that stops replicating.
    try:
        if n_samples > 2000000:
            raise RuntimeError("skipped above 2 M samples (the host-side list sharing is a numpy pass over every entry)")
        r_rep = ctx.dist2mat_fetch(want_tie=False)
        ds = synth.share_lists(d)
        tps = [pin(x) for x in (ds.spheres, ds.samples, ds.offset, ds.count, ds.prims)]
        hps = [t.numpy() for t in tps]
        ctx.dist2mat_upload(*hps)
        for _ in range(warmup):
            ctx.dist2mat_run()
        ms2 = []
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            ms2.append(ctx.dist2mat_run())
        r_sh = ctx.dist2mat_fetch(want_tie=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ctx._check(ctx.lib.mb_dist2mat(ctx._ctx, hps[0].ctypes.data, len(ds.spheres), hps[1].ctypes.data, n_samples,
                                           hps[2].ctypes.data, hps[3].ctypes.data, hps[4].ctypes.data, int(ds.prims.shape[0]),
                                           res_h.numpy().ctypes.data, cid_h.numpy().ctypes.data, None))
        torch.cuda.synchronize()
        t_sh = (time.perf_counter() - t0) / e_steps
        k2 = float(np.mean(ms2))
        out["shared_lists"] = {
            "value": n_samples / (k2 * 1e-3), "unit": "queries/s", "ms_per_step": k2,
            "distinct_list_entries": int(ds.prims.shape[0]), "replicated_list_entries": n_prims,
            "results_identical_to_replicated": bool(np.array_equal(r_rep[0].view(np.uint32), r_sh[0].view(np.uint32))
                                                    and np.array_equal(r_rep[1], r_sh[1])),
            "e2e": {"value": n_samples / t_sh, "unit": "queries/s", "h2d_bytes_per_step": int(sum(x.nbytes for x in hps)),
                    "d2h_bytes_per_step": int(8 * n_samples)}}
    except Exception as exc:  # noqa: BLE001
        out["shared_lists"] = {"error": str(exc)}
    if with_cpu:
        from oracle import oracle as O
        n_cpu = min(n_samples, 400000)
        kind = "reference" if O.ref("d2m") is not None else "port"
        _, _, sec = O.dist2mat(d, "ref" if kind == "reference" else "oracle", n=n_cpu, n_threads=host_threads())
        out["cpu_baseline"] = {"value": n_cpu / sec, "unit": "queries/s", "cores": host_threads(), "kind": kind,
                               "sample": f"first {n_cpu} samples, reference distance functions + tie rule on all host threads, {sec:.2f} s"}
        # the reference's own CUDA build on this GPU: its kernel alone (one 32-thread block per sample,
        # dist2mat.cu:301-304) on resident buffers, and its whole entry point (7 blocking H2D + kernel + 2 D2H)
        try:
            ko = O.ref_d2m_gpu(d, kernel_only=True, warmup=1, reps=3)
            wc = O.ref_d2m_gpu(d)
            if ko is not None and wc is not None:
                out["gpu_reference"] = {
                    "kernel_ms": ko[2], "queries_per_s_kernel": n_samples / (ko[2] * 1e-3),
                    "call_ms": wc[2], "queries_per_s_call": n_samples / (wc[2] * 1e-3),
                    "speedup": {"kernel": ko[2] / k_ms, "call": (wc[2] * 1e-3) / t_e2e},
                    "what": "unmodified dist2mat.cu (ClosestDistanceToLocalMat / compute_closest_dist2mat), nvcc -O2 sm_100, same input"}
        except Exception as exc:  # noqa: BLE001
            out["gpu_reference"] = {"error": str(exc)}
    return out


def run_reference_arm(args):
    """--impl reference: CPU only, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from libmat_b200 import synth
    from oracle import oracle as O

    mesh, sites, n, ns = make_workload(args.workload, args.gpus)
    knn, k, valid = synth.rt_site_lists(sites)  # what the reference's callers pass (CGAL RT, rpd_api.cxx:35,67)
    sites.flags[:] = valid.astype(np.uint32)
    # bounded sample: every `stride`-th cube of 6 tets (representative of the whole ball)
    target_tets = int(min(24576, max(1536, 2.5e8 / ns)))  # bounds the (untimed) candidate generation
    stride = max(1, mesh.n_tet // target_tets)
    cubes = np.arange(mesh.n_tet // 6)[::stride]
    sel = (cubes[:, None] * 6 + np.arange(6)[None, :]).ravel()
    sub = synth.TetMesh(mesh.vertices, mesh.indices[sel], mesh.v_adjs, mesh.e_adj6[sel], mesh.f_adjs[sel],
                        mesh.f_ids[sel], mesh.n_surf_faces)
    pt, ps = O.tet_sphere_relation(sub, sites, knn, k)  # candidate generation: not timed
    kind = "reference" if O.ref("rpd") is not None else "port"
    impl = "ref" if kind == "reference" else "oracle"
    cores = host_threads()
    cells = 0
    for _ in range(args.warmup):
        O.run_pairs(sub, sites, knn, k, pt, ps, impl=impl, n_threads=cores)
    # the thread count OpenMP really uses now (torchrun's OMP_NUM_THREADS=1 has been overridden by run_pairs)
    used = int(O.ref("rpd").ref_rpd_max_threads()) if kind == "reference" else cores
    if cores > 1 and used <= 1:
        raise SystemExit(f"reference arm would run on 1 of {cores} host threads: refusing to report it as a {cores}-core baseline")
    t_total = 0.0
    for _ in range(args.steps):
        recs, stat, sec = O.run_pairs(sub, sites, knn, k, pt, ps, impl=impl, n_threads=cores)
        cells = int((recs["status"] == 4).sum())
        t_total += sec
    cores = used
    value = cells * args.steps / t_total
    sample = (f"{len(sel)} of {mesh.n_tet} tets (every {stride}th cube), {len(pt)} candidate pairs, {cells} cells "
              f"per step; per-pair clipping loop + record copy timed, candidate generation excluded")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args, n, ns, mesh), "l2": L2_NOTE},
        "run": {"mode": f"given-neighbours (regular-triangulation lists, site_k={k}; reference semantics)",
                "omp_threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_name(args, n, ns, mesh):
    tag = {15: "cfg1 = BASELINE configs[0]", 32: "cfg2 = BASELINE configs[1]", 70: "cfg4 = BASELINE configs[3]"}.get(
        n, f"cfg2 weak-scaled to {args.gpus} GPUs")
    return (f"{tag}: synthetic Kuhn ball mesh n={n} ({mesh.n_tet} tets, {mesh.n_vert} verts), "
            f"{ns} medial spheres, neighbour cap k=80 (grid_k 96 / RT degree <= 80)")


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="grid", choices=["grid", "given"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg4", "cfg5", "d2m"])
    ap.add_argument("--samples", type=int, default=0, help="dist2mat samples (config 3 is 10 000 000)")
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--records", default="slim", choices=["full", "lean", "slim"],
                    help="transport format of the streamed legs (mb_rpd_opts.lean_records).  full: compact records with "
                         "their plane equations; lean: without them; slim (default): additionally one neighbour id per "
                         "bisector instead of three id words per plane.  The host expansion (mb_rpd_fetch_records / "
                         "mb_rpd_expand_compact) restores all of it bit-exactly from the ids; at N=1 the full-record leg is "
                         "measured as well and reported in e2e.full_records")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="N>1: how the shards reach rank 0")
    ap.add_argument("--chunks", type=int, default=0, help="e2e leg: tet spans of the streamed run (0 = automatic)")
    ap.add_argument("--equal-shards", action="store_true", help="N>1: equal-size tet shards instead of work-balanced ones")
    ap.add_argument("--sweep-chunks", default="", help="N>1: also time the streamed sinks with these span counts, e.g. 1,2,3,4")
    ap.add_argument("--grid-candidates", action="store_true", help="given mode: pairs from the grid search")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cfg2", action="store_true", help="gpu_reference: skip the reference CUDA build at config 2 (tens of GB, tens of seconds)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from libmat_b200 import synth
    from libmat_b200.rpd import Context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: whatever libraries print there (NCCL's version banner on rank 0)
    # goes to stderr instead; the line itself is written to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit_line(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmat_b200 has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version there)
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    if args.workload == "d2m":
        # dist2mat stand-alone: samples sharded across ranks, medial mesh replicated, no collective
        ctx = Context(local_rank)
        n_s = (args.samples or 10000000)
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        sub = bench_dist2mat(ctx, n_s, args.steps, args.warmup, with_cpu=(world == 1 and not args.no_cpu_baseline))
        clocks = sampler.stop() if rank == 0 else None
        vals = torch.tensor([sub["ms_per_step"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        if rank == 0:
            sub.update({"value": n_s * world / (vals.item() * 1e-3), "n_gpus": world, "steps": args.steps,
                        "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                        "dtype": "f32", "data": "synthetic", "gpu_launches": args.steps, "clocks": clocks})
            emit_line(sub)
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    if args.workload == "cfg5":
        # MATTopo-style loop (BASELINE.json configs[4]): 20 successive recomputes with sphere insertion /
        # update on the resident config-2 mesh; per-iteration end-to-end latency (H2D sites + RPD + D2H)
        from libmat_b200.loop import RpdLoop, evolve_sites
        mesh, sites, n, ns = make_workload("cfg2", 1)
        ctx = Context(local_rank)
        from libmat_b200 import capi
        loop = RpdLoop(ctx, mesh)
        fmt = {"full": 0, "lean": 1, "slim": 2}[args.records]
        for _ in range(args.warmup):
            loop.step(sites, to_host=True, lean=fmt)
        lat, dev_ms, cells = [], [], []
        inc_lat, inc_frac, inc_bytes, merge_ms = [], [], [], []
        iters = 20
        sampler = ClockSampler(local_rank)
        sampler.start()
        # the caller's copy of the result (what RPD3D_GPU::powercells / merge_convex_cells keep, rpd_api.cxx:432-479)
        r0, _, _ = loop.step_incremental(sites, to_host=True, lean=fmt)
        hb, ho = r0.host_compact()
        host_blob, host_offs = hb.copy(), ho.copy()
        for it in range(iters):
            sites, changed = evolve_sites(sites, it)  # host-side edit (the caller's fix_topo / fix_geo step), untimed
            # incremental: H2D sites + K1 + K2 on all tets + K3 on the affected tets + streamed D2H of their records
            ri, tets, dti = loop.step_incremental(sites, to_host=True, lean=fmt)
            inc_lat.append(dti * 1e3)
            inc_frac.append(len(tets) / mesh.n_tet)
            inc_bytes.append(ri.compact_bytes)
            pb, po = ri.host_compact()
            tm = time.perf_counter()
            host_blob, host_offs = capi.merge_compact(host_blob, host_offs, pb, po, tets)
            merge_ms.append(1e3 * (time.perf_counter() - tm))
            # full recompute of the same iteration (the checker, and round 1's number)
            res, dt = loop.step(sites, to_host=True, lean=fmt)  # H2D sites + K1..K4a + streamed D2H of the compact result
            lat.append(dt * 1e3)
            dev_ms.append(res.kernel_ms["total"])
            cells.append(res.n_cells)
            fb, fo = res.host_compact()
            if not (np.array_equal(fo, host_offs) and np.array_equal(fb, host_blob)):
                raise SystemExit(f"cfg5: incremental result differs from the full recompute at iteration {it}")
        line = {"metric": "rpd_loop_iteration_latency_ms", "value": float(np.median(inc_lat)), "unit": "ms", "n_gpus": 1,
                "steps": iters, "warmup": args.warmup, "ms_per_step": float(np.mean(inc_lat)), "higher_is_better": False,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "config": {"workload": f"cfg5: 20 iterations on the resident config-2 mesh ({mesh.n_tet} tets), "
                                       f"{ns} -> {sites.n_site} spheres (0.5 % inserted + 0.5 % updated per iteration), "
                                       "incremental recompute (mb_rpd_run_incremental: K2 on all tets, K3 on the tets whose candidate "
                                       f"list changed), {args.records} records streamed to pinned host memory; every iteration is "
                                       "checked byte for byte against a full recompute"},
                "incremental": {"latency_ms": {"min": float(np.min(inc_lat)), "median": float(np.median(inc_lat)), "max": float(np.max(inc_lat))},
                                "affected_tet_fraction": {"min": float(np.min(inc_frac)), "median": float(np.median(inc_frac)), "max": float(np.max(inc_frac))},
                                "d2h_bytes_median": int(np.median(inc_bytes)),
                                "host_merge_ms_median": float(np.median(merge_ms)),
                                "identical_to_full_recompute": True},
                "full_recompute": {"latency_ms": {"min": float(np.min(lat)), "median": float(np.median(lat)), "max": float(np.max(lat))},
                                   "device_ms_median": float(np.median(dev_ms)), "cells_last": int(cells[-1])},
                "e2e": {"value": float(np.median(inc_lat)), "unit": "ms", "h2d_bytes_per_step": int(16 * sites.n_site + 4 * sites.n_site),
                        "d2h_bytes_per_step": int(np.median(inc_bytes))},
                "gpu_launches": int(ctx.launch_count()), "clocks": sampler.stop()}
        emit_line(line)
        ctx.close()
        return

    mesh, sites, n, ns = make_workload(args.workload, n_gpus)
    mode = args.mode if world == 1 else "grid"
    # regular-triangulation neighbour lists (the reference's callers get them from CGAL): used by the
    # given-neighbours mode and by the CPU baseline; hidden sites (empty power cell) are unflagged
    knn, k, valid = synth.rt_site_lists(sites)
    sites.flags[:] = valid.astype(np.uint32)

    stream = torch.cuda.Stream(device=dev)
    ctx = Context(local_rank)
    ctx.set_stream(stream.cuda_stream)

    # host inputs in pinned memory (the e2e leg copies from here every step)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t

    keep = []
    h = {}
    for name, arr in (("verts", mesh.vertices), ("idx", mesh.indices), ("v_adjs", mesh.v_adjs),
                      ("e6", mesh.e_adj6), ("f_adjs", mesh.f_adjs), ("f_ids", mesh.f_ids),
                      ("site", sites.site_soa), ("w", sites.weights), ("flags", sites.flags)):
        h[name], t = pinned(arr)
        keep.append(t)
    if mode == "given":
        h["knn"], t = pinned(knn)
        keep.append(t)
    else:
        h["knn"] = None
    site_k = k if mode == "given" else 0

    my = {"range": shard(mesh.n_tet, rank, world), "balanced": False}  # replaced by work-balanced cuts below (N > 1)

    def set_mesh():
        ctx.set_tetmesh(h["verts"], h["idx"], h["v_adjs"], h["f_adjs"], h["f_ids"], e_adj6=h["e6"])
        first, count = my["range"]
        if world > 1:
            ctx.set_tet_range(first, count)

    def set_mesh_shard():
        """e2e leg, N > 1: a rank uploads only ITS tets (global vertices) and keeps global tet ids"""
        first, count = my["range"]
        ctx.set_tetmesh(h["verts"], h["idx"][first:first + count], h["v_adjs"], h["f_adjs"][first:first + count],
                        h["f_ids"][first:first + count], e_adj6=h["e6"][first:first + count])
        ctx.set_tet_id_base(first)

    def upload_sites():
        ctx.upload_sites(h["site"], h["w"], h["flags"], h["knn"], site_k)

    set_mesh()
    upload_sites()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- multi-GPU result gather on rank 0 -----------------------------------------------------------------
    # default ("p2p"): every rank's streamed run writes its ordered shard straight into rank 0's HBM through a
    # CUDA-IPC peer mapping -- tet span c crosses NVLink by copy-engine DMA while span c+1 is clipped -- and
    # the shard directory goes through a shared-memory mailbox (also the completion barrier).
    # "nccl": one-shot run, then an all-gather of the sizes + grouped NCCL send/recv (libmat_b200.dist).
    gather_buf = {"t": None}
    sink_dev = sink_host = None
    lean = {"full": 0, "lean": 1, "slim": 2}[args.records]
    gather_mode = args.gather if world > 1 else "none"
    if world > 1:
        from libmat_b200.dist import ShardSink, balanced_shards, rebalance_cuts
        if not args.equal_shards:
            # equal-SIZE slabs of a ball differ by 1.5x in work (outer-shell tets hold one or two cells): cut the tet
            # order into shards of equal estimated work from the per-tet cell counts of one untimed run -- the
            # statistics an iteration loop has from its previous iteration anyway (libmat_b200.dist.balanced_shards)
            r0 = ctx.run(lanes_per_cell=args.lanes)
            pt0, _, st0 = r0.pairs()
            r0.free()
            f0, c0 = my["range"]
            cuts, tet_w = balanced_shards(mesh.n_tet, f0, np.bincount(pt0[st0 == 4] - f0, minlength=c0), device=dev)
            for _ in range(3):
                # measured refinement (interior cells cost more than outer-shell cells): per-rank kernel time of two
                # untimed runs -> new cuts
                ctx.set_tet_range(int(cuts[rank]), int(cuts[rank + 1] - cuts[rank]))
                ms_r = []
                for _ in range(2):
                    rb = ctx.run(lanes_per_cell=args.lanes)
                    ms_r.append(sum(rb.kernel_ms[key] for key in ("candidates", "clip", "order")))
                    rb.free()
                tm = torch.zeros(world, dtype=torch.float64, device=dev)
                tm[rank] = min(ms_r)
                dist.all_reduce(tm)
                cuts = rebalance_cuts(cuts, tet_w, tm.cpu().numpy())
            my["range"] = (int(cuts[rank]), int(cuts[rank + 1] - cuts[rank]))
            my["balanced"] = True
            ctx.set_tet_range(*my["range"])
        r0 = ctx.run(lanes_per_cell=args.lanes)
        t = torch.tensor([r0.compact_bytes, r0.n_cells], dtype=torch.int64, device=dev)
        r0.free()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        cap_b, cap_c = int(t[0].item() * 1.25) + 65536, int(t[1].item() * 1.25) + 4096
        if gather_mode == "p2p":
            try:
                sink_dev = ShardSink(ctx, cap_b, cap_c, kind="device")
            except RuntimeError as exc:
                if rank == 0:
                    print(f"[bench] peer sink unavailable ({exc}); falling back to the NCCL gather", file=sys.stderr)
                gather_mode = "nccl"
        try:
            sink_host = ShardSink(ctx, cap_b, cap_c, kind="host", tag=f"mb_bench_{os.environ.get('MASTER_PORT', '0')}")
        except Exception as exc:  # noqa: BLE001  (every rank fails alike: /dev/shm or cudaHostRegister)
            if rank == 0:
                print(f"[bench] shared host sink unavailable ({exc})", file=sys.stderr)
            sink_host = None

    def gather(res):
        d_blob, n_bytes, d_off, n_cells = res.device_buffers()
        src = as_u8_tensor(d_blob, n_bytes, dev)
        if rank == 0 and gather_buf["t"] is None:
            gather_buf["t"] = torch.empty(int(n_bytes * world * 1.25) + 1024, dtype=torch.uint8, device=dev)
        out, sz = gather_varlen(src, dst=0, out=gather_buf["t"])
        gather_buf["last"] = out
        return sum(sz) if rank == 0 else 0

    def step():
        if world == 1:
            return ctx.run(lanes_per_cell=args.lanes, grid_candidates=args.grid_candidates), 0
        if gather_mode == "p2p":
            res, directory = sink_dev.run(n_chunks=args.chunks, lanes_per_cell=args.lanes, lean=lean)
            return res, int(directory[:, 0].sum())
        res = ctx.run(lanes_per_cell=args.lanes, grid_candidates=args.grid_candidates)
        return res, gather(res)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            res, _ = step()
            res.free()
        barrier()
        launches0 = ctx.launch_count()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        clip_ms, cand_ms, order_ms = [], [], []
        cells = pairs = listed = rec_bytes = n_exact_last = 0
        barrier()
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            ev[i][0].record(stream)
            res, _ = step()
            ev[i][1].record(stream)
            ev[i][1].synchronize()
            clip_ms.append(res.kernel_ms["clip"])
            cand_ms.append(res.kernel_ms["candidates"])
            order_ms.append(res.kernel_ms["order"])
            cells, pairs, rec_bytes = res.n_cells, res.n_pairs, res.compact_bytes
            listed = res.n_clips + res.n_culled
            n_exact_last = res.n_exact
            res.free()
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop() if rank == 0 else None
        launches = ctx.launch_count() - launches0
        step_ms = [a.elapsed_time(b) for a, b in ev]
        t_dev = sum(step_ms) / 1e3

        # ---- e2e: host buffers in, compact records out, every step ------------------------------
        e2e_steps = max(3, min(args.steps, 10))
        # destination buffers are allocated (pinned) once, outside the timed region
        e2e_chunks = 1
        if world == 1:
            # the library's own pinned destination is sized by an untimed first streamed run
            ctx.run_to_host(n_chunks=args.chunks, lanes_per_cell=args.lanes, grid_candidates=args.grid_candidates, lean=lean).free()
            blob_host = None
        elif rank == 0:
            blob_host = (torch.empty(int(tot_bytes_guess(rec_bytes, world)), dtype=torch.uint8).pin_memory(),)
        else:
            blob_host = None
        barrier()
        e2e_parts = np.zeros(3)
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            ta = time.perf_counter()
            set_mesh() if world == 1 else set_mesh_shard()
            tb_ = time.perf_counter()
            upload_sites()
            tc = time.perf_counter()
            if world > 1:
                if sink_host is not None:
                    # every rank streams its shard into the shared pinned host segment (all PCIe links in
                    # parallel); after the directory exchange the whole result is in rank 0's address space
                    res, directory = sink_host.run(n_chunks=args.chunks, lanes_per_cell=args.lanes, lean=lean)
                    d2h = int(directory[:, 0].sum() + 8 * (directory[:, 1].sum() + world))
                    e2e_chunks = int(res.n_spans)
                    res.free()
                    e2e_parts += (tb_ - ta, tc - tb_, time.perf_counter() - tc)
                    continue
                # no shared host segment: gather on rank 0's GPU, then one D2H from there
                if gather_mode == "p2p":
                    res, directory = sink_dev.run(n_chunks=args.chunks, lanes_per_cell=args.lanes, lean=lean)
                    gathered = int(directory[:, 0].sum())
                else:
                    res, gathered = step()
                if rank == 0:
                    if blob_host is None or blob_host[0].numel() < gathered:
                        blob_host = (torch.empty(int(gathered * 1.05) + 16, dtype=torch.uint8).pin_memory(),)
                    if gather_mode == "p2p":
                        sink_dev.read_host(directory, blob_host[0].numpy())
                    else:
                        blob_host[0][:gathered].copy_(gather_buf["last"], non_blocking=False)
                d2h = gathered
                res.free()
                e2e_parts += (tb_ - ta, tc - tb_, time.perf_counter() - tc)
                continue
            # streamed run: the D2H of tet span c overlaps the kernels of span c+1; on return the complete
            # compact result (records + offsets) is in pinned host memory
            res = ctx.run_to_host(n_chunks=args.chunks, lanes_per_cell=args.lanes, grid_candidates=args.grid_candidates, lean=lean)
            d2h = res.compact_bytes + 8 * (res.n_cells + 1)
            e2e_chunks = int(res.n_spans)
            res.free()
            e2e_parts += (tb_ - ta, tc - tb_, time.perf_counter() - tc)
        barrier()
        t_e2e = time.perf_counter() - t0
        # the same leg with FULL compact records (N = 1), reported next to the headline
        e2e_lean = None
        if world == 1 and lean:
            ctx.run_to_host(n_chunks=args.chunks, lanes_per_cell=args.lanes, grid_candidates=args.grid_candidates, lean=0).free()
            torch.cuda.synchronize()
            tl0 = time.perf_counter()
            for i in range(e2e_steps):
                set_mesh()
                upload_sites()
                res = ctx.run_to_host(n_chunks=args.chunks, lanes_per_cell=args.lanes, grid_candidates=args.grid_candidates, lean=0)
                lean_bytes = res.compact_bytes + 8 * (res.n_cells + 1)
                res.free()
            torch.cuda.synchronize()
            e2e_lean = {"value": cells * e2e_steps / (time.perf_counter() - tl0), "unit": UNIT, "d2h_bytes_per_step": int(lean_bytes),
                        "note": "full compact records (plane equations and ids stored, nothing to recompute on expansion)"}

    # ---- optional: span-count sweep of the streamed sinks (N > 1), same timing rules, one process ----------
    span_sweep = None
    if world > 1 and args.sweep_chunks:
        span_sweep = {}
        with torch.cuda.stream(stream):
            for kind, sk in (("device", sink_dev), ("host", sink_host)):
                if sk is None:
                    continue
                for nc in [int(x) for x in args.sweep_chunks.split(",")]:
                    for _ in range(2):
                        sk.run(n_chunks=nc, lanes_per_cell=args.lanes, lean=lean)[0].free()
                    barrier()
                    ts = time.perf_counter()
                    for _ in range(args.steps):
                        flush.zero_()
                        sk.run(n_chunks=nc, lanes_per_cell=args.lanes, lean=lean)[0].free()
                    barrier()
                    tt = torch.tensor([time.perf_counter() - ts], dtype=torch.float64, device=dev)
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    span_sweep[f"{kind}:{nc}"] = 1e3 * tt.item() / args.steps  # wall ms per step incl. the L2 flush (~0.1 ms)

    # ---- max over ranks, totals ------------------------------------------------------------------
    tot = torch.tensor([float(cells), float(pairs), float(rec_bytes), float(listed)], dtype=torch.float64, device=dev)
    mx = torch.tensor([t_dev, t_e2e, t_wall, float(np.mean(clip_ms))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    total_cells, total_pairs = int(tot[0].item()), int(tot[1].item())
    t_dev, t_e2e, t_wall, clip_ms_max = mx.tolist()

    if rank == 0:
        peak, peak_src = hbm_peak()
        first, count = my["range"]
        b_alg = algorithmic_bytes(count, mesh.n_vert, ns, pairs, listed, rec_bytes)  # rank 0's launch
        clip_avg_ms = float(np.mean(clip_ms))
        achieved = b_alg / (clip_avg_ms * 1e-3) / 1e9
        clip_traffic, clip_traffic_src = measured_traffic("k_clip", f"{args.workload}-{mode}") if world == 1 else (None, None)
        h2d = sum(h[x].nbytes for x in ("verts", "idx", "v_adjs", "e6", "f_adjs", "f_ids", "site", "w", "flags"))
        if h["knn"] is not None:
            h2d += h["knn"].nbytes
        line = {
            "metric": METRIC, "value": total_cells * args.steps / t_dev, "unit": UNIT, "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
            "data": "synthetic",
            "config": {"workload": workload_name(args, n, ns, mesh), "l2": L2_NOTE},
            "run": {   "mode": "grid-kNN (uniform-grid per-tet search)" if mode == "grid" else f"given-neighbours (RT lists, site_k={k})" + (", grid candidates" if args.grid_candidates else ", reference relation predicate"),
                       "parallelism": f"tet-shards x{world}" + (" (contiguous, cut for equal work: per-tet cell counts of one untimed run + 3 measured refinements)" if my["balanced"] else "") + ", sites replicated" + (
                           "" if world == 1 else (", shards streamed into rank 0's HBM over NVLink peer memory (CUDA IPC, copy-engine DMA "
                                                  "overlapped with the next tet span, %s records) + shared-memory directory mailbox" % args.records if gather_mode == "p2p"
                                                  else ", NCCL gather to rank 0 (all-gather of sizes + grouped send/recv)")),
                       "cells_per_step": total_cells, "candidate_pairs_per_step": total_pairs,
                       "pairs_per_sec": total_pairs * args.steps / t_dev},
            "stage_ms": {"candidates": float(np.mean(cand_ms)), "clip": clip_avg_ms, "order": float(np.mean(order_ms)),
                         "step_wall_ms": 1e3 * t_wall / args.steps},
            "roofline": {"kernel": "k_clip", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": clip_traffic, "traffic_source": clip_traffic_src,
                         "ncu_pipes": measured_pipes("k_clip", f"{args.workload}-{mode}") if world == 1 else None,
                         "exact_predicate_fallbacks_per_step": int(n_exact_last),
                         "measured_peaks_tflops": dict(zip(("fp32", "fp64"), ctx.measure_peaks())) if world == 1 else None,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_alg,
                         "note": "latency-bound irregular kernel (not bandwidth-bound: ncu_pipes); see DESIGN.md section 5 and profiles/"},
            "e2e": {"value": total_cells * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "records": {"full": "full compact records",
                                "lean": "lean transport format (ids without plane equations; the expansion recomputes them bit-exactly)",
                                "slim": "slim transport format (vertices, one neighbour id per bisector, edges; the host expansion "
                                        "recomputes plane equations and plane ids bit-exactly: tests/test_gpu_stream.py)"}[args.records],
                    "full_records": e2e_lean,
                    "path": ("mb_set_tetmesh + mb_rpd_upload_sites + mb_rpd_run_to_host (%d tet spans, D2H of span c "
                             "overlapped with the kernels of span c+1)" % e2e_chunks) if world == 1 else
                            ("mb_set_tetmesh + mb_rpd_upload_sites + mb_rpd_run_to_sink into a shared page-locked host segment "
                             "(every rank streams its tet shard over its own PCIe link, %d spans each%s) + shared-memory directory mailbox"
                             % (e2e_chunks, ", slabs first-touched on the GPU's NUMA node" if getattr(sink_host, "numa_pinned", False) else "")
                             if sink_host is not None else "mb_set_tetmesh + mb_rpd_upload_sites + mb_rpd_run + gather + D2H on rank 0"),
                    "stage_ms": {"set_tetmesh": 1e3 * e2e_parts[0] / e2e_steps, "upload_sites": 1e3 * e2e_parts[1] / e2e_steps,
                                 "run_to_host": 1e3 * e2e_parts[2] / e2e_steps},
                    # achieved device->host rate over the streamed run: per rank (its own PCIe link) and summed over the
                    # box (N > 1: the ranks share the host's PCIe root ports / memory controllers)
                    "d2h_gbs": {"per_rank": d2h / max(world, 1) / (e2e_parts[2] / e2e_steps) / 1e9,
                                "box_aggregate": d2h / (e2e_parts[2] / e2e_steps) / 1e9}},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if span_sweep:
            line["span_sweep_wall_ms"] = span_sweep
        if world == 1 and not args.no_cpu_baseline:
            # CPU baseline: the reference's clipping code on the candidate pairs of a given-mode run
            ctx.upload_sites(h["site"], h["w"], h["flags"], knn, k)
            rg = ctx.run()
            pt, ps, _ = rg.pairs()
            gpu_cells = rg.n_cells
            rg.free()
            kind, cpu_cells, sec = cpu_reference_run(mesh, sites, knn, k, pt, ps)
            line["cpu_baseline"] = {
                "value": cpu_cells / sec, "unit": UNIT, "cores": host_threads(), "kind": kind,
                "sample": f"all {len(pt)} candidate pairs of the workload (given-neighbours, RT lists site_k={k}), {cpu_cells} cells, "
                          f"best of 2, {sec:.2f} s; clipping loop + record copy timed, candidate generation excluded",
                "cells_match_gpu_given_mode": bool(cpu_cells == gpu_cells)}
            # second metric of BASELINE.json: dist2mat queries/s (config 3 shape at 1M samples by default;
            # `--workload d2m --samples 10000000` runs the full config)
            # the drop-in C++ call (reference signature, std::vector<ConvexCellHost> out) and the reference's own CUDA
            # build on this GPU
            try:
                line["e2e_shim"] = bench_e2e_shim(mesh, sites, knn, k)
            except Exception as exc:  # noqa: BLE001
                line["e2e_shim"] = {"error": str(exc)}
            try:
                line["dist2mat"] = bench_dist2mat(ctx, args.samples or 10000000, max(3, args.steps // 2), 3)
            except Exception as exc:  # never lose the headline line
                line["dist2mat"] = {"error": str(exc)}
            # last: the reference's CUDA build at config 2 allocates ~10 GB of dense matrices in this process and the
            # legs measured after it ran several times slower (by_face e2e 131 ms instead of 17 ms)
            try:
                line["gpu_reference"] = bench_gpu_reference(ctx, with_cfg2=not args.no_ref_cfg2)
            except Exception as exc:  # noqa: BLE001
                line["gpu_reference"] = {"error": str(exc)}
        emit_line(line)
    for sk in (sink_dev, sink_host):
        if sk is not None:
            sk.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
