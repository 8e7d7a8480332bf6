"""Summarise an .ncu-rep (read here, no GPU): key metrics + per-source-line hot spots.
usage: python scripts/ncu_summary.py gpurun_out/prof_k_clip.ncu-rep [n_lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed",
        "TPC.TriageCompute.sm__inst_executed_pipe_alu_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in data:
    print("=" * 100)
    for w in want:
        for i, h in enumerate(hdr):
            if h == w:
                print(f"{w:95s} {units[i]:14s} {r[i]}")
    # stall breakdown (pct of warp cycles per issue)
    st = [(h, r[i]) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
    st = sorted(((float(v.replace(",", "")), h) for h, v in st if v not in ("", "n/a")), reverse=True)
    for v, h in st[:8]:
        print(f"   stall {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur = None; agg = []; h2 = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": h2 = r; continue
    if r and r[0] and r[0].isdigit() and h2:
        d = dict(zip(h2, r))
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:80], int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0), int(d["Thread Instructions Executed"] or 0)))
        except Exception:
            pass
ts = sum(a[3] for a in agg) or 1; ti = sum(a[4] for a in agg) or 1
print(f"total samples {ts} total warp-inst {ti}")
for a in sorted(agg, key=lambda x: -x[3])[:topn]:
    print(f"{a[0]:22s} {a[1]:4d} {100*a[3]/ts:5.1f}% smp {100*a[4]/ti:5.1f}% inst thr {a[5]/max(a[4],1):4.1f}  {a[2]}")
