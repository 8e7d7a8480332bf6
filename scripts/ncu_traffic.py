"""profiles/traffic.json <- dram bytes per launch of one kernel from an .ncu-rep (read here, no GPU).
usage: python scripts/ncu_traffic.py REPORT.ncu-rep KERNEL_KEY WORKLOAD_KEY"""
import csv, io, json, os, subprocess, sys
rep, kernel, key = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    i = hdr.index(name)
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return sum(float(r[i].replace(",", "")) * scale for r in data) / len(data)
rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
def pct(name):
    try:
        i = hdr.index(name)
        return round(sum(float(r[i].replace(",", "")) for r in data) / len(data), 2)
    except ValueError:
        return None
pipes = {"fp64_pipe_pct": pct("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
         "fma_pipe_pct": pct("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
         "lsu_pipe_pct": pct("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
         "issue_active_pct": pct("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "warps_active_pct": pct("sm__warps_active.avg.pct_of_peak_sustained_active"),
         "threads_per_inst": pct("smsp__thread_inst_executed_per_inst_executed.ratio"),
         "dram_throughput_pct": pct("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d.setdefault(kernel, {})[key] = {"dram_bytes": int(rd + wr), "read": int(rd), "write": int(wr), "launches_averaged": len(data), "pipes": pipes,
                                "source": os.path.basename(rep) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
json.dump(d, open(path, "w"), indent=1, sort_keys=True)
print(kernel, key, d[kernel][key])
