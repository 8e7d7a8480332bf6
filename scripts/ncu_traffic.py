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
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d.setdefault(kernel, {})[key] = {"dram_bytes": int(rd + wr), "read": int(rd), "write": int(wr), "launches_averaged": len(data),
                                "source": os.path.basename(rep) + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
json.dump(d, open(path, "w"), indent=1, sort_keys=True)
print(kernel, key, d[kernel][key])
