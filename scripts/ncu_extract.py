"""Reads an .ncu-rep (ncu must be on PATH) and prints / stores the key metrics of the first profiled launch.
usage: python scripts/ncu_extract.py <report.ncu-rep> <out_raw.csv> [traffic_key]"""
import csv, json, subprocess, sys
from pathlib import Path
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
Path(out).write_text(raw)
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
def num(k):
    v, u = d[k]
    v = float(v.replace(",", ""))
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1, "ns": 1e-6, "s": 1e3}.get(u, 1)
    return v * mult
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "lts__t_sector_hit_rate.pct"]
for k in keys:
    if k in d:
        print(f"{k:65s} {d[k][0]:>16s} {d[k][1]}")
if len(sys.argv) > 3:
    tj = Path("profiles/search_kernel_traffic.json")
    cur = json.loads(tj.read_text()) if tj.exists() else {}
    cur[sys.argv[3]] = {"dram_bytes": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
                        "kernel_ms_under_ncu": num("gpu__time_duration.sum"), "source": Path(out).name}
    tj.write_text(json.dumps(cur, indent=1) + "\n")
    print("traffic ->", cur[sys.argv[3]])
