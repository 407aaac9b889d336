"""Joins an `ncu --page source --csv` dump (SASS rows, in program order) with `nvdisasm -g` line info and
aggregates executed instructions / stall samples per CUDA source line.
usage: python profiles/map_sass.py <ncu_sass.csv> <nvdisasm_-g_output_of_that_function> [top_n]"""
import csv, re, sys
from collections import defaultdict

ncu_csv, disasm = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
lines, cur = [], ("?", 0)
for ln in open(disasm, errors="replace"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
rows = list(csv.reader(open(ncu_csv)))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ins, smp = ci["Instructions Executed"], ci["# Samples"]
stall_cols = [(h, i) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for idx, r in enumerate(body):
    key = lines[idx] if idx < len(lines) else ("?", -1)
    a = agg[key]
    a[0] += int(r[ins] or 0)
    a[1] += int(r[smp] or 0)
    for h, i in stall_cols:
        a[2][h] += int(r[i] or 0)
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
print(f"sass rows {len(body)}, disasm instr {len(lines)}, total warp-instr {ti}, samples {ts}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = sorted(a[2].items(), key=lambda kv: -kv[1])[:3]
    print(f"{100*a[1]/ts:5.1f}% samples {100*a[0]/ti:5.1f}% instr  {key[0]}:{key[1]}  " + " ".join(f"{h[6:]}={v}" for h, v in st if v))
