"""Per-source-line instruction and stall shares of one kernel from an .ncu-rep (needs ncu, cuobjdump, nvdisasm).
usage: python scripts/ncu_lines.py <report.ncu-rep> <cubin-name-fragment e.g. jv_q8> <mangled-kernel-name-fragment> [top_n]"""
import csv, re, subprocess, sys, tempfile, os
from collections import defaultdict
from pathlib import Path
rep, cubin_frag, kern_frag = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
root = Path(__file__).resolve().parent.parent
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", str(root / "opensearch-jvector_b200" / "lib" / "libjvgpu.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if cubin_frag in f][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True, errors="replace").stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and kern_frag in l][0]
end = next((i for i in range(start + 1, len(dis)) if dis[i].lstrip().startswith(".section")), len(dis))
lines, cur = [], ("?", 0)
for ln in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", ln):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
ins, smp = ci["Instructions Executed"], ci["# Samples"]
stall_cols = [(h, i) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[2:] if len(r) == len(hdr)]
assert len(body) == len(lines), (len(body), len(lines), "library and report are from different builds")
agg = defaultdict(lambda: [0, 0, defaultdict(int)])
for idx, r in enumerate(body):
    a = agg[lines[idx]]
    a[0] += int(r[ins] or 0)
    a[1] += int(r[smp] or 0)
    for h, i in stall_cols:
        a[2][h] += int(r[i] or 0)
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[1] for a in agg.values()) or 1
srcs = {}
def text(key):
    f = root / "opensearch-jvector_b200" / "csrc" / key[0]
    if f.exists():
        if key[0] not in srcs:
            srcs[key[0]] = f.read_text().split("\n")
        return srcs[key[0]][key[1] - 1].strip()[:90]
    return ""
tot = defaultdict(int)
for a in agg.values():
    for h, v in a[2].items():
        tot[h] += v
print(f"sass rows {len(body)}, total warp-instr {ti}, samples {ts}")
print("stalls: " + " ".join(f"{h[6:]}={100*v/ts:.1f}%" for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
for key, a in sorted(agg.items(), key=lambda kv: -(kv[1][1] / ts + kv[1][0] / ti))[:top]:
    st = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
    print(f"{100*a[1]/ts:5.1f}% smp {100*a[0]/ti:5.1f}% ins  {key[0]}:{key[1]:<4d} " + " ".join(f"{h[6:]}={v}" for h, v in st if v) + "  | " + text(key))
