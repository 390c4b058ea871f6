"""Attribute ncu per-SASS-instruction counters to CUDA source lines.
Usage: python tools/ncu_lines.py <rep.ncu-rep> <object-or-so with -lineinfo> <kernel substring> [topN]
Joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line markers of the same function by
instruction index."""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, obj, pat = sys.argv[1:4]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
if rep.endswith(".csv.gz"):
    import gzip
    out = gzip.open(rep, "rt").read()
elif rep.endswith(".csv"):
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
sass = []
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        sass.append(r)
ix = {k: i for i, k in enumerate(hdr)}
# disassemble
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
on = False
cur = ("?", 0)
lines = []
for ln in dis.splitlines():
    if ln.startswith(".text."):
        on = pat in ln
        continue
    if ln.startswith("\t.section") or ln.startswith("//-----"):
        if on and ln.startswith("//-----") and ".text." in ln and pat not in ln:
            on = False
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
print(f"sass rows {len(sass)}, disasm instrs {len(lines)}")
n = min(len(sass), len(lines))
agg = {}
tot_i = tot_s = 0.0
for i in range(n):
    ex = float(sass[i][ix["Instructions Executed"]] or 0)
    sm = float(sass[i][ix["# Samples"]] or 0)
    a = agg.setdefault(lines[i], [0.0, 0.0, 0])
    a[0] += ex; a[1] += sm; a[2] += 1
    tot_i += ex; tot_s += sm
print(f"total executed {tot_i:.3e}, samples {tot_s:.0f}")
print(f"{'file:line':32s} {'sass':>5s} {'inst%':>7s} {'samples%':>9s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{k[0]+':'+str(k[1]):32s} {v[2]:5d} {100*v[0]/tot_i:7.2f} {100*v[1]/tot_s:9.2f}")
