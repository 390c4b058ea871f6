"""Compile one .cu with -Xptxas -v and print a table: kernel, registers, stack frame, spill stores/loads.
Usage: python tools/ptxas_table.py magi_v2_b200/csrc/sampler.cu"""
import re
import subprocess
import sys

src = sys.argv[1]
r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                    "-Xptxas", "-v", "-c", src, "-o", "/tmp/_ptxas_table.o"], capture_output=True, text=True)
txt = r.stderr + r.stdout
if r.returncode:
    print(txt)
    sys.exit(1)
rows = []
for ln in txt.splitlines():
    m = re.search(r"Compiling entry function '([^']+)'", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|\(.*", "", name).replace("void ", "")
        rows.append({"name": name})
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m and rows and "stack" not in rows[-1]:
        rows[-1].update(stack=m.group(1), st=m.group(2), ld=m.group(3))
    m = re.search(r"Used (\d+) registers", ln)
    if m and rows:
        rows[-1]["regs"] = m.group(1)
print(f"{'kernel':50s} {'regs':>5s} {'stack':>6s} {'spill_st':>8s} {'spill_ld':>8s}")
for r_ in rows:
    print(f"{r_['name'][:50]:50s} {r_.get('regs','?'):>5s} {r_.get('stack','?'):>6s} {r_.get('st','?'):>8s} {r_.get('ld','?'):>8s}")
