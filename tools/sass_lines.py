#!/usr/bin/env python3
"""Attribute ncu per-SASS-instruction counts to CUDA source lines.

usage: sass_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]
Joins `ncu --page source --csv` (per-instruction executed counts / stall samples, in
address order) with `nvdisasm -g` (the same instructions with //## File/line markers)."""
import csv
import os
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
rname = os.environ.get("NCU_KERNEL", kname)  # kernel name as ncu prints it (demangled), if it differs from the mangled substring
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# a report with several kernels repeats "Kernel Name" + header per kernel: keep the section that matches
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if len(starts) > 1:
    pick = [i for i in starts if rname in rows[i][1]]
    i0 = pick[0]
    i1 = min([j for j in starts if j > i0] + [len(rows)])
    rows = rows[i0:i1]
h = rows[1]
iex, isamp, isrc = h.index("Instructions Executed"), h.index("# Samples"), h.index("Source")
ncu_ins = [(r[isrc].strip(), int(r[iex] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) == len(h)]

dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# split per function
funcs = re.split(r"\n\s*\.text\.", dis)
body = None
for f in funcs:
    if kname in f.split("\n", 1)[0]:
        body = f
        break
assert body, "kernel not found"
cur = ("?", 0)
seq = []
for line in body.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m:
        seq.append((cur, m.group(1).strip()))
print("ncu instructions:", len(ncu_ins), "nvdisasm instructions:", len(seq))
n = min(len(ncu_ins), len(seq))
by_line = defaultdict(lambda: [0, 0])
tot = sum(x[1] for x in ncu_ins)
tots = sum(x[2] for x in ncu_ins)
for i in range(n):
    by_line[seq[i][0]][0] += ncu_ins[i][1]
    by_line[seq[i][0]][1] += ncu_ins[i][2]
print("total warp instructions: %.3fG, samples %d" % (tot / 1e9, tots))
srcs = {}
for (f, l), (ex, sm) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        try:
            srcs[f] = open(os.environ.get("ZB_SRC", "/root/repo/zippy_b200/csrc/") + f).read().split("\n")
        except Exception:
            srcs[f] = []
    text = srcs[f][l - 1].strip() if 0 < l <= len(srcs[f]) else ""
    print("%5.2f%% inst %5.2f%% stall  %s:%d  %s" % (100.0 * ex / tot, 100.0 * sm / max(1, tots), f, l, text[:100]))

# optional phase summary: ranges given as name:file:lo-hi,...
if len(sys.argv) > 5:
    print("---- phases")
    for spec in sys.argv[5].split(","):
        name, f, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        ex = sum(v[0] for (ff, l), v in by_line.items() if ff == f and lo <= l <= hi)
        sm = sum(v[1] for (ff, l), v in by_line.items() if ff == f and lo <= l <= hi)
        print("%-14s %6.2f%% inst %6.2f%% stall" % (name, 100.0 * ex / tot, 100.0 * sm / max(1, tots)))
    other = defaultdict(int)
    for (ff, l), v in by_line.items():
        other[ff] += v[0]
    for ff, v in sorted(other.items(), key=lambda kv: -kv[1]):
        print("file %-32s %6.2f%%" % (ff, 100.0 * v / tot))
