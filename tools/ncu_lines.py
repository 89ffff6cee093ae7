#!/usr/bin/env python
"""Attribute ncu stall samples / executed instructions to CUDA source lines.
usage: ncu_lines.py report.ncu-rep <kernel-regex> <cubin> <mangled-substring> [top]
The cubin must be the one the report was taken with (cuobjdump -xelf all libemplanner.so).
Caveat: nvdisasm reports one line per instruction; code inlined from a lambda or a helper is attributed to whatever
line the compiler kept (often the lambda's first statement or the call site), so a single line can collect a whole
inlined loop.  Cross-check a surprising line against the per-instruction counts of `ncu --page source --csv`
(e.g. the integer division `e / R` in sl_dp.cu's exact() shows 12.9 % here and 1.5 % by its SASS sequence)."""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, kre, cubin, sym = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, infn, files = {}, None, False, {}
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        infn = sym in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, d = rows[1], rows[2:]
ix = {n: i for i, n in enumerate(h)}
base = int(d[0][ix["Address"]], 16)
smp, ins = defaultdict(int), defaultdict(int)
for r in d:
    off = int(r[ix["Address"]], 16) - base
    key = line_of.get(off, ("?", 0))
    smp[key] += int(r[ix["# Samples"]]); ins[key] += int(r[ix["Instructions Executed"]])
ts, ti = sum(smp.values()) or 1, sum(ins.values()) or 1
srcs = {}
def text(key):
    f, l = key
    if f not in srcs:
        try:
            srcs[f] = open("/root/repo/emplanner_carla_b200/csrc/" + f).read().splitlines()
        except Exception:
            srcs[f] = []
    return srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
print("samples %d, warp instructions %d" % (ts, ti))
for key in sorted(smp, key=lambda k: -smp[k])[:top]:
    print("%5.1f%% smp %5.1f%% ins  %s:%-4d %s" % (100.0 * smp[key] / ts, 100.0 * ins[key] / ti, key[0], key[1], text(key)))
