#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + source page) into a short text report. Usage: ncu_summary.py file.ncu-rep [kernel-regex ...]"""
import csv, io, subprocess, sys
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
ki = hdr.index("Kernel Name")
for r in data:
    print("== kernel:", r[ki][:90])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("   %-62s %s %s" % (w, r[i], units[i]))
for kre in sys.argv[2:]:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = rows[1]; d = rows[2:]
    ix = {n: i for i, n in enumerate(h)}
    tot = sum(int(r[ix["# Samples"]]) for r in d) or 1
    ti = sum(int(r[ix["Instructions Executed"]]) for r in d) or 1
    print("== source page for", kre, ": samples", tot, "warp instructions", ti)
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {s: sum(int(r[ix[s]]) for r in d) for s in stalls}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]:
        print("   %-26s %5.1f%%" % (k, 100.0 * v / tot))
    c, cs = Counter(), Counter()
    for r in d:
        toks = r[ix["Source"]].split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        op = op.split(".")[0]
        c[op] += int(r[ix["Instructions Executed"]]); cs[op] += int(r[ix["# Samples"]])
    print("   opcode mix (executed / stall samples):")
    for op, v in c.most_common(14):
        print("     %-8s %5.1f%%  %5.1f%%" % (op, 100.0 * v / ti, 100.0 * cs[op] / tot))
