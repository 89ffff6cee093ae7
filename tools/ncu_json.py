#!/usr/bin/env python
"""profiles/ncu_r2_summary.json from an `ncu --set full` report of the two S-L kernels (tools/ncu_ab.sh), keyed by the
hash of the kernel sources so bench.py can refuse stale numbers.  usage: ncu_json.py report.ncu-rep [out.json]"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
rep = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "ncu_r2_summary.json")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[0], rows[2:]
ki = hdr.index("Kernel Name")
col = lambda r, name: float(r[hdr.index(name)].replace(",", "")) if name in hdr and r[hdr.index(name)] not in ("", "n/a") else None
kernels = {}
for r in data:
    name = "sl_dp_kernel" if "sl_dp_kernel" in r[ki] else ("sl_qp_kernel" if "sl_qp_kernel" in r[ki] else None)
    if name is None:
        continue
    unit = lambda m: rows[1][hdr.index(m)]
    dr, dw = col(r, "dram__bytes_read.sum"), col(r, "dram__bytes_write.sum")
    scale = lambda m, v: v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(unit(m), 1.0) if v is not None else None
    kernels[name] = {
        "duration_us": col(r, "gpu__time_duration.sum"),
        "grid": col(r, "launch__grid_size"), "block": col(r, "launch__block_size"),
        "registers_per_thread": col(r, "launch__registers_per_thread"),
        "waves_per_sm": col(r, "launch__waves_per_multiprocessor"),
        "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active": (col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active") or 0) / 100.0,
        "fp64_pipe_active": (col(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") or 0) / 100.0,
        "lsu_pipe_active": (col(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active") or 0) / 100.0,
        "warp_instructions": col(r, "smsp__inst_executed.sum"),
        "dram_bytes_per_launch": (scale("dram__bytes_read.sum", dr) or 0) + (scale("dram__bytes_write.sum", dw) or 0),
        "local_load_instructions": col(r, "smsp__sass_inst_executed_op_local_ld.sum"),
    }
d = {"source_sha": bench.source_hashes(), "report": os.path.basename(rep),
     "workload": "batch 4096, lattice C_60x11, emp_sl_order_dev schedule (tools/sched_probe.py --quick), one launch each, "
                 "ncu --set full --clock-control none (cold-cache, serialised: shares, not absolute times)",
     "kernels": kernels}
json.dump(d, open(out, "w"), indent=1, sort_keys=True)
print(json.dumps(d, indent=1)[:1500])
