"""bench.py's JSON contract: the reference arm runs anywhere (CPU oracle port); the product arm needs a GPU."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


def test_reference_arm_line_on_cpu():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-frames", "16", "--lattice", "A"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"].startswith("planning frames/sec") and d["unit"] == "frames/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["dtype"] == "f64" and "workload" in d["config"]
    # the two arms name the same workload (the driver compares config.workload)
    sys.path.insert(0, ROOT)
    import bench
    from emplanner_carla_b200 import scenes
    assert d["config"]["workload"] == bench.workload_string(4096, scenes.LATTICE_A)


@pytest.mark.gpu
def test_product_arm_line_on_gpu():
    d = _run(["--steps", "3", "--warmup", "3", "--cpu-frames", "32", "--batch", "512", "--min-seconds", "0.05"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["regions"]["count"] >= 3 and abs(d["ms_per_step"] * d["steps"] - d["regions"]["median_ms"]) < 1e-9
    assert d["regions"]["timed_total_s"] >= 0.04
    assert "flush" not in d["config"]["l2"].split(";")[0] and "rotate" in d["config"]["l2"]
    assert d["extra"]["config2"]["global_batch"] == 8192 and d["extra"]["config2"]["value"] > 0
    assert d["extra"]["config4"]["bug_compatible"]["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 512 * 22 * 8 and e["d2h_bytes_per_step"] == 512 * 224 * 8
    assert d["gpu_launches"] == 9 and d["value"] > d["cpu_baseline"]["value"]           # schedule + DP + QP per step
    assert set(d["kernels_ms"]) == {"sl_order_kernel", "sl_dp_kernel", "sl_qp_kernel", "collective"}
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert d["pipeline"]["batches_in_flight"] >= 1 and d["pipeline"]["value_serial_steps"] > 0


def test_bench_defaults_and_host_helpers():
    """Defaults the driver relies on (no flags: one GPU, minutes) and the host-side helpers that must never raise."""
    sys.path.insert(0, ROOT)
    import bench
    from emplanner_carla_b200 import api

    class _Props:                      # a device whose sysfs entry does not exist: numa_bind reports nothing, changes nothing
        pci_bus_id, pci_domain_id, pci_device_id = 0xfe, 0xffff, 0x1f

    class _Torch:
        class cuda:
            @staticmethod
            def get_device_properties(dev):
                return _Props()

    before = os.sched_getaffinity(0)
    assert bench.numa_bind(_Torch, 0) is None and os.sched_getaffinity(0) == before
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'add_argument("--pipeline", type=int, default=4' in src and api.PLAN_SLOTS >= 4
    assert 'add_argument("--gpus", type=int, default=1)' in src and 'add_argument("--warmup", type=int, default=3)' in src
    h = bench.source_hashes()
    assert set(h) == {"sl_dp.cu", "sl_qp.cu"}
    facts, path = bench.ncu_facts()
    # the committed ncu summary belongs to the committed kernels (bench.py refuses it otherwise and prints null fields)
    assert facts is not None and facts["source_sha"] == h, "profiles/ncu_r2_summary.json is stale: re-capture with tools/ncu_ab.sh + tools/ncu_json.py"
