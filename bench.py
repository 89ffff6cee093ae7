#!/usr/bin/env python
"""Benchmark of the S-L DP + QP planning step (BASELINE.json metric: planning frames/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--lattice C]

One "step" = one pass of the hot path (DP kernel -> corridor+QP kernel [-> gather of the plan records when
N > 1]) over one batch of synthetic Frenet scenes.  Default workload = BASELINE.json configs[1]:
batch 4096 per GPU, 60x11 S-L lattice, 0-8 random obstacles (weak scaling: per-GPU batch fixed).

  value     frames/s over exactly K steps, inputs resident in HBM, one start and one end CUDA event on the
            launching stream, max over ranks; --pipeline P (default 2) batches are in flight: step k runs on
            stream k % P with its own plan/gather buffer, so step k+1's DP kernel overlaps step k's QP kernel
  pipeline  the same K steps strictly one after the other (value_serial_steps); kernels_ms and roofline come
            from this serial pass (events around each kernel)
  e2e       frames/s through the host-buffer C ABI (emp_sl_plan_submit / emp_sl_plan_wait, P slots; emp_sl_plan_host
            for P = 1): page-locked host buffers in and out, H2D + D2H of every step inside the timed region
  roofline  dominant kernel's algorithmic HBM bytes / its CUDA-event duration vs MEASURED_PEAKS hbm_gbs
  compute   same launch against the FP64 FMA peak measured on this GPU (the binding roofline)
  cpu_baseline  the oracle port (numpy float64, oracle/) on a bounded prefix of the same batch, all cores
`--impl reference` times that CPU port alone (the reference itself is Python that cannot travel to
the GPU box; the port is pinned to it bit-for-bit / by golden vectors).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "planning frames/sec (S-L DP+QP batch)"
UNIT = "frames/s"
# Algorithmic (compulsory) HBM bytes per frame, float64 (DESIGN.md "Kernels").
SCENE_BYTES = 164


# Static ncu facts of the current kernels (one `ncu --set full` capture per change, see profiles/): the FP64 pipe
# utilisation is the compute-roofline fraction; dram bytes per launch (batch 4096) go into roofline.traffic.
NCU_STATIC = {
    "C_60x11": {"fp64_pipe_active": {"sl_dp_kernel": 0.213, "sl_qp_kernel": 0.381},
                "issue_active": {"sl_dp_kernel": 0.609, "sl_qp_kernel": 0.434},
                "dram_bytes_per_launch": {"sl_dp_kernel": 919040 + 107520, "sl_qp_kernel": 5506048},
                "source": "profiles/r1_ncu_sl_final_summary.txt (batch 4096, lattice C, serial pass)"},
}


def algorithmic_bytes(n_enr, n_qp):
    dp = SCENE_BYTES + (4 + n_enr * 16 + 4)                      # scene in; n_enr + (s,l) + status out
    qp = SCENE_BYTES + n_qp * 16 + (n_qp * 24 + 8)               # scene + [::2] path in; (l,dl,ddl) + status out
    return dp, qp


def dp_flops(lat, nobs_mean):
    return lat.edges * (390 + 80 * nobs_mean)                    # SURVEY.md §8(d) operation count of the reference


def qp_flops(n, iters_mean):
    return 1500.0 * n * iters_mean                               # banded IPM ~1.5 k flops per station per iteration


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index=0, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """One CPU core's share of the oracle port: DP -> [::2] -> corridor -> QP on frames [lo, hi)."""
    lat_name, config_id, lo, hi, total = args
    import warnings
    warnings.filterwarnings("ignore")
    from emplanner_carla_b200 import scenes
    from oracle import sl_dp, sl_qp
    lat = scenes.LATTICES[lat_name]
    sc = scenes.sl_scenes(total, lat, config_id)
    sl = slice(lo, hi)
    t0 = time.perf_counter()
    rows, ns, nl, _ = sl_dp.dp_batch(sc["obs_s"][sl], sc["obs_l"][sl], sc["nobs"][sl], sc["s0"][sl], sc["l0"][sl],
                                     sc["dl0"][sl], sc["ddl0"][sl], **lat.kwargs())
    t_dp = time.perf_counter() - t0
    n_opt = 0
    for i, b in enumerate(range(lo, hi)):
        K = int(sc["nobs"][b])
        es, el = sl_dp.enrich(ns[i], nl[i], sc["s0"][b], sc["l0"][b], sc["dl0"][b], sc["ddl0"][b], lat.sampling_res)
        try:
            lmin, lmax = sl_qp.corridor(es[::2], el[::2], sc["obs_s"][b, :K], sc["obs_l"][b, :K], 5, 5)
        except IndexError:
            continue
        _, _, _, info = sl_qp.quadratic_planning(lmin, lmax, sc["l0"][b], sc["dl0"][b], sc["ddl0"][b])
        n_opt += info["status"] == "optimal"
    return hi - lo, t_dp, time.perf_counter() - t0, n_opt


def cpu_port_throughput(lat_name, config_id, frames_per_core, cores):
    """frames/s of the oracle port with `cores` processes, each on its own prefix slice of the batch."""
    import multiprocessing as mp
    total = frames_per_core * cores
    jobs = [(lat_name, config_id, c * frames_per_core, (c + 1) * frames_per_core, total) for c in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    frames = sum(r[0] for r in res)
    return frames / wall, frames, wall, sum(r[1] for r in res) / max(sum(r[2] for r in res), 1e-9)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
def run_reference(args, lat_name, config_id):
    """--impl reference: the CPU port of the reference path, all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_core = max(2, args.cpu_frames // cores)
    for _ in range(min(args.warmup, 1)):
        cpu_port_throughput(lat_name, config_id, 2, cores)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(args.steps):
        _, f, _, _ = cpu_port_throughput(lat_name, config_id, per_core, cores)
        frames += f
    wall = time.perf_counter() - t0
    val = frames / wall
    sample = "%d frames/step (prefix of the batch-%d workload, %d per core), numpy float64 port of DP->corridor->QP" % (
        per_core * cores, args.batch, per_core)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batch=%d synthetic Frenet scenes, lattice %s, 0-8 obstacles" % (args.batch, lat_name),
                       "lattice": lat_name, "batch_per_gpu": args.batch},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--lattice", default="C", choices=["A", "B", "C"])
    ap.add_argument("--cpu-frames", type=int, default=1024, help="frames of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--pipeline", type=int, default=2,
                    help="batches in flight: step k+1's kernels (and copies) overlap step k's on another stream; "
                         "1 = strictly serial steps")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: p2p = plan records stored into peers' buffers by the QP kernel (CUDA IPC, NVLink); "
                         "nccl = one in-place NCCL all-gather after the kernels")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    from emplanner_carla_b200 import scenes
    lat = scenes.LATTICES[args.lattice]
    config_id = {"A": 0, "B": 1, "C": 2}[args.lattice]
    if args.impl == "reference":
        return run_reference(args, args.lattice, config_id)

    import torch
    import torch.distributed as dist
    from emplanner_carla_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    dp, qp = api.dp_params(**lat.kwargs()), api.qp_params()
    NE = api.max_enriched(dp)
    stride = api.plan_stride(NE)
    sc = scenes.sl_scenes(B * world, lat, config_id)                  # each rank plans its own contiguous shard
    shard = {k: v[rank * B:(rank + 1) * B] for k, v in sc.items()}
    rec_host = api.pack_scenes(**shard)
    scenes_d = torch.from_numpy(rec_host).to(dev)
    from emplanner_carla_b200 import sharding
    use_p2p = world > 1 and args.gather == "p2p"
    P = max(1, min(args.pipeline, api.PLAN_SLOTS))
    gbufs = [sharding.PeerGatherBuffer(B, stride, dev) if use_p2p else sharding.GatherBuffer(B, stride, dev)
             for _ in range(P)]                                        # one gather buffer per batch in flight
    gbuf = gbufs[0]
    mine = gbuf.mine                                                   # kernels write straight into the send slice
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def step(ev=None, gb=None, st=None):
        gb = gb or gbuf
        st = st or stream
        if ev:
            ev[0].record(st)
        api.dp_device(scenes_d, dp, NE, plans=gb.mine)
        if ev:
            ev[1].record(st)
        if use_p2p:
            api.plan_device_p2p_qp(scenes_d, qp, NE, gb.mine, gb.peer_array, gb.n_peers)   # QP + peer stores
        else:
            api.qp_device(scenes_d, qp, NE, gb.mine)
        if ev:
            ev[2].record(st)
        if world > 1:
            if use_p2p:
                gb.fence()                                             # 1-element all-reduce = cross-rank barrier
            else:
                gb.all_gather()                                        # one in-place NCCL all-gather of plan records
        if ev:
            ev[3].record(st)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    for k in range(args.steps):
        if flush is not None:
            flush.fill_(k & 0xFF)                                      # evict L2 (126 MB) between timed steps
        step(evs[k])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()

    t_dp = sum(e[0].elapsed_time(e[1]) for e in evs)
    t_qp = sum(e[1].elapsed_time(e[2]) for e in evs)
    t_ser = sum(e[0].elapsed_time(e[3]) for e in evs)                  # ms over the K serial steps

    # ---- the timed region of `value`: exactly K steps, P batches in flight ------------------
    # Step k runs on stream k % P with its own plan/gather buffer, so step k+1's DP kernel overlaps step k's QP
    # kernel (each kernel alone leaves the SMs half idle in its tail: profiles/).  One start and one end event on
    # the launching stream bracket the region; the L2 flush stays inside it.
    t_pipe = t_ser
    if P > 1:
        streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
        for k in range(P):
            with torch.cuda.stream(streams[k]):
                step(gb=gbufs[k], st=streams[k])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler2 = ClockSampler(local)
        sampler2.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for st_ in streams:
            st_.wait_event(e0)
        for k in range(args.steps):
            st_ = streams[k % P]
            with torch.cuda.stream(st_):
                if flush is not None:
                    flush.fill_(k & 0xFF)
                step(gb=gbufs[k % P], st=st_)
        for st_ in streams:
            stream.wait_stream(st_)
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler2.stop()
        t_pipe = e0.elapsed_time(e1)
    tt = torch.tensor([t_pipe, t_dp, t_qp, t_ser], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_all, t_dp, t_qp, t_ser = [float(v) for v in tt.tolist()]
    value = B * world * args.steps / (t_all * 1e-3)
    value_serial = B * world * args.steps / (t_ser * 1e-3)

    # ---- end-to-end through the host-buffer C ABI (what a drop-in caller gets) --------------
    planner = api.HostPlanner(max_batch=B, kmax=8, n_enr_max=NE, max_col=lat.col, device=local)
    outs = [api.pinned_empty((B, stride)) for _ in range(P)]         # page-locked host buffers, like a real ingest path
    recs = [api.pinned_empty(rec_host.shape) for _ in range(P)]
    for r in recs:
        r[:] = rec_host
    rec_host, out = recs[0], outs[0]

    def e2e_steps(n):
        if P == 1:
            for _ in range(n):
                planner.plan(rec_host, dp, qp, out=out)
            return
        for k in range(n):                                             # every step: H2D of its scenes, D2H of its plans
            if k >= P:
                planner.plan_wait(k % P)
            planner.plan_submit(k % P, recs[k % P], dp, qp, outs[k % P])
        for k in range(P):
            planner.plan_wait(k)

    e2e_steps(args.warmup)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = B * world * args.steps / float(e2e_t.item())
    planner.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    view = api.PlanView(mine.cpu().numpy(), NE)
    n_enr = float(view.n_enr.mean())
    n_qp = float(view.n_qp.mean())
    iters = float(view.qp_iters.mean())
    frac_opt = float((view.qp_status == 0).mean())
    it_arr = np.sort(view.qp_iters.astype(int))
    it_stats = {"p50": int(it_arr[len(it_arr) // 2]), "p99": int(it_arr[int(len(it_arr) * 0.99)]), "max": int(it_arr[-1]),
                "frames_over_30": int((it_arr > 30).sum())}
    b_dp, b_qp = algorithmic_bytes(n_enr, n_qp)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    dom = "sl_qp_kernel" if t_qp >= t_dp else "sl_dp_kernel"
    dom_ms = max(t_qp, t_dp) / args.steps
    dom_bytes = (b_qp if t_qp >= t_dp else b_dp) * B
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    fp64_tf, _ = api.fp64_peak(8192)
    # Reference-equivalent FP64 work (SURVEY.md §8(d) operation counts of the *reference* algorithm) per second.
    # It can exceed the FMA peak: the kernels skip work the reference performs (obstacle culling, closed-form
    # quintics, a banded instead of dense KKT solve).  The roofline fraction that bounds the kernels is the FP64
    # pipe utilisation measured by ncu (profiles/, static per round), not this number.
    ref_flops = (dp_flops(lat, float(shard["nobs"].mean())) + 25e6) * B      # DP + dense-cvxopt QP (~25 MFLOP, §8(d))
    ref_equiv = ref_flops / (t_all / args.steps * 1e-3) / 1e12
    ncu = NCU_STATIC.get(lat.name, {})

    cpu = None
    if not args.no_cpu_baseline:
        cores = host_cores()
        per_core = max(2, args.cpu_frames // cores)
        v, f, wall, dp_share = cpu_port_throughput(args.lattice, config_id, per_core, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d-frame prefix of the batch (%d per core), numpy float64 oracle port of DP->corridor->QP, "
                         "%.1f s wall, DP share of CPU time %.0f%%" % (f, per_core, wall, 100 * dp_share)}

    if cpu:
        # the unmodified reference's own DP time per frame, recorded while the golden fixtures were generated in the
        # build container (the reference is Python that cannot travel to the GPU box): context, not a measurement here
        try:
            gsec = np.load(os.path.join(ROOT, "tests", "golden", "sl_%s.npz" % args.lattice))["ref_dp_seconds"]
            cpu["reference_python"] = {"dp_seconds_per_frame_one_core": float(np.mean(gsec)), "frames": int(len(gsec)),
                                       "where": "build container, tests/golden/sl_%s.npz:ref_dp_seconds "
                                                "(DP_algorithm only; its QP needs cvxopt)" % args.lattice}
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_all / args.steps, "higher_is_better": True,
        "pipeline": {"batches_in_flight": P, "value_serial_steps": value_serial, "ms_per_serial_step": t_ser / args.steps,
                     "note": "value: step k+1's DP kernel overlaps step k's QP kernel (stream k % P, own plan buffer); "
                             "kernels_ms / roofline: the serial pass, one kernel at a time"}, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batch=%d/GPU synthetic Frenet scenes, %dx%d S-L lattice (%s), 0-8 random obstacles, "
                               "DP->[::2]->corridor->QP%s" % (B, lat.col, lat.row, lat.name,
                                                               (", all-gather of plan records fused into the QP kernel "
                                                                "(P2P stores over NVLink)" if use_p2p else
                                                                ", NCCL all-gather of plan records") if world > 1 else ""),
                   "lattice": lat.name, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": "frame-sharded x%d" % world,
                   "l2": "flushed with a 256 MiB write between timed steps" if flush is not None else "not flushed",
                   "qp_optimal_fraction": frac_opt, "qp_mean_iterations": iters, "qp_iterations": it_stats},
        "kernels_ms": {"sl_dp_kernel": t_dp / args.steps, "sl_qp_kernel": t_qp / args.steps,
                       "collective": max(t_ser - t_dp - t_qp, 0.0) / args.steps},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": ncu.get("dram_bytes_per_launch", {}).get(dom),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_frame": (b_qp if t_qp >= t_dp else b_dp),
                     "note": "path is FP64-latency bound, not HBM bound (SURVEY.md §0.5); see compute"},
        "compute": {"bound": "fp64", "kernel": dom, "peak": fp64_tf, "unit": "TFLOP/s",
                    "peak_source": "emp_fp64_peak FMA micro-benchmark, this GPU, this run",
                    "fp64_pipe_active_frac": ncu.get("fp64_pipe_active"), "issue_active_frac": ncu.get("issue_active"),
                    "ncu_source": ncu.get("source"),
                    "reference_equivalent_tflops": ref_equiv,
                    "note": "fraction of the FP64 roofline = ncu sm__pipe_fp64_cycles_active of the dominant kernel; "
                            "reference_equivalent counts the reference's operation count and may exceed the peak"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(rec_host.nbytes),
                "d2h_bytes_per_step": int(out.nbytes), "api": ("emp_sl_plan_submit/emp_sl_plan_wait via HostPlanner.plan_submit, %d slots, " % P if P > 1 else
                        "emp_sl_plan_host via HostPlanner.plan, ") + "page-locked host buffers"},
        "gpu_launches": 2 * args.steps, "collectives": (args.steps if world > 1 else 0),
        "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
