#!/usr/bin/env python
"""Benchmark of the S-L DP + QP planning step (BASELINE.json metric: planning frames/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--batch B] [--lattice C]

One "step" = one pass of the hot path (schedule -> DP kernel -> corridor+QP kernel [-> gather of the plan records when
N > 1]) over one batch of synthetic Frenet scenes.  Default workload = BASELINE.json configs[1]: batch 4096 per GPU,
60x11 S-L lattice, 0-8 random obstacles (weak scaling: per-GPU batch fixed).

  value     frames/s over exactly K steps, inputs resident in HBM, one start and one end CUDA event on the launching
            stream, barrier + synchronize on both sides, max over ranks.  --pipeline P (default 2) batches are in flight
            (step k on stream k % P).  The K-step region is repeated until --min-seconds of timed work has
            accumulated; `value` / `ms_per_step` are those of the MEDIAN region (min / max in `regions`).
            L2: steps rotate over enough scene / plan buffer sets to exceed twice the L2 capacity, so every step reads
            and writes memory that has been evicted since its last use - no flush kernel inside the timed region.
  pipeline  the same K steps strictly one after the other; kernels_ms and roofline come from this serial pass
            (CUDA events around each kernel)
  e2e       N = 1: frames/s through the host-buffer C ABI (emp_sl_plan_submit / emp_sl_plan_wait, P slots): page-locked
            host buffers in and out, H2D + D2H of every step inside the timed region (wall clock).
            N > 1: the same through the multi-GPU product path: every rank copies its shard's scenes from page-locked
            host memory, plans, gathers (so every GPU holds the whole batch) and copies its own shard of the plan
            records back to its host
  roofline  dominant kernel's algorithmic HBM bytes / its CUDA-event duration vs MEASURED_PEAKS hbm_gbs
  compute   the FP64 FMA peak measured on this GPU in this run, next to the ncu FP64-pipe utilisation of the current
            kernels (profiles/ncu_r2_summary.json, refused when the kernel sources changed since the capture)
  cpu_baseline  the oracle port (numpy float64, oracle/) on a bounded prefix of the same batch, all cores
  extra     secondary configurations of BASELINE.json measured in the same run: config2 (8192 frames per GPU = 65536
            on 8 GPUs) and config4 (full EM cycle, 2048 frames per GPU = 16384 on 8 GPUs)
`--impl reference` times that CPU port alone on the same frames (the reference itself is Python that cannot travel to
the GPU box; the port is pinned to it bit-for-bit / by golden vectors).
"""
import argparse
import hashlib
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "planning frames/sec (S-L DP+QP batch)"
UNIT = "frames/s"
SCENE_BYTES = 164                 # algorithmic (compulsory) input bytes per frame
DIVERGE_FACTOR = 2.0              # EmpQpParams.diverge_factor default; the CPU port runs the same stopping rule
NCU_SUMMARY = os.path.join(ROOT, "profiles", "ncu_r2_summary.json")
KERNEL_SOURCES = ("emplanner_carla_b200/csrc/sl_dp.cu", "emplanner_carla_b200/csrc/sl_qp.cu")


def algorithmic_bytes(n_enr, n_qp):
    dp = SCENE_BYTES + (4 + n_enr * 16 + 4)                      # scene in; n_enr + (s,l) + status out
    qp = SCENE_BYTES + n_qp * 16 + (n_qp * 24 + 8)               # scene + [::2] path in; (l,dl,ddl) + status out
    return dp, qp


def dp_flops(lat, nobs_mean):
    return lat.edges * (390 + 80 * nobs_mean)                    # SURVEY.md 8(d) operation count of the reference


def workload_string(batch, lat, world=1):
    return ("batch=%d/GPU synthetic Frenet scenes, %dx%d S-L lattice (%s), 0-8 random obstacles, "
            "DP->[::2]->corridor->QP" % (batch, lat.col, lat.row, lat.name))


def source_hashes():
    out = {}
    for rel in KERNEL_SOURCES:
        with open(os.path.join(ROOT, rel), "rb") as f:
            out[os.path.basename(rel)] = hashlib.sha256(f.read()).hexdigest()[:16]
    return out


def ncu_facts():
    """Per-kernel ncu metrics of the CURRENT kernels, or a note saying why there are none."""
    try:
        with open(NCU_SUMMARY) as f:
            d = json.load(f)
    except Exception:
        return None, "no %s" % os.path.relpath(NCU_SUMMARY, ROOT)
    if d.get("source_sha") != source_hashes():
        return None, "%s is stale: kernel sources changed since the capture" % os.path.relpath(NCU_SUMMARY, ROOT)
    return d, os.path.relpath(NCU_SUMMARY, ROOT)


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
        # same stopping rule as the kernels (the divergence exit), so both arms do the same work on every frame
        _, _, _, info = sl_qp.quadratic_planning(lmin, lmax, sc["l0"][b], sc["dl0"][b], sc["ddl0"][b],
                                                 diverge_factor=DIVERGE_FACTOR)
        n_opt += info["status"] == "optimal"
    return hi - lo, t_dp, time.perf_counter() - t0, n_opt


def cpu_pool(cores):
    """Worker pool of the CPU arm: forked once, after the oracle has been imported, and reused by every step (process
    start-up and imports are not part of the reference path and are not timed)."""
    import multiprocessing as mp
    import warnings
    warnings.filterwarnings("ignore")
    from emplanner_carla_b200 import scenes  # noqa: F401
    from oracle import sl_dp, sl_qp  # noqa: F401
    return mp.get_context("fork").Pool(cores) if cores > 1 else None


def cpu_port_throughput(lat_name, config_id, frames, cores, total, pool=None):
    """frames/s of the oracle port with `cores` processes over the first `frames` frames of the `total`-frame batch."""
    frames = max(cores, min(int(frames), total))
    bounds = [round(c * frames / cores) for c in range(cores + 1)]
    jobs = [(lat_name, config_id, bounds[c], bounds[c + 1], total) for c in range(cores) if bounds[c + 1] > bounds[c]]
    own = pool is None and len(jobs) > 1
    if own:
        pool = cpu_pool(len(jobs))
    t0 = time.perf_counter()
    if len(jobs) == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        res = pool.map(_cpu_worker, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    if own:
        pool.close()
    done = sum(r[0] for r in res)
    return done / wall, done, wall, sum(r[1] for r in res) / max(sum(r[2] for r in res), 1e-9)


_T0 = time.perf_counter()


def trace(msg):
    """Progress marks on stderr (EMP_BENCH_TRACE=1): which phase a multi-rank run is in."""
    if os.environ.get("EMP_BENCH_TRACE"):
        print("[bench %6.1fs rank %s] %s" % (time.perf_counter() - _T0, os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------
def run_reference(args, lat, config_id):
    """--impl reference: the CPU port of the reference path, all host cores, on the product arm's own frames."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    total = args.batch
    pool = cpu_pool(cores)
    # bounded sample: the whole batch per step when the run fits ~200 s, otherwise a prefix of it
    cpu_port_throughput(args.lattice, config_id, cores, cores, total, pool)                          # untimed: warm the workers
    rate, _, _, _ = cpu_port_throughput(args.lattice, config_id, 4 * cores, cores, total, pool)    # untimed probe
    budget = 200.0 / max(args.steps + min(args.warmup, 1), 1)
    frames = total if args.cpu_frames <= 0 else min(total, args.cpu_frames)
    frames = int(max(cores, min(frames, rate * budget)))
    for _ in range(min(args.warmup, 1)):
        cpu_port_throughput(args.lattice, config_id, min(frames, 4 * cores), cores, total, pool)
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        _, f, _, _ = cpu_port_throughput(args.lattice, config_id, frames, cores, total, pool)
        done += f
    wall = time.perf_counter() - t0
    if pool is not None:
        pool.close()
    val = done / wall
    if frames == total:
        sample = "all %d frames of the batch per step" % total
    else:
        sample = "the first %d frames of the batch-%d workload per step (bounded to keep the run within minutes)" % (frames, total)
    sample += ", numpy float64 port of DP->corridor->QP (oracle/), %d processes" % cores
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(args.batch, lat), "lattice": lat.name, "batch_per_gpu": args.batch,
                       "frames_per_step_timed": frames},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class SlWorkload:
    """Device buffers + the step of the S-L planning workload for one per-GPU batch size."""

    def __init__(self, torch, dist, api, sharding, lat, config_id, B, rank, world, dev, gather, P, flush):
        self.torch, self.dist, self.api, self.world, self.rank, self.dev, self.P = torch, dist, api, world, rank, dev, P
        self.B, self.lat = B, lat
        self.dp, self.qp = api.dp_params(**lat.kwargs()), api.qp_params()
        self.NE = api.max_enriched(self.dp)
        self.stride = api.plan_stride(self.NE)
        from emplanner_carla_b200 import scenes
        sc = scenes.sl_scenes(B * world, lat, config_id)               # each rank plans its own contiguous shard
        self.shard = {k: v[rank * B:(rank + 1) * B] for k, v in sc.items()}
        self.rec_host = api.pack_scenes(**self.shard)
        self.p2p = world > 1 and gather == "p2p"
        self.gather = gather
        set_bytes = world * B * self.stride * 8 + self.rec_host.nbytes
        l2 = int(torch.cuda.get_device_properties(dev).L2_cache_size)
        if flush:
            nb = P
        else:
            nb = max(P, int(math.ceil(2.0 * l2 / set_bytes)))
            nb = (nb + P - 1) // P * P                                 # buffer set b is always used on stream b % P
        self.NB, self.l2_bytes, self.set_bytes = nb, l2, set_bytes
        self.scenes = [torch.from_numpy(self.rec_host).to(dev) for _ in range(nb)]
        self.orders = [torch.empty(B, dtype=torch.int32, device=dev) for _ in range(nb)]
        mk = (lambda: sharding.PeerGatherBuffer(B, self.stride, dev)) if self.p2p else \
             (lambda: sharding.GatherBuffer(B, self.stride, dev))
        self.gbufs = [mk() for _ in range(nb)]
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None
        self.main = torch.cuda.current_stream()
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(P)] if P > 1 else [self.main]
        # N > 1, fused gather: the consumer side (wait for every rank's shard, hand the buffer back) runs on its own
        # stream per slot, so rank skew never stalls the stream that launches the next step's kernels
        self.cons = [torch.cuda.Stream(device=dev) for _ in range(P)] if self.p2p else None
        self.done_ev = [None] * nb
        self.launches_per_step = 3 + (1 if self.p2p else 0) + (1 if self.p2p else 0)   # order, DP, QP (+ credit wait, consume)

    def step(self, k, ev=None):
        """Enqueue step k on torch's current stream (buffer set k % NB)."""
        api, b = self.api, k % self.NB
        sc, od, gb = self.scenes[b], self.orders[b], self.gbufs[b]
        st = self.torch.cuda.current_stream()
        if self.flush is not None:
            self.flush.fill_(k & 0xFF)
        if ev:
            ev[0].record(st)
        if self.done_ev[b] is not None:
            st.wait_event(self.done_ev[b])                             # this rank's consumer is done with buffer set b
        api.order_device(sc, od)
        if ev:
            ev[1].record(st)
        api.dp_device(sc, self.dp, self.NE, plans=gb.mine, order=od)
        if ev:
            ev[2].record(st)
        if self.p2p:
            gb.plan_qp(sc, self.qp, self.NE, order=od)                 # QP + NVLink peer stores + arrival flag
        else:
            api.qp_device(sc, self.qp, self.NE, gb.mine, order=od)
        if ev:
            ev[3].record(st)
        if self.world > 1:
            if self.p2p and ev is None:
                torch = self.torch
                ready = torch.cuda.Event()
                ready.record(st)
                cs = self.cons[k % self.P]
                with torch.cuda.stream(cs):
                    cs.wait_event(ready)
                    gb.consume()                                       # wait for every shard, hand the buffer back
                    self.done_ev[b] = torch.cuda.Event()
                    self.done_ev[b].record(cs)
            elif self.p2p:
                gb.consume()                                           # serial pass: on the launching stream, timed
            else:
                gb.all_gather()                                        # one in-place NCCL all-gather of plan records
        if ev:
            ev[4].record(st)

    def sync(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def region(self, K, k0=0):
        """Exactly K steps, P in flight; returns the region's duration in ms on this rank (CUDA events)."""
        torch = self.torch
        self.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.main)
        if self.P > 1:
            for s in self.streams:
                s.wait_event(e0)
        for k in range(K):
            with torch.cuda.stream(self.streams[(k0 + k) % self.P]):
                self.step(k0 + k)
        if self.P > 1:
            for s in self.streams:
                self.main.wait_stream(s)
        if self.cons is not None:
            for s in self.cons:
                self.main.wait_stream(s)                               # the step is done when the gathered batch is there
        e1.record(self.main)
        self.sync()
        return e0.elapsed_time(e1)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def timed(self, K, warmup, min_seconds, max_regions=400):
        """Warm up, then repeat the K-step region until min_seconds of timed work; per-region max over ranks."""
        self.region(max(warmup, 1), 0)                                 # >= `warmup` untimed steps
        k0 = max(warmup, 1)
        t_cal = self.max_over_ranks([self.region(K, k0)])[0]           # untimed calibration region
        k0 += K
        n = int(min(max_regions, max(3, math.ceil(min_seconds * 1e3 / max(t_cal, 1e-3)))))
        times = []
        for _ in range(n):
            times.append(self.region(K, k0))
            k0 += K
        times = self.max_over_ranks(times)
        return times

    def serial(self, K):
        """K steps strictly one after the other on the main stream with events around every kernel."""
        torch = self.torch
        self.sync()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
        for k in range(K):
            self.step(k, evs[k])
        self.sync()
        seg = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) for i in range(4)]
        seg.append(sum(e[0].elapsed_time(e[4]) for e in evs))
        return self.max_over_ranks(seg)                                # order, dp, qp, collective, total (ms over K steps)

    def verify_gather(self):
        """Every buffer set's gathered batch against one NCCL all-gather of the same plan records, bit for bit."""
        if self.world == 1:
            return None
        torch, dist = self.torch, self.dist
        ok = True
        tmp = torch.empty((self.world, self.B, self.stride), dtype=torch.float64, device=self.dev)
        for gb in self.gbufs:
            dist.all_gather_into_tensor(tmp.view(-1), gb.mine.reshape(-1).clone())
            torch.cuda.synchronize()
            ok = ok and bool(torch.equal(tmp.view(torch.int64), gb.full.view(torch.int64)))
            if self.p2p:
                ok = ok and int(gb.err.item()) == 0
        flag = torch.tensor([1 if ok else 0], device=self.dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return bool(flag.item() == 1)

    def close(self):
        for gb in self.gbufs:
            if hasattr(gb, "close"):
                gb.close()


def e2e_single(api, wl, K, warmup, min_seconds, local):
    """N = 1: host buffers through the C ABI (emp_sl_plan_submit / emp_sl_plan_wait), wall clock."""
    B, P = wl.B, wl.P
    planner = api.HostPlanner(max_batch=B, kmax=8, n_enr_max=wl.NE, max_col=wl.lat.col, device=local)
    outs = [api.pinned_empty((B, wl.stride)) for _ in range(P)]       # page-locked host buffers, like a real ingest path
    recs = [api.pinned_empty(wl.rec_host.shape) for _ in range(P)]
    for r in recs:
        r[:] = wl.rec_host

    def steps(n):
        if P == 1:
            for _ in range(n):
                planner.plan(recs[0], wl.dp, wl.qp, out=outs[0])
            return
        for k in range(n):                                             # every step: H2D of its scenes, D2H of its plans
            if k >= P:
                planner.plan_wait(k % P)
            planner.plan_submit(k % P, recs[k % P], wl.dp, wl.qp, outs[k % P])
        for k in range(min(P, n)):
            planner.plan_wait(k)

    steps(max(warmup, 3))
    times, total = [], 0.0
    while len(times) < 3 or (total < min_seconds and len(times) < 400):
        t0 = time.perf_counter()
        steps(K)
        times.append(time.perf_counter() - t0)
        total += times[-1]
    planner.close()
    api_name = ("emp_sl_plan_submit/emp_sl_plan_wait via HostPlanner.plan_submit, %d slots, " % P if P > 1 else
                "emp_sl_plan_host via HostPlanner.plan, ") + "page-locked host buffers"
    return times, int(recs[0].nbytes), int(outs[0].nbytes), api_name


def numa_bind(torch, dev):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so page-locked buffers allocated afterwards are
    local to the GPU's PCIe root (first touch).  Returns a description, or None when sysfs does not say."""
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devid = torch.cuda.get_device_properties(dev).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (dom, bus, devid)
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"node": node, "cpus": len(cpus)}
    except Exception:
        return None


def e2e_multi(torch, wl, K, warmup, min_seconds, copy_streams=None):
    """N > 1: pinned host scenes -> H2D -> schedule/DP/QP with the fused gather -> consume -> D2H of the rank's own
    shard of plan records, every step; wall clock between barriers, max over ranks.
    copy_streams: the copies run on a copy-in and a copy-out stream per slot, ordered against the slot's kernels with
    events, so a slow PCIe copy never sits between two steps' kernels on the launching stream."""
    if copy_streams is None:
        copy_streams = os.environ.get("EMP_BENCH_COPY_STREAMS", "1") != "0"
    B, P = wl.B, wl.P
    pin_in = [torch.from_numpy(wl.rec_host).clone().pin_memory() for _ in range(P)]
    pin_out = [torch.empty((B, wl.stride), dtype=torch.float64).pin_memory() for _ in range(P)]
    state = {"k": 10 * wl.NB}                                          # keep buffer set / stream pairing (k % NB, k % P)
    cin = [torch.cuda.Stream(device=wl.dev) for _ in range(P)] if copy_streams else None
    cout = [torch.cuda.Stream(device=wl.dev) for _ in range(P)] if copy_streams else None
    solved = [None] * wl.NB                                            # buffer set b: its kernels are done
    copied = [None] * wl.NB                                            # buffer set b: its shard is on the host

    def steps(n):
        for _ in range(n):
            k = state["k"]
            state["k"] += 1
            b, s = k % wl.NB, k % P
            st = wl.streams[s]
            if not copy_streams:
                with torch.cuda.stream(st):
                    wl.scenes[b].copy_(pin_in[s], non_blocking=True)
                    wl.step(k)
                    pin_out[s].copy_(wl.gbufs[b].mine, non_blocking=True)
                continue
            with torch.cuda.stream(cin[s]):
                if solved[b] is not None:
                    cin[s].wait_event(solved[b])                       # the last kernels that read scenes[b]
                wl.scenes[b].copy_(pin_in[s], non_blocking=True)
                here = torch.cuda.Event()
                here.record(cin[s])
            with torch.cuda.stream(st):
                st.wait_event(here)
                if copied[b] is not None:
                    st.wait_event(copied[b])                           # the shard of set b's last use has left the GPU
                wl.step(k)
                solved[b] = torch.cuda.Event()
                solved[b].record(st)
            with torch.cuda.stream(cout[s]):
                cout[s].wait_event(solved[b])
                pin_out[s].copy_(wl.gbufs[b].mine, non_blocking=True)
                copied[b] = torch.cuda.Event()
                copied[b].record(cout[s])
        state["enqueued"] = time.perf_counter()
        torch.cuda.synchronize()

    steps(max(warmup, 3))
    times, total, n_regions = [], 0.0, None
    wl.sync()
    t0 = time.perf_counter()
    steps(K)
    cal = wl.max_over_ranks([time.perf_counter() - t0])[0]
    n_regions = int(min(400, max(3, math.ceil(min_seconds / max(cal, 1e-6)))))
    enq = []
    for _ in range(n_regions):
        wl.sync()
        t0 = time.perf_counter()
        steps(K)
        times.append(time.perf_counter() - t0)
        enq.append(state["enqueued"] - t0)
    times = wl.max_over_ranks(times)
    e2e_multi.host_enqueue_s = float(np.median(wl.max_over_ranks(enq)))   # host time to enqueue a region, slowest rank
    return times, int(pin_in[0].nbytes) * wl.world, int(pin_out[0].nbytes) * wl.world, (
        "torch pinned host tensors -> api.order_device / dp_device / sharding.PeerGatherBuffer.plan_qp + consume "
        "(or qp_device + NCCL all-gather) -> D2H of the rank's own shard; every GPU ends each step holding the gathered "
        "batch, each rank returns its shard to its host; bytes are summed over ranks" +
        ("; copies on a copy-in / copy-out stream per slot" if copy_streams else ""))


def extra_config2(torch, dist, api, sharding, lat, rank, world, dev, gather, P, K, min_seconds):
    """BASELINE.json configs[2]: 8192 frames per GPU (65536 on 8 GPUs), same lattice, same step."""
    wl = SlWorkload(torch, dist, api, sharding, lat, 3, 8192, rank, world, dev, gather, P, False)
    try:
        times = wl.timed(K, 3, min_seconds)
        ok = wl.verify_gather()
    finally:
        wl.close()
    med = float(np.median(times))
    return {"config": "BASELINE.json configs[2]: %d frames per GPU, global batch %d, lattice %s" % (8192, 8192 * world, lat.name),
            "global_batch": 8192 * world, "value": 8192 * world * K / (med * 1e-3), "unit": UNIT, "ms_per_step": med / K,
            "regions": len(times), "gather_verified": ok}


def extra_config4(torch, dist, api, lat, rank, world, dev, K, min_seconds):
    """BASELINE.json configs[4]: full EM cycle (S-L DP -> QP -> path -> S-T DP -> speed QP -> densify -> merge), 2048
    frames per GPU (16384 on 8 GPUs).  Reported twice: bug-compatible with the reference draft (a full 16-column S-T
    profile raises IndexError at dp_speed_s[16], speed_planning_test.py:428, so most frames stop before the speed QP)
    and with the documented repair R3 (EmpSpeedQpParams.repair_full_profile) under which every frame runs the QP."""
    from emplanner_carla_b200 import scenes
    B = 2048
    sc = scenes.cycle_scenes(B * world, lat, 4)
    sl = {k: v[rank * B:(rank + 1) * B] for k, v in sc["sl"].items()}
    st = {k: v[rank * B:(rank + 1) * B] for k, v in sc["st"].items()}
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sl_d = t(api.pack_scenes(**sl))
    st_d = t(api.pack_st_scenes(st["s_in"], st["s_out"], st["t_in"], st["t_out"], st["v0"]))
    start = t(np.stack([st["v0"], sc["a0"][rank * B:(rank + 1) * B]], axis=1))
    now = t(sc["now"][rank * B:(rank + 1) * B])
    out = {"config": "BASELINE.json configs[4]: full EM cycle, %d frames per GPU, global batch %d, lattice %s, "
                     "no collective (per-rank outputs)" % (B, B * world, lat.name), "unit": "EM cycles/s"}
    main = torch.cuda.current_stream()
    for tag, r3 in (("bug_compatible", False), ("repair_R3", True)):
        try:
            sq = api.speed_qp_params(repair_full_profile=r3)
        except TypeError:
            if r3:
                continue
            sq = api.speed_qp_params()
        cyc = api.EmCycle(B, api.dp_params(**lat.kwargs()), api.qp_params(), api.st_params(), sq, api.smooth_params())

        def region():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for _ in range(K):
                buf = cyc.run(sl_d, st_d, start, now)
            e1.record(main)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            return e0.elapsed_time(e1), buf

        region()
        cal = torch.tensor([region()[0]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(cal, op=dist.ReduceOp.MAX)                 # every rank must run the same number of regions
        n_regions = int(min(100, max(3, math.ceil(min_seconds * 1e3 / max(float(cal.item()), 1e-3)))))
        times = []
        for _ in range(n_regions):
            ms, buf = region()
            times.append(ms)
        tt = torch.tensor(times, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        med = float(np.median(tt.cpu().numpy()))
        fin = ((buf["merge_status"] == 0) & (buf["dense_status"] == 0) & (buf["qp_status_iters"][:, 0] == 0) &
               (buf["convex_status"] == 0)).double().mean()
        ran_qp = (buf["qp_status_iters"][:, 1] > 0).double().mean()
        stats = torch.stack([fin, ran_qp])
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
            stats = stats / world
        out[tag] = {"value": B * world * K / (med * 1e-3), "ms_per_cycle_batch": med / K, "regions": len(times),
                    "frames_finishing_whole_cycle": float(stats[0].item()), "frames_reaching_speed_qp": float(stats[1].item()),
                    "launches_per_cycle": cyc.N_LAUNCHES}
    return out


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="frames per GPU per step")
    ap.add_argument("--lattice", default="C", choices=["A", "B", "C"])
    ap.add_argument("--cpu-frames", type=int, default=0,
                    help="frames of the CPU sample (0 = default: 1024 for cpu_baseline, the whole batch for --impl reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--flush", action="store_true",
                    help="flush L2 with a 256 MiB write before every step instead of rotating over buffer sets")
    ap.add_argument("--min-seconds", type=float, default=0.25, help="timed work to accumulate by repeating the K-step region")
    ap.add_argument("--pipeline", type=int, default=4,
                    help="batches in flight: step k+1's kernels (and copies) overlap step k's on another stream; "
                         "1 = strictly serial steps")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: p2p = plan records stored into peers' buffers by the QP kernel (CUDA IPC, NVLink) with "
                         "arrival flags; nccl = one in-place NCCL all-gather after the kernels")
    ap.add_argument("--no-extras", action="store_true", help="skip the config2 / config4 secondary measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    from emplanner_carla_b200 import scenes
    lat = scenes.LATTICES[args.lattice]
    config_id = {"A": 0, "B": 1, "C": 2}[args.lattice]
    if args.impl == "reference":
        return run_reference(args, lat, config_id)

    import torch
    import torch.distributed as dist
    from emplanner_carla_b200 import api, sharding

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = numa_bind(torch, dev) if world > 1 else None              # before any page-locked allocation
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B, K = args.batch, args.steps
    P = max(1, min(args.pipeline, api.PLAN_SLOTS))
    trace("process group up")
    wl = SlWorkload(torch, dist, api, sharding, lat, config_id, B, rank, world, dev, args.gather, P, args.flush)

    trace("workload buffers ready (%d buffer sets)" % wl.NB)
    sampler = ClockSampler(local)
    sampler.start()
    times = wl.timed(K, args.warmup, args.min_seconds)                 # the timed regions of `value`
    clocks = sampler.stop()
    trace("timed regions done (%d)" % len(times))
    seg = wl.serial(K)                                                 # per-kernel events, one step at a time
    trace("serial pass done")
    gather_ok = wl.verify_gather()
    trace("gather verified: %s" % gather_ok)
    med = float(np.median(times))
    value = B * world * K / (med * 1e-3)
    t_order, t_dp, t_qp, t_coll, t_ser = seg

    if world == 1:
        e_times, h2d, d2h, e_api = e2e_single(api, wl, K, args.warmup, args.min_seconds, local)
    else:
        e_times, h2d, d2h, e_api = e2e_multi(torch, wl, K, args.warmup, args.min_seconds)
        gather_ok = gather_ok and wl.verify_gather()
    trace("e2e done")
    e_med = float(np.median(e_times))
    e2e_val = B * world * K / e_med

    extra = {}
    if not args.no_extras:
        extra["config2"] = extra_config2(torch, dist, api, sharding, lat, rank, world, dev, args.gather, P, K,
                                         args.min_seconds / 2)
        trace("extra config2 done")
        extra["config4"] = extra_config4(torch, dist, api, lat, rank, world, dev, max(2, K // 4), args.min_seconds / 2)
        trace("extra config4 done")

    if rank != 0:
        wl.close()
        if world > 1:
            dist.destroy_process_group()
        return

    view = api.PlanView(wl.gbufs[0].mine.cpu().numpy(), wl.NE)
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
    dom_ms = max(t_qp, t_dp) / K
    dom_bytes = (b_qp if t_qp >= t_dp else b_dp) * B
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    fp64_tf, _ = api.fp64_peak(8192)
    ref_flops = (dp_flops(lat, float(wl.shard["nobs"].mean())) + 25e6) * B      # DP + dense-cvxopt QP (~25 MFLOP, 8(d))
    ref_equiv = ref_flops / (med / K * 1e-3) / 1e12
    ncu, ncu_src = ncu_facts()
    nk = (ncu or {}).get("kernels", {})

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cores = host_cores()
        frames = args.cpu_frames if args.cpu_frames > 0 else 1024
        pool = cpu_pool(cores)
        cpu_port_throughput(args.lattice, config_id, cores, cores, B, pool)                       # untimed: warm the workers
        v, f, wall, dp_share = cpu_port_throughput(args.lattice, config_id, frames, cores, B, pool)
        if pool is not None:
            pool.close()
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d-frame prefix of the batch, numpy float64 oracle port of DP->corridor->QP (same stopping rule as "
                         "the kernels), %d processes, %.1f s wall, DP share of CPU time %.0f%%" % (f, cores, wall, 100 * dp_share)}
        try:
            gsec = np.load(os.path.join(ROOT, "tests", "golden", "sl_%s.npz" % args.lattice))["ref_dp_seconds"]
            cpu["reference_python"] = {"dp_seconds_per_frame_one_core": float(np.mean(gsec)), "frames": int(len(gsec)),
                                       "where": "build container, tests/golden/sl_%s.npz:ref_dp_seconds (the unmodified "
                                                "reference's DP_algorithm only; its QP needs cvxopt)" % args.lattice}
        except Exception:
            pass

    if args.flush:
        l2_note = "flushed with a 256 MiB write before every step (inside the timed region)"
    else:
        l2_note = ("steps rotate over %d scene/plan buffer sets = %.0f MB > 2 x %.0f MB L2: every step touches memory "
                   "evicted since its last use; no flush kernel in the timed region" % (
                       wl.NB, wl.NB * wl.set_bytes / 1e6, wl.l2_bytes / 1e6))
    workload = workload_string(B, lat)
    gather_note = None
    if world > 1:
        gather_note = ("all-gather of plan records fused into the QP kernel (P2P stores over NVLink + arrival flags + "
                       "credits, no collective)" if wl.p2p else "one NCCL all-gather of plan records per step")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": med / K, "higher_is_better": True,
        "regions": {"count": len(times), "median_ms": med, "min_ms": float(min(times)), "max_ms": float(max(times)),
                    "timed_total_s": float(sum(times)) * 1e-3,
                    "note": "each region = exactly `steps` steps between barrier + synchronize, max over ranks; value and "
                            "ms_per_step are the median region's"},
        "pipeline": {"batches_in_flight": P, "value_serial_steps": B * world * K / (t_ser * 1e-3),
                     "ms_per_serial_step": t_ser / K,
                     "note": "value: step k+1's DP kernel overlaps step k's QP kernel (stream k % P, own buffers); "
                             "kernels_ms / roofline: the serial pass, one kernel at a time"},
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "lattice": lat.name, "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": "frame-sharded x%d" % world, "gather": gather_note, "l2": l2_note,
                   "qp_optimal_fraction": frac_opt, "qp_mean_iterations": iters, "qp_iterations": it_stats},
        "kernels_ms": {"sl_order_kernel": t_order / K, "sl_dp_kernel": t_dp / K, "sl_qp_kernel": t_qp / K,
                       "collective": t_coll / K},
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": nk.get(dom, {}).get("dram_bytes_per_launch"),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_frame": (b_qp if t_qp >= t_dp else b_dp),
                     "note": "path is FP64-latency bound, not HBM bound (SURVEY.md 0.5); see compute"},
        "compute": {"bound": "fp64", "kernel": dom, "peak": fp64_tf, "unit": "TFLOP/s",
                    "peak_source": "emp_fp64_peak FMA micro-benchmark, this GPU, this run",
                    "fp64_pipe_active_frac": {k: v.get("fp64_pipe_active") for k, v in nk.items()} or None,
                    "issue_active_frac": {k: v.get("issue_active") for k, v in nk.items()} or None,
                    "ncu_source": ncu_src, "reference_equivalent_tflops": ref_equiv,
                    "note": "fraction of the FP64 roofline = ncu sm__pipe_fp64_cycles_active of the dominant kernel (one "
                            "capture per kernel change, keyed by source hash); reference_equivalent counts the reference's "
                            "operation count and may exceed the peak"},
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": e_api,
                "regions": len(e_times), "min_s": float(min(e_times)), "max_s": float(max(e_times))},
        "gpu_launches": wl.launches_per_step * K, "collectives": (K if world > 1 and not wl.p2p else 0),
        "clocks": clocks,
    }
    if world > 1:
        line["gather_verified"] = gather_ok
        line["e2e"]["host_enqueue_ms_per_step"] = e2e_multi.host_enqueue_s / K * 1e3
        line["e2e"]["numa_bound"] = numa
    if extra:
        line["extra"] = extra
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)
    wl.close()
    if world > 1:
        dist.destroy_process_group()
    if world > 1 and not gather_ok:
        raise SystemExit("gathered plan records differ from the NCCL all-gather of the same records")


if __name__ == "__main__":
    main()
