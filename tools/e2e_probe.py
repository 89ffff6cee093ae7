#!/usr/bin/env python
"""Where does the end-to-end number go at N > 1?  Run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
        tools/e2e_probe.py > gpurun_out/e2e_probe_n4.json
For every rank: bare pinned D2H / H2D bandwidth of one step's buffers with all ranks copying at once, then bench.py's
e2e_multi with the copies on the launching stream and on copy streams of their own - all of it twice, before and after
binding the rank to the NUMA node of its GPU (page-locked buffers are allocated inside each measurement)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def copy_bandwidth(torch, wl, n=40):
    out = torch.empty((wl.B, wl.stride), dtype=torch.float64).pin_memory()
    inp = torch.from_numpy(wl.rec_host).clone().pin_memory()
    res = {}
    for name, fn, nbytes in (("d2h", lambda: out.copy_(wl.gbufs[0].mine, non_blocking=True), out.nbytes),
                             ("h2d", lambda: wl.scenes[0].copy_(inp, non_blocking=True), inp.nbytes)):
        for _ in range(3):
            fn()
        wl.sync()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        res[name + "_GBps"] = nbytes * n / (time.perf_counter() - t0) / 1e9
    return res


def main():
    import torch
    import torch.distributed as dist
    from emplanner_carla_b200 import api, scenes, sharding
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lat = scenes.LATTICES["C"]
    K, B = 20, 4096
    wl = bench.SlWorkload(torch, dist, api, sharding, lat, 2, B, rank, world, dev, "p2p", 2, False)
    value = B * world * K / (float(np.median(wl.timed(K, 3, 0.1))) * 1e-3)
    rows = []
    for bound in (False, True):
        info = bench.numa_bind(torch, dev) if bound else None
        row = {"rank": rank, "numa_bound": info if bound else False, "affinity": len(os.sched_getaffinity(0))}
        row.update(copy_bandwidth(torch, wl))
        for cs in (False, True):
            t = bench.e2e_multi(torch, wl, K, 3, 0.1, copy_streams=cs)[0]
            row["e2e_copy_streams_%d" % int(cs)] = B * world * K / float(np.median(t))
            row["host_enqueue_ms_per_step_%d" % int(cs)] = bench.e2e_multi.host_enqueue_s / K * 1e3
        rows.append(row)
    ok = wl.verify_gather()
    allrows = [None] * world
    dist.all_gather_object(allrows, rows)
    if rank == 0:
        print(json.dumps({"n_gpus": world, "value": value, "gather_verified": ok, "ranks": allrows}, indent=1))
    wl.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
