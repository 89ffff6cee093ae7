#!/usr/bin/env python
"""Full EM cycle benchmark (BASELINE.json configs[4]): S-L DP -> QP -> path smoothing -> S-T DP -> convex space ->
speed QP -> densify -> merge, api.EmCycle, frame-sharded over the GPUs of one box with one NCCL all-gather of the
per-frame result records (S-L plan record + speed-QP profile).  Also the S-T DP alone (configs[3]).

    python tools/cycle_bench.py [--batch 16384] [--steps 10] [--lattice C]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/cycle_bench.py

Prints one JSON line per workload (rank 0).  CUDA events on the launching stream, max over ranks."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16384, help="global batch (split over the ranks)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--lattice", default="C", choices=["A", "B", "C"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from emplanner_carla_b200 import api, scenes
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lat = scenes.LATTICES[args.lattice]
    B = args.batch // world
    sc = scenes.cycle_scenes(B * world, lat)
    lo, hi = rank * B, (rank + 1) * B
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
    sl = {k: v[lo:hi] for k, v in sc["sl"].items()}
    st = {k: v[lo:hi] for k, v in sc["st"].items()}
    sl_d = cu(api.pack_scenes(**sl))
    st_d = cu(api.pack_st_scenes(st["s_in"], st["s_out"], st["t_in"], st["t_out"], st["v0"]))
    start_d, now_d = cu(np.stack([st["v0"], sc["a0"][lo:hi]], 1)), cu(sc["now"][lo:hi])
    cyc = api.EmCycle(B, api.dp_params(**lat.kwargs()), api.qp_params(), api.st_params(), api.speed_qp_params(),
                      api.smooth_params())
    stride = cyc.buf["plans"].shape[1] + 4 * api.ST_QP_POINTS + 4
    gathered = torch.zeros((world, B, stride), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()

    def pack_and_gather(b):
        mine = gathered[rank]
        n = b["plans"].shape[1]
        mine[:, :n] = b["plans"]
        mine[:, n:n + 4 * api.ST_QP_POINTS] = b["qp"].reshape(B, -1)
        mine[:, -4:-2] = b["qp_status_iters"].to(torch.float64)
        mine[:, -2] = b["convex_status"].to(torch.float64)
        mine[:, -1] = b["merge_status"].to(torch.float64)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), mine.reshape(-1))

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for k, (e0, e1) in enumerate(evs):
            flush.fill_(k & 0xFF)
            e0.record(stream)
            fn()
            e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / args.steps

    ms_cycle = timed(lambda: pack_and_gather(cyc.run(sl_d, st_d, start_d, now_d)))
    stages = []                                            # per-stage share of one cycle (CUDA events between the stages)
    ev = [torch.cuda.Event(enable_timing=True)]
    ev[0].record(stream)
    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        stages.append(name)
        ev.append(e)
    cyc.run(sl_d, st_d, start_d, now_d, mark=mark)
    torch.cuda.synchronize()
    stage_ms = {n: round(ev[i].elapsed_time(ev[i + 1]), 4) for i, n in enumerate(stages)}
    st_out = torch.empty((B, api.ST_OUT_STRIDE), dtype=torch.float64, device=dev)
    ms_st = timed(lambda: api.st_dp_device(st_d, cyc.st, st_out))
    if rank == 0:
        b = cyc.buf
        h = {k: b[k].cpu().numpy() for k in ("qp_status_iters", "convex_status", "merge_status", "dense_status")}
        view = api.PlanView(b["plans"].cpu().numpy(), cyc.NE)
        done = (h["merge_status"] == 0) & (h["qp_status_iters"][:, 0] == 0) & (h["dense_status"] == 0) & (view.qp_status == 0)
        base = {"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
                "dtype": "f64", "data": "synthetic", "unit": "frames/s"}
        print(json.dumps(dict(base, metric="EM cycles/sec (S-L DP->QP->S-T DP->QP, full cycle)", value=B * world / ms_cycle * 1e3,
                              ms_per_step=ms_cycle,
                              config={"workload": "BASELINE configs[4]: batch=%d full EM cycle, lattice %s, %d-way frame shard, "
                                      "NCCL all-gather of result records" % (B * world, lat.name, world),
                                      "gpu_launches_per_step": api.EmCycle.N_LAUNCHES, "stage_ms_rank0": stage_ms,
                                      "frames_finishing_whole_cycle": float(done.mean()),
                                      "sl_qp_optimal": float((view.qp_status == 0).mean()),
                                      "st_full_profile_index_error": float((h["qp_status_iters"][:, 0] == api.ST_INDEX_ERROR).mean())})))
        print(json.dumps(dict(base, metric="S-T speed DP frames/sec", value=B * world / ms_st * 1e3, ms_per_step=ms_st,
                              config={"workload": "BASELINE configs[3]-style: batch=%d S-T DP, 0-8 dynamic obstacles, %d-way frame "
                                      "shard" % (B * world, world)})))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
