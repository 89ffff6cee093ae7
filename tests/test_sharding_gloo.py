"""World-size-2 CPU test (gloo) of the multi-GPU host logic: contiguous sharding + the single in-place
all-gather of plan records reproduce the global frame order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from emplanner_carla_b200 import scenes, sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_plan(frame_ids, stride):
    # a deterministic "plan record" per global frame id
    return torch.from_numpy(np.outer(frame_ids.astype(np.float64) + 1.0, np.arange(1, stride + 1, dtype=np.float64)))


def _worker(rank, world, port, total, stride, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_bounds(total, world, rank)
        buf = sharding.GatherBuffer(hi - lo, stride, "cpu")
        # the rank's scenes are a contiguous slice of the global synthetic batch
        sc_all = scenes.sl_scenes(total, scenes.LATTICE_A, 0)
        sc_mine = scenes.sl_scenes(total, scenes.LATTICE_A, 0)["s0"][lo:hi]
        assert np.array_equal(sc_mine, sc_all["s0"][lo:hi])
        buf.mine.copy_(_fake_plan(np.arange(lo, hi), stride))         # stands in for the kernels' output
        full = buf.all_gather()
        expect = _fake_plan(np.arange(total), stride)
        ret[rank] = bool(torch.equal(full, expect))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_the_batch():
    for total, world in [(4096, 8), (10, 3), (7, 8), (65536, 8)]:
        spans = [sharding.shard_bounds(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_all_gather_of_plan_records_world2():
    world, total, stride = 2, 64, 12
    port = _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, port, total, stride, ret), nprocs=world, join=True)
    assert ret[0] and ret[1]
