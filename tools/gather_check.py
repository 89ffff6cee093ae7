"""torchrun --nproc-per-node N tools/gather_check.py : the fused all-gather against NCCL, on hardware.

Every rank plans its shard of STEPS different batches through two PeerGatherBuffer slots on two streams (the same
schedule bench.py uses): DP -> QP with NVLink peer stores + arrival flags -> consumer (wait for every shard, snapshot the
gathered batch, hand the buffer back).  Each snapshot must equal, bit for bit, what one NCCL all_gather_into_tensor of
the same plan records produces.  A producer that overwrote a buffer before its consumer was done (write-after-read), a
flag published before the data, or a lost record would all show up as a mismatch; a lost flag as a timeout.
Writes gpurun_out/gather_check_n<N>.json on rank 0; exit code 0 only if every rank verified every step."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from emplanner_carla_b200 import api, scenes, sharding

STEPS, SLOTS = 8, 2
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lat = scenes.LATTICE_C
B = int(os.environ.get("EMP_CHECK_BATCH", "2048"))
dp, qp = api.dp_params(**lat.kwargs()), api.qp_params()
NE = api.max_enriched(dp)
stride = api.plan_stride(NE)
recs = []
for k in range(STEPS):                                                 # a different batch per step
    sc = scenes.sl_scenes(B * world, lat, 50 + k)
    recs.append(torch.from_numpy(api.pack_scenes(**{n: v[rank * B:(rank + 1) * B] for n, v in sc.items()})).to(dev))

ref = sharding.GatherBuffer(B, stride, dev)
want = []
for k in range(STEPS):
    api.plan_device(recs[k], dp, qp, NE, plans=ref.mine)
    want.append(ref.all_gather().clone())
torch.cuda.synchronize()

bufs = [sharding.PeerGatherBuffer(B, stride, dev) for _ in range(SLOTS)]
streams = [torch.cuda.Stream(device=dev) for _ in range(SLOTS)]
got = [None] * STEPS
dist.barrier()
for k in range(STEPS):
    gb, st = bufs[k % SLOTS], streams[k % SLOTS]
    with torch.cuda.stream(st):
        if rank == world - 1 and k == 3:
            torch.cuda._sleep(20_000_000)                               # ~10 ms of skew: one rank falls behind
        api.dp_device(recs[k], dp, NE, plans=gb.mine)
        gb.plan_qp(recs[k], qp, NE)
        gb.wait_arrival()
        got[k] = gb.gathered().clone()                                  # the consumer's read ...
        gb.release()                                                    # ... completes before the buffer is handed back
torch.cuda.synchronize()
ok = True
for gb in bufs:
    ok = ok and int(gb.err.item()) == 0
bad_steps = [k for k in range(STEPS) if not torch.equal(got[k].view(torch.int64), want[k].view(torch.int64))]
ok = ok and not bad_steps
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
print("rank %d: fused gather == NCCL all-gather on %d/%d steps, flag timeouts: %s" % (
    rank, STEPS - len(bad_steps), STEPS, [int(gb.err.item()) for gb in bufs]), flush=True)
if rank == 0:
    out = {"world": world, "steps": STEPS, "slots": SLOTS, "frames_per_rank": B, "record_bytes": stride * 8,
           "all_ranks_verified": bool(flag.item() == 1), "what": "bit-for-bit vs NCCL all_gather_into_tensor, every step, every rank"}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/gather_check_n%d.json" % world, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out), flush=True)
dist.barrier()
for gb in bufs:
    gb.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
