"""torchrun -n N tools/p2p_check.py : verify the fused P2P gather equals the NCCL all-gather, bit for bit."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
from emplanner_carla_b200 import api, scenes, sharding
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lat = scenes.LATTICE_C; B = 1024
dp, qp = api.dp_params(**lat.kwargs()), api.qp_params(); NE = api.max_enriched(dp); stride = api.plan_stride(NE)
sc = scenes.sl_scenes(B * world, lat, 2)
rec = torch.from_numpy(api.pack_scenes(**{k: v[rank * B:(rank + 1) * B] for k, v in sc.items()})).to(dev)
g1 = sharding.GatherBuffer(B, stride, dev)
api.plan_device(rec, dp, qp, NE, plans=g1.mine); full1 = g1.all_gather().clone()
g2 = sharding.PeerGatherBuffer(B, stride, dev)
api.dp_device(rec, dp, NE, plans=g2.mine)
api.plan_device_p2p_qp(rec, qp, NE, g2.mine, g2.peer_array, g2.n_peers)
g2.fence(); torch.cuda.synchronize()
full2 = g2.gathered()
same = torch.equal(full1.view(torch.int64), full2.view(torch.int64))
print("rank", rank, "p2p gather identical to NCCL all-gather:", same, flush=True)
dist.barrier(); g2.close(); dist.destroy_process_group()
sys.exit(0 if same else 1)
