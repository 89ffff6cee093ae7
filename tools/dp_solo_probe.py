"""One S-L DP launch with one heavy frame (K = 8) per SM: the solo latency of a frame, for `ncu --set full` source-level
sampling (where does a single warp spend its time?).  usage: python tools/dp_solo_probe.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from emplanner_carla_b200 import api, scenes
lat = scenes.LATTICE_C
n = int(sys.argv[1]) if len(sys.argv) > 1 else 148
sc = scenes.sl_scenes(4096, lat, 2, k_fixed=8)
d = torch.from_numpy(api.pack_scenes(**sc)[:n].copy()).cuda()
dp = api.dp_params(**lat.kwargs())
NE = api.max_enriched(dp)
plans = torch.zeros((n, api.plan_stride(NE)), dtype=torch.float64, device="cuda")
for _ in range(4):
    api.dp_device(d, dp, NE, plans=plans)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); api.dp_device(d, dp, NE, plans=plans); b.record(); torch.cuda.synchronize()
print("dp solo, %d frames K=8: %.4f ms" % (n, a.elapsed_time(b)))
