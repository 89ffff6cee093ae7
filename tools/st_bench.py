import sys, time, numpy as np
sys.path.insert(0, '.')
import torch
from emplanner_carla_b200 import api, scenes
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sc = scenes.st_scenes(B, 3)
rec = torch.from_numpy(api.pack_st_scenes(sc["s_in"], sc["s_out"], sc["t_in"], sc["t_out"], sc["v0"])).cuda()
st = api.st_params()
out = api.st_dp_device(rec, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    api.st_dp_device(rec, st, out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print("S-T DP batch %d K=8: %.3f ms/step  %.0f frames/s" % (B, ms, B / ms * 1e3))
