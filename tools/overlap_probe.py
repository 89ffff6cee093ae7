"""Does overlapping step k's QP with step k+1's DP (two streams, double-buffered plan records) pay?"""
import sys, numpy as np
sys.path.insert(0, '.')
import torch
from emplanner_carla_b200 import api, scenes
lat = scenes.LATTICE_C
B, K = 4096, 20
dp, qp = api.dp_params(**lat.kwargs()), api.qp_params()
NE = api.max_enriched(dp)
sc = scenes.sl_scenes(B, lat, 2)
sd = torch.from_numpy(api.pack_scenes(**sc)).cuda()
for ns in (1, 2, 3, 4):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    plans = [torch.zeros((B, api.plan_stride(NE)), dtype=torch.float64, device="cuda") for _ in range(ns)]
    def run(n):
        for k in range(n):
            with torch.cuda.stream(streams[k % ns]):
                api.dp_device(sd, dp, NE, plans=plans[k % ns])
                api.qp_device(sd, qp, NE, plans[k % ns])
    run(4); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams: s.wait_event(e0)
    run(K)
    for s in streams: main.wait_stream(s)
    e1.record(main); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print("streams=%d: %.3f ms/step  %.0f frames/s" % (ns, ms, B / ms * 1e3))
