"""Where does the time of the two S-L kernels go: throughput or the slowest frame?  (round-2 scheduling probe)

Times sl_dp_kernel and sl_qp_kernel alone (CUDA events, median of 20) on lattice C for
  * the benchmark batch (4096 frames, K ~ U{0..8}) in its natural order, sorted by K, and snake-dealt by K;
  * one frame per SM (148 frames) for K = 1, 4, 8: the solo latency of a frame = lower bound of any single-wave launch;
  * 4096 frames with fixed K.
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from emplanner_carla_b200 import api, scenes

lat = scenes.LATTICE_C
dp, qp = api.dp_params(**lat.kwargs()), api.qp_params()
NE = api.max_enriched(dp)
dev = torch.device("cuda", 0)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def run(name, rec, scheduled=False):
    d = torch.from_numpy(np.ascontiguousarray(rec)).to(dev)
    plans = torch.zeros((d.shape[0], api.plan_stride(NE)), dtype=torch.float64, device=dev)
    order = api.order_device(d) if scheduled else None
    if scheduled:
        print("%-44s order kernel %.4f ms" % (name, timeit(lambda: api.order_device(d, order))))
    t_dp = timeit(lambda: api.dp_device(d, dp, NE, plans=plans, order=order))
    t_qp = timeit(lambda: api.qp_device(d, qp, NE, plans, order=order))
    it = plans[:, 4].cpu().numpy()
    print("%-44s frames %5d  dp %.4f ms  qp %.4f ms   (qp iterations mean %.2f max %d)" % (name, d.shape[0], t_dp, t_qp, it.mean(), it.max()))


sc = scenes.sl_scenes(4096, lat, 2)
rec = api.pack_scenes(**sc)
K = sc["nobs"]
run("benchmark batch, natural order", rec)
run("benchmark batch, emp_sl_order_dev schedule", rec, scheduled=True)
if "--quick" in sys.argv:
    sys.exit(0)
order = np.argsort(-K, kind="stable")
run("sorted by K descending", rec[order])
N = 1024
snake = np.empty(4096, dtype=np.int64)
for r in range(4):
    for c in range(N):
        snake[c * 4 + r] = order[r * N + (c if r % 2 == 0 else N - 1 - c)]
run("snake-dealt by K (4 frames per CTA balanced)", rec[snake])
for k in (0, 1, 4, 8):
    sck = scenes.sl_scenes(4096, lat, 2, k_fixed=k)
    rk = api.pack_scenes(**sck)
    run("fixed K = %d" % k, rk)
    run("fixed K = %d, one frame per SM" % k, rk[:148])
    run("fixed K = %d, one CTA (4 frames) per SM" % k, rk[:592])
