"""Single-frame latency of the drop-in functions (what the reference's CARLA loop sees, test_9.py:180-201):
DP_algorithm -> [::2] -> cal_lmin_lmax -> Quadratic_planning by their reference names, one frame per call."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "emplanner_carla_b200"))
from planner import path_planning            # the drop-in package
from emplanner_carla_b200 import scenes

sc = scenes.sl_scenes(64, scenes.LATTICE_A, 0, k_fixed=3)
def frame(b):
    K = int(sc["nobs"][b])
    obs_s, obs_l = list(sc["obs_s"][b, :K]), list(sc["obs_l"][b, :K])
    st = (sc["s0"][b], sc["l0"][b], sc["dl0"][b], sc["ddl0"][b])
    t0 = time.perf_counter()
    s, l = path_planning.DP_algorithm(obs_s, obs_l, *st)
    t1 = time.perf_counter()
    try:
        lmin, lmax = path_planning.cal_lmin_lmax(s[::2], l[::2], obs_s, obs_l, 5, 5)
        t2 = time.perf_counter()
        path_planning.Quadratic_planning(lmin, lmax, st[1], st[2], st[3])
    except IndexError:
        t2 = time.perf_counter()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2
for b in range(8):
    frame(b)                                 # warm-up (CUDA context, planner handle)
t = np.array([frame(b) for b in range(8, 64)]) * 1e3
print("drop-in, one frame per call, default 12x6 lattice, K=3 (ms, median / p95):")
for name, col in (("DP_algorithm", 0), ("cal_lmin_lmax", 1), ("Quadratic_planning", 2)):
    print("  %-20s %.3f / %.3f" % (name, np.median(t[:, col]), np.percentile(t[:, col], 95)))
print("  %-20s %.3f / %.3f   (reference on one CPU core: DP alone 76-118 ms, SURVEY.md §6)" % (
    "S-L step total", np.median(t.sum(1)), np.percentile(t.sum(1), 95)))
