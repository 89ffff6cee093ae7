"""Small pass over every kernel through the host-buffer C ABI (ctypes + numpy only, no torch) - the workload for
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_run.py
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from emplanner_carla_b200 import api, scenes

g = lambda n: np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", n))
pl = api.HostPlanner(max_batch=64, n_enr_max=128, max_col=64)
for lat, B in ((scenes.LATTICE_A, 24), (scenes.LATTICE_C, 8)):
    dp, qp = api.dp_params(**lat.kwargs()), api.qp_params()
    view, rows = pl.plan(api.pack_scenes(**scenes.sl_scenes(B, lat, 90)), dp, qp, want_rows=True)
    print("S-L", lat.name, "qp ok", int((view.qp_status == 0).sum()), "/", B)
# general-size paths: 46 QP stations (64-thread CTA), 100 stations (128 threads), 100-point smoother (64 threads),
# any-length cal_heading_kappa; the schedule kernel runs inside every plan() above
lat1 = scenes.Lattice("A_res1", 12, 6, 15, 1.5, 1)
dp1 = api.dp_params(**lat1.kwargs())
pl1 = api.HostPlanner(max_batch=8, n_enr_max=api.max_enriched(dp1), max_col=6)
v1, _ = pl1.plan(api.pack_scenes(**scenes.sl_scenes(8, lat1, 91)), dp1, api.qp_params())
print("S-L 46 stations: qp ok", int((v1.qp_status == 0).sum()), "/ 8")
x, st, it = pl1.qp(np.full((4, 100), -10.0), np.full((4, 100), 10.0), np.zeros((4, 3)), api.qp_params())
print("QP 100 stations status", st.tolist())
pl1.close()
xy = scenes.reference_line(100, 2.0)[None] + 0.01
o, st, it = pl.smooth_lines(xy, api.smooth_params())
hk, _, _ = pl.smooth_lines(scenes.reference_line(400, 2.0)[None], api.smooth_params(heading_only=True))
print("smoother 100 points status", st.tolist(), "heading of 400 points", hk.shape)
gs = g("st_back.npz")
n = 12
sc = api.pack_st_scenes(gs["s_in"][:n], gs["s_out"][:n], gs["t_in"][:n], gs["t_out"][:n], gs["v0"][:n])
r = pl.st_dp(sc, api.st_params(), want_tables=True)
st_out = np.full((n, api.ST_OUT_STRIDE), np.nan)
st_out[:, 8:24], st_out[:, 24:40] = r["dp_s"], r["dp_t"]
for r3 in (False, True):
    pl.st_back(api.speed_qp_params(repair_full_profile=r3), st_out=st_out, scenes=sc, path=gs["path"][:n],
               start=np.stack([gs["v0"][:n], gs["a0"][:n]], 1), want=("bounds", "qp"))
b = pl.st_back(api.speed_qp_params(), st_out=st_out, scenes=sc, path=gs["path"][:n], start=np.stack([gs["v0"][:n], gs["a0"][:n]], 1),
               current_time=gs["current_time"][:n], path5=gs["path5"][:n], want=("bounds", "qp", "dense", "traj"))
print("S-T: qp status", b["qp_status"].tolist())
gf = g("frenet.npz")
out, status, iters = pl.smooth_lines(gf["local"][:4, :, :2], api.smooth_params())
print("smooth status", status.tolist())
fr = pl.frenet_project(gf["line"][:4], origin=gf["line"][:4, 10, :2], points=gf["line"][:4, 5:9, :2] + 0.3,
                       want=("s_map", "match", "sl", "proj"))
print("frenet s_map[0,:3]", fr["s_map"][0, :3])
print("quintic", pl.quintic([[0.1, 0.0, 0.0, 1.5, 0, 0, 2.0, 17.0]])[0, :3])
pl.close()
print("done")
