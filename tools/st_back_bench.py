"""Per-kernel CUDA-event timings of the S-T half: speed DP -> convex space -> speed QP -> increase_points -> merge.
usage: python tools/st_back_bench.py [batch] [reps]"""
import sys, numpy as np
sys.path.insert(0, '.')
import torch
from emplanner_carla_b200 import api, scenes, _lib
import ctypes as C
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sc = scenes.st_scenes(B, 5, k_fixed=None)
rec = torch.from_numpy(api.pack_st_scenes(sc["s_in"], sc["s_out"], sc["t_in"], sc["t_out"], sc["v0"])).cuda()
xs = np.arange(0, 96, 1.0)
f = np.arange(B)[:, None]
ys = (0.001 + 0.0005 * (f % 7)) * xs[None] ** 2
s_cum = np.concatenate([np.zeros((B, 1)), np.cumsum(np.hypot(np.diff(xs)[None], np.diff(ys, axis=1)), axis=1)], axis=1)
kap = 2 * (0.001 + 0.0005 * (f % 7)) / (1 + (2 * (0.001 + 0.0005 * (f % 7)) * xs[None]) ** 2) ** 1.5
nanpad = np.full((B, 4), np.nan)
path = torch.from_numpy(np.ascontiguousarray(np.stack([np.concatenate([s_cum, 0 * nanpad], 1), np.concatenate([kap, 0 * nanpad + 0], 1)], 1))).cuda()
path5 = torch.from_numpy(np.ascontiguousarray(np.stack([np.concatenate([s_cum, nanpad], 1), np.concatenate([np.tile(xs, (B, 1)), nanpad], 1),
                                   np.concatenate([ys, nanpad], 1), np.concatenate([kap * 0, nanpad], 1),
                                   np.concatenate([kap, nanpad], 1)], 1))).cuda()
path = torch.nan_to_num(path, nan=0.0)
start = torch.from_numpy(np.stack([sc["v0"], np.zeros(B)], 1)).cuda()
now = torch.zeros(B, dtype=torch.float64, device="cuda")
st, sq = api.st_params(), api.speed_qp_params()
out = api.st_dp_device(rec, st)
r = api.st_back_device(out, rec, path, start, sq, now, path5)
torch.cuda.synchronize()
L, sp = _lib.lib(), api._stream_ptr()
p = api._dev_ptr
steps = {
    "st_dp": lambda: api.st_dp_device(rec, st, out),
    "convex_space": lambda: L.emp_st_convex_space_dev(p(out), p(rec), 8, p(path), path.shape[2], C.byref(sq), B, p(r["bounds"]), p(r["convex_status"]), sp),
    "speed_qp": lambda: L.emp_st_speed_qp_dev(p(start), p(out), p(r["bounds"]), B, C.byref(sq), p(r["qp"]), p(r["qp_status_iters"]), sp),
    "increase_points": lambda: L.emp_st_increase_points_dev(p(r["qp"]), B, p(r["dense"]), p(r["dense_status"]), sp),
    "path_speed_merge": lambda: L.emp_st_path_speed_merge_dev(p(r["dense"]), p(now), p(path5), path5.shape[2], B, p(r["traj"]), p(r["merge_status"]), sp),
}
tot = 0.0
for name, fn in steps.items():
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tot += ms
    print("%-18s %8.3f ms / %d frames  (%.0f frames/s)" % (name, ms, B, B / ms * 1e3))
si = r["qp_status_iters"].cpu().numpy()
ran = si[:, 0] < 0x10
print("total %.3f ms; QP ran on %d/%d frames (rest: full 16-column profile -> IndexError in the reference), optimal %d, mean iters %.1f"
      % (tot, ran.sum(), B, (si[:, 0] == 0).sum(), si[ran, 1].mean() if ran.any() else 0))
print("convex status counts:", np.bincount(r["convex_status"].cpu().numpy() // 16))
