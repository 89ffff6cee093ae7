"""The drop-in boundary, proven with the reference's own caller.

CPU tests (need /root/reference, skipped elsewhere): the header imports of the driver script (test_9.py:24-26)
resolve under the INTEGRATION.md recipe; the restated caller equals the reference's function text (AST call sequence
and bit-identical payloads on the reference planner); the committed golden was made from the current text.
GPU test: the planning process of test_9.py (Pipe + forked child, CUDA initialised after the fork) replayed through
the drop-in and compared with the payloads the UNMODIFIED reference produced for the same messages
(tests/golden/motion_planning.npz, oracle/make_golden.py::make_motion_planning).
Tolerances: path_s length and values exact-length / 1e-9; l and x, y <= 1e-4 m (north_star); match index exact.
"""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden
import motion_planning_caller as mpc

REF = os.environ.get("EMP_REFERENCE_ROOT", "/root/reference")
needs_reference = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present (build container only)")

_STUBS = r'''
import sys, types
class _Any(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})
for _m in ("carla", "cvxopt", "cvxopt.solvers", "networkx"):
    sys.modules.setdefault(_m, _Any(_m))
'''


@needs_reference
def test_driver_header_imports_resolve_under_the_integration_recipe():
    """test_9.py:24-26 executed verbatim with PYTHONPATH = <repo>/emplanner_carla_b200:<reference>:
    planning_utils / path_planning come from the drop-in, global_planning and the controller from the reference."""
    src = open(os.path.join(REF, "test_9.py"), encoding="utf-8").read().splitlines()
    k = next(i for i, ln in enumerate(src) if ln.startswith("from planner import"))
    lines = src[k:k + 3]                                      # test_9.py:24-26
    assert "controller.controller" in lines[1] and "global_planning" in lines[2]
    code = _STUBS + "\n".join(lines) + r'''
import planner, planner.global_planning as gp
print(planning_utils.__file__); print(path_planning.__file__); print(gp.__file__)
for name in ("Vector_fun", "waypoint_list_2_target_path", "predict_block", "find_match_points", "sampling",
             "smooth_reference_line", "cal_s_map_fun", "cal_s_l_fun", "cal_s_l_deri_fun"):
    assert callable(getattr(planning_utils, name)), name
assert gp.planning_utils is planning_utils
'''
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "emplanner_carla_b200"), REF]))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=120)
    assert r.returncode == 0, r.stderr
    pu_file, pp_file, gp_file = r.stdout.split()[:3]
    assert pu_file.startswith(ROOT) and pp_file.startswith(ROOT) and gp_file.startswith(REF)


@needs_reference
def test_golden_was_made_from_the_current_reference_text():
    g = load_golden("motion_planning.npz")
    assert hashlib.sha256(mpc.reference_function_text().encode()).hexdigest() == str(g["caller_sha256"])


@needs_reference
def test_restated_caller_is_equivalent_to_the_reference_text():
    from oracle import refshim
    ref_text, my_text = mpc.reference_function_text(), mpc.restated_function_text()
    assert mpc.planner_call_sequence(my_text) == mpc.planner_call_sequence(ref_text)
    pp, pu, _ = refshim.modules()
    g = load_golden("motion_planning.npz")
    msgs = mpc.messages_from_golden(g)[:12]
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        a = mpc.replay(ref_text, pu, pp, msgs)
        b = mpc.replay(my_text, pu, pp, msgs)
    assert len(a) == len(b) == len(msgs)
    for f, (x, y) in enumerate(zip(a, b)):
        assert x[1] == y[1] and len(x[0]) == len(y[0])
        assert np.array_equal(np.array(x[0]), np.array(y[0])), "Cartesian payload differs at frame %d" % f
        assert np.array_equal(np.array(x[2]), np.array(y[2])) and np.array_equal(np.array(x[3]).reshape(-1),
                                                                                 np.array(y[3]).reshape(-1))
        n = int(g["ref_cart_n"][f])
        assert np.array_equal(np.array(x[0]), g["ref_cart"][f, :n]), "golden payload differs at frame %d" % f


@needs_reference
def test_actor_helpers_match_the_reference():
    """Vector_fun / predict_block (planning_utils.py:14, :591) on duck-typed actors: identical to the reference."""
    from oracle import refshim
    from emplanner_carla_b200.planner import planning_utils as mine
    _, pu, _ = refshim.modules()
    V = lambda x, y, z: type("V", (), dict(x=x, y=y, z=z))()

    class Car:
        def __init__(self, rng):
            self.loc, self.vel, self.w = V(*rng.uniform(-50, 50, 3)), V(*rng.uniform(-12, 12, 3)), V(*rng.uniform(-30, 30, 3))
            self.yaw = rng.uniform(-180, 180)
        get_location = lambda s: s.loc
        get_velocity = lambda s: s.vel
        get_angular_velocity = lambda s: s.w
        get_transform = lambda s: type("T", (), dict(rotation=type("R", (), dict(yaw=s.yaw))()))()
    rng = np.random.default_rng(3)
    for _ in range(50):
        car = Car(rng)
        assert mine.predict_block(car, ts=0.2) == pu.predict_block(car, ts=0.2)
        a, b = V(*rng.uniform(-9, 9, 3)), V(*rng.uniform(-9, 9, 3))
        assert np.array_equal(mine.Vector_fun(a, b), pu.Vector_fun(a, b))
    z = V(1.0, 2.0, 3.0)
    assert np.array_equal(mine.Vector_fun(z, z), pu.Vector_fun(z, z))


@pytest.mark.gpu
def test_waypoint_list_2_target_path_on_the_gpu():
    """planning_utils.py:29-46 through the drop-in: heading / curvature of the 400-node route (any length)."""
    from emplanner_carla_b200.planner import planning_utils as mine
    g = load_golden("motion_planning.npz")
    W = lambda x, y: (type("W", (), dict(transform=type("T", (), dict(location=type("L", (), dict(x=x, y=y))()))()))(), None)
    path = mine.waypoint_list_2_target_path([W(float(r[0]), float(r[1])) for r in g["route"]])
    got = np.array(path)
    assert got.shape == g["route"].shape and np.array_equal(got[:, :2], g["route"][:, :2])
    assert np.abs(got[:, 2] - g["route"][:, 2]).max() <= 1e-14 and np.abs(got[:, 3] - g["route"][:, 3]).max() <= 1e-13


# Per-call tolerances of the forced replay (every call on exactly the inputs the reference saw):
#   list slicing and integer results: exact; Frenet projections 1e-9; DP s exact, l 1e-9; corridor exact;
#   QP solves (smooth_reference_line x/y, Quadratic_planning l/dl/ddl, the trailing smoother of
#   frenet_2_x_y_theta_kappa) 1e-4 m (north_star), observed ~1e-9; theta 1e-6 rad, kappa 1e-5 1/m of a smoothed line.
FORCED_TOL = {"find_match_points": 1e-9, "sampling": 0.0, "smooth_reference_line": 1e-4, "cal_s_map_fun": 1e-9,
              "cal_s_l_fun": 1e-9, "cal_s_l_deri_fun": 1e-8, "DP_algorithm": 1e-9, "cal_lmin_lmax": 0.0,
              "Quadratic_planning": 1e-4, "frenet_2_x_y_theta_kappa": 1e-4}


@pytest.mark.gpu
def test_planning_process_replay_through_the_dropin(tmp_path):
    """test_9.py's planning process (Pipe + forked child, CUDA initialised after the fork) on the drop-in.
    Forced replay: every planner call of every frame against the reference's recorded answer (see
    motion_planning_caller.ForcingProxy) - this is the parity assertion.  Free replay: payloads compared frame by
    frame where no discrete event of the reference's own (Q10 length flip, lattice near-tie) separates the runs; the
    number of such frames is recorded next to the reference's own rate under a 1e-9 m input move."""
    import json
    g = load_golden("motion_planning.npz")
    out_path, report_path = str(tmp_path / "replay.npz"), str(tmp_path / "forced.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "replay_test9.py"),
                        os.path.join(ROOT, "tests", "golden", "motion_planning.npz"), out_path, report_path],
                       capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    o = np.load(out_path)
    assert int(o["child_pid"][0]) != int(o["parent_pid"][0])
    F = len(g["ego"])
    # ---- forced replay: the parity assertion ------------------------------------------------------------------
    rep = json.load(open(report_path))
    assert rep["calls_done"] == int(o["recorded_calls"][0]), "the forced replay stopped early"
    per = rep["per_function"]
    assert set(per) == set(FORCED_TOL)
    for name, tol in FORCED_TOL.items():
        assert per[name]["errors"] == [] and per[name]["shape_mismatch"] == 0, (name, per[name])
        assert per[name]["worst"] <= tol, (name, per[name])
    assert per["smooth_reference_line"]["worst_theta"] <= 1e-6 and per["smooth_reference_line"]["worst_kappa"] <= 1e-5
    for f in range(F):                                         # forced: the payloads ARE the reference's
        m = int(g["refd_cart_n"][f])
        assert int(o["forced_cart_n"][f]) == m and np.array_equal(o["forced_cart"][f, :m], g["refd_cart"][f, :m])
    # ---- free replay -------------------------------------------------------------------------------------------
    same, flipped, worst = 0, [], dict(path_s=0.0, l=0.0, xy=0.0)
    for f in range(F):
        n, m = int(g["refd_path_n"][f]), int(g["refd_cart_n"][f])
        if int(o["path_n"][f]) != n or int(o["cart_n"][f]) != m or int(o["match"][f]) != int(g["refd_match"][f]):
            flipped.append(f)
            continue
        dl = float(np.abs(o["path_l"][f, :n] - g["refd_path_l"][f, :n]).max())
        ds = float(np.abs(o["path_s"][f, :n] - g["refd_path_s"][f, :n]).max())
        if dl > 1e-3 or ds > 1e-3:                             # other stations (Q10) or another lattice path: discrete
            flipped.append(f)
            continue
        same += 1
        worst["path_s"] = max(worst["path_s"], float(np.abs(o["path_s"][f, :n] - g["refd_path_s"][f, :n]).max()))
        worst["l"] = max(worst["l"], dl)
        worst["xy"] = max(worst["xy"], float(np.abs(o["cart"][f, :m, :2] - g["refd_cart"][f, :m, :2]).max()))
    ref_self = int((g["refd_moved_path_n"] != g["refd_path_n"]).sum())
    summary = dict(caller=str(o["source"]), frames=F, forced_per_function=per, free_frames_compared=same,
                   free_frames_with_discrete_event=flipped, free_worst=worst,
                   reference_vs_itself_inputs_moved_1e_9_m_length_flips=ref_self)
    print("replay:", json.dumps(summary)[:3000])
    try:
        d = os.path.join(ROOT, "gpurun_out", "parity")
        os.makedirs(d, exist_ok=True)
        json.dump(summary, open(os.path.join(d, "motion_planning_replay.json"), "w"), indent=1)
    except OSError:
        pass
    assert same >= F // 4, "too few comparable frames in the free replay"
    assert worst["path_s"] <= 1e-8 and worst["l"] <= 1e-4 and worst["xy"] <= 1e-4
