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


@pytest.mark.gpu
def test_planning_process_replay_through_the_dropin(tmp_path):
    g = load_golden("motion_planning.npz")
    out_path = str(tmp_path / "replay.npz")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "replay_test9.py"),
                        os.path.join(ROOT, "tests", "golden", "motion_planning.npz"), out_path],
                       capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    o = np.load(out_path)
    assert int(o["child_pid"][0]) != int(o["parent_pid"][0])
    F = len(g["ego"])
    worst = dict(path_s=0.0, l_opt=0.0, xy_opt=0.0, l_unknown=0.0, xy_unknown=0.0)
    for f in range(F):
        optimal = g["refd_qp_status"][f, 1] == 1
        assert o["match"][f] == g["refd_match"][f]
        n = int(g["refd_path_n"][f])
        assert int(o["path_n"][f]) == n, "frame %d: path length %d != %d" % (f, o["path_n"][f], n)
        worst["path_s"] = max(worst["path_s"], np.abs(o["path_s"][f, :n] - g["refd_path_s"][f, :n]).max())
        m = int(g["refd_cart_n"][f])
        assert int(o["cart_n"][f]) == m, "frame %d: %d Cartesian points != %d" % (f, o["cart_n"][f], m)
        dl = np.abs(o["path_l"][f, :n] - g["refd_path_l"][f, :n]).max()
        dxy = np.abs(o["cart"][f, :m, :2] - g["refd_cart"][f, :m, :2]).max()
        k = "opt" if optimal else "unknown"
        worst["l_" + k] = max(worst["l_" + k], dl)
        worst["xy_" + k] = max(worst["xy_" + k], dxy)
        if optimal:                                            # cvxopt-rule golden == exit-rule golden on these frames
            assert np.array_equal(g["ref_cart"][f], g["refd_cart"][f], equal_nan=True)
    print("replay (%s): %d frames, worst" % (str(o["source"]), F), {k: "%.2e" % v for k, v in worst.items()})
    assert worst["path_s"] <= 1e-9
    assert worst["l_opt"] <= 1e-4 and worst["xy_opt"] <= 1e-4
    assert worst["l_unknown"] <= 1e-4 and worst["xy_unknown"] <= 1e-4
