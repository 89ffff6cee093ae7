"""CPU tests of the host-side logic: C-ABI library loads and exports every declared symbol (no compute
calls without a GPU), record packing, scene generation, the kernel's arithmetic model vs the oracle."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def test_library_loads_and_exports_every_declared_symbol():
    from emplanner_carla_b200 import _lib
    lib = _lib.lib()                       # must not touch CUDA: works on a box without a GPU
    hdr = open(os.path.join(ROOT, "include", "emplanner.h")).read()
    declared = set(re.findall(r"\b(emp_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.emp_version() == 100
    assert lib.emp_scene_stride(8) == 22 and lib.emp_plan_stride(61) == 224


def test_param_defaults_match_reference_keyword_defaults():
    from emplanner_carla_b200 import _lib
    lib = _lib.lib()
    d, q = _lib.DpParams(), _lib.QpParams()
    lib.emp_dp_params_default(ctypes.byref(d))
    lib.emp_qp_params_default(ctypes.byref(q))
    py_d, py_q = _lib.dp_params(), _lib.qp_params()
    for f, _t in _lib.DpParams._fields_:
        a, b = getattr(d, f), getattr(py_d, f)
        assert (list(a) == list(b)) if f == "w_smooth" else (a == b), f
    for f, _t in _lib.QpParams._fields_:
        assert getattr(q, f) == getattr(py_q, f), f
    assert (d.row, d.col, d.sample_s, d.sample_l, d.w_collision) == (12, 6, 15, 1.5, 1e12)   # path_planning.py:276-279
    assert lib.emp_max_enriched(ctypes.byref(d)) == 49


def test_argument_errors_are_reported_not_crashed():
    from emplanner_carla_b200 import _lib
    lib = _lib.lib()
    d = _lib.dp_params(row=40)
    rc = lib.emp_sl_dp_dev(None, 8, 4, ctypes.byref(d), 49, None, None, None, None)
    assert rc < 0 and b"row" in lib.emp_last_error()
    with pytest.raises(_lib.EmpError):
        _lib.check(rc)
    # weights the exact branch and bound cannot take are argument errors, not silently different answers
    d = _lib.dp_params(w_collision_cost=-1.0)
    assert lib.emp_sl_dp_dev(None, 8, 4, ctypes.byref(d), 49, None, None, None, None) < 0
    assert b"w_collision" in lib.emp_last_error()
    st = _lib.st_params(w_cost_obs=0.0)
    assert lib.emp_st_dp_dev(None, 8, 4, ctypes.byref(st), None, None, None, None, None) < 0
    assert b"w_cost_obs" in lib.emp_last_error()


def test_pack_scenes_layout_and_scene_prefix_property():
    from emplanner_carla_b200 import api, scenes
    a = scenes.sl_scenes(64, scenes.LATTICE_C, 2)
    b = scenes.sl_scenes(16, scenes.LATTICE_C, 2)
    for k in a:
        assert np.array_equal(a[k][:16], b[k], equal_nan=True), k
    rec = api.pack_scenes(**b)
    assert rec.shape == (16, 22)
    assert np.array_equal(rec[:, 0], b["s0"]) and np.array_equal(rec[:, 4], b["nobs"])
    k = int(b["nobs"][3])
    assert np.array_equal(rec[3, 6:6 + k], b["obs_s"][3, :k]) and np.array_equal(rec[3, 14:14 + k], b["obs_l"][3, :k])
    assert np.all(np.isfinite(rec))
    # obstacle generator keeps cal_lmin_lmax inside the path
    assert np.nanmax(a["obs_s"]) < scenes.LATTICE_C.horizon - 5


@pytest.mark.parametrize("tag", ["A", "C"])
def test_kernel_arithmetic_model_matches_oracle_qp(tag):
    """The algorithm the CUDA QP kernel implements (tests/qp_kernel_model.py) reproduces the oracle's
    coneqp iterates: same status, same iteration count, l within 1e-8 m."""
    from qp_kernel_model import solve_qp
    g = load_golden("sl_%s.npz" % tag)
    worst, worst_unknown, n_checked, n_unknown = 0.0, 0.0, 0, 0
    for b in range(0, len(g["s0"]), max(4, len(g["s0"]) // 64)):
        if g["cor_index_error"][b]:
            continue
        n = (g["ref_n"][b] + 1) // 2
        start = [g["l0"][b], g["dl0"][b], g["ddl0"][b]]
        # with the documented divergence exit EVERY frame is comparable: status, iteration count and iterate
        x, status, it = solve_qp(g["ref_lmin"][b, :n], g["ref_lmax"][b, :n], start,
                                 diverge_factor=float(g["diverge_factor"]))
        assert (status == "optimal") == bool(g["qpd_optimal"][b])
        assert it == g["qpd_iters"][b]
        err = np.abs(x[:, 0] - g["qpd_l"][b, :n]).max()
        if status == "optimal":
            worst, n_checked = max(worst, err), n_checked + 1
        else:
            worst_unknown, n_unknown = max(worst_unknown, err), n_unknown + 1
            # cvxopt's rule on the same frame: both sides end 'unknown' (the iterates are no longer comparable)
            _, status0, _ = solve_qp(g["ref_lmin"][b, :n], g["ref_lmax"][b, :n], start)
            assert status0 == "unknown" and not g["qp_optimal"][b]
    assert n_checked >= 8 and worst < 1e-8, worst
    assert n_unknown >= 2 and worst_unknown < 1e-5, worst_unknown


def test_wire_format_round_trip_and_errors():
    """Packed Pipe messages (reference test_9.py:390-392, :220): lossless round trip incl. empty lists and the
    route-less form, the reference's tuple shape back out, truncation / wrong magic rejected."""
    from emplanner_carla_b200 import wire
    rng = np.random.default_rng(5)
    route = rng.normal(size=(300, 4))
    static = [(1.0, 2.0, 11.5), (3.5, -4.25, 30.0)]
    dyn = [(9.0, 8.0, 20.0, 3.5)]
    msg = wire.pack_request(static, dyn, (0.1, 0.2), (0.3, 0.4), (5.0, 0.5), (0.01, -0.02), route, [17], route_id=7)
    assert len(msg) == 32 + 8 * (2 * 3 + 4 + 8 + 300 * 4) + 8
    r = wire.unpack_request(msg)
    assert r["route_id"] == 7 and np.array_equal(r["global_frenet_path"], route)
    tup = wire.as_reference_tuple(r)
    assert tup[0] == static and tup[1] == dyn and tup[2] == (0.1, 0.2) and tup[5] == (0.01, -0.02) and tup[7] == [17]
    assert len(tup[6]) == 300 and tup[6][4] == tuple(route[4])
    lean = wire.pack_request([], [], (0, 0), (1, 1), (2, 2), (3, 3), None, [], route_id=7)
    assert len(lean) == 32 + 64
    r2 = wire.unpack_request(lean)
    assert r2["global_frenet_path"] is None and len(r2["static_obs_info"]) == 0
    assert len(wire.as_reference_tuple(r2, route=route)[6]) == 300
    with pytest.raises(wire.WireError):
        wire.as_reference_tuple(r2)
    with pytest.raises(wire.WireError):
        wire.unpack_request(msg[:-16])
    with pytest.raises(wire.WireError):
        wire.unpack_reply(msg)
    path = rng.normal(size=(23, 4))
    rep = wire.unpack_reply(wire.pack_reply(path, [17, 3], np.arange(26.0), -np.arange(26.0)))
    assert np.array_equal(rep["path"], path) and list(rep["match_point_list"]) == [17, 3] and rep["path_l"][5] == -5.0
    xy, n = wire.stack_obstacles([r, r2], kmax=4)
    assert xy.shape == (2, 4, 2) and list(n) == [2, 0] and xy[0, 1, 1] == -4.25 and np.isnan(xy[1]).all()
    assert wire.stack_obstacles([r, r2])[0].shape == (2, 8, 2)
    with pytest.raises(wire.WireError):                       # never drop an obstacle silently
        wire.stack_obstacles([r, r2], kmax=1)


def test_speed_qp_kernel_linear_algebra_model_reproduces_the_oracle_iterates():
    """csrc/st_qp.cu's structured KKT solve (three tridiagonal chains + dense K), modelled in numpy and plugged into
    the oracle's coneqp iteration, gives the oracle's own status, iteration count and solution on golden frames."""
    from oracle import coneqp, st_back
    from speed_qp_model import ChainKkt
    g = load_golden("st_back.npz")
    frames = np.nonzero(g["matrices_identical"] == 1)[0][:10]
    assert len(frames) >= 8
    for f in frames:
        m = st_back.speed_qp_matrices(g["v0"][f], g["a0"][f], g["dp_s"][f], g["dp_t"][f], *g["ref_bounds"][f])
        P, q, G, h, A, b = st_back.speed_qp_cone_form(m)
        want = coneqp.qp(P, q, G, h, A, b)
        got = coneqp.qp(P, q, G, h, A, b, kkt=ChainKkt(P, G, A))
        assert got["status"] == want["status"]
        if want["status"] == "optimal":
            assert got["iterations"] == want["iterations"]
            assert np.abs(got["x"] - want["x"]).max() <= 1e-7 * max(1.0, np.abs(want["x"]).max())
