"""GPU parity tests for the Cartesian<->Frenet kernels (csrc/frenet.cu) through the C ABI, against golden
outputs of the unmodified reference (tests/golden/frenet.npz) and the oracle.  Tolerance: 1e-9 absolute
(SURVEY.md §8(c)); indices exact."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def planner():
    from emplanner_carla_b200 import api
    pl = api.HostPlanner(max_batch=64)
    yield pl
    pl.close()


def test_batched_projection_against_reference_golden(planner):
    g = load_golden("frenet.npz")
    F = len(g["ego"])
    r = planner.frenet_project(g["line"], origin=g["ego"], points=g["obs_xy"], want=("s_map", "match", "sl"))
    assert np.abs(r["s_map"] - g["smap"]).max() < TOL
    assert np.abs(r["sl"][:, :, 0] - g["obs_s"]).max() < TOL and np.abs(r["sl"][:, :, 1] - g["obs_l"]).max() < TOL
    # plan-start derivatives: cal_s_l_deri_fun([ego], [v], [a], line, ego)
    d = planner.frenet_project(g["line"], origin=g["ego"], points=g["ego"][:, None, :], vel=g["vel"][:, None, :],
                               acc=g["acc"][:, None, :], want=("deriv",))
    assert np.abs(d["deriv"][:, 0, :] - g["deriv"]).max() < TOL
    # find_match_points(is_first_run=False) on the long global path, one scene per golden frame
    gp = np.broadcast_to(g["global_path"], (F,) + g["global_path"].shape)
    m = planner.frenet_project(gp, points=g["ego"][:, None, :], pre_match=g["pre_idx"], want=("match",))
    assert np.array_equal(m["match"][:, 0], g["match_idx"])
    assert F >= 8


def test_frenet_to_cartesian_against_reference_golden(planner):
    g = load_golden("frenet.npz")
    cart, ncart = planner.frenet_to_cart(g["line"], g["smap"], g["start_sl"], g["path_s"], g["path_l"])
    for f in range(len(ncart)):
        ref = g["cart_raw"][f]
        n_ref = int(np.sum(~np.isnan(ref[:, 0])))
        assert ncart[f] == n_ref                                   # truncation at s_map[-1] (Q14)
        assert np.abs(cart[f, 0, :2] - ref[0, :2]).max() < TOL
        assert np.abs(cart[f, 1:n_ref] - ref[1:n_ref]).max() < TOL
    assert (ncart < g["path_s"].shape[1] + 1).any(), "golden set must exercise the truncation"


def test_quirk_q12_first_point_index_and_ragged_lists(planner):
    """Obstacle #i>0 is projected through obstacle #0's match node (planning_utils.py:413); ragged point
    counts; empty lists."""
    from oracle import frenet as ofr
    g = load_golden("frenet.npz")
    line = [tuple(r) for r in g["line"][0]]
    smap = list(g["smap"][0])
    pts = [tuple(p) for p in g["obs_xy"][0]]
    s_o, l_o = ofr.s_l(pts, line, smap)
    s_rev, l_rev = ofr.s_l(pts[::-1], line, smap)
    assert abs(l_o[1] - l_rev[2]) > 1e-6, "scene must make the quirk observable"
    npts = np.array([4, 2, 0], dtype=np.int32)
    r = planner.frenet_project(g["line"][:3], points=g["obs_xy"][:3], npts=npts, s_map_in=g["smap"][:3], want=("sl",))
    assert np.abs(r["sl"][0, :, 0] - s_o).max() < TOL and np.abs(r["sl"][0, :, 1] - l_o).max() < TOL
    s1, l1 = ofr.s_l([tuple(p) for p in g["obs_xy"][1, :2]], [tuple(v) for v in g["line"][1]], list(g["smap"][1]))
    assert np.abs(r["sl"][1, :2, 0] - s1).max() < TOL and np.abs(r["sl"][1, :2, 1] - l1).max() < TOL
    assert np.all(r["sl"][1, 2:] == 0) and np.all(r["sl"][2] == 0)


def test_dropin_planning_utils_sequence(planner):
    """The reference's call sequence test_9.py:99-134,172,213 with the reference's function names."""
    pkg_dir = os.path.join(ROOT, "emplanner_carla_b200")
    sys.path.insert(0, pkg_dir)
    for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
        del sys.modules[k]
    try:
        from planner import path_planning, planning_utils
        g = load_golden("frenet.npz")
        gp = [tuple(r) for r in g["global_path"]]
        for f in range(3):
            ego = tuple(g["ego"][f])
            mi, _ = planning_utils.find_match_points(xy_list=[ego], frenet_path_node_list=gp, is_first_run=False,
                                                     pre_match_index=int(g["pre_idx"][f]))
            assert mi[0] == g["match_idx"][f]
            local = planning_utils.sampling(mi[0], gp, back_length=10, forward_length=50)
            assert np.array_equal(np.array(local), g["local"][f])
            line = [tuple(r) for r in g["line"][f]]          # smooth_reference_line output of the reference
            s_map = planning_utils.cal_s_map_fun(line, origin_xy=ego)
            assert isinstance(s_map, list) and np.abs(np.array(s_map) - g["smap"][f]).max() < TOL
            obs_s, obs_l = planning_utils.cal_s_l_fun([tuple(p) for p in g["obs_xy"][f]], line, s_map)
            assert np.abs(np.array(obs_s) - g["obs_s"][f]).max() < TOL
            assert np.abs(np.array(obs_l) - g["obs_l"][f]).max() < TOL
            der = planning_utils.cal_s_l_deri_fun(xy_list=[ego], V_xy_list=[tuple(g["vel"][f])],
                                                  a_xy_list=[tuple(g["acc"][f])], local_path_xy_opt=line, origin_xy=ego)
            assert len(der) == 7 and np.abs(np.array([d[0] for d in der]) - g["deriv"][f]).max() < TOL
            idx, proj = planning_utils.match_projection_points([ego], line)
            raw = path_planning.frenet_2_x_y_theta_kappa(g["start_sl"][f, 0], g["start_sl"][f, 1], list(g["path_s"][f]),
                                                         list(g["path_l"][f]), line, s_map, smooth=False)
            n_ref = int(np.sum(~np.isnan(g["cart_raw"][f][:, 0])))
            assert len(raw) == n_ref and np.abs(np.array(raw[1:]) - g["cart_raw"][f, 1:n_ref]).max() < TOL
            x, y, th, kap, m = path_planning.cal_proj_point(g["path_s"][f, 3], 0, line, s_map)
            assert abs(th - g["cart_raw"][f, 4, 2]) < TOL and kap == g["cart_raw"][f, 4, 3]
    finally:
        sys.path.remove(pkg_dir)
        for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
            del sys.modules[k]


def test_smooth_reference_line_against_reference_golden(planner):
    """QP#1 (51-node local line -> smoothed line with heading/curvature) and QP#3 (Cartesian path smoothing) against
    the reference's own outputs (reference matrix construction + restated coneqp).  Tolerance 1e-6 m on x, y;
    1e-6 rad on theta; 1e-6 1/m on kappa (the QP stops at cvxopt's default tolerances, the kernel's linear algebra
    differs from the dense oracle's only by rounding)."""
    from emplanner_carla_b200 import api
    g = load_golden("frenet.npz")
    out, status, iters = planner.smooth_lines(g["local"][:, :, :2], api.smooth_params())
    assert np.all(status == 0)
    assert np.abs(out[:, :, :2] - g["line"][:, :, :2]).max() < 1e-6
    assert np.abs(out[:, :, 2] - g["line"][:, :, 2]).max() < 1e-6
    assert np.abs(out[:, :, 3] - g["line"][:, :, 3]).max() < 1e-6
    assert iters.min() >= 2 and iters.max() <= 8
    # QP#3: ragged lengths (the Frenet->Cartesian output is truncated at s_map[-1])
    F = len(g["cart_n"])
    M = g["cart_raw"].shape[1]
    xy = np.where(np.isnan(g["cart_raw"][:, :, :2]), 0.0, g["cart_raw"][:, :, :2])
    out3, status3, _ = planner.smooth_lines(xy, api.smooth_params(), npts=g["cart_n"])
    for f in range(F):
        n = g["cart_n"][f]
        assert status3[f] == 0
        assert np.abs(out3[f, :n] - g["cart"][f, :n]).max() < 1e-6
    # cal_heading_kappa alone
    hk, _, _ = planner.smooth_lines(g["line"][:, :, :2], api.smooth_params(heading_only=True))
    assert np.abs(hk[:, :, 2:] - g["line"][:, :, 2:]).max() < 1e-9


def test_full_frame_pipeline_dropin(planner):
    """test_9.motion_planning steps 1-14 with the drop-in modules only (no reference, no cvxopt)."""
    pkg_dir = os.path.join(ROOT, "emplanner_carla_b200")
    sys.path.insert(0, pkg_dir)
    for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
        del sys.modules[k]
    try:
        from planner import path_planning, planning_utils
        g = load_golden("frenet.npz")
        gp = [tuple(r) for r in g["global_path"]]
        f = 2
        ego = tuple(g["ego"][f])
        mi, _ = planning_utils.find_match_points([ego], gp, False, int(g["pre_idx"][f]))
        local = planning_utils.sampling(mi[0], gp)
        line = planning_utils.smooth_reference_line(local)
        assert np.abs(np.array(line) - g["line"][f]).max() < 1e-6
        s_map = planning_utils.cal_s_map_fun(line, origin_xy=ego)
        obs_s, obs_l = planning_utils.cal_s_l_fun([tuple(p) for p in g["obs_xy"][f][:2]], line, s_map)
        bs, bl = planning_utils.cal_s_l_fun([ego], line, s_map)
        l_list, _, _, _, l_ds, _, l_dds = planning_utils.cal_s_l_deri_fun([ego], [tuple(g["vel"][f])], [tuple(g["acc"][f])],
                                                                          line, ego)
        dp_s, dp_l = path_planning.DP_algorithm(obs_s, obs_l, plan_start_s=bs[0], plan_start_l=l_list[0],
                                                plan_start_dl=l_ds[0], plan_start_ddl=l_dds[0])
        dp_s, dp_l = dp_s[::2], dp_l[::2]
        l_min, l_max = path_planning.cal_lmin_lmax(dp_s, dp_l, obs_s, obs_l, 5, 5)
        ql, qdl, qddl = path_planning.Quadratic_planning(l_min, l_max, l_list[0], l_ds[0], l_dds[0])
        path_s = [dp_s[0]] + [(dp_s[i] + dp_s[i - 1]) / 2 for i in range(1, len(ql))] + [dp_s[-1]]
        path_l = [ql[0]] + [(ql[i] + ql[i - 1]) / 2 for i in range(1, len(ql))] + [ql[-1]]
        cart = path_planning.frenet_2_x_y_theta_kappa(bs[0], bl[0], path_s, path_l, line, s_map)
        assert 10 < len(cart) <= len(path_s) + 1 and all(len(c) == 4 for c in cart)
        assert np.hypot(cart[0][0] - ego[0], cart[0][1] - ego[1]) < 0.5      # starts at the ego position
        # the same frame through the CPU oracle chain
        from oracle import frenet as ofr, sl_dp, sl_qp
        o_line, _ = ofr.smooth_reference_line(local)
        o_smap = ofr.s_map(o_line, ego)
        o_os, o_ol = ofr.s_l([tuple(p) for p in g["obs_xy"][f][:2]], o_line, o_smap)
        o_bs, o_bl = ofr.s_l([ego], o_line, o_smap)
        o_der = ofr.s_l_deriv([ego], [tuple(g["vel"][f])], [tuple(g["acc"][f])], o_line, ego)
        o_s, o_l, _ = sl_dp.dp_algorithm(o_os, o_ol, o_bs[0], o_der[0][0], o_der[4][0], o_der[6][0])
        o_lmin, o_lmax = sl_qp.corridor(o_s[::2], o_l[::2], o_os, o_ol, 5, 5)
        o_ql, _, _, info = sl_qp.quadratic_planning(o_lmin, o_lmax, o_der[0][0], o_der[4][0], o_der[6][0])
        o_ps, o_pl = sl_qp.midpoint_path(o_s[::2], o_ql)
        o_cart = ofr.frenet_to_cartesian(o_bs[0], o_bl[0], list(o_ps), list(o_pl), o_line, o_smap)
        assert info["status"] == "optimal" and len(o_cart) == len(cart)
        err = np.abs(np.array(cart) - np.array(o_cart))
        assert err[:, :2].max() < 1e-4 and err[:, 2].max() < 1e-4 and err[:, 3].max() < 1e-4, err.max(axis=0)
    finally:
        sys.path.remove(pkg_dir)
        for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
            del sys.modules[k]


def test_smooth_reference_line_general_sizes(planner):
    """Polylines longer than the 64 points of the one-warp path (planning_utils.py:262 takes any length): 100 and 128
    points on the 64-thread CTA path, ragged lengths, against the oracle (reference matrix construction, pinned by
    golden, + coneqp).  Same tolerances as the 51-point line; 129 points are reported as TOO_LARGE; cal_heading_kappa
    alone takes any length (the 400-node route)."""
    from emplanner_carla_b200 import api, scenes
    from oracle import frenet as ofr
    xy = scenes.reference_line(128, 2.0)
    rng = np.random.default_rng(9)
    B = 6
    lines = np.stack([xy + rng.normal(0, 0.05, xy.shape) for _ in range(B)])
    npts = np.array([128, 100, 65, 127, 97, 71], dtype=np.int32)
    out, status, iters = planner.smooth_lines(lines, api.smooth_params(), npts=npts)
    for b in range(B):
        n = int(npts[b])
        want, info = ofr.smooth_reference_line([tuple(p) for p in lines[b, :n]], diverge_factor=2.0)
        want = np.array(want)
        assert status[b] == 0 and info["status"] == "optimal" and iters[b] == info["iterations"]
        assert np.abs(out[b, :n, :2] - want[:, :2]).max() < 1e-6
        assert np.abs(out[b, :n, 2] - want[:, 2]).max() < 1e-6 and np.abs(out[b, :n, 3] - want[:, 3]).max() < 1e-6
    # the one-warp path when the polyline fills every lane (63, 64 points: the last lane is a stage of its own level)
    out64, st64, it64 = planner.smooth_lines(lines[:, :64], api.smooth_params(), npts=np.array([64, 63, 62, 64, 63, 33], np.int32))
    for b, n in enumerate([64, 63, 62, 64, 63, 33]):
        want, info = ofr.smooth_reference_line([tuple(p) for p in lines[b, :n]], diverge_factor=2.0)
        assert st64[b] == 0 and it64[b] == info["iterations"] and np.abs(out64[b, :n] - np.array(want)).max() < 1e-6
    big = np.zeros((1, 129, 2))
    big[0, :, 0] = np.arange(129)
    _, st, _ = planner.smooth_lines(big, api.smooth_params())
    assert st[0] == api.QP_TOO_LARGE
    route = scenes.reference_line(400, 2.0)[None]
    hk, st, _ = planner.smooth_lines(route, api.smooth_params(heading_only=True))
    th, kp = ofr.heading_kappa([tuple(p) for p in route[0]])
    assert st[0] == 0 and np.abs(hk[0, :, 2] - th).max() < 1e-12 and np.abs(hk[0, :, 3] - kp).max() < 1e-12
