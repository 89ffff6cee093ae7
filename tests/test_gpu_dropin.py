"""GPU tests of the drop-in `planner` package: the reference's own call sequence (test_9.py:180-201)
with the reference's function names, against golden outputs of the unmodified reference."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def planner_pkg():
    pkg_dir = os.path.join(ROOT, "emplanner_carla_b200")
    sys.path.insert(0, pkg_dir)
    for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
        del sys.modules[k]
    from planner import path_planning      # exactly the reference's import line (test_9.py:26)
    yield path_planning
    sys.path.remove(pkg_dir)
    for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
        del sys.modules[k]


def test_reference_call_sequence_lattice_a(planner_pkg):
    pp = planner_pkg
    g = load_golden("sl_A.npz")
    done = 0
    for b in range(0, 40):
        K = int(g["nobs"][b])
        obs_s, obs_l = list(g["obs_s"][b, :K]), list(g["obs_l"][b, :K])
        dp_s, dp_l = pp.DP_algorithm(obs_s, obs_l, plan_start_s=g["s0"][b], plan_start_l=g["l0"][b],
                                     plan_start_dl=g["dl0"][b], plan_start_ddl=g["ddl0"][b])
        n = g["ref_n"][b]
        assert isinstance(dp_s, list) and len(dp_s) == n
        assert np.array_equal(dp_s, g["ref_s"][b, :n]) and np.abs(np.array(dp_l) - g["ref_l"][b, :n]).max() < 1e-9
        dp_l2, dp_s2 = dp_l[::2], dp_s[::2]
        if g["cor_index_error"][b]:
            with pytest.raises(IndexError):
                pp.cal_lmin_lmax(dp_path_s=dp_s2, dp_path_l=dp_l2, obs_s_list=obs_s, obs_l_list=obs_l,
                                 obs_length=5, obs_width=5)
            continue
        l_min, l_max = pp.cal_lmin_lmax(dp_path_s=dp_s2, dp_path_l=dp_l2, obs_s_list=obs_s, obs_l_list=obs_l,
                                        obs_length=5, obs_width=5)
        m = len(l_min)
        assert np.array_equal(l_min, g["ref_lmin"][b, :m]) and np.array_equal(l_max, g["ref_lmax"][b, :m])
        ql, qdl, qddl = pp.Quadratic_planning(l_min, l_max, plan_start_l=g["l0"][b], plan_start_dl=g["dl0"][b],
                                              plan_start_ddl=g["ddl0"][b])
        assert isinstance(ql, list) and len(ql) == m == len(qdl) == len(qddl)
        # every frame, optimal or 'unknown': the drop-in runs the kernels' documented divergence exit (golden qpd_*)
        assert (pp.Quadratic_planning.last_status[0] == 0) == bool(g["qpd_optimal"][b])
        assert pp.Quadratic_planning.last_status[1] == g["qpd_iters"][b]
        assert np.sqrt(np.mean((np.array(ql) - g["qpd_l"][b, :m]) ** 2)) < 1e-4
        assert np.abs(np.array(qdl) - g["qpd_dl"][b, :m]).max() < 1e-4
        done += 1
    assert done >= 38


def test_non_default_parameters_against_oracle(planner_pkg):
    from oracle import sl_dp, sl_qp
    pp = planner_pkg
    kw = dict(sampling_res=1, w_collision_cost=1e10, w_smooth_cost=[100, 2000, 3000], w_reference_cost=35,
              row=9, col=8, sample_s=10, sample_l=1.0)
    obs_s, obs_l = [18.0, 40.5, 66.0], [0.4, -1.1, 1.9]
    s, l = pp.DP_algorithm(obs_s, obs_l, 1.3, 0.2, 0.01, -0.004, **kw)
    so, lo, _ = sl_dp.dp_algorithm(obs_s, obs_l, 1.3, 0.2, 0.01, -0.004, **kw)
    assert len(s) == len(so) and np.array_equal(s, so) and np.abs(np.array(l) - lo).max() < 1e-12
    lmin = np.linspace(-10, -8, 17)
    lmax = np.linspace(9, 10, 17)
    qkw = dict(dp_sampling_res=1, w_cost_l=700, w_cost_dl=11, w_cost_ddl=1300, w_cost_dddl=170, w_cost_centre=190,
               w_cost_end_l=23, w_cost_end_dl=29, w_cost_end_ddl=31, host_d1=2.5, host_d2=1.2, host_w=2)
    ql, qdl, qddl = pp.Quadratic_planning(lmin, lmax, 0.3, 0.01, -0.02, **qkw)
    ol, odl, oddl, info = sl_qp.quadratic_planning(lmin, lmax, 0.3, 0.01, -0.02, **qkw)
    assert info["status"] == "optimal" and pp.Quadratic_planning.last_status == (0, info["iterations"])
    assert np.abs(np.array(ql) - ol).max() < 1e-6 and np.abs(np.array(qddl) - oddl).max() < 1e-6


def test_more_than_eight_obstacles_and_empty_list(planner_pkg):
    from oracle import sl_dp
    pp = planner_pkg
    rng = np.random.default_rng(4)
    obs_s = sorted(rng.uniform(8, 80, 11))
    obs_l = list(rng.uniform(-3, 3, 11))
    s, l = pp.DP_algorithm(obs_s, obs_l, 0.7, -0.1, 0.0, 0.0)
    so, lo, _ = sl_dp.dp_algorithm(obs_s, obs_l, 0.7, -0.1, 0.0, 0.0)
    assert np.array_equal(s, so) and np.abs(np.array(l) - lo).max() < 1e-12
    s, l = pp.DP_algorithm([], [], 0.7, -0.1, 0.0, 0.0)        # no obstacle: straight branch (:362-363)
    assert all(v == 0.0 for v in l[8:])


def test_standalone_dp_pieces(planner_pkg):
    """cal_start_cost / cal_neighbor_cost / cal_obs_cost / enrich_DP_s_l / cal_quintic_coefficient by their
    reference names: bit-exact against the closed-form oracle, and within the reference's own inversion noise
    of the reference-faithful ('inv') oracle."""
    import sys
    from oracle import sl_dp
    pp = planner_pkg
    from planner import planning_utils as pu
    obs_s, obs_l = [20.0, 33.0, 47.5], [0.5, -1.0, 2.0]
    w = (1e12, [300, 1000, 5000], 20)
    # neighbour edges at several stations, incl. collisions and the near zone
    for (s0, lk, li) in [(15.3, 0.75, -0.75), (30.3, -2.25, 2.25), (45.3, 0.75, 0.75), (0.3, -8.25, 8.25)]:
        got = pp.cal_neighbor_cost(obs_s, obs_l, s0, lk, s0 + 15, li, 15, *w)
        assert got.shape == (1, 1)
        exact = sl_dp.edge_cost_closed(obs_s, obs_l, s0, lk, 0.0, 0.0, s0 + 15, li, 15, w[0], w[1], w[2])
        faithful = sl_dp._edge_cost_inv(obs_s, obs_l, s0, lk, 0, 0, s0 + 15, li, 15, w[0], w[1], w[2])
        assert got[0, 0] == exact
        assert abs(got[0, 0] - faithful) <= 1e-9 * abs(faithful)
    got = pp.cal_start_cost(obs_s, obs_l, 0.4, 0.2, 0.01, -0.003, 3, 12, 15, 1.5, *w)
    end_l = ((12 + 1) / 2 - 1 - 3) * 1.5
    assert got[0, 0] == sl_dp.edge_cost_closed(obs_s, obs_l, 0.4, 0.2, 0.01, -0.003, 15.4, end_l, 15, w[0], w[1], w[2])
    assert abs(got[0, 0] - sl_dp._edge_cost_inv(obs_s, obs_l, 0.4, 0.2, 0.01, -0.003, 15.4, end_l, 15, *w)) <= 1e-9 * got[0, 0]
    # cal_obs_cost: Q8 order dependence
    d2 = np.array([[40.0], [30.0], [20.0], [9.0], [20.0], [50.0], [50.0], [50.0], [50.0], [50.0]])
    assert pp.cal_obs_cost(1e12, d2) == 5000 / 30.0 + 5000 / 20.0 + 1e12
    assert pp.cal_obs_cost(7.0, np.full((10, 1), 36.0)) == 0.0
    # enrich_DP_s_l with non-uniform nodes and resolution 1
    S, L = [14.2, 30.0, 44.9, 61.0], [0.75, -0.75, -0.75, 2.25]
    es, el = pp.enrich_DP_s_l(S, L, 0.2, 0.1, 0.02, -0.001, resolution=1)
    os_, ol_ = sl_dp.enrich(S, L, 0.2, 0.1, 0.02, -0.001, resolution=1)
    assert len(es) == len(os_) and np.array_equal(es, os_) and np.abs(np.array(el) - ol_).max() < 1e-12
    # cal_quintic_coefficient: boundary conditions reproduced, agrees with the reference's inverse within its noise
    c = pu.cal_quintic_coefficient(0.3, 0.02, -0.004, -1.5, 0.01, 0.002, 12.5, 27.5)
    ref = sl_dp._quintic_inv(0.3, 0.02, -0.004, -1.5, 0.01, 0.002, 12.5, 27.5)
    p = np.polynomial.polynomial
    for s, want in ((12.5, (0.3, 0.02, -0.004)), (27.5, (-1.5, 0.01, 0.002))):
        for k in range(3):
            assert abs(p.polyval(s, p.polyder(c, k)) - want[k]) < 1e-9
    assert np.allclose(c, ref, rtol=1e-6, atol=1e-9)
    assert c == sl_dp.quintic_closed(0.3, 0.02, -0.004, -1.5, 0.01, 0.002, 12.5, 27.5)
    # cal_projection_s_fun with explicit match indices
    g = load_golden("frenet.npz")
    from oracle import frenet as ofr
    line = [tuple(r) for r in g["line"][0]]
    smap = list(g["smap"][0])
    pts = [tuple(p) for p in g["obs_xy"][0]]
    mi, _ = ofr.match_projection(pts, line)
    want = ofr.projection_s(line, mi, pts, smap)
    got = pu.cal_projection_s_fun(line, mi, pts, smap)
    assert np.abs(np.array(got) - np.array(want)).max() < 1e-9


def test_predict_block_based_on_frenet_and_cal_proj_point_1(planner_pkg):
    """planning_utils.predict_block_based_on_frenet (:617) / cal_proj_point_1 (:647) against the oracle's proj_point
    on a golden reference line; tolerance 1e-9."""
    import types
    from oracle import frenet as ofr
    from planner import planning_utils
    g = load_golden("frenet.npz")
    f = 3
    line = [tuple(r) for r in g["line"][f]]
    ego = g["ego"][f]
    loc = types.SimpleNamespace(x=float(ego[0]), y=float(ego[1]))
    vel = types.SimpleNamespace(x=6.0, y=1.5, z=0.0)
    cur_s = list(np.arange(0.0, 40.0, 2.0))
    cur_l = list(0.3 * np.sin(np.arange(20) / 3.0))
    got = planning_utils.predict_block_based_on_frenet(loc, vel, line, cur_s, cur_l, ts=0.2)
    smap = ofr.s_map(line, (loc.x, loc.y))
    s = max(cur_s[0], 0) + np.hypot(6.0, 1.5) * 0.2
    l = cur_l[int(np.argmin(abs(np.array(cur_s) - s)))]
    px, py, pth, _, idx = ofr.proj_point(s, 0, line, smap)
    want = (px - l * np.sin(pth), py + l * np.cos(pth))
    assert np.allclose(got, want, rtol=0, atol=1e-9)
    r1 = planning_utils.cal_proj_point_1(s, 0, line, smap)
    assert np.allclose(r1[:4], ofr.proj_point(s, 0, line, smap)[:4], rtol=0, atol=1e-9) and r1[4] == idx


def test_speed_planning_stub_dropin(planner_pkg):
    """`from planner import speed_planning` resolves to the drop-in: cal_start_cost / cal_obs_cost against outputs of
    the unmodified reference (tests/golden/speed_stub.npz; 1e-9 relative = the noise of the reference's 6x6
    inverse, and bit-exact against the oracle's closed form), the entry point raises what the reference raises."""
    from oracle import sl_dp
    from planner import speed_planning as sp
    g = load_golden("speed_stub.npz")
    w = (1e10, [300, 2000, 10000], 20)
    for i in range(len(g["cost"])):
        b, e = g["begin"][i], g["end"][i]
        got = sp.cal_start_cost(list(g["obs_s"][i]), list(g["obs_t"][i]), b[0], b[1], b[2], e[0], e[1], g["sample_t"][i], *w)
        assert got.shape == (1, 1) and abs(got[0, 0] - g["cost"][i]) <= 1e-9 * abs(g["cost"][i])
        assert got[0, 0] == sl_dp.edge_cost_closed(list(g["obs_s"][i]), list(g["obs_t"][i]), b[0], b[1], b[2], 0.0, e[0],
                                                   e[1], g["sample_t"][i], w[0], w[1], w[2], danger2=9.0, safe2=25.0)
    d2 = g["obs_cost_d2"].reshape(10, 1)
    assert sp.cal_obs_cost(1e10, d2) == g["obs_cost"][0] and sp.cal_obs_cost(7.0, np.full((10, 1), 25.0)) == g["obs_cost"][1]
    assert sp.cal_obs_cost(3.0, d2, danger_dis=1, safe_dis=6) == g["obs_cost"][2]
    with pytest.raises(TypeError):
        sp.construct_ST_graph_and_planning([0.0], 0, 5, 0, [12.0], [3.0])
    with pytest.raises(NameError):
        sp.construct_ST_graph_and_planning([0.0], 0, 5, 0, [], [])
