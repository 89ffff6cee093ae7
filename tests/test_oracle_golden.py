"""CPU tests: the oracle restatement against the golden outputs of the unmodified reference
(tests/golden/*.npz, made by oracle/make_golden.py from 6Lackiu/EMplanner_Carla @ 1f115a5)."""
import numpy as np
import pytest

from conftest import lattice_from_golden, load_golden
from oracle import coneqp, frenet, sl_dp, sl_qp, st_dp


def _frame(g, b):
    K = int(g["nobs"][b])
    return (list(g["obs_s"][b, :K]), list(g["obs_l"][b, :K]), g["s0"][b], g["l0"][b], g["dl0"][b], g["ddl0"][b])


@pytest.mark.parametrize("tag,frames", [("A", range(0, 12)), ("B", [0, 3]), ("C", [1, 5])])
def test_dp_inv_mode_is_bit_exact_to_reference(tag, frames):
    """mode='inv' evaluates the quintic like the reference (6x6 inverse): identical bits."""
    g = load_golden("sl_%s.npz" % tag)
    lat = lattice_from_golden(g)
    for b in frames:
        s, l, rows = sl_dp.dp_algorithm(*_frame(g, b), mode="inv", **lat.kwargs())
        n = g["ref_n"][b]
        assert len(s) == n
        assert np.array_equal(s, g["ref_s"][b, :n]) and np.array_equal(l, g["ref_l"][b, :n])
        assert np.array_equal(rows, g["rows"][b]) or g["nobs"][b] == 0


@pytest.mark.parametrize("tag,tol,frames", [("A", 1e-10, 512), ("B", 1e-9, 128), ("C", 1e-9, 256)])
def test_dp_closed_form_matches_reference(tag, tol, frames):
    """mode='closed' (what the CUDA kernel implements) against the unmodified reference: same rows and length, l within
    SURVEY 8(c)'s 1e-9 m (1e-10 on lattice A).  The golden generator evaluated this on EVERY golden frame (512 / 128 /
    1024) and stored the result (`closed_*`); the first `frames` frames are recomputed here with the batched oracle."""
    g = load_golden("sl_%s.npz" % tag)
    assert np.all(g["closed_rows_equal"] == 1) and np.all(g["closed_n_equal"] == 1), "row / length flip vs the reference"
    assert np.nanmax(g["closed_dl"]) <= tol, np.nanmax(g["closed_dl"])
    g = {k: (g[k][:frames] if g[k].ndim and len(g[k]) == len(g["s0"]) else g[k]) for k in g.files}
    lat = lattice_from_golden(g)
    rows, ns, nl, _ = sl_dp.dp_batch(g["obs_s"], g["obs_l"], g["nobs"], g["s0"], g["l0"], g["dl0"], g["ddl0"],
                                     **lat.kwargs())
    has = g["nobs"] > 0
    assert np.array_equal(rows[has], g["rows"][has])
    worst = 0.0
    for b in range(len(has)):
        s, l = sl_dp.enrich(ns[b], nl[b], g["s0"][b], g["l0"][b], g["dl0"][b], g["ddl0"][b], lat.sampling_res)
        n = g["ref_n"][b]
        assert len(s) == n and np.array_equal(s, g["ref_s"][b, :n])
        worst = max(worst, np.abs(l - g["ref_l"][b, :n]).max())
    assert worst <= tol, worst


def test_config1_single_frame_60x9():
    """BASELINE.json configs[0]: single ego frame, 60x9 lattice, 3 static obstacles."""
    g = load_golden("sl_B.npz")
    assert g["nobs"][0] == 3 and np.allclose(g["obs_s"][0, :3], [22, 47, 71])
    lat = lattice_from_golden(g)
    s, l, rows = sl_dp.dp_algorithm(*_frame(g, 0), **lat.kwargs())
    assert len(s) == 61 and np.array_equal(rows, g["rows"][0])
    assert np.abs(l - g["ref_l"][0, :61]).max() < 1e-4


def test_quirk_ledger_dp():
    """SURVEY.md Appendix A quirks visible in DP outputs."""
    g = load_golden("sl_A.npz")
    lat = lattice_from_golden(g)
    # Q9: even row count -> no lattice row at l = 0; K = 0 frames run straight at l = 0 anyway
    assert 0.0 not in sl_dp.node_l(12, 1.5)
    k0 = np.where(g["nobs"] == 0)[0][0]
    assert np.all(g["ref_l"][k0, 8:g["ref_n"][k0]] == 0.0)
    # Q10: the densified length depends on float rounding of int(end_s - start_s)
    lengths = set(int(v) for v in np.unique(g["ref_n"]))
    assert len(lengths) > 1 and lengths <= {47, 48, 49}
    # Q1: `dddl` uses s*2 instead of s**2, so the cost of one and the same edge geometry depends on absolute s
    e0 = sl_dp._edge_cost_inv([], [], 0.0, 0.75, 0, 0, 15.0, -0.75, 15.0, 1e12, (300, 1000, 5000), 20)
    e1 = sl_dp._edge_cost_inv([], [], 60.0, 0.75, 0, 0, 75.0, -0.75, 15.0, 1e12, (300, 1000, 5000), 20)
    assert e1 > 10 * e0
    # Q8: an obstacle's cost stops at the first colliding sample, after the near-miss terms before it
    c = sl_dp._obs_cost(1e12, [np.array([30.0]), np.array([20.0]), np.array([9.0]), np.array([20.0])])
    assert c[0] == 5000 / 30.0 + 5000 / 20.0 + 1e12


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_corridor_and_qp_against_golden(tag):
    g = load_golden("sl_%s.npz" % tag)
    checked = 0
    for b in range(0, len(g["s0"]), max(3, len(g["s0"]) // 40)):
        n = g["ref_n"][b]
        K = int(g["nobs"][b])
        s, l = g["ref_s"][b, :n:2], g["ref_l"][b, :n:2]
        if g["cor_index_error"][b]:
            with pytest.raises(IndexError):
                sl_qp.corridor(s, l, g["obs_s"][b, :K], g["obs_l"][b, :K], 5, 5)
            continue
        lmin, lmax = sl_qp.corridor(s, l, g["obs_s"][b, :K], g["obs_l"][b, :K], 5, 5)
        m = len(lmin)
        assert np.array_equal(lmin, g["ref_lmin"][b, :m]) and np.array_equal(lmax, g["ref_lmax"][b, :m])
        ql, qdl, qddl, info = sl_qp.quadratic_planning(lmin, lmax, g["l0"][b], g["dl0"][b], g["ddl0"][b])
        assert (info["status"] == "optimal") == bool(g["qp_optimal"][b])
        assert info["iterations"] == g["qp_iters"][b]
        assert np.array_equal(ql, g["qp_l"][b, :m])
        # the same solve with the kernels' documented divergence exit (golden qpd_*): identical on converging frames
        dl_, _, _, info_d = sl_qp.quadratic_planning(lmin, lmax, g["l0"][b], g["dl0"][b], g["ddl0"][b],
                                                      diverge_factor=float(g["diverge_factor"]))
        assert (info_d["status"] == "optimal") == bool(g["qpd_optimal"][b]) == bool(g["qp_optimal"][b])
        assert info_d["iterations"] == g["qpd_iters"][b] and np.array_equal(dl_, g["qpd_l"][b, :m])
        if info["status"] == "optimal":
            assert np.array_equal(dl_, ql)
        else:
            assert info_d["iterations"] <= info["iterations"]
        checked += 1
    assert checked > 3


def test_coneqp_known_answer():
    """The 2-variable QP of the reference's own cvxopt smoke script (test.py:8-27, printed not asserted there):
    min 2x1^2 + x2^2 + x1x2 + x1 + x2  s.t. x >= 0, x1 + x2 = 1  ->  x = (0.25, 0.75) analytically."""
    P = 2 * np.array([[2, .5], [.5, 1]])
    q = np.array([1.0, 1.0])
    G = np.array([[-1.0, 0.0], [0.0, -1.0]])
    h = np.array([0.0, 0.0])
    A = np.array([[1.0, 1.0]])
    b = np.array([1.0])
    r = coneqp.qp(P, q, G, h, A, b)
    assert r["status"] == "optimal"
    assert np.allclose(r["x"], [0.25, 0.75], atol=1e-6)
    # KKT check on a planner QP: optimal status implies small residuals by construction
    lmin, lmax = -10 * np.ones(12), 10 * np.ones(12)
    lmax[5:8] = 1.0
    l, dl, ddl, info = sl_qp.quadratic_planning(lmin, lmax, 0.2, 0.01, 0.0)
    assert info["status"] == "optimal" and abs(l[0] - 0.2) < 1e-6 and abs(l[-1]) < 1e-6
    assert info["pres"] <= 1e-7 and info["dres"] <= 1e-7


def test_frenet_against_golden():
    g = load_golden("frenet.npz")
    gp = [tuple(r) for r in g["global_path"]]
    for f in range(len(g["ego"])):
        ego = tuple(g["ego"][f])
        mi, _ = frenet.find_match_points([ego], gp, False, int(g["pre_idx"][f]))
        assert mi[0] == g["match_idx"][f]
        local = frenet.sampling(mi[0], gp)
        assert np.array_equal(np.array(local), g["local"][f])
        line, _ = frenet.smooth_reference_line(local)
        assert np.array_equal(np.array(line), g["line"][f])          # Q13 heading/kappa quirk included
        smap = frenet.s_map(line, ego)
        assert np.array_equal(smap, g["smap"][f])
        os_, ol_ = frenet.s_l([tuple(p) for p in g["obs_xy"][f]], line, smap)
        assert np.array_equal(os_, g["obs_s"][f]) and np.array_equal(ol_, g["obs_l"][f])   # Q12 index-0 quirk
        der = frenet.s_l_deriv([ego], [tuple(g["vel"][f])], [tuple(g["acc"][f])], line, ego)
        assert np.array_equal([d[0] for d in der], g["deriv"][f])
        raw = frenet.frenet_to_cartesian_raw(g["start_sl"][f, 0], g["start_sl"][f, 1], list(g["path_s"][f]),
                                             list(g["path_l"][f]), line, smap)
        assert np.array_equal(np.array(raw[1:]), g["cart_raw"][f, 1:len(raw)])            # Q14 truncation
        cart = frenet.frenet_to_cartesian(g["start_sl"][f, 0], g["start_sl"][f, 1], list(g["path_s"][f]),
                                          list(g["path_l"][f]), line, smap)
        assert len(cart) == g["cart_n"][f] and np.array_equal(np.array(cart), g["cart"][f, :len(cart)])


def test_st_dp_against_reference_tables():
    """Forward tables harvested from the reference's speed_DP frame before its back-track raises."""
    g = load_golden("st_dp.npz")
    for f in range(len(g["v0"])):
        o = st_dp.speed_dp(g["s_in"][f], g["s_out"][f], g["t_in"][f], g["t_out"][f], g["v0"][f])
        assert np.array_equal(o["node"], g["node"][f])
        assert o["min_row"] == g["min_row"][f] and o["min_col"] == g["min_col"][f]
        assert np.array_equal(o["s_dot"], g["s_dot"][f])
        fin = np.isfinite(g["cost"][f])
        assert np.allclose(o["cost"][fin], g["cost"][f][fin], rtol=1e-12, atol=0)
        # repaired back-track: a monotone s(t) profile ending at (min_row, min_col)
        c = o["min_col"]
        assert np.all(np.diff(o["dp_s"][:c + 1]) >= 0) and np.all(np.isnan(o["dp_s"][c + 1:]))


def test_st_misc_oracle_against_reference_golden():
    from oracle import st_misc
    g = load_golden("st_misc.npz")
    for f in range(len(g["dyn_s"])):
        got = st_misc.st_graph(g["dyn_s"][f], g["dyn_l"][f], g["dyn_sd"][f], g["dyn_ld"][f])
        assert np.array_equal(np.array(got), g["stg"][f], equal_nan=True)
        assert np.array_equal(st_misc.start_condition(*g["sc_in"][f]), g["sc_out"][f])
        d = st_misc.dy_obs_deri(g["dyn_l"][f], g["dyn_sd"][f], g["dyn_ld"][f], g["deri_head"][f], g["deri_kap"][f])
        assert np.array_equal(np.array(d), g["deri_out"][f], equal_nan=True)
    assert np.array_equal(st_misc.trajectory_index2s(g["traj_x"], g["traj_y"]), g["traj_s"])
    xs, ys = g["traj_x"][:60], g["traj_y"][:60]
    f2c = st_misc.frenet2cartesian(g["f2c_s"], g["f2c_l"], g["f2c_dl"], g["f2c_ddl"], xs, ys, g["path_heading"],
                                   g["path_kappa"], g["traj_s"][:60])
    assert np.allclose(np.array([a[:, 0] for a in f2c]), g["f2c_out"], rtol=0, atol=1e-12, equal_nan=True)


def test_st_back_oracle_against_reference_golden():
    """oracle/st_back.py vs outputs of the unmodified reference (generate_convex_space incl. its exceptions,
    increase_points, path_speed_merge: bit-exact) and vs the stored repaired speed_QP solve."""
    from oracle import st_back
    g = load_golden("st_back.npz")
    EXC = {IndexError: 0x20, ValueError: 0x10}
    n_qp = 0
    for f in range(len(g["v0"])):
        args = (g["dp_s"][f], g["dp_t"][f], g["path"][f, 0], g["s_in"][f], g["s_out"][f], g["t_in"][f], g["t_out"][f],
                g["path"][f, 1])
        try:
            cs = st_back.convex_space(*args)
            assert g["ref_convex_exc"][f] == 0 and np.array_equal(np.array(cs), g["ref_bounds"][f])
        except (IndexError, ValueError) as e:
            assert g["ref_convex_exc"][f] == EXC[type(e)]
            continue
        if g["matrices_identical"][f] and n_qp < 6:
            r = st_back.speed_qp(g["v0"][f], g["a0"][f], g["dp_s"][f], g["dp_t"][f], *cs)
            assert np.array_equal(np.array(r[:4]), g["qp"][f], equal_nan=True)
            assert r[4]["iterations"] == g["qp_iters"][f] and (r[4]["status"] == "optimal") == bool(g["qp_optimal"][f])
            n_qp += 1
        elif g["qp_exc"][f]:
            with pytest.raises(IndexError):
                st_back.speed_qp_matrices(g["v0"][f], g["a0"][f], g["dp_s"][f], g["dp_t"][f], *cs)
    assert n_qp == 6
    for k, f in enumerate(g["dense_frames"]):
        ip = st_back.increase_points(*g["qp"][f])
        assert np.array_equal(np.array(ip), g["ref_dense"][k])
        mg = st_back.path_speed_merge(*ip, g["current_time"][f], *g["path5"][f])
        assert np.array_equal(np.array(mg), g["ref_traj"][k], equal_nan=True)


def test_speed_qp_repair_is_a_valid_qp_solution():
    """KKT check of the repaired speed_QP on a golden frame: primal feasibility of Aeq' x = 0, bounds, and that no
    feasible perturbation along the null space lowers the objective (first-order optimality)."""
    from oracle import st_back
    g = load_golden("st_back.npz")
    f = int(np.nonzero((g["matrices_identical"] == 1) & (g["qp_optimal"] == 1))[0][0])
    m = st_back.speed_qp_matrices(g["v0"][f], g["a0"][f], g["dp_s"][f], g["dp_t"][f], *g["ref_bounds"][f])
    P, q, G, h, A, b = st_back.speed_qp_cone_form(m)
    r = st_back.speed_qp(g["v0"][f], g["a0"][f], g["dp_s"][f], g["dp_t"][f], *g["ref_bounds"][f])
    x, info = r[4]["x"], r[4]
    assert np.abs(A @ x - b).max() < 1e-6 and (G @ x - h).max() < 1e-6
    assert np.abs(P @ x + q + A.T @ info["y"] + G.T @ info["z"]).max() < 1e-4 * max(1.0, np.abs(q).max())
    assert info["z"].min() > -1e-9 and abs(info["z"] @ (h - G @ x)) < 1e-5 * abs(info["pcost"])   # reltol 1e-6


def test_speed_planning_stub_oracle_against_reference_golden():
    """planner/speed_planning.py (the stub the north_star names): cal_start_cost is the S-L edge cost on an (s, t)
    quintic with 3 m / 5 m distances - oracle closed form vs the unmodified reference (whose 6x6 inverse is noisy:
    1e-9 relative); cal_obs_cost exact; the entry point raises TypeError / NameError in the reference."""
    g = load_golden("speed_stub.npz")
    w = (1e10, [300, 2000, 10000], 20)
    hit = 0
    for i in range(len(g["cost"])):
        b, e = g["begin"][i], g["end"][i]
        got = sl_dp.edge_cost_closed(list(g["obs_s"][i]), list(g["obs_t"][i]), b[0], b[1], b[2], 0.0, e[0], e[1],
                                     g["sample_t"][i], w[0], w[1], w[2], danger2=9.0, safe2=25.0)
        assert abs(got - g["cost"][i]) <= 1e-9 * abs(g["cost"][i]), i
        hit += g["cost"][i] >= 1e10
    assert hit >= 2                                            # collisions are exercised
    assert list(g["entry_point_exceptions"]) == ["TypeError", "NameError"]
    d2 = g["obs_cost_d2"]
    assert g["obs_cost"][0] == 5000 / 24.0 + 5000 / 16.0 + 1e10 and g["obs_cost"][1] == 0.0
