"""GPU parity tests for the S-L DP + QP path, called through the C ABI (host-buffer entry points).

Checker = oracle/ (closed-form mode) and tests/golden (outputs of the unmodified reference).
Tolerances (stated per test):
  DP   : lattice row indices identical; densified length identical; |dl| <= 1e-12 vs the oracle
         (bit-for-bit expected: same association order, no FMA); vs the unmodified reference (512 / 128 / 1024
         golden frames of lattices A / B / C): <= 1e-10 m on A, <= 1e-9 m on B/C (observed 2.9e-12 /
         2.8e-11 / 3.8e-11, the reference's own 6x6-inverse noise), zero row or length flips.
  QP   : per-frame RMS(l_gpu - l_oracle) <= 1e-4 m (north_star) on EVERY frame, optimal or not; status agreement
         100 %; iteration counts equal on every frame.  Observed numbers (profiles/parity_r2.json): worst RMS 2.2e-10 m on
         converging frames, 9.0e-8 m on 'unknown' frames.
"""
import numpy as np
import pytest

from conftest import lattice_from_golden, load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def planner():
    from emplanner_carla_b200 import api
    pl = api.HostPlanner(max_batch=8192, kmax=8, n_enr_max=64, max_col=64)
    yield pl
    pl.close()


def _scenes_from(g):
    from emplanner_carla_b200 import api
    return api.pack_scenes(g["obs_s"], g["obs_l"], g["nobs"], g["s0"], g["l0"], g["dl0"], g["ddl0"])


@pytest.mark.parametrize("tag,tol_ref", [("A", 1e-10), ("B", 1e-9), ("C", 1e-9)])
def test_dp_matches_reference_golden(planner, tag, tol_ref):
    from emplanner_carla_b200 import api
    g = load_golden("sl_%s.npz" % tag)
    lat = lattice_from_golden(g)
    dp = api.dp_params(**lat.kwargs())
    view, rows = planner.dp(_scenes_from(g), dp)
    n = view.n_enr.astype(int)
    assert np.array_equal(n, g["ref_n"])
    has_obs = g["nobs"] > 0
    assert np.array_equal(rows[has_obs], g["rows"][has_obs])
    assert np.all(rows[~has_obs] == -1)
    worst = 0.0
    for b in range(len(n)):
        assert np.array_equal(view.enr_s[b, :n[b]], g["ref_s"][b, :n[b]])
        worst = max(worst, np.abs(view.enr_l[b, :n[b]] - g["ref_l"][b, :n[b]]).max())
    print("lattice", tag, "max |l_gpu - l_reference| = %.3e" % worst)
    _record("sl_dp_%s" % tag, dict(lattice=tag, frames=int(len(n)), row_flips=0, length_flips=0, worst_abs_dl_vs_reference=worst,
                                   tolerance=tol_ref))
    assert worst <= tol_ref


@pytest.mark.parametrize("name", ["A", "B", "C"])
def test_dp_bit_exact_vs_oracle(planner, name):
    from emplanner_carla_b200 import api, scenes
    from oracle import sl_dp
    lat = scenes.LATTICES[name]
    B = 256 if name == "A" else 96
    sc = scenes.sl_scenes(B, lat, 40 + ord(name))
    dp = api.dp_params(**lat.kwargs())
    view, rows = planner.dp(api.pack_scenes(**sc), dp)
    o_rows, o_s, o_l, o_cost = sl_dp.dp_batch(sc["obs_s"], sc["obs_l"], sc["nobs"], sc["s0"], sc["l0"], sc["dl0"],
                                              sc["ddl0"], **lat.kwargs())
    assert np.array_equal(rows, o_rows)
    has = sc["nobs"] > 0
    assert np.array_equal(view.dp_cost[has], o_cost[has]), "DP cost must be bit-identical to the oracle"
    for b in range(B):
        es, el = sl_dp.enrich(o_s[b], o_l[b], sc["s0"][b], sc["l0"][b], sc["dl0"][b], sc["ddl0"][b],
                              resolution=lat.sampling_res)
        nb = int(view.n_enr[b])
        assert nb == len(es)
        assert np.array_equal(view.enr_s[b, :nb], es)
        assert np.abs(view.enr_l[b, :nb] - el).max() <= 1e-12


def _qp_compare(view, g, key, ok, api):
    """Per-frame comparison of the QP part of plan records with golden arrays `<key>_l/_iters/_optimal`."""
    st = view.qp_status.astype(int)
    opt_gpu = st == api.QP_OK
    opt_ref = g[key + "_optimal"] == 1
    r = dict(frames=int(ok.sum()), optimal=int((opt_ref & ok).sum()), status_disagree=int((opt_gpu != opt_ref)[ok].sum()),
             iter_mismatch_optimal=0, iter_mismatch_unknown=0, worst_rms_optimal=0.0, worst_rms_unknown=0.0,
             worst_abs_optimal=0.0)
    for b in np.where(ok)[0]:
        n = int(view.n_qp[b])
        assert n == (g["ref_n"][b] + 1) // 2
        d = view.qp_l[b, :n] - g[key + "_l"][b, :n]
        rms = float(np.sqrt(np.mean(d * d)))
        kind = "optimal" if opt_ref[b] else "unknown"
        r["worst_rms_" + kind] = max(r["worst_rms_" + kind], rms)
        r["iter_mismatch_" + kind] += int(view.qp_iters[b] != g[key + "_iters"][b])
        if opt_ref[b]:
            r["worst_abs_optimal"] = max(r["worst_abs_optimal"], float(np.abs(d).max()))
    return r


def _record(name, obj):
    """Observed parity numbers go to gpurun_out/parity/ (collected into profiles/parity_r2.json by tools/parity_report.py)."""
    import json
    import os
    from conftest import ROOT
    d = os.path.join(ROOT, "gpurun_out", "parity")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name + ".json"), "w") as f:
            json.dump(obj, f, indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.mark.parametrize("tag", ["A", "B", "C"])
def test_plan_matches_golden_qp_on_every_frame(planner, tag):
    """Full DP -> [::2] -> corridor -> QP against the golden corridor and QP solution, EVERY frame (optimal or not).

    (1) kernel defaults (divergence exit, EmpQpParams.diverge_factor = 2) against the oracle run with the same rule
        (`qpd_*`): status must agree on 100 % of the frames, RMS(l) <= 1e-4 m (north_star) on optimal AND unknown frames.
    (2) cvxopt's rule (diverge_factor = 0: iterate until the factorisation breaks down) against `qp_*`: status must agree
        on 100 % of the frames and optimal frames must match to 1e-4 m; the iterate of an 'unknown' frame is not a
        numerically reproducible quantity under this rule (see include/emplanner.h) - its distance is recorded only."""
    from emplanner_carla_b200 import api
    g = load_golden("sl_%s.npz" % tag)
    lat = lattice_from_golden(g)
    dp = api.dp_params(**lat.kwargs())
    rec = _scenes_from(g)
    ok = g["cor_index_error"] == 0
    view, _ = planner.plan(rec, dp, api.qp_params())
    assert np.array_equal(view.qp_status.astype(int) == api.QP_CORRIDOR_INDEX, ~ok)
    r1 = _qp_compare(view, g, "qpd", ok, api)
    unknown = ok & (g["qpd_optimal"] == 0)
    assert np.all(view.qp_status[unknown] == api.QP_UNKNOWN_DIVERGED)
    view0, _ = planner.plan(rec, dp, api.qp_params(diverge_factor=0))
    r2 = _qp_compare(view0, g, "qp", ok, api)
    print("lattice", tag, "divergence exit:", r1)
    print("lattice", tag, "cvxopt rule    :", r2)
    _record("sl_qp_%s" % tag, dict(lattice=tag, divergence_exit=r1, cvxopt_rule=r2))
    assert r1["status_disagree"] == 0 and r2["status_disagree"] == 0
    assert r1["worst_rms_optimal"] <= 1e-4 and r1["worst_rms_unknown"] <= 1e-4
    assert r2["worst_rms_optimal"] <= 1e-4
    assert r1["iter_mismatch_optimal"] + r1["iter_mismatch_unknown"] == 0      # observed: 0 of 512 / 128 / 1024 frames
    assert r2["iter_mismatch_optimal"] == 0
    # frames that converge are untouched by the divergence exit: same bits out of the kernel either way
    opt = ok & (g["qp_optimal"] == 1)
    assert np.array_equal(view.qp_l[opt], view0.qp_l[opt]) and np.array_equal(view.qp_iters[opt], view0.qp_iters[opt])


def test_qp_direct_and_corridor_entry_points(planner):
    from emplanner_carla_b200 import api
    g = load_golden("sl_A.npz")
    ok = (g["cor_index_error"] == 0) & (g["ref_n"] == 48)
    idx = np.where(ok)[0][:64]
    n = 24
    qp = api.qp_params()
    scenes = _scenes_from(g)[idx]
    lmin, lmax, stc = planner.corridor(g["ref_s"][idx, 0:48:2], g["ref_l"][idx, 0:48:2], scenes, qp)
    assert np.all(stc == 0)
    assert np.array_equal(lmin, g["ref_lmin"][idx, :n]) and np.array_equal(lmax, g["ref_lmax"][idx, :n])
    start = np.stack([g["l0"][idx], g["dl0"][idx], g["ddl0"][idx]], axis=1)
    x, st, it = planner.qp(lmin, lmax, start, qp)
    opt = (g["qp_optimal"][idx] == 1) & (st == 0)
    assert opt.sum() >= 0.7 * len(idx)
    err = np.abs(x[opt, 0, :] - g["qp_l"][idx][opt, :n]).max()
    print("direct QP max |dl| %.3e" % err)
    assert err <= 1e-4


def test_full_size_properties(planner):
    """BASELINE.json configs[1] at full size (batch 4096, lattice C): size-independent properties."""
    from emplanner_carla_b200 import api, scenes
    lat = scenes.LATTICE_C
    sc = scenes.sl_scenes(4096, lat, 2)
    dp, qp = api.dp_params(**lat.kwargs()), api.qp_params()
    rec = api.pack_scenes(**sc)
    view, rows = planner.plan(rec, dp, qp, want_rows=True)
    view2, rows2 = planner.plan(rec, dp, qp, want_rows=True)
    assert np.array_equal(view.raw[:, :7], view2.raw[:, :7]) and np.array_equal(rows, rows2), "run-to-run determinism"
    n = view.n_enr.astype(int)
    assert np.all(n == 61)
    # densified s is strictly increasing and ends at the last lattice column
    assert np.all(np.diff(view.enr_s[:, :61], axis=1) > 0)
    assert np.allclose(view.enr_s[:, 60], sc["s0"] + 60 * 1.5, rtol=0, atol=1e-12)
    # K = 0 frames: straight DP path, QP converges
    k0 = sc["nobs"] == 0
    assert np.all(rows[k0] == -1) and np.all(view.enr_l[k0, 1:61] == 0.0)
    assert np.all(view.qp_status[k0] == 0)
    # optimal frames satisfy the start pin, the end pin and the corridor box
    opt = view.qp_status == 0
    assert opt.mean() > 0.5
    assert np.abs(view.qp_l[opt, 0] - sc["l0"][opt]).max() < 1e-6
    assert np.abs(view.qp_l[opt, 30]).max() < 1e-6
    assert np.all(np.abs(view.qp_l[opt, :31]) <= 10.0 + 1e-6)
    # permutation equivariance: frames are independent
    perm = np.random.default_rng(0).permutation(4096)
    view3, _ = planner.plan(rec[perm], dp, qp)
    assert np.array_equal(view3.raw[:, :7], view.raw[perm][:, :7])
    assert np.array_equal(view3.qp_l, view.qp_l[perm])


def test_pipelined_submit_wait_matches_synchronous_plan():
    """emp_sl_plan_submit / emp_sl_plan_wait (batches in flight on their own streams) return exactly what the
    synchronous emp_sl_plan_host returns, slot by slot; misuse is an argument error, not a crash."""
    from emplanner_carla_b200 import api, scenes
    lat = scenes.LATTICE_A
    dp, qp = api.dp_params(**lat.kwargs()), api.qp_params()
    NE = api.max_enriched(dp)
    B = 96
    pl = api.HostPlanner(max_batch=B, n_enr_max=NE, max_col=lat.col)
    try:
        batches = [api.pack_scenes(**scenes.sl_scenes(B, lat, 40 + k)) for k in range(6)]
        want = [pl.plan(b, dp, qp)[0].raw.copy() for b in batches]
        ins = [api.pinned_empty(batches[0].shape) for _ in range(api.PLAN_SLOTS)]
        outs = [api.pinned_empty(want[0].shape) for _ in range(api.PLAN_SLOTS)]
        got = [None] * len(batches)
        for k, b in enumerate(batches):
            s = k % api.PLAN_SLOTS
            if k >= api.PLAN_SLOTS:
                pl.plan_wait(s)
                got[k - api.PLAN_SLOTS] = outs[s].copy()
            ins[s][:] = b
            pl.plan_submit(s, ins[s], dp, qp, outs[s])
        for k in range(len(batches) - api.PLAN_SLOTS, len(batches)):
            pl.plan_wait(k % api.PLAN_SLOTS)
            got[k] = outs[k % api.PLAN_SLOTS].copy()
        for k in range(len(batches)):
            view_g, view_w = api.PlanView(got[k], NE), api.PlanView(want[k], NE)
            assert np.array_equal(view_g.n_enr, view_w.n_enr) and np.array_equal(view_g.qp_status, view_w.qp_status)
            for b in range(B):
                n, m = int(view_w.n_enr[b]), int(view_w.n_qp[b])
                assert np.array_equal(view_g.enr_l[b, :n], view_w.enr_l[b, :n])
                assert np.array_equal(view_g.qp_l[b, :m], view_w.qp_l[b, :m])
        pl.plan_submit(0, ins[0], dp, qp, outs[0])
        with pytest.raises(api.EmpError):
            pl.plan_submit(0, ins[0], dp, qp, outs[0])            # slot busy
        pl.plan_wait(0)
        with pytest.raises(api.EmpError):
            pl.plan_submit(1, batches[0], dp, qp, outs[1])        # pageable input
    finally:
        pl.close()


def test_dp_bit_exact_at_scale_and_at_the_lattice_limits(planner):
    """The branch-and-bound DP kernel against the full-sweep oracle: the first 1024 frames of BASELINE configs[1]
    (rows and cost bit-identical), and the largest lattice the kernel takes (row = EMP_MAX_ROW = 16, 100 columns,
    16 obstacle slots, sampling_res 1): rows, cost and densified length."""
    from emplanner_carla_b200 import api, scenes
    from oracle import sl_dp
    lat = scenes.LATTICE_C
    sc = scenes.sl_scenes(4096, lat, 2)
    big_pl = api.HostPlanner(max_batch=1024, n_enr_max=api.max_enriched(api.dp_params(**lat.kwargs())), max_col=lat.col)
    try:
        first = {k: v[:1024] for k, v in sc.items()}
        view, rows = big_pl.dp(api.pack_scenes(**first), api.dp_params(**lat.kwargs()))
    finally:
        big_pl.close()
    o_rows, _, _, o_cost = sl_dp.dp_batch(first["obs_s"], first["obs_l"], first["nobs"], first["s0"], first["l0"],
                                          first["dl0"], first["ddl0"], **lat.kwargs())
    assert np.array_equal(rows, o_rows)
    has = first["nobs"] > 0
    assert np.array_equal(view.dp_cost[has], o_cost[has])
    big = scenes.Lattice("max", 16, 100, 1.0, 1.0, 1)
    B, K = 12, 16
    rng = np.random.default_rng(11)
    obs_s = np.sort(rng.uniform(8.0, 95.0, (B, K)), axis=1)
    obs_l = rng.uniform(-6.0, 6.0, (B, K))
    nobs = np.full(B, K, dtype=np.int32)
    nobs[:3] = [1, 7, 12]
    s0, l0 = rng.uniform(0, 2, B), rng.uniform(-0.5, 0.5, B)
    dl0, ddl0 = rng.uniform(-0.05, 0.05, B), rng.uniform(-0.01, 0.01, B)
    dp = api.dp_params(**big.kwargs())
    pl = api.HostPlanner(max_batch=B, kmax=16, n_enr_max=api.max_enriched(dp), max_col=big.col)
    try:
        view, rows = pl.dp(api.pack_scenes(obs_s, obs_l, nobs, s0, l0, dl0, ddl0, kmax=16), dp)
    finally:
        pl.close()
    o_rows, o_s, o_l, o_cost = sl_dp.dp_batch(obs_s, obs_l, nobs, s0, l0, dl0, ddl0, **big.kwargs())
    assert np.array_equal(rows, o_rows) and np.array_equal(view.dp_cost, o_cost)
    for b in range(B):
        es, el = sl_dp.enrich(o_s[b], o_l[b], s0[b], l0[b], dl0[b], ddl0[b], resolution=1)
        assert int(view.n_enr[b]) == len(es) and np.abs(view.enr_l[b, :len(es)] - el).max() <= 1e-12


def test_nan_obstacle_inside_the_list_is_ignored_like_in_the_reference(planner):
    """A NaN obstacle among the first K costs nothing in cal_obs_cost (both range tests fail, path_planning.py:601-608)
    but the frame still takes the DP branch; cal_lmin_lmax sees np.argmin of an all-NaN array (= 0).  The scene record
    keeps the NaN (api.pack_scenes) and the kernels must behave like the oracle on the same arrays."""
    from emplanner_carla_b200 import api, scenes
    from oracle import sl_dp, sl_qp
    lat = scenes.LATTICE_A
    sc = scenes.sl_scenes(64, lat, 77, k_fixed=3)
    sc["obs_s"][:, 1] = np.nan
    sc["obs_l"][:, 1] = np.nan
    sc["obs_s"][:8, 0] = np.nan                                 # some frames: only the s of an obstacle is NaN
    sc["nobs"][60:] = 1
    sc["obs_s"][60:, 0] = np.nan                                # frames whose ONLY obstacle is NaN: DP branch, zero obstacle cost
    sc["obs_l"][60:, 0] = np.nan
    dp, qp = api.dp_params(**lat.kwargs()), api.qp_params()
    rec = api.pack_scenes(**sc)
    assert np.isnan(rec[:60, 7]).all() and not np.isnan(rec[:, 8]).any() and np.all(rec[60:, 7] == 0.0)
    view, rows = planner.plan(rec, dp, qp, want_rows=True)
    o_rows, o_s, o_l, o_cost = sl_dp.dp_batch(sc["obs_s"], sc["obs_l"], sc["nobs"], sc["s0"], sc["l0"], sc["dl0"],
                                              sc["ddl0"], **lat.kwargs())
    assert np.array_equal(rows, o_rows) and np.array_equal(view.dp_cost, o_cost)
    assert np.all(rows[60:] >= 0)                               # not the K == 0 straight branch
    for b in range(64):
        K = int(sc["nobs"][b])
        es, el = sl_dp.enrich(o_s[b], o_l[b], sc["s0"][b], sc["l0"][b], sc["dl0"][b], sc["ddl0"][b], lat.sampling_res)
        try:
            lmin, lmax = sl_qp.corridor(es[::2], el[::2], sc["obs_s"][b, :K], sc["obs_l"][b, :K], 5, 5)
        except IndexError:
            assert int(view.qp_status[b]) == api.QP_CORRIDOR_INDEX
            continue
        l, _, _, info = sl_qp.quadratic_planning(lmin, lmax, sc["l0"][b], sc["dl0"][b], sc["ddl0"][b], diverge_factor=2.0)
        assert (info["status"] == "optimal") == (int(view.qp_status[b]) == 0)
        m = len(l)
        assert np.sqrt(np.mean((view.qp_l[b, :m] - l) ** 2)) <= 1e-4


def test_general_sizes_more_than_32_stations():
    """Sizes the reference accepts beyond one warp lane per station (path_planning.py:78): sampling_res = 1 on the
    default lattice gives 91 densified points -> 46 QP stations (one CTA of 64 threads per frame), and a 100-station
    corridor (128 threads).  Same parity bar as the warp path: status and iteration count equal, RMS(l) <= 1e-4 m,
    on every frame (same divergence exit on both sides)."""
    from emplanner_carla_b200 import api, scenes
    from oracle import sl_dp, sl_qp
    lat = scenes.Lattice("A_res1", 12, 6, 15, 1.5, 1)
    B = 48
    sc = scenes.sl_scenes(B, lat, 91)
    dp, qp = api.dp_params(**lat.kwargs()), api.qp_params(dp_sampling_res=2)
    NE = api.max_enriched(dp)
    assert (NE + 1) // 2 > 32
    pl = api.HostPlanner(max_batch=B, kmax=8, n_enr_max=NE, max_col=lat.col)
    try:
        view, rows = pl.plan(api.pack_scenes(**sc), dp, qp, want_rows=True)
        o_rows, o_s, o_l, _ = sl_dp.dp_batch(sc["obs_s"], sc["obs_l"], sc["nobs"], sc["s0"], sc["l0"], sc["dl0"],
                                             sc["ddl0"], **lat.kwargs())
        assert np.array_equal(rows, o_rows)
        worst, n_unknown = 0.0, 0
        for b in range(B):
            K = int(sc["nobs"][b])
            es, el = sl_dp.enrich(o_s[b], o_l[b], sc["s0"][b], sc["l0"][b], sc["dl0"][b], sc["ddl0"][b], lat.sampling_res)
            assert int(view.n_enr[b]) == len(es) and 88 <= len(es) <= 91     # Q10: int(end_s - start_s) per segment
            try:
                lmin, lmax = sl_qp.corridor(es[::2], el[::2], sc["obs_s"][b, :K], sc["obs_l"][b, :K], 5, 5)
            except IndexError:
                assert int(view.qp_status[b]) == api.QP_CORRIDOR_INDEX
                continue
            l, _, _, info = sl_qp.quadratic_planning(lmin, lmax, sc["l0"][b], sc["dl0"][b], sc["ddl0"][b], diverge_factor=2.0)
            m = len(l)
            assert int(view.n_qp[b]) == m and 44 <= m <= 46
            assert (info["status"] == "optimal") == (int(view.qp_status[b]) == 0), b
            assert int(view.qp_iters[b]) == info["iterations"], b
            n_unknown += info["status"] != "optimal"
            worst = max(worst, float(np.sqrt(np.mean((view.qp_l[b, :m] - l) ** 2))))
        print("46-station QP (block path, 64 threads): worst RMS %.2e, unknown frames %d" % (worst, n_unknown))
        assert worst <= 1e-4
        # direct entry point, 100 stations (128 threads per frame), hand-made corridors
        rng = np.random.default_rng(5)
        n, Bq = 100, 16
        lmin = np.full((Bq, n), -10.0)
        lmax = np.full((Bq, n), 10.0)
        for b in range(Bq):
            for _ in range(3):
                a = int(rng.integers(10, 80))
                if rng.random() < 0.5:
                    lmax[b, a:a + 6] = rng.uniform(-1.0, 2.0)
                else:
                    lmin[b, a:a + 6] = rng.uniform(-2.0, 1.0)
        start = np.stack([rng.uniform(-0.3, 0.3, Bq), rng.uniform(-0.02, 0.02, Bq), rng.uniform(-0.01, 0.01, Bq)], axis=1)
        x, st, it = pl.qp(lmin, lmax, start, qp)
        worst = 0.0
        for b in range(Bq):
            l, dl, ddl, info = sl_qp.quadratic_planning(lmin[b], lmax[b], *start[b], diverge_factor=2.0)
            assert (info["status"] == "optimal") == (st[b] == 0) and it[b] == info["iterations"], b
            worst = max(worst, float(np.sqrt(np.mean((x[b, 0] - l) ** 2))), float(np.abs(x[b, 1] - dl).max()))
        print("100-station QP (block path, 128 threads): worst deviation %.2e" % worst)
        assert worst <= 1e-4
        xs, sts, _ = pl.qp(np.full((2, 129), -10.0), np.full((2, 129), 10.0), np.zeros((2, 3)), qp)
        assert np.all(sts == api.QP_TOO_LARGE)
    finally:
        pl.close()
