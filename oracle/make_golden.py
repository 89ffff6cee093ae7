"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE - see oracle/__init__.py.   Usage:  python -m oracle.make_golden [A B C frenet st]

The reference (6Lackiu/EMplanner_Carla @ 1f115a5, /root/reference) is imported
through oracle/refshim.py (stub `carla`, `cvxopt` shim -> oracle/coneqp.py).
Everything stored with prefix `ref_` is the reference's own output; `qp_*`
values come from the reference's matrix construction solved by the restated
coneqp (parity unpinned at that boundary).  While generating, the script
asserts that oracle.sl_dp in mode "inv" reproduces the reference DP output
bit-for-bit, and stores the oracle's row indices (the reference does not
return them).
"""
import contextlib
import io
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, sl_dp, st_dp  # noqa: E402
from emplanner_carla_b200 import scenes  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_COMMIT = "1f115a5"


def _quiet(fn, *a, **kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)


DIVERGE_FACTOR = 2.0          # default of EmpQpParams.diverge_factor (include/emplanner.h)
_STATUS_CODE = {"optimal": 0}


def _sl_frame(job):
    """One golden frame: reference DP -> [::2] -> reference corridor -> reference Quadratic_planning (matrices by the
    unmodified reference, solve by oracle/coneqp.py through the cvxopt shim), then the same QP once more with the
    kernels' divergence exit."""
    from oracle import sl_qp, coneqp
    sc, b, kw = job
    pp, pu, _ = refshim.modules()
    K = int(sc["nobs"][b])
    os_, ol_ = list(sc["obs_s"][b, :K]), list(sc["obs_l"][b, :K])
    st = (sc["s0"][b], sc["l0"][b], sc["dl0"][b], sc["ddl0"][b])
    r = dict(b=b)
    t0 = time.perf_counter()
    rs, rl = _quiet(pp.DP_algorithm, list(os_), list(ol_), *st, **kw)
    r["ref_dp_seconds"] = time.perf_counter() - t0
    rs = np.array(rs, dtype=np.float64)
    rl = np.array(rl, dtype=np.float64).reshape(-1)
    s_i, l_i, rows_i = sl_dp.dp_algorithm(os_, ol_, *st, mode="inv", **kw)
    assert np.array_equal(rs, s_i) and np.array_equal(rl, l_i), "oracle(inv) != reference at frame %d" % b
    s_c, l_c, rows_c = sl_dp.dp_algorithm(os_, ol_, *st, mode="closed", **kw)
    r.update(ref_s=rs, ref_l=rl, rows=rows_i, closed_rows_equal=int(np.array_equal(rows_i, rows_c)),
             closed_n_equal=int(len(s_c) == len(rs)),
             closed_dl=float(np.abs(l_c - rl).max()) if len(s_c) == len(rs) else np.nan)
    ds_, dl_ = rs[::2], rl[::2]
    try:
        lmin, lmax = pp.cal_lmin_lmax(ds_, dl_, os_, ol_, 5, 5)
    except IndexError:
        r["cor_index_error"] = 1
        return r
    r["cor_index_error"] = 0
    r["lmin"], r["lmax"] = np.array(lmin), np.array(lmax)
    refshim.qp_log_enabled = True
    refshim.qp_log.clear()
    ql, qdl, qddl = pp.Quadratic_planning(lmin, lmax, st[1], st[2], st[3])
    log = refshim.qp_log[-1]
    res = log["result"]
    r.update(qp_l=np.array(ql), qp_dl=np.array(qdl), qp_ddl=np.array(qddl), qp_optimal=int(res["status"] == "optimal"),
             qp_iters=res["iterations"])
    tr = []
    rd = coneqp.qp(log["P"], log["q"], log["G"], log["h"], log["A"], log["b"], diverge_factor=DIVERGE_FACTOR, trace=tr)
    x = rd["x"]
    gaps = np.array([t["gap"] for t in tr])
    r.update(qpd_l=x[0::3].copy(), qpd_dl=x[1::3].copy(), qpd_ddl=x[2::3].copy(), qpd_iters=rd["iterations"],
             qpd_optimal=int(rd["status"] == "optimal"),
             gap_ratio_max=float(max(gaps[k] / gaps[:k].min() for k in range(1, len(gaps)))) if len(gaps) > 1 else 0.0)
    return r


def make_sl(tag, lattice, config_id, batch, extra=None, procs=None):
    import multiprocessing as mp
    refshim.modules()
    sc = scenes.sl_scenes(batch, lattice, config_id)
    if extra is not None:                                   # prepend hand-written frames
        for k in sc:
            sc[k] = np.concatenate([extra[k], sc[k]])
    B = len(sc["s0"])
    nmax = sl_dp.max_enriched(lattice.col, lattice.sample_s, lattice.sampling_res)
    nq = (nmax + 1) // 2
    nanq = lambda: np.full((B, nq), np.nan)
    out = dict(ref_n=np.zeros(B, np.int32), ref_s=np.full((B, nmax), np.nan), ref_l=np.full((B, nmax), np.nan),
               rows=np.zeros((B, lattice.col), np.int32), cor_index_error=np.zeros(B, np.int32),
               ref_lmin=nanq(), ref_lmax=nanq(), qp_l=nanq(), qp_dl=nanq(), qp_ddl=nanq(),
               qp_optimal=np.zeros(B, np.int32), qp_iters=np.zeros(B, np.int32),
               qpd_l=nanq(), qpd_dl=nanq(), qpd_ddl=nanq(), qpd_optimal=np.zeros(B, np.int32),
               qpd_iters=np.zeros(B, np.int32), gap_ratio_max=np.zeros(B), ref_dp_seconds=np.zeros(B),
               closed_rows_equal=np.zeros(B, np.int32), closed_n_equal=np.zeros(B, np.int32), closed_dl=np.zeros(B))
    kw = lattice.kwargs()
    procs = procs or len(os.sched_getaffinity(0))
    with mp.get_context("fork").Pool(procs) as pool:
        for i, r in enumerate(pool.imap_unordered(_sl_frame, [(sc, b, kw) for b in range(B)], chunksize=4)):
            b = r["b"]
            n = len(r["ref_s"])
            out["ref_n"][b] = n
            out["ref_s"][b, :n] = r["ref_s"]
            out["ref_l"][b, :n] = r["ref_l"]
            for k in ("rows", "ref_dp_seconds", "closed_rows_equal", "closed_n_equal", "closed_dl", "cor_index_error"):
                out[k][b] = r[k]
            if not r["cor_index_error"]:
                m = len(r["lmin"])
                out["ref_lmin"][b, :m], out["ref_lmax"][b, :m] = r["lmin"], r["lmax"]
                for k in ("qp_l", "qp_dl", "qp_ddl", "qpd_l", "qpd_dl", "qpd_ddl"):
                    out[k][b, :m] = r[k]
                for k in ("qp_optimal", "qp_iters", "qpd_optimal", "qpd_iters", "gap_ratio_max"):
                    out[k][b] = r[k]
            if i % 64 == 0:
                print(tag, i, "/", B, flush=True)
    meta = dict(ref_commit=REF_COMMIT, lattice=np.array([lattice.row, lattice.col, lattice.sample_s, lattice.sample_l,
                                                          lattice.sampling_res], dtype=np.float64),
                config_id=config_id, numpy=np.__version__, diverge_factor=DIVERGE_FACTOR)
    np.savez_compressed(os.path.join(GOLDEN, "sl_%s.npz" % tag), **sc, **out, **meta)
    ok = out["cor_index_error"] == 0
    print(tag, "done: optimal", int(out["qp_optimal"].sum()), "/", B, "index errors", int(out["cor_index_error"].sum()),
          "mean ref DP s/frame %.3f" % out["ref_dp_seconds"].mean(),
          "closed-form vs reference: row flips %d, length flips %d, worst |dl| %.3e" % (
              int((out["closed_rows_equal"] == 0).sum()), int((out["closed_n_equal"] == 0).sum()),
              float(np.nanmax(out["closed_dl"]))),
          "optimal disagreement with/without divergence exit:", int((out["qp_optimal"] != out["qpd_optimal"])[ok].sum()),
          "worst gap ratio on optimal frames %.4f" % float(out["gap_ratio_max"][ok & (out["qp_optimal"] == 1)].max()))


def make_frenet():
    pp, pu, _ = refshim.modules()
    xy = scenes.reference_line()
    th, k = pu.cal_heading_kappa([tuple(p) for p in xy])
    gp = [(xy[i, 0], xy[i, 1], th[i], k[i]) for i in range(len(xy))]
    rng = np.random.default_rng(scenes.BASE_SEED + 100)
    F = 12
    rec = dict(global_path=np.array(gp), ego=np.zeros((F, 2)), pre_idx=np.zeros(F, np.int32), match_idx=np.zeros(F, np.int32),
               local=np.zeros((F, 51, 4)), line=np.zeros((F, 51, 4)), smap=np.zeros((F, 51)),
               obs_xy=np.zeros((F, 4, 2)), obs_s=np.zeros((F, 4)), obs_l=np.zeros((F, 4)),
               vel=np.zeros((F, 2)), acc=np.zeros((F, 2)), deriv=np.zeros((F, 7)),
               path_s=np.zeros((F, 26)), path_l=np.zeros((F, 26)), start_sl=np.zeros((F, 2)),
               cart_raw=np.full((F, 27, 4), np.nan), cart=np.full((F, 27, 4), np.nan), cart_n=np.zeros(F, np.int32))
    for f in range(F):
        x = rng.uniform(60, 600)
        ego = (x, 8 * np.sin(x / 60) + rng.uniform(-1, 1))
        pre = int(x // 2) - int(rng.integers(0, 4))
        mi, _ = pu.find_match_points([ego], gp, False, pre)
        local = pu.sampling(mi[0], gp)
        line = pu.smooth_reference_line(local)
        smap = pu.cal_s_map_fun(line, ego)
        obs = [(ego[0] + d, 8 * np.sin((ego[0] + d) / 60) + rng.uniform(-3, 3)) for d in rng.uniform(8, 28, 4)]
        os_, ol_ = pu.cal_s_l_fun(obs, line, smap)
        v, a = (rng.uniform(3, 12), rng.uniform(-1, 1)), (rng.uniform(-1, 1), rng.uniform(-0.5, 0.5))
        der = pu.cal_s_l_deri_fun([ego], [v], [a], line, ego)
        bs, bl = pu.cal_s_l_fun([ego], line, smap)
        ps = bs[0] + np.arange(26) * 3.6
        pl = 0.8 * np.sin(ps / 9.0)
        from oracle import frenet as ofr
        raw = ofr.frenet_to_cartesian_raw(bs[0], bl[0], list(ps), list(pl), line, smap)
        cart = pp.frenet_2_x_y_theta_kappa(bs[0], bl[0], list(ps), list(pl), line, smap)
        rec["ego"][f] = ego; rec["pre_idx"][f] = pre; rec["match_idx"][f] = mi[0]
        rec["local"][f] = np.array(local); rec["line"][f] = np.array(line); rec["smap"][f] = smap
        rec["obs_xy"][f] = obs; rec["obs_s"][f] = os_; rec["obs_l"][f] = ol_
        rec["vel"][f] = v; rec["acc"][f] = a; rec["deriv"][f] = [d[0] for d in der]
        rec["path_s"][f] = ps; rec["path_l"][f] = pl; rec["start_sl"][f] = (bs[0], bl[0])
        rec["cart_n"][f] = len(cart)
        rec["cart"][f, :len(cart)] = np.array(cart)
        rec["cart_raw"][f, 0, :2] = raw[0]
        rec["cart_raw"][f, 1:len(raw)] = np.array(raw[1:])
    np.savez_compressed(os.path.join(GOLDEN, "frenet.npz"), ref_commit=REF_COMMIT, **rec)
    print("frenet done")


def make_st():
    _, _, spt = refshim.modules()
    F = 16
    sc = scenes.st_scenes(F, 3, k_fixed=None)
    sc["nobs"][:4] = [8, 3, 0, 5]                           # the rest: the generator's own 0..8
    act = np.arange(scenes.KMAX)[None, :] < sc["nobs"][:, None]
    for k in ("s_in", "s_out", "t_in", "t_out"):
        full = scenes.st_scenes(F, 3)[k]
        sc[k] = np.where(act, full, np.nan)
    out = dict(cost=np.zeros((F, 40, 16)), s_dot=np.zeros((F, 40, 16)), node=np.zeros((F, 40, 16), np.int32),
               min_row=np.zeros(F, np.int32), min_col=np.zeros(F, np.int32), ref_seconds=np.zeros(F))
    for f in range(F):
        t0 = time.perf_counter()
        loc = None
        try:
            spt.speed_DP(sc["s_in"][f], sc["s_out"][f], sc["t_in"][f], sc["t_out"][f], sc["v0"][f])
        except IndexError as e:                      # the reference's back-track always raises
            tb = e.__traceback__
            while tb is not None:
                if tb.tb_frame.f_code.co_name == "speed_DP":
                    loc = tb.tb_frame.f_locals
                tb = tb.tb_next
        out["ref_seconds"][f] = time.perf_counter() - t0
        out["cost"][f] = loc["dp_st_cost"]
        out["s_dot"][f] = loc["dp_st_s_dot"]
        out["node"][f] = loc["dp_st_node"].astype(np.int32)
        out["min_row"][f] = loc["min_row"]
        out["min_col"][f] = loc["min_col"]
        o = st_dp.speed_dp(sc["s_in"][f], sc["s_out"][f], sc["t_in"][f], sc["t_out"][f], sc["v0"][f])
        assert np.array_equal(o["node"], out["node"][f]) and o["min_row"] == out["min_row"][f]
        print("st", f, "ref seconds %.1f" % out["ref_seconds"][f], flush=True)
    np.savez_compressed(os.path.join(GOLDEN, "st_dp.npz"), ref_commit=REF_COMMIT, **sc, **out)
    print("st done")


def make_st_misc():
    """Small S-T / speed-draft helpers that do run in the reference."""
    _, pu, spt = refshim.modules()
    from oracle import st_misc
    rng = np.random.default_rng(scenes.BASE_SEED + 200)
    F, n = 16, 8
    rec = dict(dyn_s=rng.uniform(5, 40, (F, n)), dyn_l=rng.uniform(-6, 6, (F, n)), dyn_sd=rng.uniform(0, 8, (F, n)),
               dyn_ld=rng.uniform(-2, 2, (F, n)), stg=np.zeros((F, 4, n)),
               sc_in=rng.uniform(-3, 9, (F, 5)), sc_out=np.zeros((F, 2)),
               deri_head=rng.uniform(-1, 1, (F, n)), deri_kap=rng.uniform(-0.02, 0.02, (F, n)),
               deri_out=np.zeros((F, 3, 128)))
    rec["dyn_ld"][:, 1] *= 0.1                                 # some slow lateral movers
    for f in range(F):
        k = int(rng.integers(2, n + 1))
        rec["dyn_s"][f, k:] = np.nan
        rec["dyn_l"][f, k:] = np.nan
        rec["stg"][f] = np.array(spt.generate_st_graph(rec["dyn_s"][f], rec["dyn_l"][f], rec["dyn_sd"][f], rec["dyn_ld"][f]))
        assert np.array_equal(rec["stg"][f], np.array(st_misc.st_graph(rec["dyn_s"][f], rec["dyn_l"][f], rec["dyn_sd"][f],
                                                                        rec["dyn_ld"][f])), equal_nan=True)
        rec["sc_out"][f] = spt.calc_speed_planning_start_condition(*rec["sc_in"][f])
        rec["deri_out"][f] = np.array(pu.cal_dy_obs_deri(rec["dyn_l"][f], rec["dyn_sd"][f], rec["dyn_ld"][f],
                                                         rec["deri_head"][f], rec["deri_kap"][f]))
    # trajectory helpers on a curved path with NaN padding
    xs = np.arange(0, 60, 1.0)
    ys = 0.02 * xs ** 2
    tx, ty = np.append(xs, [np.nan] * 4), np.append(ys, [np.nan] * 4)
    th, kp = pu.cal_heading_kappa(list(zip(xs, ys)))
    idx2s = pu.trajectory_index2s(tx, ty)
    ss = np.append(np.linspace(0.7, 55, 20), [np.nan, np.nan])
    ll = np.append(0.5 * np.sin(np.linspace(0, 3, 20)), [0, 0])
    dl = np.append(0.05 * np.cos(np.linspace(0, 3, 20)), [0, 0])
    ddl = np.append(-0.01 * np.sin(np.linspace(0, 3, 20)), [0, 0])
    f2c = pu.Frenet2Cartesian(ss, ll, dl, ddl, xs, ys, np.array(th), np.array(kp), idx2s[:60])
    rec.update(traj_x=tx, traj_y=ty, traj_s=idx2s, path_heading=np.array(th), path_kappa=np.array(kp),
               f2c_s=ss, f2c_l=ll, f2c_dl=dl, f2c_ddl=ddl, f2c_out=np.array([a[:, 0] for a in f2c]))
    # single-edge S-T costs
    s_list, t_list = st_dp.S_LIST, st_dp.T_LIST
    sc = scenes.st_scenes(1, 77)
    obs = [sc[k][0] for k in ("s_in", "s_out", "t_in", "t_out")]
    sdot_tab = rng.uniform(2, 12, (40, 16))
    edges = [(0, 0, 5, 0), (3, 2, 5, 3), (0, 4, 7, 5), (20, 7, 12, 8), (39, 1, 30, 2), (10, 10, 10, 11)]
    rec["edge_idx"] = np.array(edges)
    rec["edge_cost"] = np.array([spt.CalcDpCost(a, b, c, d, *obs, 4000, 50, 100, 1e7, 6.5, s_list, t_list, sdot_tab)
                                 for a, b, c, d in edges])
    rec["edge_obs"] = np.array(obs)
    rec["edge_sdot"] = sdot_tab
    rec["edge_obs_cost"] = np.array([float(spt.CalcObsCost(3.0 + i, 1.0, 6.0 + 2 * i, 1.5, *obs, 1e7)) for i in range(6)])
    rec["coll_d"] = np.array([0.0, 0.3, 0.5, 0.9, 1.4999, 1.5, 2.0, -0.7])
    rec["coll_cost"] = np.array([spt.CalcCollisionCost(1e7, d) for d in rec["coll_d"]])
    np.savez_compressed(os.path.join(GOLDEN, "st_misc.npz"), ref_commit=REF_COMMIT, **rec)
    print("st_misc done")


def st_back_path(f, n_pad=3):
    """Per-frame synthetic planned path for the S-T back half: 1 m spaced parabola with frame-dependent curvature,
    NaN padded like the reference's fixed-size trajectory buffers (test_10.py:228-262)."""
    n_valid = 45 if f % 8 == 3 else 90                     # short paths: interp1d ValueError at :341
    xs = np.arange(0, 90, 1.0)
    ys = (0.001 + 0.0005 * (f % 7)) * xs ** 2 * (1 if f % 2 else -1)
    return xs[:n_valid], ys[:n_valid], n_pad + 90 - n_valid


def make_st_back(n_frames=128, n_dense=6):
    """S-T back half: generate_convex_space / increase_points / path_speed_merge outputs of the unmodified
    reference, everything speed_QP builds before its (failing) solver call, and the repaired solve (oracle)."""
    _, pu, spt = refshim.modules()
    from oracle import st_back
    sc = scenes.st_scenes(n_frames, 5, k_fixed=None)
    F = n_frames
    sc["t_out"][5::16, 0] = np.where(np.isnan(sc["t_out"][5::16, 0]), np.nan, 7.6)   # late exit: dp_speed_t[16], :391
    rec = dict(dp_s=np.zeros((F, 16)), dp_t=np.zeros((F, 16)), a0=_rng_a0(F), ref_bounds=np.full((F, 4, 16), np.nan),
               ref_convex_exc=np.zeros(F, np.int32), qp=np.full((F, 4, 17), np.nan), qp_optimal=np.zeros(F, np.int32),
               qp_iters=np.zeros(F, np.int32), qp_exc=np.zeros(F, np.int32), matrices_identical=np.zeros(F, np.int32))
    dense_frames, ref_dense, ref_traj, paths, path5s = [], [], [], [], []
    EXC = {IndexError: 0x20, ValueError: 0x10}

    class Captured(Exception):
        pass

    cap = {}

    def fake_qp(*a, **k):
        cap["args"] = [np.array(np.asarray(x)) for x in a]
        raise Captured()

    for f in range(F):
        o = st_dp.speed_dp(sc["s_in"][f], sc["s_out"][f], sc["t_in"][f], sc["t_out"][f], sc["v0"][f])
        dp_s, dp_t = o["dp_s"], o["dp_t"]
        rec["dp_s"][f], rec["dp_t"][f] = dp_s, dp_t
        xs, ys, n_pad = st_back_path(f)
        th, kp = pu.cal_heading_kappa(list(zip(xs, ys)))
        pad = [np.nan] * n_pad
        tx, ty = np.append(xs, pad), np.append(ys, pad)
        idx2s = np.asarray(pu.trajectory_index2s(tx, ty))
        kap = np.append(np.array(kp), [0.0] * n_pad)
        paths.append(np.array([idx2s, kap]))
        path5s.append(np.array([idx2s, tx, ty, np.append(th, pad), np.append(kp, pad)]))
        args = (dp_s, dp_t, idx2s, sc["s_in"][f], sc["s_out"][f], sc["t_in"][f], sc["t_out"][f], kap)
        try:
            cs = spt.generate_convex_space(*args)
            rec["ref_bounds"][f] = np.array(cs)
        except (IndexError, ValueError) as e:
            rec["ref_convex_exc"][f] = EXC[type(e)]
            cs = None
        try:
            mine = st_back.convex_space(*args)
            assert cs is not None and all(np.array_equal(a, b) for a, b in zip(cs, mine)), f
        except (IndexError, ValueError) as e:
            assert rec["ref_convex_exc"][f] == EXC[type(e)], f
        if cs is None:
            continue
        a0 = rec["a0"][f]
        old = spt.cvxopt.solvers.qp
        spt.cvxopt.solvers.qp = fake_qp
        try:
            spt.speed_QP(sc["v0"][f], a0, dp_s, dp_t, *cs)
            raise AssertionError("reference speed_QP returned")
        except Captured:
            m = st_back.speed_qp_matrices(sc["v0"][f], a0, dp_s, dp_t, *cs)
            same = all(np.array_equal(x, m[k].reshape(x.shape)) for x, k in zip(cap["args"], ("H", "f", "A", "b", "Aeq", "beq")))
            assert same, f
            rec["matrices_identical"][f] = 1
        except IndexError:
            rec["qp_exc"][f] = EXC[IndexError]
            try:
                st_back.speed_qp_matrices(sc["v0"][f], a0, dp_s, dp_t, *cs)
                raise AssertionError("oracle did not raise")
            except IndexError:
                pass
            continue
        finally:
            spt.cvxopt.solvers.qp = old
        r = st_back.speed_qp(sc["v0"][f], a0, dp_s, dp_t, *cs)
        rec["qp"][f] = np.array(r[:4])
        rec["qp_optimal"][f] = int(r[4]["status"] == "optimal")
        rec["qp_iters"][f] = r[4]["iterations"]
        if len(dense_frames) < n_dense:
            ip = spt.increase_points(*r[:4])
            assert all(np.array_equal(a, b) for a, b in zip(ip, st_back.increase_points(*r[:4])))
            mg = spt.path_speed_merge(*ip, 12.5 + f, *path5s[-1])
            assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(mg, st_back.path_speed_merge(*ip, 12.5 + f, *path5s[-1])))
            dense_frames.append(f)
            ref_dense.append(np.array(ip))
            ref_traj.append(np.array(mg))
        print("st_back", f, "q=%d" % m["qp_size"], r[4]["status"], r[4]["iterations"], flush=True)
    rec.update(path=np.array(paths), path5=np.array(path5s), dense_frames=np.array(dense_frames),
               ref_dense=np.array(ref_dense), ref_traj=np.array(ref_traj), current_time=12.5 + np.arange(F))
    np.savez_compressed(os.path.join(GOLDEN, "st_back.npz"), ref_commit=REF_COMMIT, **sc, **rec)
    print("st_back done: %d convex ok, %d QPs (%d optimal)" % (int((rec["ref_convex_exc"] == 0).sum()),
          int(rec["matrices_identical"].sum()), int(rec["qp_optimal"].sum())))


def make_speed_stub():
    """planner/speed_planning.py: what runs (cal_start_cost, cal_obs_cost) and how the entry point fails."""
    sp = refshim.speed_planning_module()
    rng = np.random.default_rng(scenes.BASE_SEED + 400)
    n = 24
    rec = dict(begin=np.stack([rng.uniform(0, 20, n), rng.uniform(0, 1, n), rng.uniform(0, 0.2, n)], 1),
               end=np.zeros((n, 2)), sample_t=rng.uniform(0.5, 2.0, n), obs_s=rng.uniform(2, 40, (n, 3)),
               obs_t=rng.uniform(0, 6, (n, 3)), cost=np.zeros(n))
    rec["end"][:, 0] = rec["begin"][:, 0] + rng.uniform(3, 12, n)
    rec["end"][:, 1] = rec["begin"][:, 1] + rng.uniform(0.5, 3, n)
    rec["obs_s"][:6] = rec["begin"][:6, :1] + rng.uniform(0, 3, (6, 3))       # some inside the 3 m / 5 m zones
    rec["obs_t"][:6] = rec["begin"][:6, 1:2] + rng.uniform(0, 2, (6, 3))
    w = (1e10, [300, 2000, 10000], 20)                                         # the stub's own weights (:83-85)
    for i in range(n):
        rec["cost"][i] = float(sp.cal_start_cost(list(rec["obs_s"][i]), list(rec["obs_t"][i]), *rec["begin"][i],
                                                 *rec["end"][i], rec["sample_t"][i], *w))
    d2 = np.array([30.0, 24.0, 16.0, 9.0, 4.0, 50.0, 20.0, 20.0, 20.0, 20.0])
    rec["obs_cost_d2"] = d2
    rec["obs_cost"] = np.array([sp.cal_obs_cost(1e10, d2.reshape(10, 1)), sp.cal_obs_cost(7.0, np.full((10, 1), 25.0)),
                                sp.cal_obs_cost(3.0, d2.reshape(10, 1), danger_dis=1, safe_dis=6)])
    exc = []
    for obs_pos in ([12.0], []):
        try:
            sp.construct_ST_graph_and_planning([0.0], 0, 5, 0, obs_pos, [3.0] * len(obs_pos))
            exc.append("returned")
        except Exception as e:
            exc.append(type(e).__name__)
    rec["entry_point_exceptions"] = np.array(exc)
    np.savez_compressed(os.path.join(GOLDEN, "speed_stub.npz"), ref_commit=REF_COMMIT, **rec)
    print("speed_stub done:", exc)


def motion_planning_messages(n_frames=40):
    """Recorded inputs of test_9.motion_planning (the tuple the main loop sends, test_9.py:390-392) for an ego car
    driving along the synthetic global route: static obstacles ahead (some frames none, some beyond the 30 m gate of
    test_9.py:116), a slower dynamic obstacle on every 4th frame (test_9.py:137-169)."""
    _, pu, _ = refshim.modules()
    xy = scenes.reference_line()
    th, k = pu.cal_heading_kappa([tuple(p) for p in xy])
    route = [(float(xy[i, 0]), float(xy[i, 1]), float(th[i]), float(k[i])) for i in range(len(xy))]
    rng = np.random.default_rng(scenes.BASE_SEED + 900)
    on_route = lambda x, off: (float(x), float(8 * np.sin(x / 60) + off))
    msgs = []
    x0 = 60.0
    first_match, _ = pu.find_match_points([on_route(x0, 0.0)], route, True, 0)
    match = list(first_match)
    for f in range(n_frames):
        x = x0 + 6.5 * f
        ego = on_route(x, rng.uniform(-0.4, 0.4))
        heading = float(np.arctan2(8 / 60 * np.cos(x / 60), 1.0))
        speed = rng.uniform(7.0, 12.0)
        v = (float(speed * np.cos(heading) + rng.uniform(-0.2, 0.2)), float(speed * np.sin(heading) + rng.uniform(-0.2, 0.2)))
        a = (float(rng.uniform(-1, 1)), float(rng.uniform(-0.5, 0.5)))
        pred = (ego[0] + v[0] * 0.2, ego[1] + v[1] * 0.2)
        K = int(rng.integers(0, 4))
        far = f % 9 == 5                                      # nearest obstacle beyond 30 m: the list is ignored
        stat = []
        for d in np.sort(rng.uniform(31.0, 40.0, K) if far else rng.uniform(10.0, 28.0, K)):
            ox, oy = on_route(x + d, rng.uniform(-3.0, 3.0))
            stat.append((ox, oy, float(np.hypot(ox - ego[0], oy - ego[1]))))
        stat.sort(key=lambda t: t[2])
        dyn = []
        if f % 4 == 2:
            gap = float(rng.uniform(18.0, 30.0))
            ox, oy = on_route(x + gap, 0.0)
            dyn.append((ox, oy, gap, float(rng.uniform(2.0, 5.0))))
        msgs.append((stat, dyn, ego, pred, v, a, route, list(match)))
    return msgs


def make_motion_planning(n_frames=40):
    """Replay the UNMODIFIED source text of test_9.motion_planning (test_9.py:92-221) against a fake Pipe on the
    reference's own planner modules and store every conn.send payload.  Two runs: with the cvxopt shim behaving like
    cvxopt (`ref_*`), and with the kernels' documented divergence exit (`refd_*`); they differ only on frames whose
    path QP ends 'unknown' (infeasible corridor), where the former is not numerically reproducible."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import motion_planning_caller as mpc
    pp, pu, _ = refshim.modules()
    text = mpc.reference_function_text()
    msgs = motion_planning_messages(n_frames)
    F = len(msgs)
    rec = dict(route=np.array(msgs[0][6]), first_match=np.array(msgs[0][7], np.int32),
               ego=np.array([m[2] for m in msgs]), pred=np.array([m[3] for m in msgs]),
               vel=np.array([m[4] for m in msgs]), acc=np.array([m[5] for m in msgs]),
               n_static=np.array([len(m[0]) for m in msgs], np.int32), static=np.full((F, 3, 3), np.nan),
               n_dynamic=np.array([len(m[1]) for m in msgs], np.int32), dynamic=np.full((F, 1, 4), np.nan))
    for f, m in enumerate(msgs):
        for j, o in enumerate(m[0]):
            rec["static"][f, j] = o
        for j, o in enumerate(m[1]):
            rec["dynamic"][f, j] = o
    import hashlib
    rec["caller_sha256"] = hashlib.sha256(text.encode()).hexdigest()
    for tag, df in (("ref", None), ("refd", DIVERGE_FACTOR)):
        refshim.diverge_factor = df
        refshim.qp_log_enabled = True
        refshim.qp_log.clear()
        calls = []
        sent = _quiet(mpc.replay, text, mpc.RecordingProxy(pu, calls), mpc.RecordingProxy(pp, calls), msgs)
        refshim.diverge_factor = None
        if tag == "refd":                                     # every planner call of the caller, for the forced replay
            import json
            rec["refd_calls_json"] = np.array(json.dumps(calls))
        assert len(sent) == F and len(refshim.qp_log) == 3 * F
        rec[tag + "_cart"] = np.full((F, 28, 4), np.nan)
        rec[tag + "_cart_n"] = np.zeros(F, np.int32)
        rec[tag + "_path_s"] = np.full((F, 27), np.nan)
        rec[tag + "_path_l"] = np.full((F, 27), np.nan)
        rec[tag + "_path_n"] = np.zeros(F, np.int32)
        rec[tag + "_match"] = np.zeros(F, np.int32)
        rec[tag + "_qp_status"] = np.zeros((F, 3), np.int32)   # 1 = optimal, for the frame's QP#1, QP#2, QP#3
        rec[tag + "_qp_iters"] = np.zeros((F, 3), np.int32)
        for f, (cart, match, ps, pl) in enumerate(sent):
            rec[tag + "_cart_n"][f] = len(cart)
            rec[tag + "_cart"][f, :len(cart)] = np.array(cart)
            rec[tag + "_path_n"][f] = len(ps)
            rec[tag + "_path_s"][f, :len(ps)] = ps
            rec[tag + "_path_l"][f, :len(pl)] = np.array(pl, dtype=np.float64).reshape(-1)
            rec[tag + "_match"][f] = match[0]
            for q in range(3):
                r = refshim.qp_log[3 * f + q]["result"]
                rec[tag + "_qp_status"][f, q] = int(r["status"] == "optimal")
                rec[tag + "_qp_iters"][f, q] = r["iterations"]
    # Q10 sensitivity of the reference itself: the same frames with the ego / plan-start positions moved by 1e-9 m.
    # `int(end_s - start_s)` (path_planning.py:405) turns the last bits of plan_start_s into a path that is one point
    # longer or shorter, so the reference does not reproduce its own output length under input noise of that size.
    moved = []
    for m in msgs:
        m = list(m)
        m[2] = (m[2][0] + 1e-9, m[2][1] - 1e-9)
        m[3] = (m[3][0] + 1e-9, m[3][1] - 1e-9)
        moved.append(tuple(m))
    refshim.diverge_factor = DIVERGE_FACTOR
    sent = _quiet(mpc.replay, text, pu, pp, moved)
    refshim.diverge_factor = None
    rec["refd_moved_path_n"] = np.array([len(p[2]) for p in sent], np.int32)
    rec["refd_moved_path_l"] = np.full((F, 27), np.nan)
    rec["refd_moved_path_s"] = np.full((F, 27), np.nan)
    for f, p in enumerate(sent):
        rec["refd_moved_path_s"][f, :len(p[2])] = p[2]
        rec["refd_moved_path_l"][f, :len(p[3])] = np.array(p[3], dtype=np.float64).reshape(-1)
    np.savez_compressed(os.path.join(GOLDEN, "motion_planning.npz"), ref_commit=REF_COMMIT, **rec)
    print("reference vs itself with inputs moved by 1e-9 m: path length differs on %d of %d frames"
          % (int((rec["refd_moved_path_n"] != rec["refd_path_n"]).sum()), F))
    unk = (rec["ref_qp_status"][:, 1] == 0)
    same = [f for f in range(F) if np.array_equal(rec["ref_cart"][f], rec["refd_cart"][f], equal_nan=True)]
    print("motion_planning done: %d frames, path QP 'unknown' on %d, payload identical with/without the divergence exit on %d"
          % (F, int(unk.sum()), len(same)))


def _rng_a0(n):
    return np.random.default_rng(scenes.BASE_SEED + 300).uniform(-1.0, 1.0, n)


def main(argv):
    what = argv or ["A", "B", "C", "frenet", "st", "st_misc", "st_back", "speed_stub", "motion_planning"]
    os.makedirs(GOLDEN, exist_ok=True)
    if "A" in what:
        make_sl("A", scenes.LATTICE_A, 0, 512)
    if "B" in what:
        # frame 0 = BASELINE.json configs[0]: single ego frame, 60x9 lattice, 3 static obstacles
        nanpad = lambda v: np.array([v + [np.nan] * (scenes.KMAX - len(v))])
        extra = dict(obs_s=nanpad([22.0, 47.0, 71.0]), obs_l=nanpad([0.6, -1.2, 1.8]), nobs=np.array([3], np.int32),
                     s0=np.array([0.5]), l0=np.array([0.1]), dl0=np.array([0.01]), ddl0=np.array([0.0]))
        make_sl("B", scenes.LATTICE_B, 1, 127, extra=extra)
    if "C" in what:
        make_sl("C", scenes.LATTICE_C, 2, 1024)
    if "frenet" in what:
        make_frenet()
    if "st" in what:
        make_st()
    if "st_misc" in what:
        make_st_misc()
    if "st_back" in what:
        make_st_back()
    if "speed_stub" in what:
        make_speed_stub()
    if "motion_planning" in what:
        make_motion_planning()


if __name__ == "__main__":
    main(sys.argv[1:])
