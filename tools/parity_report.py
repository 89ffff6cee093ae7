"""Assemble profiles/parity_r2.json: every observed parity number of the S-L path in one file.

CPU part (this script): golden-set statistics stored by oracle/make_golden.py (closed-form DP vs the unmodified
reference: row flips, length flips, worst |dl|; gap monotonicity of converging QP solves), how reproducible the QP
iterate of an 'unknown' (infeasible-corridor) frame is under cvxopt's rule and under the kernels' divergence exit
(oracle/coneqp.py vs the numpy model of the kernel's arithmetic, tests/qp_kernel_model.py), and the distance of
every 'optimal' golden QP answer to the independently certified optimum (oracle/qp_optimum.py).
GPU part: the numbers the `-m gpu` tests wrote to gpurun_out/parity/*.json on the B200 (kernel vs golden).

usage: python tools/parity_report.py [--optimum-frames N] [--out profiles/parity_r2.json]
"""
import argparse
import glob
import json
import multiprocessing as mp
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.filterwarnings("ignore")


def _frame_qp(g, b):
    from oracle import sl_qp
    m = int((g["ref_n"][b] + 1) // 2)
    return m, sl_qp.qp_matrices(g["ref_lmin"][b, :m], g["ref_lmax"][b, :m], g["l0"][b], g["dl0"][b], g["ddl0"][b])


def _optimum_job(job):
    from oracle import qp_optimum
    tag, b = job
    g = np.load(os.path.join(ROOT, "tests", "golden", "sl_%s.npz" % tag))
    m, (P, q, G, h, A, bb) = _frame_qp(g, b)
    x = np.stack([g["qp_l"][b, :m], g["qp_dl"][b, :m], g["qp_ddl"][b, :m]], axis=1).reshape(-1)
    try:
        xs, cert = qp_optimum.solve(P, q, G, h, A, bb)
    except ValueError as e:
        return dict(frame=int(b), failed=str(e))
    obj = 0.5 * x @ P @ x + q @ x
    return dict(frame=int(b), dist_l=float(np.abs(x[0::3] - xs[0::3]).max()),
                rel_obj=float(abs(obj - cert["objective"]) / max(1.0, abs(cert["objective"]))),
                stationarity=cert["stationarity"], primal=cert["primal"])


def _unknown_job(job):
    from oracle import coneqp
    import qp_kernel_model as km
    tag, b, df = job
    g = np.load(os.path.join(ROOT, "tests", "golden", "sl_%s.npz" % tag))
    m, (P, q, G, h, A, bb) = _frame_qp(g, b)
    r = coneqp.qp(P, q, G, h, A, bb, diverge_factor=df)
    x, status, it = km.solve_qp(g["ref_lmin"][b, :m], g["ref_lmax"][b, :m], [g["l0"][b], g["dl0"][b], g["ddl0"][b]],
                                diverge_factor=df)
    return dict(frame=int(b), rms=float(np.sqrt(np.mean((x[:, 0] - r["x"][0::3]) ** 2))),
                iters_equal=bool(it == r["iterations"]), iters=int(r["iterations"]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--optimum-frames", type=int, default=128, help="optimal golden frames per lattice checked against the certified optimum")
    ap.add_argument("--unknown-frames", type=int, default=64)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "parity_r2.json"))
    args = ap.parse_args()
    rep = {"reference_commit": "1f115a5", "what": __doc__.split("\n")[0], "lattices": {}}
    pool = mp.get_context("fork").Pool(len(os.sched_getaffinity(0)))
    for tag in "ABC":
        g = np.load(os.path.join(ROOT, "tests", "golden", "sl_%s.npz" % tag))
        ok = g["cor_index_error"] == 0
        opt = ok & (g["qp_optimal"] == 1)
        unk = ok & (g["qp_optimal"] == 0)
        L = {"golden_frames": int(len(ok)), "corridor_index_errors": int((~ok).sum()),
             "qp_optimal": int(opt.sum()), "qp_unknown": int(unk.sum()),
             "dp_closed_form_vs_reference": {"row_flips": int((g["closed_rows_equal"] == 0).sum()),
                                             "length_flips": int((g["closed_n_equal"] == 0).sum()),
                                             "worst_abs_dl_m": float(np.nanmax(g["closed_dl"]))},
             "reference_dp_seconds_per_frame_mean": float(g["ref_dp_seconds"].mean()),
             "converging_solves_worst_gap_ratio": float(g["gap_ratio_max"][opt].max()),
             "diverge_factor": float(g["diverge_factor"]),
             "optimal_frames_changed_by_divergence_exit": int(((g["qp_optimal"] != g["qpd_optimal"]) & ok).sum()),
             "unknown_frames_iterations_cvxopt_rule_max": int(g["qp_iters"][unk].max()) if unk.any() else 0,
             "unknown_frames_iterations_divergence_exit_max": int(g["qpd_iters"][unk].max()) if unk.any() else 0}
        idx = np.where(opt & (g["nobs"] > 0))[0]
        idx = idx[np.linspace(0, len(idx) - 1, min(args.optimum_frames, len(idx))).astype(int)]
        res = pool.map(_optimum_job, [(tag, int(b)) for b in idx])
        good = [r for r in res if "failed" not in r]
        L["coneqp_vs_certified_optimum"] = {
            "frames": len(res), "certified": len(good),
            "worst_dist_l_m": max(r["dist_l"] for r in good), "median_dist_l_m": float(np.median([r["dist_l"] for r in good])),
            "worst_rel_objective_gap": max(r["rel_obj"] for r in good),
            "worst_certificate_stationarity": max(r["stationarity"] for r in good),
            "note": "coneqp stops at cvxopt's default tolerances (reltol 1e-6); the certified optimum is exact"}
        uidx = np.where(unk)[0][:args.unknown_frames]
        for name, df in (("cvxopt_rule", None), ("divergence_exit", float(g["diverge_factor"]))):
            res = pool.map(_unknown_job, [(tag, int(b), df) for b in uidx])
            L["unknown_frames_oracle_vs_kernel_model_" + name] = {
                "frames": len(res), "worst_rms_l_m": max(r["rms"] for r in res) if res else 0.0,
                "median_rms_l_m": float(np.median([r["rms"] for r in res])) if res else 0.0,
                "iteration_mismatches": sum(not r["iters_equal"] for r in res)}
        rep["lattices"][tag] = L
        print(tag, json.dumps(L, indent=1), flush=True)
    pool.close()
    gm = np.load(os.path.join(ROOT, "tests", "golden", "motion_planning.npz"))
    rep["motion_planning_replay_golden"] = {"frames": int(len(gm["ego"])), "caller_sha256": str(gm["caller_sha256"]),
                                            "path_qp_unknown_frames": int((gm["ref_qp_status"][:, 1] == 0).sum())}
    rep["gpu"] = {}
    for f in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "parity", "*.json"))):
        rep["gpu"][os.path.splitext(os.path.basename(f))[0]] = json.load(open(f))
    if not rep["gpu"]:
        rep["gpu"] = "no gpurun_out/parity/*.json found: run `pytest -m gpu` on the B200 first"
    with open(args.out, "w") as f:
        json.dump(rep, f, indent=1, sort_keys=True)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
