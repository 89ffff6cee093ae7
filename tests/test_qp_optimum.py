"""Independent check of the QP answers (no code or algorithm shared with oracle/coneqp.py).

oracle/qp_optimum.py finds the exact minimiser of the reference's own (P, q, G, h, A, b) with an active-set method
and certifies it through the KKT conditions (sufficient for a convex QP).  The coneqp restatement - "parity unpinned"
because cvxopt cannot be installed here - must land where an interior-point method stopped at cvxopt's default
tolerances (abstol 1e-7, reltol 1e-6, feastol 1e-7) is expected to land: feasible to feastol, objective within
reltol of the certified optimum, l within 2e-3 m of it (up to 1.6e-3 m, median 1e-7)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import qp_optimum, sl_qp


@pytest.mark.parametrize("tag,count", [("A", 10), ("C", 6)])
def test_coneqp_answers_sit_at_the_certified_optimum(tag, count):
    g = load_golden("sl_%s.npz" % tag)
    idx = np.where((g["cor_index_error"] == 0) & (g["qp_optimal"] == 1) & (g["nobs"] > 0))[0]
    idx = idx[np.linspace(0, len(idx) - 1, count).astype(int)]
    dist = []
    for b in idx:
        m = int((g["ref_n"][b] + 1) // 2)
        P, q, G, h, A, bb = sl_qp.qp_matrices(g["ref_lmin"][b, :m], g["ref_lmax"][b, :m], g["l0"][b], g["dl0"][b], g["ddl0"][b])
        x_star, cert = qp_optimum.solve(P, q, G, h, A, bb)
        assert cert["primal"] <= 1e-7 and cert["stationarity"] <= 1e-9
        x = np.stack([g["qp_l"][b, :m], g["qp_dl"][b, :m], g["qp_ddl"][b, :m]], axis=1).reshape(-1)
        obj = 0.5 * x @ P @ x + q @ x
        assert (h - G @ x).min() >= -1e-6 and np.abs(A @ x - bb).max() <= 1e-6          # feasible to feastol
        assert abs(obj - cert["objective"]) <= 2e-6 * max(1.0, abs(cert["objective"]))      # reltol
        dist.append(np.abs(x[0::3] - x_star[0::3]).max())
    print("lattice", tag, "max |l_coneqp - l*| over %d frames: %.2e (median %.1e)" % (count, max(dist), np.median(dist)))
    assert max(dist) <= 2e-3


def test_certificate_rejects_a_wrong_point():
    g = load_golden("sl_A.npz")
    b = int(np.where((g["cor_index_error"] == 0) & (g["qp_optimal"] == 1) & (g["nobs"] > 0))[0][0])
    m = int((g["ref_n"][b] + 1) // 2)
    P, q, G, h, A, bb = sl_qp.qp_matrices(g["ref_lmin"][b, :m], g["ref_lmax"][b, :m], g["l0"][b], g["dl0"][b], g["ddl0"][b])
    x = np.stack([g["qp_l"][b, :m], g["qp_dl"][b, :m], g["qp_ddl"][b, :m]], axis=1).reshape(-1)
    good = qp_optimum.certificate(P, q, G, h, A, bb, x, act_tol=1e-5)
    bad = qp_optimum.certificate(P, q, G, h, A, bb, x + 1e-3, act_tol=1e-5)
    assert bad["stationarity"] > 100 * max(good["stationarity"], 1e-12) or bad["primal"] > 1e-4
