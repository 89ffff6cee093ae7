"""Independent optimum certificate for the planner's QPs.

TEST INFRASTRUCTURE - see oracle/__init__.py.

oracle/coneqp.py restates cvxopt's interior-point method from its published algorithm and is
"parity unpinned" (cvxopt cannot be installed offline).  This module shares no code and no
algorithm with it: it finds the exact minimiser of

    minimize 1/2 x'Px + q'x   s.t.  Gx <= h,  Ax = b

with an active-set method (scipy SLSQP for the first guess of the active set, then equality-
constrained KKT solves with constraints added / dropped until the set is stable) and then
*certifies* the point: for a convex QP the KKT conditions are sufficient, so a feasible x with
multipliers lambda >= 0 on its active rows and a vanishing stationarity residual IS the optimum,
however it was found.  The multipliers of the certificate come from a non-negative least-squares
fit (the constraint rows of Quadratic_planning are linearly dependent: duplicated corner rows and
opposing pin pairs, path_planning.py:133-166).

Used by tests/test_qp_optimum.py and tools/parity_report.py to record how far the coneqp answer
(which stops at cvxopt's default tolerances) is from the true optimum on every golden frame.
"""
import numpy as np
from scipy.linalg import qr
from scipy.optimize import minimize, nnls


def _eq_qp(P, q, Ga, ha, A, b):
    """Minimiser of the QP with the rows (Ga, ha) held at equality; least squares handles dependent rows."""
    n = P.shape[0]
    C = np.vstack([A, Ga]) if len(Ga) else A
    d = np.concatenate([b, ha]) if len(Ga) else b
    m = C.shape[0]
    K = np.zeros((n + m, n + m))
    K[:n, :n] = P
    K[:n, n:] = C.T
    K[n:, :n] = C
    rhs = np.concatenate([-q, d])
    Kp = np.linalg.pinv(K, rcond=1e-13)
    sol = Kp @ rhs
    for _ in range(3):                                   # iterative refinement: the KKT matrix mixes scales 1 .. 1e4
        sol = sol + Kp @ (rhs - K @ sol)
    return sol[:n], sol[n + A.shape[0]:]


def certificate(P, q, G, h, A, b, x, act_tol=1e-6):
    """KKT certificate of x: dict(primal = worst constraint violation, stationarity = |Px+q+G_A'lam+A'nu| with
    lam >= 0 fitted by NNLS on the rows within act_tol of their bound, objective)."""
    slack = h - G @ x
    act = slack <= act_tol
    g = P @ x + q
    M = np.hstack([G[act].T, A.T, -A.T])
    _, rnorm = nnls(M, -g, maxiter=50 * M.shape[1])
    return dict(primal=float(max(0.0, -slack.min(), np.abs(A @ x - b).max() if len(b) else 0.0)),
                stationarity=float(rnorm / max(1.0, np.linalg.norm(g))), n_active=int(act.sum()),
                objective=float(0.5 * x @ P @ x + q @ x))


def _presolve_pins(G, h, A, b):
    """Opposing inequality pairs (a'x <= c and -a'x <= -c: the start / end pins of Quadratic_planning,
    path_planning.py:148-160) are equalities; moving them to (A, b) leaves the active rows linearly independent, so the
    multipliers of the working set are unique.  Returns (G', h', A', b')."""
    key = {}
    for i in range(len(h)):
        key.setdefault((tuple(np.round(G[i], 12)), round(float(h[i]), 12)), []).append(i)
    pins, drop = [], set()
    for i in range(len(h)):
        if i in drop:
            continue
        for j in key.get((tuple(np.round(-G[i], 12)), round(float(-h[i]), 12)), []):
            if j not in drop and j != i:
                pins.append(i)
                drop.update((i, j))
                break
    keep = [i for i in range(len(h)) if i not in drop]
    if pins:
        A = np.vstack([A, G[pins]])
        b = np.concatenate([b, h[pins]])
    return G[keep], h[keep], A, b


def solve(P, q, G, h, A=None, b=None, x0=None, max_rounds=200):
    """Exact optimum by an active-set method.  Returns (x, info) with info = certificate(...) on the ORIGINAL
    problem + rounds; raises ValueError when no certified optimum is found (e.g. the problem is infeasible)."""
    P, q, G, h = (np.asarray(v, dtype=np.float64) for v in (P, q, G, h))
    n = P.shape[0]
    if A is None:
        A, b = np.zeros((0, n)), np.zeros(0)
    A, b = np.asarray(A, dtype=np.float64), np.asarray(b, dtype=np.float64).reshape(-1)
    G0, h0, A0, b0 = G, h, A, b
    G, h, A, b = _presolve_pins(G, h, A, b)
    if x0 is None:
        cons = [{"type": "ineq", "fun": lambda x: h - G @ x, "jac": lambda x: -G}]
        if len(b):
            cons.append({"type": "eq", "fun": lambda x: A @ x - b, "jac": lambda x: A})
        r = minimize(lambda x: 0.5 * x @ P @ x + q @ x, np.zeros(n), jac=lambda x: P @ x + q, constraints=cons,
                     method="SLSQP", options={"maxiter": 500, "ftol": 1e-12})
        x0 = r.x
    active = (h - G @ x0) <= 1e-5
    x = x0
    for rounds in range(max_rounds):
        idx = np.where(active)[0]
        if len(idx):                                     # keep the working set linearly independent of A and of itself
            rows = np.vstack([A, G[idx]])
            _, rr, piv = qr(rows.T, mode="economic", pivoting=True)
            d = np.abs(np.diag(rr))
            rank = int((d > 1e-10 * max(d.max(), 1.0)).sum())
            dependent = [p - len(b) for p in piv[rank:] if p >= len(b)]
            if dependent:
                active[idx[dependent]] = False
                idx = np.where(active)[0]
        x, lam = _eq_qp(P, q, G[idx], h[idx], A, b)
        slack = h - G @ x
        viol = slack < -1e-9
        viol[idx] = False
        if viol.any():                                   # blocked: add the most violated row and re-solve
            active[np.argmin(np.where(viol, slack, np.inf))] = True
            continue
        if len(lam) == 0 or lam.min() >= -1e-10:
            break
        active[idx[np.argmin(lam)]] = False              # drop the row with the most negative multiplier
    cert = certificate(P, q, G0, h0, A0, b0, x)
    cert["rounds"] = rounds + 1
    if cert["primal"] > 1e-7 or cert["stationarity"] > 1e-9:
        raise ValueError("no certified optimum (primal %.2e, stationarity %.2e)" % (cert["primal"], cert["stationarity"]))
    return x, cert
