"""Oracle: corridor + QP path smoother (numpy float64, one frame at a time).

TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates reference planner/path_planning.py:
  cal_lmin_lmax       :222-273 -> corridor()
  Quadratic_planning  :78-219  -> qp_matrices() + quadratic_planning()
The solve goes through oracle/coneqp.py (parity unpinned at the cvxopt boundary).
"""
import math
import numpy as np
from oracle import coneqp


def corridor(dp_s, dp_l, obs_s, obs_l, obs_length, obs_width):
    """cal_lmin_lmax.  Raises IndexError exactly where the reference does
    (index + 2 beyond the path, path_planning.py:240-241,265-272)."""
    dp_s = np.asarray(dp_s, dtype=np.float64)
    n = len(dp_s)
    lmin = -10 * np.ones(n)
    lmax = 10 * np.ones(n)
    offset = 2
    for o in range(len(obs_s)):
        lo = int(np.argmin(np.abs(dp_s - (obs_s[o] - obs_length / 2)))) + offset
        hi = int(np.argmin(np.abs(dp_s - (obs_s[o] + obs_length / 2)))) + offset
        ctr = int(np.argmin(np.abs(dp_s - obs_s[o])))
        if dp_l[ctr] < obs_l[o]:
            for j in range(lo, hi + 1):
                lmax[j] = min(lmax[j], obs_l[o] - obs_width / 2)   # IndexError if j >= n
        else:
            for j in range(lo, hi + 1):
                lmin[j] = max(lmin[j], obs_l[o] + obs_width / 2)
    return lmin, lmax


def qp_matrices(l_min, l_max, start_l, start_dl, start_ddl, dp_sampling_res=2,
                w_cost_l=1000, w_cost_dl=10000, w_cost_ddl=3000, w_cost_dddl=150, w_cost_centre=250,
                w_cost_end_l=40, w_cost_end_dl=40, w_cost_end_ddl=40, host_d1=3, host_d2=3, host_w=3):
    """Dense P,q,G,h,A,b of Quadratic_planning (path_planning.py:103-205).

    Variable order (l,dl,ddl) per point.  Notes on the reference's construction:
      * `w_cost_dl*(H_DL.T @ H_L)` is identically zero (:193) -> dl is unweighted;
      * ds = dp_sampling_res regardless of the real spacing (:108);
      * the 8 corner rows use l_max[min(i+ceil(d1/ds), n-1)] and
        l_min[max(i-ceil(d2/ds), 0)] (:126-141), rows 1,3,5,7 being redundant;
      * start/end pins are opposing inequality pairs, other boxes +-1e5 (:145-166).
    """
    l_min = np.asarray(l_min, dtype=np.float64)
    l_max = np.asarray(l_max, dtype=np.float64)
    n = len(l_min)
    ds = dp_sampling_res
    A = np.zeros((2 * n - 2, 3 * n))
    b = np.zeros(2 * n - 2)
    e_rows = np.array([[1, ds, ds ** 2 / 3, -1, 0, ds ** 2 / 6],
                       [0, 1, ds / 2, 0, -1, ds / 2]], dtype=np.float64)
    for i in range(n - 1):
        A[2 * i:2 * i + 2, 3 * i:3 * i + 6] = e_rows
    Gc = np.zeros((8 * n, 3 * n))
    hc = np.zeros(8 * n)
    sgn = [1, 1, 1, 1, -1, -1, -1, -1]
    lever = [host_d1, host_d1, -host_d2, -host_d2, -host_d1, -host_d1, host_d2, host_d2]
    fi = math.ceil(host_d1 / ds)
    bi = math.ceil(host_d2 / ds)
    for i in range(n):
        for r in range(8):
            Gc[8 * i + r, 3 * i] = sgn[r]
            Gc[8 * i + r, 3 * i + 1] = lever[r]
        up = l_max[min(i + fi, n - 1)]
        lo = l_min[max(i - bi, 0)]
        hc[8 * i:8 * i + 8] = [up - host_w / 2, up + host_w / 2, up - host_w / 2, up + host_w / 2,
                               -lo + host_w / 2, -lo - host_w / 2, -lo + host_w / 2, -lo - host_w / 2]
    lb = -100000.0 * np.ones(3 * n)
    ub = 100000.0 * np.ones(3 * n)
    lb[0:3] = ub[0:3] = [start_l, start_dl, start_ddl]
    lb[3 * n - 3:] = 0.0
    ub[3 * n - 3:] = 0.0
    G = np.concatenate((Gc, np.identity(3 * n), -np.identity(3 * n)))
    h = np.concatenate((hc, ub, -lb))
    # objective
    diag = np.zeros(3 * n)
    diag[0::3] = w_cost_l + w_cost_centre
    diag[2::3] = w_cost_ddl
    H = np.diag(diag)
    for i in range(n - 1):                   # w_dddl * (ddl_{i+1} - ddl_i)^2
        H[3 * i + 2, 3 * i + 2] += w_cost_dddl
        H[3 * i + 5, 3 * i + 5] += w_cost_dddl
        H[3 * i + 2, 3 * i + 5] -= w_cost_dddl
        H[3 * i + 5, 3 * i + 2] -= w_cost_dddl
    H[3 * n - 3, 3 * n - 3] += w_cost_end_l
    H[3 * n - 2, 3 * n - 2] += w_cost_end_dl
    H[3 * n - 1, 3 * n - 1] += w_cost_end_ddl
    P = 2 * H
    q = np.zeros(3 * n)
    q[0::3] = w_cost_centre * (-2 * ((l_min + l_max) / 2))
    return P, q, G, h, A, b


def quadratic_planning(l_min, l_max, start_l, start_dl, start_ddl, diverge_factor=None, **kw):
    """Quadratic_planning (path_planning.py:78-219).  Returns (l, dl, ddl, info).
    diverge_factor: None = cvxopt's behaviour; a number = the kernels' documented early exit (oracle/coneqp.py)."""
    P, q, G, h, A, b = qp_matrices(l_min, l_max, start_l, start_dl, start_ddl, **kw)
    res = coneqp.qp(P, q, G, h, A, b, diverge_factor=diverge_factor)
    x = res["x"]
    return x[0::3].copy(), x[1::3].copy(), x[2::3].copy(), res


def midpoint_path(dp_s, qp_l):
    """Mid-point averaging done by the caller between QP and Frenet->Cartesian (test_9.py:204-210)."""
    path_s = [dp_s[0]]
    path_l = [qp_l[0]]
    for i in range(1, len(qp_l)):
        path_s.append((dp_s[i] + dp_s[i - 1]) / 2)
        path_l.append((qp_l[i] + qp_l[i - 1]) / 2)
    path_s.append(dp_s[-1])
    path_l.append(qp_l[-1])
    return np.array(path_s), np.array(path_l)
