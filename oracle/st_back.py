"""Oracle: S-T back half - convex space, speed QP, densify, path/speed merge (numpy float64).

TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates reference planner/speed_planning_test.py
  generate_convex_space :308-400 -> convex_space()     (runs in the reference; pinned by tests/golden/st_back.npz)
  speed_QP              :403-511 -> speed_qp_matrices() + speed_qp()
  increase_points       :514-562 -> increase_points()  (runs in the reference on NaN-terminated profiles; pinned)
  path_speed_merge      :565-611 -> path_speed_merge() (runs in the reference; pinned)

speed_QP cannot run in the reference: it hands cvxopt an equality matrix of the wrong shape (`Aeq` is built as
3q x (2q-2) and passed untransposed, :435,:502 - its own comment at :434 says "Aeq'*X == beq") and never passes
`lb`/`ub` (:438-439 alias one array, :502 omits both).  Everything it builds BEFORE the solver call is pinned
bit-for-bit (H, f, the no-reversing rows A/b, Aeq, beq: captured from the unmodified reference in
oracle/make_golden.py).  The solve itself follows the documented repair

  R1  the equality constraint is Aeq' x = beq;
  R2  `lb` and `ub` are two arrays; every FINITE bound becomes one inequality row (-x_k <= -lb_k, x_k <= ub_k),
      infinite bounds are dropped - the MATLAB call this draft was ported from, quadprog(H,f,A,b,Aeq',beq,lb,ub);
      the start pins lb = ub stay an opposing inequality pair, as in Quadratic_planning (path_planning.py:148-166);

  R3  (opt-in, `repair_full_profile`) a DP profile that fills all 16 columns ends at index 15.  The draft sets
      `dp_speed_end = 16` / `dp_speed_end_index = len(dp_speed_s)` when it finds no NaN (:330,:416) - the 1-based count of
      the MATLAB original - while the NaN branch sets the 0-based index of the last valid entry (`i - 1`); the count is
      then used as an index (`dp_speed_s[16]`, :428; `dp_speed_t[16]`, :391) and raises IndexError.  With the default
      reference_speed = 50 m/s ~75 % of the synthetic profiles are full, so without R3 most frames never reach the QP;

and uses oracle/coneqp.py for the QP (parity unpinned at the cvxopt boundary, see that file).
Exceptions of the reference that are part of the contract: IndexError when the DP profile has no NaN terminator
(`dp_speed_s[16]`, :428; `relative_time_init[17]`, :538; `dp_speed_t[16]`, :391), ValueError from scipy's interp1d
when the query lies outside the table (:341, :357).
"""
import numpy as np

from oracle import coneqp

N_ST = 16           # len(t_list), speed_planning_test.py:116
N_QP = 17           # output length of speed_QP, :422
N_DENSE = 401       # increase_points / path_speed_merge, :541, :571


def interp1d(x, y, x_new):
    """scipy.interpolate.interp1d(x, y)(x_new): kind='linear', bounds_error=True, assume_sorted=False."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    if len(x) != len(y):
        raise ValueError("x and y arrays must be equal in length along interpolation axis.")
    if len(x) < 1:
        raise ValueError("x and y arrays must have at least 1 entries")
    ind = np.argsort(x, kind="mergesort")
    x, y = x[ind], y[ind]
    if x_new < x[0]:
        raise ValueError("A value in x_new is below the interpolation range.")
    if x_new > x[-1]:
        raise ValueError("A value in x_new is above the interpolation range.")
    hi = int(np.clip(np.searchsorted(x, x_new), 1, len(x) - 1))
    lo = hi - 1
    with np.errstate(divide="ignore", invalid="ignore"):
        slope = (y[hi] - y[lo]) / (x[hi] - x[lo])
        return slope * (x_new - x[lo]) + y[lo]


def _time_index(dp_t, t):
    """:360-370 / :373-381: the interval of dp_speed_t that holds t (0 when none does)."""
    for j in range(len(dp_t) - 1):
        if dp_t[0] > t:
            return j
        if dp_t[j] <= t < dp_t[j + 1]:
            return j
    return 0


def convex_space(dp_s, dp_t, index2s, s_in, s_out, t_in, t_out, kappa, max_lateral_accel=0.2 * 9.8,
                 repair_full_profile=False):
    """generate_convex_space.  Returns s_lb, s_ub, s_dot_lb, s_dot_ub (16 each).  repair_full_profile: repair R3."""
    n = N_ST
    s_lb, s_ub = np.ones(n) * -np.inf, np.ones(n) * np.inf
    sd_lb, sd_ub = np.ones(n) * -np.inf, np.ones(n) * np.inf
    idx_end = len(index2s)
    dp_end = len(dp_s) - 1 if repair_full_profile else len(dp_s)
    for k in range(1, len(index2s)):
        if index2s[k] == 0 and index2s[k - 1] != 0:
            idx_end = k - 1
            break
        idx_end = k
    for k in range(len(dp_s)):
        if np.isnan(dp_s[k]):
            dp_end = k - 1
            break
    for i in range(n):
        if np.isnan(dp_s[i]):
            break
        cur_kappa = interp1d(index2s[0:idx_end], kappa[0:idx_end], dp_s[i])
        sd_lb[i] = 0
        sd_ub[i] = np.sqrt(max_lateral_accel / (abs(cur_kappa) + 1e-10))
    for i in range(len(s_in)):
        if np.isnan(s_in[i]):
            continue
        obs_t = (t_in[i] + t_out[i]) / 2
        obs_s = (s_in[i] + s_out[i]) / 2
        obs_speed = (s_out[i] - s_in[i]) / (t_out[i] - t_in[i])
        dp_at = interp1d([0] + list(dp_t[0:dp_end]), [0] + list(dp_s[0:dp_end]), obs_t)
        lb_i = max(_time_index(dp_t, t_in[i]) - 2, 3)
        ub_i = min(_time_index(dp_t, t_out[i]) + 2, dp_end)
        for m in range(lb_i, ub_i + 1):
            line = s_in[i] + obs_speed * (dp_t[m] - t_in[i])            # IndexError for m == 16 (:391)
            if obs_s > dp_at:
                s_ub[m] = min(s_ub[m], line)
            else:
                s_lb[m] = max(s_lb[m], line)
    return s_lb, s_ub, sd_lb, sd_ub


def speed_qp_setup(dp_s, dp_t, repair_full_profile=False):
    """:415-442: (dp_speed_end, qp_size, dt).  repair_full_profile = repair R3 (opt-in): a profile without a NaN
    terminator ends at its last index (15) instead of the reference's `dp_speed_end = 16` -> IndexError (:428)."""
    end = 15 if repair_full_profile else 16
    for i in range(len(dp_s)):
        if np.isnan(dp_s[i]):
            end = i - 1
            break
    _ = dp_s[end]                                                       # IndexError without a NaN terminator (:428)
    T = dp_t[end]
    with np.errstate(divide="ignore", invalid="ignore"):
        dt = T / np.float64(end)
    return end, end + 1, dt


def speed_qp_matrices(v0, a0, dp_s, dp_t, s_lb, s_ub, sd_lb, sd_ub, w_cost_s_dot2=10, w_cost_v_ref=50,
                      w_cost_jerk=500, reference_speed=50, repair_full_profile=False):
    """What speed_QP builds (:434-499) plus the separate lb/ub of repair R2.
    Returns dict(H, f, A, b, Aeq, beq, lb, ub, dt, qp_size); Aeq is 3q x (2q-2) as in the reference."""
    end, q, dt = speed_qp_setup(dp_s, dp_t, repair_full_profile)
    nv = 3 * q
    Aeq = np.zeros((nv, 2 * q - 2))
    beq = np.zeros((2 * q - 2, 1))
    A_sub = np.array([[1, 0], [dt, 1], [(1 / 3) * dt ** 2, (1 / 2) * dt], [-1, 0], [0, -1], [(1 / 6) * dt ** 2, dt / 2]])
    for i in range(q - 1):
        Aeq[3 * i:3 * i + 6, 2 * i:2 * i + 2] = A_sub
    A = np.zeros((q - 1, nv))
    b = np.zeros((q - 1, 1))
    for i in range(q - 1):
        A[i, 3 * i] = 1
        A[i, 3 * i + 3] = -1
    lb, ub = np.ones(nv), np.ones(nv)
    for i in range(1, q):
        lb[3 * i], lb[3 * i + 1], lb[3 * i + 2] = s_lb[i - 1], sd_lb[i - 1], -6
        ub[3 * i], ub[3 * i + 1], ub[3 * i + 2] = s_ub[i - 1], sd_ub[i - 1], 4
    lb[0:3] = (0, v0, a0)
    ub[0:3] = lb[0:3]
    A_s_dot2, A_ref = np.zeros((nv, nv)), np.zeros((nv, nv))
    A_jerk = np.zeros((nv, q - 1))
    A4_sub = np.array([[0], [0], [1], [0], [0], [-1]])
    for i in range(1, q + 1):
        A_s_dot2[3 * i - 1, 3 * i - 1] = 1
        A_ref[3 * i - 2, 3 * i - 2] = 1
    for i in range(1, q):
        A_jerk[3 * i - 3:3 * i + 3, i - 1:i] = A4_sub
    H = w_cost_s_dot2 * (A_s_dot2 @ A_s_dot2.T) + w_cost_jerk * (A_jerk @ A_jerk.T) + w_cost_v_ref * (A_ref @ A_ref.T)
    H = 2 * H
    f = np.zeros((nv, 1))
    for i in range(1, q + 1):
        f[3 * i - 2] = -2 * w_cost_v_ref * reference_speed
    return dict(H=H, f=f, A=A, b=b, Aeq=Aeq, beq=beq, lb=lb, ub=ub, dt=dt, qp_size=q)


def speed_qp_cone_form(m):
    """Repairs R1/R2: (P, q, G, h, A, b) for coneqp.  Row order of G: no-reversing rows, then the finite lower
    bounds in variable order, then the finite upper bounds in variable order."""
    nv = m["H"].shape[0]
    rows, rhs = [m["A"]], [m["b"].reshape(-1)]
    eye = np.eye(nv)
    lo = np.isfinite(m["lb"])
    hi = np.isfinite(m["ub"])
    rows.append(-eye[lo])
    rhs.append(-m["lb"][lo])
    rows.append(eye[hi])
    rhs.append(m["ub"][hi])
    return m["H"], m["f"].reshape(-1), np.vstack(rows), np.concatenate(rhs), m["Aeq"].T.copy(), m["beq"].reshape(-1)


def speed_qp(v0, a0, dp_s, dp_t, s_lb, s_ub, sd_lb, sd_ub, diverge_factor=None, **kw):
    """speed_QP with repairs R1/R2.  Returns (qp_s, qp_s_dot, qp_s_dot2, relative_time) of length 17, NaN padded,
    and the coneqp info dict."""
    m = speed_qp_matrices(v0, a0, dp_s, dp_t, s_lb, s_ub, sd_lb, sd_ub, **kw)
    info = coneqp.qp(*speed_qp_cone_form(m), diverge_factor=diverge_factor)
    X = info["x"]
    out = [np.ones(N_QP) * np.nan for _ in range(4)]
    for i in range(1, m["qp_size"] + 1):
        out[0][i - 1] = X[3 * i - 3]
        out[1][i - 1] = X[3 * i - 2]
        out[2][i - 1] = X[3 * i - 1]
        out[3][i - 1] = (i - 1) * m["dt"]
    return out[0], out[1], out[2], out[3], info


def increase_points(s_init, s_dot_init, s_dot2_init, t_init, n=N_DENSE):
    """increase_points: 401 samples; the first one is at t = -dt (`(i - 1) * dt`, :550)."""
    t_end = len(t_init)
    for i in range(len(t_init)):
        if np.isnan(t_init[i]):
            t_end = i - 1
            break
    T = t_init[t_end]                                                   # IndexError without a NaN terminator (:538)
    dt = T / (n - 1)
    s, sd, sdd, rt = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
    tmp = 0
    for i in range(n):
        t = (i - 1) * dt
        for j in range(t_end - 1):
            if t_init[j] <= t < t_init[j + 1]:
                tmp = j
                break
        x = t - t_init[tmp]
        s[i] = s_init[tmp] + s_dot_init[tmp] * x + (1 / 3) * s_dot2_init[tmp] * x ** 2 + (1 / 6) * s_dot2_init[tmp + 1] * x ** 2
        sd[i] = s_dot_init[tmp] + 0.5 * s_dot2_init[tmp] * x + 0.5 * s_dot2_init[tmp + 1] * x
        sdd[i] = s_dot2_init[tmp] + (s_dot2_init[tmp + 1] - s_dot2_init[tmp]) * x / (t_init[tmp + 1] - t_init[tmp])
        rt[i] = t
    return s, sd, sdd, rt


def np_interp(x, xp, fp):
    """numpy.interp for one query (compiled arr_interp: clamp outside, exact hit returns fp[j])."""
    n = len(xp)
    if x > xp[n - 1]:
        return fp[n - 1]
    if x < xp[0]:
        return fp[0]
    j = int(np.searchsorted(xp, x, side="right")) - 1
    if j == n - 1 or xp[j] == x:
        return fp[j]
    slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j])
    r = slope * (x - xp[j]) + fp[j]
    if np.isnan(r):
        r = slope * (x - xp[j + 1]) + fp[j + 1]
        if np.isnan(r) and fp[j] == fp[j + 1]:
            r = fp[j]
    return r


def path_speed_merge(s, s_dot, s_dot2, rel_t, current_time, path_s, tx, ty, th, tk, n=N_DENSE):
    """path_speed_merge: tables cut at `index` = (first NaN of trajectory_x_init) - 1, the last valid point excluded
    (`[:index]`, :594); the last output row copies the LAST ELEMENT of the init arrays (:603-606)."""
    index = 0
    while not np.isnan(tx[index]):                                      # IndexError without NaN padding (:581)
        index += 1
    index -= 1
    out = [np.zeros(n) for _ in range(7)]
    for i in range(n - 1):
        for k, tab in enumerate((tx, ty, th, tk)):
            out[k][i] = np_interp(s[i], path_s[:index], tab[:index])
        out[4][i] = s_dot[i]
        out[5][i] = s_dot2[i]
        out[6][i] = rel_t[i] + current_time
    for k, tab in enumerate((tx, ty, th, tk)):
        out[k][-1] = tab[-1]
    out[4][-1] = s_dot[-1]
    out[5][-1] = s_dot2[-1]
    out[6][-1] = rel_t[-1] + current_time
    return tuple(out)
