"""Oracle: S-T speed DP (numpy float64, one frame at a time, vectorised over edges).

TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates reference planner/speed_planning_test.py:
  speed_DP          :101-188 -> speed_dp()
  CalcDpCost        :191-231 -> _edge_costs()
  CalcObsCost       :234-271 -> _obs_cost()
  CalcCollisionCost :274-284
  CalcSTCoordinate  :287-305 (row 0 = largest s)
The reference's back-track (:181-186) cannot run: `dp_st_node` is a float
array used as an index (:122,:182-184) and `dp_speed_t = dp_speed_s` aliases
one array (:155-156).  The forward tables (dp_st_cost, dp_st_s_dot,
dp_st_node, min_row, min_col) ARE produced by the reference before it raises
and are what tests/golden pins; the back-track here is the documented
two-line repair (int cast of the parent row; separate t array).
planner/speed_planning.py::construct_ST_graph_and_planning (:17-93) is a
non-running stub (TypeError at :57) and has no oracle.
"""
import numpy as np

S_LIST = np.concatenate((np.arange(0, 5, 0.5), np.arange(5.5, 15, 1), np.arange(16, 30, 1.5), np.arange(32, 55, 2.5)))
T_LIST = np.arange(0.5, 8.5, 0.5)
NROW, NCOL = 40, 16
assert len(S_LIST) == NROW and len(T_LIST) == NCOL


def node_s(row):
    return S_LIST[NROW - np.asarray(row) - 1]


def _obs_cost(s_start, t_start, s_end, t_end, obs, w_obs):
    """CalcObsCost for arrays of edges.  5 samples at t_start + (i-1)*dt (sic, :251-252)."""
    s_in, s_out, t_in, t_out = obs
    cost = np.zeros(np.broadcast(s_start, s_end).shape)
    n = 5
    dt = (t_end - t_start) / (n - 1)
    k = (s_end - s_start) / (t_end - t_start)
    for i in range(n):
        t = t_start + (i - 1) * dt
        s = s_start + k * (i - 1) * dt
        for j in range(len(s_in)):
            if np.isnan(s_in[j]):
                continue
            v1s, v1t = s_in[j] - s, t_in[j] - t
            v2s, v2t = s_out[j] - s, t_out[j] - t
            v3s, v3t = v2s - v1s, v2t - v1t
            dis1 = np.sqrt(v1s * v1s + v1t * v1t)
            dis2 = np.sqrt(v2s * v2s + v2t * v2t)
            dis3 = np.abs(v1s * v3t - v1t * v3s) / np.sqrt(v3s * v3s + v3t * v3t)
            d13 = v1s * v3s + v1t * v3t
            d23 = v2s * v3s + v2t * v3t
            ends = ((d13 > 0) & (d23 > 0)) | ((d13 < 0) & (d23 < 0))
            dmin = np.where(ends, np.minimum(dis1, dis2), dis3)
            ad = np.abs(dmin)
            with np.errstate(over="ignore", invalid="ignore"):
                soft = np.power(float(w_obs), (0.5 - dmin) + 1)
            cost = cost + np.where(ad < 0.5, float(w_obs), np.where((0.5 < ad) & (ad < 1.5), soft, 0.0))
    return cost


def speed_dp(obs_s_in, obs_s_out, obs_t_in, obs_t_out, plan_start_s_dot, reference_speed=50,
             w_cost_ref_speed=4000, w_cost_accel=100, w_cost_obs=10000000):
    """Returns dict(cost[40,16], s_dot[40,16], node[40,16] int, min_row, min_col, dp_s[16], dp_t[16])."""
    obs = [np.asarray(a, dtype=np.float64) for a in (obs_s_in, obs_s_out, obs_t_in, obs_t_out)]
    cost = np.ones((NROW, NCOL)) * np.inf
    sdot = np.zeros((NROW, NCOL))
    node = np.zeros((NROW, NCOL), dtype=np.int64)
    rows = np.arange(NROW)
    s_of = node_s(rows)

    def edge(s_st, t_st, sd_st, s_en, t_en):
        cur = (s_en - s_st) / (t_en - t_st)
        acc = (cur - sd_st) / (t_en - t_st)
        c_ref = w_cost_ref_speed * (cur - reference_speed) ** 2
        ok = (4 > acc) & (acc > -6)
        c_acc = np.where(ok, w_cost_accel * acc ** 2, 100000 * w_cost_accel * acc ** 2)
        c_obs = _obs_cost(s_st, t_st, s_en, t_en, obs, w_cost_obs)
        return c_obs + c_acc + c_ref

    # column 0 from the origin (:125-131)
    cost[:, 0] = edge(0.0, 0.0, plan_start_s_dot, s_of, T_LIST[0])
    sdot[:, 0] = s_of / T_LIST[0]
    for c in range(1, NCOL):
        # source k (axis 1) in column c-1 -> target j (axis 0) in column c
        k_is0 = (rows == 0)[None, :]
        s_st = np.where(k_is0, 0.0, s_of[None, :])                  # row_start == 0 quirk (:208-212)
        t_st = np.where(k_is0, 0.0, T_LIST[c - 1])
        sd_st = np.where(k_is0, plan_start_s_dot, sdot[None, :, c - 1])
        s_en = s_of[:, None]
        e = edge(s_st, t_st, sd_st, s_en, T_LIST[c])
        tot = e + cost[None, :, c - 1]
        kbest = np.argmin(tot, axis=1)                              # strict '<' ascending k (:145)
        best = tot[rows, kbest]
        upd = best < cost[:, c]
        cost[:, c] = np.where(upd, best, cost[:, c])
        node[:, c] = np.where(upd, kbest, 0)
        sdot[:, c] = np.where(upd, (s_of - s_of[kbest]) / (T_LIST[c] - T_LIST[c - 1]), 0.0)
    # end node: right boundary then top boundary, '<=' so the last minimum wins (:162-172)
    min_cost, min_row, min_col = np.inf, -1, -1
    for i in range(NROW):
        if cost[i, NCOL - 1] <= min_cost:
            min_cost, min_row, min_col = cost[i, NCOL - 1], i, NCOL - 1
    for j in range(NCOL):
        if cost[0, j] <= min_cost:
            min_cost, min_row, min_col = cost[0, j], 0, j
    dp_s = np.ones(NCOL) * np.nan
    dp_t = np.ones(NCOL) * np.nan
    r, c = min_row, min_col
    dp_s[c], dp_t[c] = node_s(r), T_LIST[c]
    while c != 0:                                                   # repaired back-track
        r = int(node[r, c])
        c -= 1
        dp_s[c], dp_t[c] = node_s(r), T_LIST[c]
    return dict(cost=cost, s_dot=sdot, node=node, min_row=min_row, min_col=min_col, dp_s=dp_s, dp_t=dp_t)
