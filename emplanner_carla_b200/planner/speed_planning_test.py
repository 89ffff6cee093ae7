"""GPU drop-in for the S-T speed DP of the reference's planner/speed_planning_test.py (a draft that cannot
run to completion in the reference).  speed_DP keeps the reference's signature (:101-102);
its forward tables are bit-compatible with the reference's, and the back-track is the documented repair
(int parent index, separate dp_speed_t array), so the function RETURNS where the reference raises IndexError.
generate_convex_space (:308), increase_points (:514) and path_speed_merge (:565) reproduce the reference bit for
bit, including the exceptions it raises (IndexError on a profile without NaN terminator, ValueError from scipy's
interp1d outside its table).  speed_QP (:403) cannot run in the reference at all (Aeq is handed to cvxopt
untransposed and lb/ub are never passed, :435-439,:502); here it solves the problem the draft states, with the two
repairs documented in oracle/st_back.py and DESIGN.md.
"""
import numpy as np

from emplanner_carla_b200 import api, _lib
from . import _runtime


def speed_DP(obs_st_s_in_set, obs_st_s_out_set, obs_st_t_in_set, obs_st_t_out_set, plan_start_s_dot,
             reference_speed=50, w_cost_ref_speed=4000, w_cost_accel=100, w_cost_obs=10000000):
    """S-T lattice DP (reference :101-188).  Returns (dp_speed_s, dp_speed_t), 16-element arrays, NaN beyond
    the optimal end column."""
    k = len(obs_st_s_in_set)
    kmax = max(2, k + (k & 1))
    if kmax > 16:
        raise ValueError("at most 16 dynamic obstacles are supported (got %d)" % k)
    rec = api.pack_st_scenes(np.asarray(obs_st_s_in_set, dtype=np.float64)[None], np.asarray(obs_st_s_out_set)[None],
                             np.asarray(obs_st_t_in_set)[None], np.asarray(obs_st_t_out_set)[None],
                             [float(plan_start_s_dot)], kmax=kmax)
    st = api.st_params(reference_speed, w_cost_ref_speed, w_cost_accel, w_cost_obs)
    r = _runtime.planner().st_dp(rec, st)
    return r["dp_s"][0], r["dp_t"][0]


def calc_speed_planning_start_condition(plan_start_vx, plan_start_vy, plan_start_ax, plan_start_ay, plan_start_heading):
    """Tangential speed and acceleration of the plan start (reference :23-35)."""
    o = _runtime.planner().misc(_lib.OP_START_CONDITION, [[plan_start_vx, plan_start_vy, plan_start_ax, plan_start_ay,
                                                            plan_start_heading]], 5, 1, (1, 2))
    return o[0, 0], o[0, 1]


def generate_st_graph(dynamic_obs_s_set, dynamic_obs_l_set, dynamic_obs_s_dot_set, dynamic_obs_l_dot_set):
    """Entry/exit (s, t) of every dynamic obstacle through the +-2 m lateral band (reference :38-98).
    Returns (s_in, s_out, t_in, t_out) arrays, NaN where the obstacle is ignored."""
    inp = np.array([dynamic_obs_s_set, dynamic_obs_l_set, dynamic_obs_s_dot_set, dynamic_obs_l_dot_set], dtype=np.float64)
    n = inp.shape[1]
    o = _runtime.planner().misc(_lib.OP_ST_GRAPH, inp[None], n, 1, (1, 4, n))
    return o[0, 0], o[0, 1], o[0, 2], o[0, 3]


def CalcSTCoordinate(row, col, s_list, t_list):
    """(s, t) of a lattice node; row 0 is the largest s (reference :287-305).  Pure indexing."""
    return s_list[len(s_list) - row - 1], t_list[col]


def CalcCollisionCost(w_cost_obs, min_dis):
    """Reference :274-284."""
    o = _runtime.planner().misc(_lib.OP_ST_COLLISION, [float(min_dis)], 1, 1, (1,), obs4=np.zeros((4, 0)),
                                st=api.st_params(w_cost_obs=w_cost_obs))
    return float(o[0])


def _obs4(a, b, c, d):
    o = np.array([a, b, c, d], dtype=np.float64)
    if o.shape[1] > 16:
        raise ValueError("at most 16 dynamic obstacles are supported")
    return o


def CalcObsCost(s_start, t_start, s_end, t_end, obs_st_s_in_set, obs_st_s_out_set, obs_st_t_in_set, obs_st_t_out_set,
                w_cost_obs):
    """Obstacle cost of one S-T edge (reference :234-271)."""
    o = _runtime.planner().misc(_lib.OP_ST_OBS_COST, [[s_start, t_start, s_end, t_end]], 1, 1, (1,),
                                obs4=_obs4(obs_st_s_in_set, obs_st_s_out_set, obs_st_t_in_set, obs_st_t_out_set),
                                st=api.st_params(w_cost_obs=w_cost_obs))
    return float(o[0])


def CalcDpCost(row_start, col_start, row_end, col_end, obs_st_s_in_set, obs_st_s_out_set, obs_st_t_in_set,
               obs_st_t_out_set, w_cost_ref_speed, reference_speed, w_cost_accel, w_cost_obs, plan_start_s_dot,
               s_list, t_list, dp_st_s_dot):
    """Cost of one S-T lattice edge (reference :191-231).  s_list / t_list must be the reference's own lattice
    (:114-116); they are regenerated on the device."""
    o = _runtime.planner().misc(_lib.OP_ST_DP_COST, [[row_start, col_start, row_end, col_end]], 1, 1, (1,),
                                obs4=_obs4(obs_st_s_in_set, obs_st_s_out_set, obs_st_t_in_set, obs_st_t_out_set),
                                sdot_tab=np.asarray(dp_st_s_dot, dtype=np.float64), v0=plan_start_s_dot,
                                st=api.st_params(reference_speed, w_cost_ref_speed, w_cost_accel, w_cost_obs))
    return float(o[0])


def _raise_st(status, where):
    """Re-raise what the reference raises for a per-frame status word of the S-T back half."""
    status = int(status)
    if status & api.ST_INDEX_ERROR:
        raise IndexError("%s: index out of bounds (S-T profile without NaN terminator, or index 16)" % where)
    if status & api.ST_VALUE_ERROR:
        raise ValueError("%s: a value in x_new is outside the interpolation range" % where)
    if status & api.ST_BAD_PROFILE:
        raise ValueError("%s: S-T profile holds fewer than 2 valid points" % where)


def _st_out(dp_speed_s, dp_speed_t):
    rec = np.full((1, api.ST_OUT_STRIDE), np.nan)
    s = np.asarray(dp_speed_s, dtype=np.float64).reshape(-1)
    t = np.asarray(dp_speed_t, dtype=np.float64).reshape(-1)
    if len(s) != 16 or len(t) != 16:
        raise ValueError("dp_speed_s / dp_speed_t must hold 16 entries (speed_DP output)")
    rec[0, 8:24], rec[0, 24:40] = s, t
    return rec


def generate_convex_space(dp_speed_s, dp_speed_t, path_index2s, obs_st_s_in_set, obs_st_s_out_set, obs_st_t_in_set,
                          obs_st_t_out_set, trajectory_kappa_init, max_lateral_accel=0.2 * 9.8):
    """Bounds on s and s_dot for the speed QP (reference :308-400).  Returns (s_lb, s_ub, s_dot_lb, s_dot_ub)."""
    k = len(obs_st_s_in_set)
    kmax = max(2, k + (k & 1))
    if kmax > 16:
        raise ValueError("at most 16 dynamic obstacles are supported (got %d)" % k)
    sc = api.pack_st_scenes(np.asarray(obs_st_s_in_set, dtype=np.float64)[None], np.asarray(obs_st_s_out_set)[None],
                            np.asarray(obs_st_t_in_set)[None], np.asarray(obs_st_t_out_set)[None], [0.0], kmax=kmax)
    n = min(len(path_index2s), len(trajectory_kappa_init))
    path = np.array([np.asarray(path_index2s, dtype=np.float64)[:n], np.asarray(trajectory_kappa_init, dtype=np.float64)[:n]])
    r = _runtime.planner().st_back(api.speed_qp_params(max_lateral_accel=max_lateral_accel), st_out=_st_out(dp_speed_s, dp_speed_t),
                                   scenes=sc, path=path[None], want=("bounds",))
    _raise_st(r["convex_status"][0], "generate_convex_space")
    b = r["bounds"][0]
    return b[0], b[1], b[2], b[3]


def speed_QP(plan_start_s_dot, plan_start_s_dot2, dp_speed_s, dp_speed_t, s_lb, s_ub, s_dot_lb, s_dot_ub,
             w_cost_s_dot2=10, w_cost_v_ref=50, w_cost_jerk=500, reference_speed=50):
    """Speed QP (reference :403-511).  Returns (qp_s_init, qp_s_dot_init, qp_s_dot2_init, relative_time_init),
    17 entries each, NaN beyond qp_size."""
    bounds = np.array([s_lb, s_ub, s_dot_lb, s_dot_ub], dtype=np.float64)
    r = _runtime.planner().st_back(api.speed_qp_params(w_cost_s_dot2, w_cost_v_ref, w_cost_jerk, reference_speed),
                                   st_out=_st_out(dp_speed_s, dp_speed_t), bounds_in=bounds[None],
                                   start=[[float(plan_start_s_dot), float(plan_start_s_dot2)]], want=("qp",))
    _raise_st(r["qp_status"][0], "speed_QP")
    q = r["qp"][0]
    return q[0], q[1], q[2], q[3]


def increase_points(s_init, s_dot_init, s_dot2_init, relative_time_init):
    """Densify the QP profile to 401 samples (reference :514-562)."""
    qp = np.array([s_init, s_dot_init, s_dot2_init, relative_time_init], dtype=np.float64)
    if qp.shape[1] != 17:
        raise ValueError("profiles must hold 17 entries (speed_QP output)")
    r = _runtime.planner().st_back(api.speed_qp_params(), qp_in=qp[None], want=("dense",))
    _raise_st(r["dense_status"][0], "increase_points")
    d = r["dense"][0]
    return d[0], d[1], d[2], d[3]


def path_speed_merge(s, s_dot, s_dot2, relative_time, current_time, path_s, trajectory_x_init, trajectory_y_init,
                     trajectory_heading_init, trajectory_kappa_init):
    """Merge the path with the speed profile (reference :565-611).  Returns (trajectory_x, trajectory_y,
    trajectory_heading, trajectory_kappa, trajectory_speed, trajectory_accel, trajectory_time), 401 entries each."""
    dense = np.array([s, s_dot, s_dot2, relative_time], dtype=np.float64)
    if dense.shape[1] != 401:
        raise ValueError("profiles must hold 401 entries (increase_points output)")
    tabs = [np.asarray(a, dtype=np.float64).reshape(-1) for a in (path_s, trajectory_x_init, trajectory_y_init,
                                                                  trajectory_heading_init, trajectory_kappa_init)]
    n = len(tabs[1])
    if any(len(a) < n for a in tabs):
        raise IndexError("path tables shorter than trajectory_x_init")
    path5 = np.array([a[:n] for a in tabs])
    r = _runtime.planner().st_back(api.speed_qp_params(), dense_in=dense[None], current_time=[float(current_time)],
                                   path5=path5[None], want=("traj",))
    _raise_st(r["merge_status"][0], "path_speed_merge")
    t = r["traj"][0]
    return t[0], t[1], t[2], t[3], t[4], t[5], t[6]
