"""GPU drop-in for the S-T speed DP of the reference's planner/speed_planning_test.py (a draft that cannot
run to completion in the reference).  speed_DP keeps the reference's signature (:101-102);
its forward tables are bit-compatible with the reference's, and the back-track is the documented repair
(int parent index, separate dp_speed_t array), so the function RETURNS where the reference raises IndexError.
Not provided: speed_QP (:403) cannot run in the reference at all (Aeq is built transposed and lb/ub are never
passed to the solver, :435-439,:502), increase_points / path_speed_merge raise IndexError on a full profile
(:538) and generate_convex_space depends on them; there is no behaviour to be compatible with.
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
