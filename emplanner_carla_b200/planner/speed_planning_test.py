"""GPU drop-in for the S-T speed DP of the reference's planner/speed_planning_test.py (a draft that cannot
run to completion in the reference).  speed_DP keeps the reference's signature (:101-102);
its forward tables are bit-compatible with the reference's, and the back-track is the documented repair
(int parent index, separate dp_speed_t array), so the function RETURNS where the reference raises IndexError.
The QP/merge half of the draft (generate_convex_space, speed_QP, increase_points, path_speed_merge) is not
provided yet.
"""
import numpy as np

from emplanner_carla_b200 import api
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
