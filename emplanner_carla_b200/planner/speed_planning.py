"""GPU drop-in for the reference's planner/speed_planning.py - the S-T planning STUB the north_star names.

In the reference this module does not work: construct_ST_graph_and_planning (:17-93) multiplies an
int by the `obs_speed` list (TypeError at :57) or, without obstacles, reaches undefined names (NameError at :88) and
returns nothing; the working S-T draft is planner/speed_planning_test.py (drop-in: speed_planning_test.py here).
What does run in it are the two pure cost functions, which this module serves from the GPU under their own names:
  cal_start_cost (:96-169)  the S-L edge cost on an (s, t) quintic - same smooth / reference terms, same `s * 2`
                            slip in the jerk term - with cal_obs_cost's 3 m / 5 m distances
  cal_obs_cost   (:172-197) danger_dis = 3, safe_dis = 5
"""
import numpy as np

from emplanner_carla_b200 import api
from . import _runtime


def construct_ST_graph_and_planning(path_s: list, ego_pos: int, ego_speed: int, ego_acc: int,
                                    obs_pos: list, obs_speed: list, time_interval=8, time_slot=1):
    """Raises exactly where the reference raises (it never returns a plan): TypeError when there is an obstacle
    (`obs_pos[n] + t * obs_speed` adds a number to a list, :57), NameError otherwise (`obs_s_list`, :88)."""
    if len(obs_pos) > 0:
        raise TypeError("unsupported operand type(s) for +: '%s' and 'list'" % type(obs_pos[0]).__name__)
    raise NameError("name 'obs_s_list' is not defined")


def cal_start_cost(obs_s_list, obs_t_list, begin_s, begin_t, begin_dt, end_s, end_t, sample_t,
                   w_cost_collision, w_cost_smooth, w_cost_ref):
    """Cost of the (s, t) quintic from (begin_s, begin_t, begin_dt, 0) to (end_s, end_t, 0, 0), 10 samples spaced
    sample_t / 10 in s (reference :96-169).  Returns a (1, 1) ndarray like the reference."""
    dp = api.dp_params(w_collision_cost=w_cost_collision, w_smooth_cost=w_cost_smooth, w_reference_cost=w_cost_ref,
                       sample_s=sample_t)
    c = _runtime.planner().edge_cost([[begin_s, begin_t, begin_dt, 0.0, end_s, end_t]], obs_s_list, obs_t_list, dp,
                                     danger_dis=3, safe_dis=5)
    return np.array([[c[0]]])


def cal_obs_cost(w_cost_collision, square_d: np.ndarray, danger_dis=3, safe_dis=5):
    """Obstacle cost from the squared distances of an edge's samples (reference :172-197)."""
    return float(_runtime.planner().obs_cost(np.asarray(square_d, dtype=np.float64).reshape(1, -1), w_cost_collision,
                                             danger_dis, safe_dis)[0])
