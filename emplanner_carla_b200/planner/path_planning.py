"""GPU drop-in for the reference's planner/path_planning.py (same names, arguments, defaults, return types).

Every function forwards to libemplanner.so through ctypes (emplanner_carla_b200.api.HostPlanner);
nothing is computed on the CPU.  Batch-of-one calls are the latency path; the throughput path is the
batched API (emplanner_carla_b200.api).  Reference line numbers cite path_planning.py @ 1f115a5.
"""
import numpy as np

from emplanner_carla_b200 import api
from . import _runtime


def DP_algorithm(obs_s_list: list, obs_l_list: list,
                 plan_start_s, plan_start_l, plan_start_dl, plan_start_ddl, sampling_res=2,
                 w_collision_cost=1e12, w_smooth_cost=[300, 1000, 5000], w_reference_cost=20,
                 row=12, col=6, sample_s=15, sample_l=1.5):
    """S-L lattice DP + densify (reference :276-375).  Returns (enriched_s_list, enriched_l_list),
    plan start excluded, like the reference.  Prints the reference's banner when no collision-free
    path exists (:351-353)."""
    dp = api.dp_params(sampling_res, w_collision_cost, w_smooth_cost, w_reference_cost, row, col, sample_s, sample_l)
    kmax = _runtime.kmax_for(len(obs_s_list))
    pl = _runtime.planner(kmax=kmax, n_enr_max=api.max_enriched(dp), max_col=int(col))
    rec = _runtime.scene_record(obs_s_list, obs_l_list, plan_start_s, plan_start_l, plan_start_dl, plan_start_ddl,
                                pl.kmax)
    view, _ = pl.dp(rec, dp, want_rows=False)
    if int(view.dp_status[0]) & api.DP_OVERFLOW:
        raise RuntimeError("densified path longer than the planner buffer")
    if int(view.dp_status[0]) & api.DP_ALL_COLLIDE:
        print("********************     can't find a feasible path      ********************")
    n = int(view.n_enr[0])
    return list(view.enr_s[0, :n]), list(view.enr_l[0, :n])


def cal_lmin_lmax(dp_path_s, dp_path_l, obs_s_list, obs_l_list, obs_length, obs_width):
    """QP corridor (reference :222-273).  Returns (lmin, lmax) ndarrays; raises IndexError where the
    reference does (obstacle index + 2 beyond the path, :265-272)."""
    dp_s = np.asarray(dp_path_s, dtype=np.float64).reshape(1, -1)
    dp_l = np.asarray(dp_path_l, dtype=np.float64).reshape(1, -1)
    n = dp_s.shape[1]
    if n > 128:
        raise NotImplementedError("cal_lmin_lmax on the GPU handles at most 128 path points (got %d)" % n)
    qp = api.qp_params(obs_length=obs_length, obs_width=obs_width)
    kmax = _runtime.kmax_for(len(obs_s_list))
    pl = _runtime.planner(kmax=kmax)
    rec = _runtime.scene_record(obs_s_list, obs_l_list, 0.0, 0.0, 0.0, 0.0, pl.kmax)
    lmin, lmax, status = pl.corridor(dp_s, dp_l, rec, qp)
    if int(status[0]) == api.QP_CORRIDOR_INDEX:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (n, n))
    return lmin[0], lmax[0]


def Quadratic_planning(l_min, l_max, plan_start_l, plan_start_dl, plan_start_ddl, dp_sampling_res=2,
                       w_cost_l=1000, w_cost_dl=10000, w_cost_ddl=3000, w_cost_dddl=150, w_cost_centre=250,
                       w_cost_end_l=40, w_cost_end_dl=40, w_cost_end_ddl=40,
                       host_d1=3, host_d2=3, host_w=3):
    """QP path smoother (reference :78-219, cvxopt.solvers.qp at :211).  Returns
    (qp_path_l, qp_path_dl, qp_path_ddl) lists.  Like the reference, the solver status is not
    inspected: a non-'optimal' solve returns its last iterate."""
    lmin = np.asarray(l_min, dtype=np.float64).reshape(1, -1)
    lmax = np.asarray(l_max, dtype=np.float64).reshape(1, -1)
    n = lmin.shape[1]
    if n > 128 or n < 2:
        raise NotImplementedError("Quadratic_planning on the GPU handles 2..128 stations (got %d)" % n)
    qp = api.qp_params(dp_sampling_res, w_cost_l, w_cost_dl, w_cost_ddl, w_cost_dddl, w_cost_centre, w_cost_end_l,
                       w_cost_end_dl, w_cost_end_ddl, host_d1, host_d2, host_w)
    pl = _runtime.planner()
    start = np.array([[plan_start_l, plan_start_dl, plan_start_ddl]], dtype=np.float64)
    x, status, iters = pl.qp(lmin, lmax, start, qp)
    Quadratic_planning.last_status = (int(status[0]), int(iters[0]))
    return list(x[0, 0]), list(x[0, 1]), list(x[0, 2])


Quadratic_planning.last_status = None


def cal_proj_point(s, pre_match_index, frenet_path_opt: list, s_map: list):
    """Projection of arc length s on the reference line (reference :52-75).  Returns
    (proj_x, proj_y, proj_theta, proj_kappa, match_index); raises IndexError beyond the map like the
    reference's unguarded walk (:63)."""
    line = np.asarray([tuple(p)[:4] for p in frenet_path_opt], dtype=np.float64)
    smap = np.asarray(s_map, dtype=np.float64)
    # the walk never moves backwards: start it at pre_match_index by hiding the nodes before it
    k = int(pre_match_index)
    cart, ncart = _runtime.planner().frenet_to_cart(line[None, k:], smap[None, k:], np.array([[s, 0.0]]),
                                                    np.array([[s]]), np.array([[0.0]]))
    if ncart[0] < 0 or len(smap) - k < 2:
        raise IndexError("list index out of range")
    if ncart[0] >= 2:
        x, y, th, kap = cart[0, 1]
        # recover the node index from theta = theta_m + kappa_m * (s - s_map[m]) is ambiguous; use the s-map
        m = k
        while m + 1 < len(smap) - 1 and smap[m + 1] < s:
            m += 1
        return x, y, th, kap, m
    raise IndexError("list index out of range")


def frenet_2_x_y_theta_kappa(plan_start_s, plan_start_l, enriched_s_list: list, enriched_l_list: list,
                             frenet_path_opt: list, s_map: list, smooth=True):
    """(s, l) path -> Cartesian [(x, y, theta, kappa), ...] (reference :15-49): projection on the GPU
    (truncated where s > s_map[-1], :40-41), then planning_utils.smooth_reference_line like the reference (:47)."""
    line = np.asarray([tuple(p)[:4] for p in frenet_path_opt], dtype=np.float64)[None]
    smap = np.asarray(s_map, dtype=np.float64)[None]
    ps = np.asarray(enriched_s_list, dtype=np.float64)[None]
    pl_ = np.asarray(enriched_l_list, dtype=np.float64)[None]
    cart, ncart = _runtime.planner().frenet_to_cart(line, smap, np.array([[plan_start_s, plan_start_l]]), ps, pl_)
    if ncart[0] < 0:
        raise IndexError("list index out of range")
    n = int(ncart[0])
    target_xy = [(cart[0, 0, 0], cart[0, 0, 1])] + [tuple(cart[0, i]) for i in range(1, n)]
    if not smooth:
        return target_xy
    from . import planning_utils
    return planning_utils.smooth_reference_line(target_xy)



def enrich_DP_s_l(DP_s_list, DP_l_list, plan_start_s, plan_start_l, plan_start_dl, plan_start_ddl, resolution=1):
    """Densify the DP nodes along their quintic connections (reference :378-432)."""
    s = np.asarray(DP_s_list, dtype=np.float64)[None]
    l = np.asarray(DP_l_list, dtype=np.float64)[None]
    span = max(float(s[0, -1] - plan_start_s), 0.0)
    n_max = int(span / float(resolution)) + s.shape[1] + 2
    os_, ol_, n = _runtime.planner().enrich(s, l, [[plan_start_s, plan_start_l, plan_start_dl, plan_start_ddl]],
                                            resolution, n_max)
    return list(os_[0, :n[0]]), list(ol_[0, :n[0]])


def cal_start_cost(obs_s_list, obs_l_list, begin_s, begin_l, begin_dl, begin_ddl, cur_node_row, row,
                   sample_s, sample_l, w_cost_collision, w_cost_smooth, w_cost_ref):
    """Cost of the edge plan start -> lattice node (row cur_node_row, first column) (reference :435-514).
    Returns a (1, 1) ndarray like the reference."""
    end_l = ((row + 1) / 2 - 1 - cur_node_row) * sample_l
    dp = api.dp_params(w_collision_cost=w_cost_collision, w_smooth_cost=w_cost_smooth, w_reference_cost=w_cost_ref,
                       sample_s=sample_s, sample_l=sample_l)
    c = _runtime.planner().edge_cost([[begin_s, begin_l, begin_dl, begin_ddl, begin_s + sample_s, end_l]],
                                     obs_s_list, obs_l_list, dp)
    return np.array([[c[0]]])


def cal_neighbor_cost(obs_s_list, obs_l_list, pre_node_s, pre_node_l, cur_node_s, cur_node_l, sample_s,
                      w_cost_collision, w_cost_smooth, w_cost_ref):
    """Cost of the edge between two lattice nodes of adjacent columns (reference :517-585).
    Returns a (1, 1) ndarray like the reference."""
    dp = api.dp_params(w_collision_cost=w_cost_collision, w_smooth_cost=w_cost_smooth, w_reference_cost=w_cost_ref,
                       sample_s=sample_s)
    c = _runtime.planner().edge_cost([[pre_node_s, pre_node_l, 0.0, 0.0, cur_node_s, cur_node_l]],
                                     obs_s_list, obs_l_list, dp)
    return np.array([[c[0]]])


def cal_obs_cost(w_cost_collision, square_d: np.ndarray, danger_dis=4, safe_dis=6):
    """Obstacle cost of one edge from the squared distances of its samples (reference :588-609)."""
    return float(_runtime.planner().obs_cost(np.asarray(square_d, dtype=np.float64).reshape(1, -1), w_cost_collision,
                                             danger_dis, safe_dis)[0])
