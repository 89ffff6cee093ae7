"""GPU drop-in for the reference's planner/planning_utils.py (same names, arguments, defaults, return types).

The Cartesian<->Frenet arithmetic runs in libemplanner.so (csrc/frenet.cu) through
emplanner_carla_b200.api.HostPlanner; `sampling` is pure list slicing and stays in Python.
Reference line numbers cite planning_utils.py @ 1f115a5.  The three CARLA-actor helpers (Vector_fun,
waypoint_list_2_target_path, predict_block) are accessor code around simulator objects: they are duck-typed here
(anything with .x/.y/.z, .transform.location, .get_location() ... works; `carla` itself is never imported).
"""
import math
import numpy as np

from emplanner_carla_b200 import api, _lib
from . import _runtime


def _line(path):
    a = np.asarray([tuple(p)[:4] for p in path], dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError("reference line must be [(x, y, theta, kappa), ...]")
    return a[None]


def _pts(xy_list):
    return np.asarray([(p[0], p[1]) for p in xy_list], dtype=np.float64)[None]


def Vector_fun(loc_1, loc_2):
    """Unit vector from loc_1 to loc_2, rounded to 4 decimals (reference :14-26; used by global_planning.py:132
    while it builds the route graph).  Three subtractions and a norm on simulator objects: host code."""
    d = [loc_2.x - loc_1.x, loc_2.y - loc_1.y, loc_2.z - loc_1.z]
    norm = np.linalg.norm(d) + np.finfo(float).eps
    return np.round([d[0] / norm, d[1] / norm, d[2] / norm], 4)


def waypoint_list_2_target_path(pathway):
    """[(waypoint, road_option), ...] -> [(x, y, theta, kappa), ...] (reference :29-46); heading and curvature
    of the whole route come from cal_heading_kappa on the GPU."""
    xy = [(w[0].transform.location.x, w[0].transform.location.y) for w in pathway]
    theta, kappa = cal_heading_kappa(xy)
    return [(xy[i][0], xy[i][1], theta[i], kappa[i]) for i in range(len(theta))]


def predict_block(ego_vehicle, ts=0.1):
    """Ego pose one prediction horizon ahead from its velocity, side-slip and yaw rate (reference :591-614).
    Returns (x, y, yaw).  A dozen scalar operations on actor accessors, once per plan: host code."""
    loc = ego_vehicle.get_location()
    yaw = ego_vehicle.get_transform().rotation.yaw * (math.pi / 180)
    vel = ego_vehicle.get_velocity()
    speed = math.sqrt(vel.x * vel.x + vel.y * vel.y + vel.z * vel.z)
    slip = math.atan2(vel.y, vel.x) - yaw                     # angle between velocity and body axis
    v_lat = speed * math.sin(slip)
    v_lon = speed * math.cos(slip)
    x = loc.x + v_lon * ts * math.cos(yaw) - v_lat * ts * math.sin(yaw)
    y = loc.y + v_lat * ts * math.cos(yaw) + v_lon * ts * math.sin(yaw)
    yaw_rate = ego_vehicle.get_angular_velocity().z * (math.pi / 180)
    return x, y, yaw + yaw_rate * ts


def match_projection_points(xy_list: list, frenet_path_node_list: list):
    """Nearest node + projection for each point (reference :364-426; every projection goes through the
    match index of the first point, :413).  Returns (index list, [(x_r, y_r, theta_r, kappa_r), ...])."""
    if len(xy_list) == 0:
        return [], []
    r = _runtime.planner().frenet_project(_line(frenet_path_node_list), points=_pts(xy_list), want=("match", "proj"))
    return list(r["match"][0]), [tuple(v) for v in r["proj"][0]]


def find_match_points(xy_list: list, frenet_path_node_list: list, is_first_run: bool, pre_match_index: int):
    """Reference :49-182: first run scans from node 0 (early exit after 50); later runs scan from
    pre_match_index forward or backward (early exit after 5)."""
    if len(xy_list) == 0:
        return [], []
    pre = None if is_first_run is True else np.array([int(pre_match_index)], dtype=np.int32)
    r = _runtime.planner().frenet_project(_line(frenet_path_node_list), points=_pts(xy_list), pre_match=pre,
                                          want=("match", "proj"))
    return list(r["match"][0]), [tuple(v) for v in r["proj"][0]]


def sampling(match_point_index: int, frenet_path_node_list: list, back_length=10, forward_length=50):
    """Local reference-line window (reference :231-259).  The arguments are ignored in favour of 10 back /
    40 forward exactly like the reference (:244-245); pure list slicing, no arithmetic."""
    back_length, forward_length = 10, 40
    total = back_length + forward_length
    if match_point_index < back_length:
        back_length = match_point_index
        forward_length = total - back_length
    if (len(frenet_path_node_list) - match_point_index) - 1 < forward_length:
        forward_length = len(frenet_path_node_list) - match_point_index - 1
        back_length = total - forward_length
    return list(frenet_path_node_list[match_point_index - back_length: match_point_index]) \
        + list(frenet_path_node_list[match_point_index: match_point_index + forward_length + 1])


def cal_s_map_fun(local_path_opt: list, origin_xy: tuple):
    """s of every reference-line node relative to the ego projection (reference :448-472)."""
    r = _runtime.planner().frenet_project(_line(local_path_opt), origin=np.array([[origin_xy[0], origin_xy[1]]]),
                                          want=("s_map",))
    return list(r["s_map"][0])


def cal_s_l_fun(obs_xy_list: list, local_path_opt: list, s_map: list):
    """Cartesian -> (s, l) for a list of points with the caller's s_map (reference :475-509)."""
    if len(obs_xy_list) == 0:
        return [], []
    r = _runtime.planner().frenet_project(_line(local_path_opt), points=_pts(obs_xy_list),
                                          s_map_in=np.asarray(s_map, dtype=np.float64)[None], want=("sl",))
    return list(r["sl"][0, :, 0]), list(r["sl"][0, :, 1])


def cal_s_l_deri_fun(xy_list: list, V_xy_list: list, a_xy_list: list, local_path_xy_opt: list, origin_xy: tuple):
    """l, dl/dt, ds/dt, ddl/dt2, dl/ds, dds/dt2, ddl/ds2 lists (reference :512-588)."""
    if len(xy_list) == 0:
        return tuple([] for _ in range(7))
    r = _runtime.planner().frenet_project(_line(local_path_xy_opt), origin=np.array([[origin_xy[0], origin_xy[1]]]),
                                          points=_pts(xy_list), vel=_pts(V_xy_list), acc=_pts(a_xy_list),
                                          want=("deriv",))
    d = r["deriv"][0]
    return tuple(list(d[:, k]) for k in range(7))


def smooth_reference_line(local_frenet_path_xy: list,
                          w_cost_smooth=0.4, w_cost_length=0.3, w_cost_ref=0.3,
                          x_thre=0.2, y_thre=0.2):
    """Box-constrained QP smoothing of a polyline, then heading and curvature (reference :262-361).
    Accepts [(x, y), ...] or [(x, y, theta, kappa), ...] (and, like the reference, a first element that is
    only (x, y)); returns [(x_opt, y_opt, theta, kappa), ...]."""
    n = len(local_frenet_path_xy)
    if n < 2 or n > 128:
        raise NotImplementedError("smooth_reference_line on the GPU handles 2..128 points (got %d)" % n)
    xy = _pts(local_frenet_path_xy)
    sp = api.smooth_params(w_cost_smooth, w_cost_length, w_cost_ref, x_thre, y_thre)
    out, status, iters = _runtime.planner().smooth_lines(xy, sp)
    smooth_reference_line.last_status = (int(status[0]), int(iters[0]))
    return [tuple(r) for r in out[0]]


smooth_reference_line.last_status = None


def cal_heading_kappa(frenet_path_xy_list: list):
    """Heading and curvature of a polyline by central differences (reference :185-228).
    Returns (theta list, kappa list)."""
    n = len(frenet_path_xy_list)
    if n < 2:
        raise IndexError("list index out of range")            # the reference indexes dx_[0] of an empty list (:213)
    sp = api.smooth_params(heading_only=True)
    out, _, _ = _runtime.planner().smooth_lines(_pts(frenet_path_xy_list), sp)
    return list(out[0, :, 2]), list(out[0, :, 3])


def cal_quintic_coefficient(start_l, start_dl, start_ddl, end_l, end_dl, end_ddl, start_s, end_s):
    """Quintic l(s) through the two boundary states; coefficients of the absolute-s power basis
    [a0..a5] (reference :671-703).  Closed form on the GPU instead of the reference's 6x6 matrix inverse: the
    two agree to the inverse's own rounding noise (its condition number reaches 1e16 at s ~ 75 m)."""
    c = _runtime.planner().quintic([[start_l, start_dl, start_ddl, end_l, end_dl, end_ddl, start_s, end_s]])
    return list(c[0])


def cal_projection_s_fun(local_path_opt: list, match_index_list: list, xy_list: list, s_map: list):
    """s of the projections of points whose match indices are given (reference :429-445)."""
    if len(match_index_list) == 0:
        return []
    line = _line(local_path_opt)
    out = []
    # one GPU call per distinct match index keeps the kernel's "scan from the given node" semantics trivial:
    # a 1-node window makes the nearest-node scan return exactly that node
    for mi, pt in zip(match_index_list, xy_list):
        mi = int(mi)
        win = np.repeat(line[:, mi:mi + 1, :], 2, axis=1)            # the kernel needs >= 2 nodes
        smap = np.array([[s_map[mi], s_map[mi]]], dtype=np.float64)
        r = _runtime.planner().frenet_project(win, points=_pts([pt]), s_map_in=smap, want=("sl",))
        out.append(r["sl"][0, 0, 0])
    return out


def trajectory_index2s(trajectory_x, trajectory_y):
    """Cumulative arc length of a NaN-padded trajectory (reference :758-780)."""
    xy = np.array([trajectory_x, trajectory_y], dtype=np.float64)
    n = xy.shape[1]
    return _runtime.planner().misc(_lib.OP_INDEX2S, xy[None], n, 1, (1, n))[0]


def cal_dy_obs_deri(l_set, vx_set, vy_set, proj_heading_set, proj_kappa_set):
    """s_dot, l_dot and dl/ds of dynamic obstacles, 128-slot NaN-padded arrays (reference :783-808)."""
    inp = np.array([l_set, vx_set, vy_set, proj_heading_set, proj_kappa_set], dtype=np.float64)
    o = _runtime.planner().misc(_lib.OP_DY_OBS_DERI, inp[None], inp.shape[1], 1, (1, 3, 128))
    return o[0, 0], o[0, 1], o[0, 2]


def _path4(frenet_path_x, frenet_path_y, frenet_path_heading, frenet_path_kappa):
    return np.array([frenet_path_x, frenet_path_y, frenet_path_heading, frenet_path_kappa], dtype=np.float64)


def CalcProjPoint(s, frenet_path_x, frenet_path_y, frenet_path_heading, frenet_path_kappa, s_map):
    """Projection of arc length s on a path given as separate arrays (reference :736-755)."""
    o = _runtime.planner().misc(_lib.OP_CALC_PROJ_POINT, [[s], [0.0], [0.0], [0.0]], 1, 1, (4, 1),
                                path=_path4(frenet_path_x, frenet_path_y, frenet_path_heading, frenet_path_kappa),
                                s_map=np.asarray(s_map, dtype=np.float64))
    return o[0, 0], o[1, 0], o[2, 0], o[3, 0]


def Frenet2Cartesian(s_set, l_set, dl_set, ddl_set, frenet_path_x, frenet_path_y, frenet_path_heading,
                     frenet_path_kappa, index2s):
    """(s, l, dl, ddl) -> (x, y, heading, kappa), 600-slot NaN-padded column arrays (reference :706-733)."""
    inp = np.array([s_set, l_set, dl_set, ddl_set], dtype=np.float64)
    n = inp.shape[1]
    o = _runtime.planner().misc(_lib.OP_FRENET2CART, inp, n, 1, (4, n),
                                path=_path4(frenet_path_x, frenet_path_y, frenet_path_heading, frenet_path_kappa),
                                s_map=np.asarray(index2s, dtype=np.float64))
    out = [np.ones((600, 1)) * np.nan for _ in range(4)]
    for c in range(4):
        out[c][:n, 0] = o[c]
    return tuple(out)


def cal_proj_point_1(s, pre_match_index, frenet_path_opt: list, s_map: list):
    """Projection of arc length s on the reference line (reference :647-668; the same walk as
    path_planning.cal_proj_point).  Returns (proj_x, proj_y, proj_theta, proj_kappa, match_index)."""
    from . import path_planning
    return path_planning.cal_proj_point(s, pre_match_index, frenet_path_opt, s_map)


def predict_block_based_on_frenet(vehicle_loc, vehicle_velocity, local_frenet_path_opt, cur_path_s, cur_path_l, ts=0.1):
    """Ego position one planning period ahead, along the current path (reference :617-644).  vehicle_loc /
    vehicle_velocity only need .x/.y(/.z) attributes (carla.Location / carla.Vector3D in the reference)."""
    speed = math.sqrt(vehicle_velocity.x * vehicle_velocity.x + vehicle_velocity.y * vehicle_velocity.y +
                      vehicle_velocity.z * vehicle_velocity.z)
    s_map = cal_s_map_fun(local_frenet_path_opt, origin_xy=(vehicle_loc.x, vehicle_loc.y))
    s = max(cur_path_s[0], 0) + speed * ts
    index = int(np.argmin(abs(np.array(cur_path_s) - s)))          # nearest sample of the current path: a selection
    l = cur_path_l[index]
    proj_x, proj_y, proj_theta, _, _ = cal_proj_point_1(s, 0, local_frenet_path_opt, s_map)
    return proj_x + l * -math.sin(proj_theta), proj_y + l * math.cos(proj_theta)
