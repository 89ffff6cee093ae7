"""Oracle: Cartesian<->Frenet helpers (numpy float64).

TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates reference planner/planning_utils.py:
  find_match_points        :49-182   -> find_match_points()
  match_projection_points  :364-426  -> match_projection()
  cal_projection_s_fun     :429-445  -> projection_s()
  cal_s_map_fun            :448-472  -> s_map()
  cal_s_l_fun              :475-509  -> s_l()
  cal_s_l_deri_fun         :512-588  -> s_l_deriv()
  cal_heading_kappa        :185-228  -> heading_kappa()
  sampling                 :231-259  -> sampling()
  smooth_reference_line    :262-361  -> smooth_matrices() / smooth_reference_line()
and planner/path_planning.py:
  cal_proj_point           :52-75    -> proj_point()
  frenet_2_x_y_theta_kappa :15-49    -> frenet_to_cartesian()
A reference line is an [N,4] array of (x, y, theta, kappa).
"""
import math
import numpy as np
from oracle import coneqp


def _scan(px, py, line, start, stop, step, limit):
    """Nearest-node scan with early exit after `limit` consecutive non-improvements."""
    best, best_i, inc = float("inf"), 0, 0
    for i in range(start, stop, step):
        d = math.sqrt((line[i][0] - px) ** 2 + (line[i][1] - py) ** 2)
        if d < best:
            best, best_i, inc = d, i, 0
        else:
            inc += 1
            if inc >= limit:
                break
    return best_i


def _project(px, py, node):
    """Projection of (px,py) through match node (x,y,theta,kappa) (planning_utils.py:414-424)."""
    xm, ym, th, km = node
    d = np.dot(np.array([px - xm, py - ym]), np.array([np.cos(th), np.sin(th)]))
    xr, yr = np.array([xm, ym]) + d * np.array([np.cos(th), np.sin(th)])
    return (xr, yr, th + km * d, km)


def match_projection(xy_list, line):
    """match_projection_points.  Quirk Q12: every projection uses the match
    index of point 0 (planning_utils.py:413)."""
    idx = np.zeros(len(xy_list), dtype=np.int32)
    proj = []
    for p, pt in enumerate(xy_list):
        idx[p] = _scan(pt[0], pt[1], line, 0, len(line), 1, 50)
        proj.append(_project(pt[0], pt[1], line[idx[0]]))
    return list(idx), proj


def find_match_points(xy_list, line, is_first_run, pre_match_index):
    """find_match_points (planning_utils.py:49-182): first run scans from 0 with
    limit 50; later runs scan from pre_match_index forward or backward (sign of
    the along-track offset) with limit 5.  Same index-0 projection quirk (:103,:169)."""
    idx = np.zeros(len(xy_list), dtype=np.int32)
    proj = []
    for p, (px, py) in enumerate(xy_list):
        if is_first_run is True:
            idx[p] = _scan(px, py, line, 0, len(line), 1, 50)
        else:
            st = pre_match_index
            th = line[st][2]
            flag = np.dot(np.array([px - line[st][0], py - line[st][1]]), np.array([np.cos(th), np.sin(th)]))
            if flag > 0:
                idx[p] = _scan(px, py, line, st, len(line), 1, 5)
            else:
                idx[p] = _scan(px, py, line, st, -1, -1, 5)
        proj.append(_project(px, py, line[idx[0]]))
    return list(idx), proj


def projection_s(line, match_idx, xy_list, smap):
    out = []
    for i, mi in enumerate(match_idx):
        x, y, th, _ = line[mi]
        d = np.dot(np.array([xy_list[i][0] - x, xy_list[i][1] - y]), np.array([math.cos(th), math.sin(th)]))
        out.append(smap[mi] + d)
    return out


def s_map(line, origin_xy):
    """cal_s_map_fun: cumulative polyline length, shifted so the ego projection is s=0."""
    mi, _ = match_projection([origin_xy], line)
    ref = [0]
    for i in range(1, len(line)):
        ref.append(math.sqrt((line[i][0] - line[i - 1][0]) ** 2 + (line[i][1] - line[i - 1][1]) ** 2) + ref[-1])
    s0 = projection_s(line, [mi[0]], [origin_xy], ref)
    return list(np.array(ref) - s0[0])


def s_l(xy_list, line, smap):
    """cal_s_l_fun."""
    mi, proj = match_projection(xy_list, line)
    s_list = projection_s(line, mi, xy_list, smap)
    l_list = []
    for i, pt in enumerate(xy_list):
        px, py, th, _ = proj[i]
        l_list.append(np.dot(np.array([pt[0], pt[1]]) - np.array([px, py]), np.array([-math.sin(th), math.cos(th)])))
    return s_list, l_list


def s_l_deriv(xy_list, v_list, a_list, line, origin_xy):
    """cal_s_l_deri_fun -> (l, dl/dt, ds/dt, ddl/dt2, dl/ds, dds/dt2, ddl/ds2) lists.
    Note l uses origin_xy for every point (planning_utils.py:542)."""
    _, proj = match_projection(xy_list, line)
    out = [[] for _ in range(7)]
    for i in range(len(xy_list)):
        x, y, th, kap = proj[i]
        nor = np.array([-math.sin(th), math.cos(th)])
        tou = np.array([math.cos(th), math.sin(th)])
        l = np.dot(np.array([origin_xy[0], origin_xy[1]]) - np.array([x, y]), nor)
        vh = np.array([v_list[i][0], v_list[i][1]])
        ah = np.array([a_list[i][0], a_list[i][1]])
        dl = np.dot(vh, nor)
        ds = np.dot(vh, tou) / (1 - kap * l)
        ddl = np.dot(ah, nor) - kap * (1 - kap * l) * (ds ** 2)
        l_ds = 0 if abs(ds) < 1e-6 else dl / ds
        dds = (np.dot(ah, tou) + 2 * (ds ** 2 * kap * l_ds) + ds ** 2 * 0 * l) / (1 - kap * l)
        l_dds = 0 if abs(ds) < 1e-6 else (ddl - l_ds * dds) / (ds ** 2)
        for k, v in enumerate((l, dl, ds, ddl, l_ds, dds, l_dds)):
            out[k].append(v)
    return tuple(out)


def heading_kappa(xy):
    """cal_heading_kappa.  Quirk Q13: np.insert(d, -1, d[-1]) inserts before the
    last element (planning_utils.py:223)."""
    xs = np.array([p[0] for p in xy], dtype=np.float64)
    ys = np.array([p[1] for p in xy], dtype=np.float64)
    dx_ = np.diff(xs)
    dy_ = np.diff(ys)
    dx = (np.concatenate(([dx_[0]], dx_)) + np.concatenate((dx_, [dx_[-1]]))) / 2
    dy = (np.concatenate(([dy_[0]], dy_)) + np.concatenate((dy_, [dy_[-1]]))) / 2
    theta = np.arctan2(dy, dx)
    dth = np.diff(theta)
    pre = np.insert(dth, 0, dth[0])
    aft = np.insert(dth, -1, dth[-1])
    k = np.sin((pre + aft) / 2) / np.sqrt(dx ** 2 + dy ** 2)
    return list(theta), list(k)


def sampling(match_index, line, back_length=10, forward_length=50):
    """sampling.  Quirk Q14: arguments are overridden by 10/40 (planning_utils.py:244-245)."""
    back, fwd = 10, 40
    total = back + fwd
    if match_index < back:
        back = match_index
        fwd = total - back
    if (len(line) - match_index) - 1 < fwd:
        fwd = len(line) - match_index - 1
        back = total - fwd
    return list(line[match_index - back: match_index]) + list(line[match_index: match_index + fwd + 1])


def smooth_matrices(xy, w_smooth=0.4, w_length=0.3, w_ref=0.3, x_thre=0.2, y_thre=0.2):
    """P,q,G,h of smooth_reference_line (planning_utils.py:300-349)."""
    n = len(xy)
    ref = np.zeros(2 * n)
    ref[0::2] = [p[0] for p in xy]
    ref[1::2] = [p[1] for p in xy]
    thr = np.zeros(2 * n)
    thr[0::2] = x_thre
    thr[1::2] = y_thre
    A1 = np.zeros((2 * n - 4, 2 * n))
    for r in range(2 * n - 4):
        A1[r, r], A1[r, r + 2], A1[r, r + 4] = 1, -2, 1
    A2 = np.zeros((2 * n - 2, 2 * n))
    for r in range(2 * n - 2):
        A2[r, r], A2[r, r + 2] = 1, -1
    P = 2 * (w_smooth * np.dot(A1.T, A1) + w_length * np.dot(A2.T, A2) + w_ref * np.identity(2 * n))
    q = -2 * w_ref * ref
    G = np.concatenate((np.identity(2 * n), -np.identity(2 * n)))
    h = np.concatenate((ref + thr, -(ref - thr)))
    return P, q, G, h


def smooth_reference_line(xy, diverge_factor=None, **kw):
    P, q, G, h = smooth_matrices(xy, **kw)
    res = coneqp.qp(P, q, G, h, diverge_factor=diverge_factor)
    x = res["x"]
    pts = [(x[i], x[i + 1]) for i in range(0, len(x), 2)]
    th, k = heading_kappa(pts)
    return [pts[i] + (th[i], k[i]) for i in range(len(pts))], res


def proj_point(s, pre_idx, line, smap):
    """cal_proj_point (path_planning.py:52-75).  Unguarded walk (Q15) -> IndexError past the map."""
    i = pre_idx
    while smap[i + 1] < s:
        i += 1
    x, y, th, kap = line[i]
    ds = s - smap[i]
    px, py = np.array([x, y]) + ds * np.array([math.cos(th), math.sin(th)])
    return px, py, th + kap * ds, kap, i


def frenet_to_cartesian_raw(start_s, start_l, s_list, l_list, line, smap):
    """frenet_2_x_y_theta_kappa up to (not including) the trailing smoother (path_planning.py:29-46)."""
    px, py, th, kap, idx = proj_point(start_s, 0, line, smap)
    nx, ny = -math.sin(th), math.cos(th)
    out = [(px + start_l * nx, py + start_l * ny)]
    for i in range(len(l_list)):
        if s_list[i] > smap[-1]:
            break
        px, py, th, kap, idx = proj_point(s_list[i], idx, line, smap)
        nx, ny = -math.sin(th), math.cos(th)
        out.append((px + l_list[i] * nx, py + l_list[i] * ny, th, kap))
    return out


def frenet_to_cartesian(start_s, start_l, s_list, l_list, line, smap):
    raw = frenet_to_cartesian_raw(start_s, start_l, s_list, l_list, line, smap)
    path, _ = smooth_reference_line(raw)
    return path
