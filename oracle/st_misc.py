"""Oracle: the small S-T / speed-draft helpers that do run in the reference (numpy float64).

TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates reference planner/speed_planning_test.py
  calc_speed_planning_start_condition :23-35 -> start_condition()
  generate_st_graph                   :38-98 -> st_graph()
  CalcDpCost :191-231 / CalcObsCost :234-271 / CalcCollisionCost :274-284 -> see oracle/st_dp.py (edge form)
and planner/planning_utils.py
  trajectory_index2s :758-780 -> trajectory_index2s()
  cal_dy_obs_deri    :783-808 -> dy_obs_deri()
  CalcProjPoint      :736-755 -> calc_proj_point()
  Frenet2Cartesian   :706-733 -> frenet2cartesian()
Pinned by tests/golden/st_misc.npz (outputs of the unmodified reference).  Not restated because they cannot
run in the reference: speed_QP (:403, transposed Aeq, lb/ub never passed), and increase_points /
path_speed_merge on a full 17-point profile (IndexError at :538).
"""
import numpy as np


def start_condition(vx, vy, ax, ay, heading):
    tor = np.array([np.cos(heading), np.sin(heading)])
    return np.dot(tor.T, np.array([vx, vy])), np.dot(tor.T, np.array([ax, ay]))


def st_graph(s_set, l_set, s_dot_set, l_dot_set):
    n = len(s_set)
    out = [np.ones(n) * np.nan for _ in range(4)]           # s_in, s_out, t_in, t_out
    for i in range(n):
        if np.isnan(s_set[i]):
            break
        if abs(l_dot_set[i]) < 0.3:                          # slow lateral motion: ignored either way (:53-66)
            continue
        t_zero = -l_set[i] / l_dot_set[i]
        b1 = 2 / l_dot_set[i] + t_zero
        b2 = -2 / l_dot_set[i] + t_zero
        t_max, t_min = (b1, b2) if b1 > b2 else (b2, b1)
        if t_max < 1 or t_min > 8:
            continue
        if t_min < 0 and t_max > 0:
            out[0][i] = s_set[i]
            out[1][i] = s_set[i] + s_dot_set[i] * t_max
            out[2][i] = 0
            out[3][i] = t_max
        else:
            out[0][i] = s_set[i] + s_dot_set[i] * t_min
            out[1][i] = s_set[i] + s_dot_set[i] * t_max
            out[2][i] = t_min
            out[3][i] = t_max
    return tuple(out)


def trajectory_index2s(x, y):
    n = len(x)
    out = np.zeros(n)
    s = 0
    for i in range(1, n):
        if np.isnan(x[i]):
            break
        s += np.sqrt((x[i] - x[i - 1]) ** 2 + (y[i] - y[i - 1]) ** 2)
        out[i] = s
    return out


def dy_obs_deri(l_set, vx_set, vy_set, heading_set, kappa_set):
    n = 128
    s_dot, l_dot, dl = [np.ones(n) * np.nan for _ in range(3)]
    for i in range(len(l_set)):
        if np.isnan(l_set[i]):
            break
        v = np.array([vx_set[i], vy_set[i]])
        nr = np.array([-np.sin(heading_set[i]), np.cos(heading_set[i])])
        tr = np.array([np.cos(heading_set[i]), np.sin(heading_set[i])])
        l_dot[i] = np.dot(v, nr)
        s_dot[i] = np.dot(v, tr) / (1 - kappa_set[i] * l_set[i])
        dl[i] = 0 if abs(s_dot[i]) < 1e-6 else l_dot[i] / s_dot[i]
    return s_dot, l_dot, dl


def calc_proj_point(s, px, py, heading, kappa, s_map):
    i = 1
    while s_map[i] < s:                                      # starts at 1 and uses node i itself (:742-748)
        i += 1
    ds = s - s_map[i]
    return px[i] + ds * np.cos(heading[i]), py[i] + ds * np.sin(heading[i]), heading[i] + ds * kappa[i], kappa[i]


def frenet2cartesian(s_set, l_set, dl_set, ddl_set, px, py, heading, kappa, s_map):
    out = [np.ones((600, 1)) * np.nan for _ in range(4)]
    for i in range(len(s_set)):
        if np.isnan(s_set[i]):
            break
        x, y, h, k = calc_proj_point(s_set[i], px, py, heading, kappa, s_map)
        out[0][i] = x + l_set[i] * -np.sin(h)
        out[1][i] = y + l_set[i] * np.cos(h)
        out[2][i] = h + np.arctan(dl_set[i] / (1 - k * l_set[i]))
        d = out[2][i] - h
        out[3][i] = ((ddl_set[i] + k * dl_set[i] * np.tan(d)) * (np.cos(d) ** 2) / (1 - k * l_set[i]) + k) * \
            np.cos(d) / (1 - k * l_set[i])
    return tuple(out)
