"""Per-process lazy GPU planner handle shared by the drop-in modules.

The reference calls the planner from a forked multiprocessing child (test_9.py:226-227), so CUDA
must not be initialised at import time: the EmpPlanner handle (stream, device workspace) is created
on the first planner call of the *calling* process and re-created after a fork.
"""
import os

import numpy as np

from emplanner_carla_b200 import api

_state = {"pid": None, "planner": None, "cfg": None}

MAX_BATCH = 64       # single-frame callers; the batched API (api.HostPlanner) is the throughput path
KMAX_DEFAULT = 8


def planner(kmax=KMAX_DEFAULT, n_enr_max=128, max_col=128):
    cfg = (kmax, n_enr_max, max_col)
    st = _state
    if st["pid"] != os.getpid() or st["planner"] is None or st["cfg"][0] < kmax or st["cfg"][1] < n_enr_max \
            or st["cfg"][2] < max_col:
        if st["planner"] is not None and st["pid"] == os.getpid():
            st["planner"].close()
            cfg = (max(kmax, st["cfg"][0]), max(n_enr_max, st["cfg"][1]), max(max_col, st["cfg"][2]))
        st["planner"] = api.HostPlanner(MAX_BATCH, kmax=cfg[0], n_enr_max=cfg[1], max_col=cfg[2])
        st["pid"] = os.getpid()
        st["cfg"] = cfg
    return st["planner"]


def kmax_for(k):
    """Obstacle slots needed for k obstacles (even, >= 8, <= 32)."""
    k = max(int(k), KMAX_DEFAULT)
    k += k & 1
    if k > 32:
        raise ValueError("at most 32 obstacles per frame are supported by the scene record (got %d)" % k)
    return k


def scene_record(obs_s_list, obs_l_list, s0, l0, dl0, ddl0, kmax):
    K = len(obs_s_list)
    obs_s = np.zeros((1, kmax))
    obs_l = np.zeros((1, kmax))
    obs_s[0, :K] = np.asarray(obs_s_list, dtype=np.float64).reshape(-1)
    obs_l[0, :K] = np.asarray(obs_l_list, dtype=np.float64).reshape(-1)
    return api.pack_scenes(obs_s, obs_l, [K], [float(s0)], [float(l0)], [float(dl0)], [float(ddl0)], kmax=kmax)
