"""ctypes binding of libemplanner.so (include/emplanner.h).  No fallback: if the
CUDA library is missing or a call fails, an exception is raised."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libemplanner.so")


class EmpError(RuntimeError):
    pass


class DpParams(C.Structure):
    """EmpDpParams: DP_algorithm keyword defaults (reference path_planning.py:276-279)."""
    _fields_ = [("sampling_res", C.c_double), ("w_collision", C.c_double), ("w_smooth", C.c_double * 3),
                ("w_ref", C.c_double), ("sample_s", C.c_double), ("sample_l", C.c_double),
                ("row", C.c_int32), ("col", C.c_int32)]


class QpParams(C.Structure):
    """EmpQpParams: Quadratic_planning / cal_lmin_lmax / cvxopt defaults (path_planning.py:78-81,222)."""
    _fields_ = [("dp_sampling_res", C.c_double), ("w_l", C.c_double), ("w_dl", C.c_double), ("w_ddl", C.c_double),
                ("w_dddl", C.c_double), ("w_centre", C.c_double), ("w_end_l", C.c_double), ("w_end_dl", C.c_double),
                ("w_end_ddl", C.c_double), ("host_d1", C.c_double), ("host_d2", C.c_double), ("host_w", C.c_double),
                ("obs_length", C.c_double), ("obs_width", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double),
                ("feastol", C.c_double), ("maxiters", C.c_int32), ("reserved", C.c_int32), ("diverge_factor", C.c_double)]


class StParams(C.Structure):
    """EmpStParams: speed_DP keyword defaults (reference speed_planning_test.py:101-102)."""
    _fields_ = [("reference_speed", C.c_double), ("w_ref_speed", C.c_double), ("w_accel", C.c_double),
                ("w_obs", C.c_double)]


class SmoothParams(C.Structure):
    """EmpSmoothParams: smooth_reference_line keyword defaults (reference planning_utils.py:262-264)."""
    _fields_ = [("w_smooth", C.c_double), ("w_length", C.c_double), ("w_ref", C.c_double), ("x_thre", C.c_double),
                ("y_thre", C.c_double), ("abstol", C.c_double), ("reltol", C.c_double), ("feastol", C.c_double),
                ("maxiters", C.c_int32), ("heading_only", C.c_int32), ("diverge_factor", C.c_double)]


class SpeedQpParams(C.Structure):
    """EmpSpeedQpParams: speed_QP / generate_convex_space keyword defaults (speed_planning_test.py:403-404, :309)."""
    _fields_ = [("w_s_dot2", C.c_double), ("w_v_ref", C.c_double), ("w_jerk", C.c_double),
                ("reference_speed", C.c_double), ("max_lateral_accel", C.c_double), ("abstol", C.c_double),
                ("reltol", C.c_double), ("feastol", C.c_double), ("maxiters", C.c_int32), ("repair_full_profile", C.c_int32),
                ("diverge_factor", C.c_double)]


class MiscAux(C.Structure):
    """EmpMiscAux: side inputs of emp_misc_host."""
    _fields_ = [("obs4", C.c_void_p), ("n_obs", C.c_int32), ("n_nodes", C.c_int32), ("sdot_tab", C.c_void_p),
                ("path", C.c_void_p), ("s_map", C.c_void_p), ("v0", C.c_double), ("st", StParams)]


OP_ST_GRAPH, OP_START_CONDITION, OP_DY_OBS_DERI, OP_INDEX2S = 0, 1, 2, 3
OP_ST_DP_COST, OP_ST_OBS_COST, OP_ST_COLLISION, OP_FRENET2CART, OP_CALC_PROJ_POINT = 4, 5, 6, 7, 8

_P = C.c_void_p
_DPP = C.POINTER(DpParams)
_QPP = C.POINTER(QpParams)
_SIGNATURES = {
    "emp_dp_params_default": (None, [_DPP]),
    "emp_qp_params_default": (None, [_QPP]),
    "emp_scene_stride": (C.c_int, [C.c_int]),
    "emp_plan_stride": (C.c_int, [C.c_int]),
    "emp_max_enriched": (C.c_int, [_DPP]),
    "emp_version": (C.c_int, []),
    "emp_last_error": (C.c_char_p, []),
    "emp_sl_order_dev": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "emp_sl_dp_dev": (C.c_int, [_P, C.c_int, C.c_int, _DPP, C.c_int, _P, _P, _P, _P]),
    "emp_sl_qp_dev": (C.c_int, [_P, C.c_int, C.c_int, _QPP, C.c_int, _P, _P, _P, _P]),
    "emp_qp_dev": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _QPP, _P, _P, _P]),
    "emp_corridor_dev": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, _QPP, _P, _P, _P]),
    "emp_sl_plan_dev": (C.c_int, [_P, C.c_int, C.c_int, _DPP, _QPP, C.c_int, _P, _P, _P, _P]),
    "emp_frenet_project_dev": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P, _P]),
    "emp_frenet_to_cart_dev": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "emp_frenet_project_host": (C.c_int, [_P, _P, C.c_int, _P, _P, _P, C.c_int, _P, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P]),
    "emp_frenet_to_cart_host": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "emp_smooth_params_default": (None, [C.POINTER(SmoothParams)]),
    "emp_smooth_line_dev": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(SmoothParams), _P, _P, _P]),
    "emp_smooth_line_host": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.POINTER(SmoothParams), _P, _P]),
    "emp_st_params_default": (None, [C.POINTER(StParams)]),
    "emp_st_dp_dev": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(StParams), _P, _P, _P, _P, _P]),
    "emp_st_dp_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(StParams), _P, _P, _P, _P]),
    "emp_speed_qp_params_default": (None, [C.POINTER(SpeedQpParams)]),
    "emp_st_convex_space_dev": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.POINTER(SpeedQpParams), C.c_int, _P, _P, _P]),
    "emp_st_speed_qp_dev": (C.c_int, [_P, _P, _P, C.c_int, C.POINTER(SpeedQpParams), _P, _P, _P]),
    "emp_st_increase_points_dev": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "emp_st_path_speed_merge_dev": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P, _P, _P]),
    "emp_st_back_host": (C.c_int, [_P, C.c_int, C.POINTER(SpeedQpParams), _P, _P, C.c_int, _P, C.c_int, _P, _P, _P, _P,
                                   _P, _P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "emp_cycle_path_points_dev": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "emp_cycle_path_tables_dev": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "emp_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "emp_planner_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "emp_planner_destroy": (None, [_P]),
    "emp_sl_plan_host": (C.c_int, [_P, _P, C.c_int, _DPP, _QPP, _P, _P]),
    "emp_sl_plan_submit": (C.c_int, [_P, C.c_int, _P, C.c_int, _DPP, _QPP, _P, _P]),
    "emp_sl_plan_wait": (C.c_int, [_P, C.c_int]),
    "emp_sl_dp_host": (C.c_int, [_P, _P, C.c_int, _DPP, _P, _P]),
    "emp_qp_host": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _QPP, _P, _P]),
    "emp_corridor_host": (C.c_int, [_P, _P, _P, C.c_int, _P, C.c_int, _QPP, _P, _P]),
    "emp_sl_plan_p2p_dev": (C.c_int, [_P, C.c_int, C.c_int, _DPP, _QPP, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "emp_sl_qp_p2p_dev": (C.c_int, [_P, C.c_int, C.c_int, _QPP, C.c_int, _P, C.c_int, _P, _P, _P]),
    "emp_sl_qp_gather_dev": (C.c_int, [_P, C.c_int, C.c_int, _QPP, C.c_int, _P, C.c_int, _P, C.c_uint64, _P, _P, _P, _P]),
    "emp_gather_wait_signal_dev": (C.c_int, [_P, C.c_int, C.c_uint64, _P, C.c_int, _P, C.c_int, _P]),
    "emp_device_alloc": (_P, [C.c_size_t]),
    "emp_device_free": (None, [_P]),
    "emp_ipc_export": (C.c_int, [_P, _P]),
    "emp_ipc_open": (C.c_int, [_P, C.POINTER(_P)]),
    "emp_ipc_close": (C.c_int, [_P]),
    "emp_quintic_host": (C.c_int, [_P, _P, C.c_int, _P]),
    "emp_edge_cost_host": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _DPP, _P]),
    "emp_edge_cost_ex_host": (C.c_int, [_P, _P, C.c_int, _P, _P, C.c_int, _DPP, C.c_double, C.c_double, _P]),
    "emp_obs_cost_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, _P]),
    "emp_enrich_host": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_double, C.c_int, _P, _P, _P]),
    "emp_misc_host": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.POINTER(MiscAux), _P]),
    "emp_host_alloc": (_P, [C.c_size_t]),
    "emp_host_free": (None, [_P]),
    "emp_planner_last_kernel_ms": (C.c_double, [_P]),
}

_lib = None


def exported_symbols():
    """Names include/emplanner.h declares (a CPU test checks the .so exports every one of them)."""
    return sorted(_SIGNATURES)


def lib():
    """Load the shared library once.  Loading does NOT touch CUDA (fork-safe, reference test_9.py:226-227)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EmpError("libemplanner.so not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `python emplanner_carla_b200/csrc/build.py`. There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise EmpError("libemplanner error %d: %s" % (rc, lib().emp_last_error().decode()))


def dp_params(sampling_res=2, w_collision_cost=1e12, w_smooth_cost=(300, 1000, 5000), w_reference_cost=20,
              row=12, col=6, sample_s=15, sample_l=1.5):
    p = DpParams()
    p.sampling_res = float(sampling_res)
    p.w_collision = float(w_collision_cost)
    p.w_smooth = (C.c_double * 3)(*[float(v) for v in w_smooth_cost])
    p.w_ref = float(w_reference_cost)
    p.sample_s = float(sample_s)
    p.sample_l = float(sample_l)
    p.row = int(row)
    p.col = int(col)
    return p


def qp_params(dp_sampling_res=2, w_cost_l=1000, w_cost_dl=10000, w_cost_ddl=3000, w_cost_dddl=150,
              w_cost_centre=250, w_cost_end_l=40, w_cost_end_dl=40, w_cost_end_ddl=40, host_d1=3, host_d2=3,
              host_w=3, obs_length=5, obs_width=5, abstol=1e-7, reltol=1e-6, feastol=1e-7, maxiters=100,
              diverge_factor=2.0):
    p = QpParams()
    p.dp_sampling_res = float(dp_sampling_res)
    p.w_l, p.w_dl, p.w_ddl, p.w_dddl = float(w_cost_l), float(w_cost_dl), float(w_cost_ddl), float(w_cost_dddl)
    p.w_centre = float(w_cost_centre)
    p.w_end_l, p.w_end_dl, p.w_end_ddl = float(w_cost_end_l), float(w_cost_end_dl), float(w_cost_end_ddl)
    p.host_d1, p.host_d2, p.host_w = float(host_d1), float(host_d2), float(host_w)
    p.obs_length, p.obs_width = float(obs_length), float(obs_width)
    p.abstol, p.reltol, p.feastol = float(abstol), float(reltol), float(feastol)
    p.maxiters = int(maxiters)
    p.diverge_factor = float(diverge_factor or 0.0)
    return p


def st_params(reference_speed=50, w_cost_ref_speed=4000, w_cost_accel=100, w_cost_obs=10000000):
    p = StParams()
    p.reference_speed, p.w_ref_speed = float(reference_speed), float(w_cost_ref_speed)
    p.w_accel, p.w_obs = float(w_cost_accel), float(w_cost_obs)
    return p


def speed_qp_params(w_cost_s_dot2=10, w_cost_v_ref=50, w_cost_jerk=500, reference_speed=50,
                    max_lateral_accel=0.2 * 9.8, abstol=1e-7, reltol=1e-6, feastol=1e-7, maxiters=100,
                    diverge_factor=100.0, repair_full_profile=False):
    p = SpeedQpParams()
    p.w_s_dot2, p.w_v_ref, p.w_jerk = float(w_cost_s_dot2), float(w_cost_v_ref), float(w_cost_jerk)
    p.reference_speed, p.max_lateral_accel = float(reference_speed), float(max_lateral_accel)
    p.abstol, p.reltol, p.feastol, p.maxiters = float(abstol), float(reltol), float(feastol), int(maxiters)
    p.diverge_factor = float(diverge_factor or 0.0)
    p.repair_full_profile = 1 if repair_full_profile else 0
    return p


def smooth_params(w_cost_smooth=0.4, w_cost_length=0.3, w_cost_ref=0.3, x_thre=0.2, y_thre=0.2,
                  abstol=1e-7, reltol=1e-6, feastol=1e-7, maxiters=100, heading_only=False, diverge_factor=2.0):
    p = SmoothParams()
    p.heading_only = 1 if heading_only else 0
    p.w_smooth, p.w_length, p.w_ref = float(w_cost_smooth), float(w_cost_length), float(w_cost_ref)
    p.x_thre, p.y_thre = float(x_thre), float(y_thre)
    p.abstol, p.reltol, p.feastol, p.maxiters = float(abstol), float(reltol), float(feastol), int(maxiters)
    p.diverge_factor = float(diverge_factor or 0.0)
    return p
