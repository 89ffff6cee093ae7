/*
 * emplanner.h - C ABI of libemplanner.so: the B200 (sm_100a) implementation of the
 * Frenet EM-planner hot path of 6Lackiu/EMplanner_Carla @ 1f115a5.
 *
 * The reference has no FFI: its boundary is the set of module-level Python
 * functions in planner/path_planning.py, planner/planning_utils.py and
 * planner/speed_planning_test.py.  Each entry point below cites the reference
 * function(s) it replaces; emplanner_carla_b200/planner/*.py binds them with
 * ctypes under the reference's own function names.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - `*_dev` entry points take DEVICE pointers and a cudaStream_t (as void*);
 *     they only enqueue work.  `*_host` entry points take HOST pointers, own a
 *     device workspace inside an EmpPlanner handle, and return after the result
 *     is in the host buffers.
 *   - return value: 0 = ok, <0 = error (emp_last_error() gives the text).
 *     Per-frame algorithmic conditions are reported in per-frame status words,
 *     never through the return value.
 *   - all arithmetic is IEEE float64, like the reference (numpy float64).
 */
#ifndef EMPLANNER_H_
#define EMPLANNER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMP_VERSION 100          /* 0.1.0 */
#define EMP_NSAMPLE 10           /* samples per lattice edge, path_planning.py:486,559 */
#define EMP_MAX_ROW 16           /* lattice rows supported by the DP kernel (9, 11 and 12 rows run instantiations with
                                    the row count fixed at compile time; same results, ~20 % faster) */
#define EMP_MAX_KMAX 32          /* obstacle slots per scene record */
#define EMP_MAX_QP_POINTS 128    /* QP stations per frame: <= 32 one warp per frame (the throughput path), 33..64 and
                                    65..128 one CTA of 64 / 128 threads per frame (same arithmetic, exchanges through
                                    shared memory; e.g. sampling_res = 1 gives 46 stations) */
#define EMP_MAX_PEERS 8          /* GPUs of one NVSwitch box */
#define EMP_PLAN_SLOTS 4         /* batches in flight through emp_sl_plan_submit */
#define EMP_ST_KMAX 16           /* dynamic-obstacle slots per S-T scene record */
#define EMP_ST_ROWS 40           /* s samples of the S-T lattice, speed_planning_test.py:114 */
#define EMP_ST_COLS 16           /* t samples (0.5 s), speed_planning_test.py:116 */
#define EMP_ST_OUT_STRIDE 40     /* words per S-T result record */
#define EMP_ST_QP_POINTS 17      /* output length of speed_QP, speed_planning_test.py:422 */
#define EMP_ST_DENSE 401         /* output length of increase_points / path_speed_merge, :541, :571 */

/* ---- per-frame status bits ------------------------------------------------ */
#define EMP_DP_NO_OBSTACLE      0x1   /* K == 0: straight branch, path_planning.py:362-363 */
#define EMP_DP_ALL_COLLIDE      0x2   /* final cost > w_collision: the reference prints its banner, :351-353 */
#define EMP_DP_OVERFLOW         0x4   /* densified path longer than n_enr_max (output truncated) */
#define EMP_QP_OK               0     /* 'optimal' */
#define EMP_QP_UNKNOWN_MAXITER  1     /* cvxopt status 'unknown' after maxiters */
#define EMP_QP_UNKNOWN_NUMERIC  2     /* 'unknown': factorisation breakdown / non-finite iterate */
#define EMP_QP_UNKNOWN_DIVERGED 3     /* 'unknown': duality gap grew above diverge_factor x its smallest earlier value */
#define EMP_QP_CORRIDOR_INDEX   4     /* cal_lmin_lmax would raise IndexError, path_planning.py:265-272 */
#define EMP_QP_TOO_LARGE        8     /* more than EMP_MAX_QP_POINTS (128) stations, or fewer than 2 */
#define EMP_ST_VALUE_ERROR      0x10  /* the reference raises ValueError (scipy interp1d out of range, empty table) */
#define EMP_ST_INDEX_ERROR      0x20  /* the reference raises IndexError (S-T profile without NaN terminator, index 16) */
#define EMP_ST_BAD_PROFILE      0x40  /* S-T profile with fewer than 2 valid points: nothing to smooth */

/* ---- parameters (defaults = the reference's keyword defaults) -------------- */
typedef struct EmpDpParams {          /* DP_algorithm, path_planning.py:276-279 */
    double sampling_res;              /* 2 */
    double w_collision;               /* 1e12 */
    double w_smooth[3];               /* 300, 1000, 5000 */
    double w_ref;                     /* 20 */
    double sample_s;                  /* 15 */
    double sample_l;                  /* 1.5 */
    int32_t row;                      /* 12 */
    int32_t col;                      /* 6 */
} EmpDpParams;

typedef struct EmpQpParams {          /* Quadratic_planning :78-81, cal_lmin_lmax :222, cvxopt defaults */
    double dp_sampling_res;           /* 2 */
    double w_l, w_dl, w_ddl, w_dddl;  /* 1000, 10000 (unused: :193), 3000, 150 */
    double w_centre;                  /* 250 */
    double w_end_l, w_end_dl, w_end_ddl; /* 40, 40, 40 */
    double host_d1, host_d2, host_w;  /* 3, 3, 3 */
    double obs_length, obs_width;     /* 5, 5 (test_9.py:190-193) */
    double abstol, reltol, feastol;   /* 1e-7, 1e-6, 1e-7 */
    int32_t maxiters;                 /* 100 */
    int32_t reserved;
    double diverge_factor;            /* 2: see "divergence exit" below; 0 = iterate like cvxopt until breakdown */
} EmpQpParams;

/* Divergence exit (all three QP kernels).  cvxopt's coneqp has no infeasibility certificate: on an infeasible
 * problem (empty corridor) it iterates until its Cholesky factorisation breaks down or maxiters and returns
 * status 'unknown' with whatever iterate it has; the reference never looks at the status (path_planning.py:211-219).
 * Those last iterates are not a numerically well defined function of the input (two exact-arithmetic-equivalent
 * factorisations differ by O(1 m), profiles/parity_r2.json).  With diverge_factor > 0 a solve ends with
 * EMP_QP_UNKNOWN_DIVERGED as soon as the duality gap exceeds diverge_factor x the smallest gap of the earlier
 * iterations - the last iterate both sides still agree on to < 1e-6 m.  The gap of a converging solve decreases
 * monotonically (worst ratio 0.98 over 3349 optimal frames), so 'optimal' frames are untouched.  oracle/coneqp.py
 * implements the same rule behind the same parameter. */

typedef struct EmpStParams {          /* speed_DP, speed_planning_test.py:101-102 */
    double reference_speed;           /* 50 */
    double w_ref_speed;               /* 4000 */
    double w_accel;                   /* 100 */
    double w_obs;                     /* 10000000 */
} EmpStParams;

typedef struct EmpSmoothParams {      /* smooth_reference_line, planning_utils.py:262-264 + cvxopt defaults */
    double w_smooth, w_length, w_ref; /* 0.4, 0.3, 0.3 */
    double x_thre, y_thre;            /* 0.2, 0.2 */
    double abstol, reltol, feastol;   /* 1e-7, 1e-6, 1e-7 */
    int32_t maxiters;                 /* 100 */
    int32_t heading_only;             /* 1: skip the QP, output cal_heading_kappa (:185-228) of the input polyline */
    double diverge_factor;            /* 2, as in EmpQpParams */
} EmpSmoothParams;

typedef struct EmpSpeedQpParams {     /* speed_QP :403-404, generate_convex_space :309, cvxopt defaults */
    double w_s_dot2, w_v_ref, w_jerk; /* 10, 50, 500 */
    double reference_speed;           /* 50 */
    double max_lateral_accel;         /* 0.2 * 9.8 */
    double abstol, reltol, feastol;   /* 1e-7, 1e-6, 1e-7 */
    int32_t maxiters;                 /* 100 */
    int32_t repair_full_profile;      /* 0.  1 = repair R3: a DP speed profile that fills all 16 columns (no NaN
                                         terminator) ends at index 15 instead of raising the reference's IndexError at
                                         dp_speed_s[16] (speed_planning_test.py:415-428: `dp_speed_end = 16` is the
                                         MATLAB original's 1-based count, every other branch sets the 0-based index) */
    double diverge_factor;            /* 100: same rule as EmpQpParams, looser - the gap of a converging speed QP is not
                                         monotone in its first iterations (ratios up to 11 on the golden frames) */
} EmpSpeedQpParams;

void emp_speed_qp_params_default(EmpSpeedQpParams *p);
void emp_smooth_params_default(EmpSmoothParams *p);
void emp_st_params_default(EmpStParams *p);
void emp_dp_params_default(EmpDpParams *p);
void emp_qp_params_default(EmpQpParams *p);

/* ---- data layouts ----------------------------------------------------------
 * Scene record (input), float64 words, stride = emp_scene_stride(kmax):
 *   [0] plan_start_s  [1] plan_start_l  [2] plan_start_dl  [3] plan_start_ddl
 *   [4] number of obstacles K (0..kmax)   [5] reserved
 *   [6 .. 6+kmax)         obs_s   (entries >= K ignored)
 *   [6+kmax .. 6+2*kmax)  obs_l
 * kmax must be even (records stay 16-byte aligned for the bulk copy).
 *
 * Plan record (output), float64 words, stride = emp_plan_stride(n_enr_max):
 *   [0] n_enr   [1] n_qp   [2] dp status bits   [3] qp status   [4] qp iterations
 *   [5] DP final cost   [6] final duality gap   [7] reserved
 *   [8 .. 8+NE)                 enriched s   (DP_algorithm output, first n_enr valid)
 *   [8+NE .. 8+2NE)             enriched l
 *   [8+2NE .. 8+2NE+NQ)         qp_path_l    (first n_qp valid), NQ = (NE+1)/2
 *   [.. +NQ)                    qp_path_dl
 *   [.. +NQ)                    qp_path_ddl
 * with NE = n_enr_max; stride rounded up to an even word count.
 */
int emp_scene_stride(int kmax);
int emp_plan_stride(int n_enr_max);
int emp_max_enriched(const EmpDpParams *p);   /* upper bound on n_enr for these lattice parameters */

int emp_version(void);
const char *emp_last_error(void);

typedef struct EmpPlanner EmpPlanner;   /* host-buffer planner handle, see below */

/* ---- device-pointer entry points -------------------------------------------*/

/* S-L lattice DP + densify for a batch of frames, one launch.
 * Replaces path_planning.DP_algorithm (:276-375) with cal_start_cost (:435),
 * cal_neighbor_cost (:517), cal_obs_cost (:588), enrich_DP_s_l (:378) and
 * planning_utils.cal_quintic_coefficient (:671).
 * scenes: [batch, emp_scene_stride(kmax)]; plans: [batch, emp_plan_stride(n_enr_max)]
 * (words 0,2,5 and the enriched arrays are written); rows: optional [batch, col]
 * int32 lattice row per column (-1 for K == 0 frames), may be NULL.
 * order: optional [batch] schedule from emp_sl_order_dev (work slot -> frame); NULL = frames in their own order.
 * The schedule only decides which warp works on which frame; every result is written at the frame's own index. */
int emp_sl_dp_dev(const double *scenes, int kmax, int batch, const EmpDpParams *p,
                  int n_enr_max, double *plans, int32_t *rows, const int32_t *order, void *stream);

/* Schedule for the two kernels: order[0..batch) = frame indices by obstacle count, descending (one small launch).
 * Both kernels' cost grows with the obstacle count and one heavy frame alone takes ~2/3 of the time of a 4096-frame
 * launch, so the heavy frames are started first (DP) / handed out first by the ticket queue (QP). */
int emp_sl_order_dev(const double *scenes, int kmax, int batch, int32_t *order, void *stream);

/* Corridor + QP smoother for a batch, one launch; reads the enriched path that
 * emp_sl_dp_dev left in `plans`, takes every 2nd point (test_9.py:187-188).
 * Replaces path_planning.cal_lmin_lmax (:222-273) and Quadratic_planning
 * (:78-219, incl. the cvxopt.solvers.qp call :211).  Writes words 1,3,4,6 and
 * the qp arrays of each plan record.  lmin_lmax: optional [batch, 2, NQ] copy of
 * the corridor, may be NULL. */
int emp_sl_qp_dev(const double *scenes, int kmax, int batch, const EmpQpParams *p,
                  int n_enr_max, double *plans, double *lmin_lmax, const int32_t *order, void *stream);

/* Quadratic_planning alone on caller-supplied corridors (l_min/l_max: [batch, n]). */
int emp_qp_dev(const double *l_min, const double *l_max, const double *start_l_dl_ddl /*[batch,3]*/,
               int n, int batch, const EmpQpParams *p, double *out_l_dl_ddl /*[batch,3,n]*/,
               int32_t *status_iters /*[batch,2]*/, void *stream);

/* cal_lmin_lmax alone (:222-273) on caller-supplied paths dp_s/dp_l: [batch, n]; obstacles come
 * from the scene records.  lmin_lmax: [batch, 2, n]; status_iters[2b] = 0 or EMP_QP_CORRIDOR_INDEX. */
int emp_corridor_dev(const double *dp_s, const double *dp_l, int n, const double *scenes, int kmax, int batch,
                     const EmpQpParams *p, double *lmin_lmax, int32_t *status_iters, void *stream);

/* DP -> [::2] -> corridor -> QP: the S-L planning step (two launches on `stream`; three with order_ws != NULL:
 * a [batch] int32 workspace that receives the emp_sl_order_dev schedule both kernels then follow). */
int emp_sl_plan_dev(const double *scenes, int kmax, int batch, const EmpDpParams *dp,
                    const EmpQpParams *qp, int n_enr_max, double *plans, int32_t *rows, int32_t *order_ws,
                    void *stream);

/* ---- Cartesian <-> Frenet (K4) ---------------------------------------------------------
 * lines: [batch, n_nodes, 4] reference line per scene as (x, y, theta, kappa) rows - the output of
 * planning_utils.sampling (:231, 51 nodes) after smooth_reference_line (:262).  n_nodes*32 bytes must be
 * a multiple of 16 (always true). */

/* planning_utils.cal_s_map_fun (:448-472) for every scene, and - when `points` is not NULL -
 * cal_s_l_fun (:475-509; match_projection_points :364, cal_projection_s_fun :429) for a list of up to
 * max_points points per scene (npts[batch], NULL = all max_points) and, when vel/acc/deriv are not NULL,
 * cal_s_l_deri_fun (:512-588) for the same points.
 *   origin [batch,2]; points/vel/acc [batch,max_points,2];
 *   s_map [batch,n_nodes]; match [batch,max_points] int32; sl [batch,max_points,2] = (s,l);
 *   deriv [batch,max_points,7] = (l, dl/dt, ds/dt, ddl/dt2, dl/ds, dds/dt2, ddl/ds2);
 *   proj [batch,max_points,4] = projection node (x_r, y_r, theta_r, kappa_r) of match_projection_points.
 * s_map_in [batch,n_nodes] (optional): use this s-map instead of computing one from `origin` (cal_s_l_fun is
 * called with the caller's s_map).  pre_match [batch] (optional): planning_utils.find_match_points (:49-182)
 * with is_first_run=False - scan from that node, forward or backward, early exit after 5.  Outputs may be NULL. */
int emp_frenet_project_dev(const double *lines, int n_nodes, const double *origin, const double *points,
                           const int32_t *npts, int max_points, const double *vel, const double *acc,
                           const double *s_map_in, const int32_t *pre_match, int batch,
                           double *s_map, int32_t *match, double *proj, double *sl, double *deriv, void *stream);

/* path_planning.frenet_2_x_y_theta_kappa (:15-46, everything before the trailing smooth_reference_line) with
 * cal_proj_point (:52-75).  start_sl [batch,2]; path_s/path_l [batch,max_points]; npath[batch] (NULL = all);
 * cart [batch, max_points+1, 4]: row 0 = (x, y, NaN, NaN) of the plan start, then (x, y, theta, kappa);
 * ncart[batch] = rows written (truncated where s > s_map[-1], :40-41), -1 if the start is beyond the map. */
int emp_frenet_to_cart_dev(const double *lines, const double *s_map, int n_nodes, const double *start_sl,
                           const double *path_s, const double *path_l, const int32_t *npath, int max_points,
                           int batch, double *cart, int32_t *ncart, void *stream);

/* Host-buffer versions (temporary device buffers on the planner's stream). */
int emp_frenet_project_host(EmpPlanner *pl, const double *lines, int n_nodes, const double *origin,
                            const double *points, const int32_t *npts, int max_points, const double *vel,
                            const double *acc, const double *s_map_in, const int32_t *pre_match, int batch,
                            double *s_map, int32_t *match, double *proj, double *sl, double *deriv);
int emp_frenet_to_cart_host(EmpPlanner *pl, const double *lines, const double *s_map, int n_nodes,
                            const double *start_sl, const double *path_s, const double *path_l,
                            const int32_t *npath, int max_points, int batch, double *cart, int32_t *ncart);

/* ---- reference-line smoother (K5) ----------------------------------------------------------
 * Replaces planning_utils.smooth_reference_line (:262-361, incl. cvxopt.solvers.qp :353) and
 * cal_heading_kappa (:185-228) for a batch of polylines, one launch.
 * xy: [batch, max_points, 2]; npts[batch] (NULL = max_points), 2 <= n <= 128 (max_points <= 64: one warp per
 * polyline, the throughput path; 65..128: one CTA of 64 threads per polyline); cal_heading_kappa alone: any length;
 * out: [batch, max_points, 4] = (x_opt, y_opt, theta, kappa); status_iters [batch,2] like emp_qp_dev. */
int emp_smooth_line_dev(const double *xy, const int32_t *npts, int max_points, int batch, const EmpSmoothParams *p,
                        double *out, int32_t *status_iters, void *stream);
int emp_smooth_line_host(EmpPlanner *pl, const double *xy, const int32_t *npts, int max_points, int batch,
                         const EmpSmoothParams *p, double *out, int32_t *status_iters);

/* ---- S-T speed DP (K3) -----------------------------------------------------------------
 * Replaces speed_planning_test.speed_DP (:101-188) with CalcDpCost (:191), CalcObsCost (:234),
 * CalcCollisionCost (:274), CalcSTCoordinate (:287), for a batch of frames, one launch.
 * scenes: [batch, 2 + 4*kmax] words = [plan_start_s_dot, -, s_in[kmax], s_out[kmax], t_in[kmax], t_out[kmax]],
 *   NaN s_in = absent obstacle (like the reference, :255); kmax <= EMP_ST_KMAX, even.
 * out: [batch, EMP_ST_OUT_STRIDE] words = [min_row, min_col, min_cost, -,-,-,-,-, dp_speed_s[16], dp_speed_t[16]]
 *   (NaN beyond min_col, like the reference's output initialisation :155).
 * cost_tab / sdot_tab [batch,40,16] float64 and node_tab [batch,40,16] int32: optional copies of the
 *   reference's dp_st_cost / dp_st_s_dot / dp_st_node tables (NULL to skip). */
int emp_st_dp_dev(const double *scenes, int kmax, int batch, const EmpStParams *p, double *out,
                  double *cost_tab, double *sdot_tab, int32_t *node_tab, void *stream);
int emp_st_dp_host(EmpPlanner *pl, const double *scenes, int kmax, int batch, const EmpStParams *p, double *out,
                   double *cost_tab, double *sdot_tab, int32_t *node_tab);

/* ---- S-T back half (K8, K9) ---------------------------------------------------------------
 * st_out: the [batch, EMP_ST_OUT_STRIDE] records emp_st_dp_dev wrote (dp_speed_s at word 8, dp_speed_t at word 24);
 * scenes: the same S-T scene records.  Per-frame status: 0, or EMP_ST_VALUE_ERROR / EMP_ST_INDEX_ERROR /
 * EMP_ST_BAD_PROFILE where the reference would raise (the drop-in wrappers re-raise).
 *
 * emp_st_convex_space_dev: speed_planning_test.generate_convex_space (:308-400).
 *   path [batch,2,n_path] = (path_index2s, trajectory_kappa_init), index2s zero padded (planning_utils.trajectory_index2s);
 *   bounds [batch,4,16] = (s_lb, s_ub, s_dot_lb, s_dot_ub), +-inf where unbounded; status [batch].
 * emp_st_speed_qp_dev: speed_planning_test.speed_QP (:403-511, incl. cvxopt.solvers.qp :502) with the two repairs
 *   the draft needs to run at all (equality matrix transposed; lb/ub as two arrays whose finite entries become
 *   inequality rows - oracle/st_back.py).  start [batch,2] = (plan_start_s_dot, plan_start_s_dot2);
 *   qp [batch,4,17] = (qp_s, qp_s_dot, qp_s_dot2, relative_time), NaN beyond qp_size; status_iters [batch,2].
 * emp_st_increase_points_dev: speed_planning_test.increase_points (:514-562).  qp [batch,4,17] -> dense [batch,4,401]
 *   = (s, s_dot, s_dot2, relative_time).
 * emp_st_path_speed_merge_dev: speed_planning_test.path_speed_merge (:565-611).  path5 [batch,5,n_path] = (path_s,
 *   trajectory_x/y/heading/kappa_init), x NaN padded; current_time [batch]; traj [batch,7,401] = (x, y, heading,
 *   kappa, speed, accel, time). */
int emp_st_convex_space_dev(const double *st_out, const double *scenes, int kmax, const double *path, int n_path,
                            const EmpSpeedQpParams *p, int batch, double *bounds, int32_t *status, void *stream);
int emp_st_speed_qp_dev(const double *start, const double *st_out, const double *bounds, int batch,
                        const EmpSpeedQpParams *p, double *qp, int32_t *status_iters, void *stream);
int emp_st_increase_points_dev(const double *qp, int batch, double *dense, int32_t *status, void *stream);
int emp_st_path_speed_merge_dev(const double *dense, const double *current_time, const double *path5, int n_path,
                                int batch, double *traj, int32_t *status, void *stream);
/* Host-buffer versions.  Any output pointer may be NULL; a stage whose inputs are missing is an argument error.
 * emp_st_back_host chains convex space -> speed QP -> increase_points -> path_speed_merge for the stages whose
 * outputs (or successors' outputs) are requested, starting from caller-supplied intermediate results when given:
 *   bounds_in [batch,4,16] (skip convex space), qp_in [batch,4,17] (skip the QP), dense_in [batch,4,401]. */
int emp_st_back_host(EmpPlanner *pl, int batch, const EmpSpeedQpParams *p,
                     const double *st_out, const double *scenes, int kmax, const double *path, int n_path,
                     const double *start, const double *bounds_in, const double *qp_in, const double *dense_in,
                     const double *current_time, const double *path5, int n_path5,
                     double *bounds, int32_t *convex_status, double *qp, int32_t *qp_status_iters,
                     double *dense, int32_t *dense_status, double *traj, int32_t *merge_status);

/* ---- glue between the S-L and the S-T half of an EM cycle (K10) ---------------------------------
 * In the reference this is list handling inside the driver scripts, not planner functions.
 * emp_cycle_path_points_dev: test_9.py:204-210 - mid-point averaging of (dp_path_s[::2], qp_path_l) of every plan
 *   record, behind the plan-start point of the frame's scene record (path_planning.py:31-34), into the polyline handed
 *   to smooth_reference_line (QP#3), planned along the s-axis (x = s, y = l).
 *   xy [batch, max_points, 2]; npts [batch] = n_qp + 2, or 0 when the frame has no QP path.
 * emp_cycle_path_tables_dev: test_10.py:234-244 - (x, y, theta, kappa) rows cart [batch, max_points, 4] (output
 *   of emp_smooth_line_dev) -> path2 [batch,2,n_path] = (trajectory_index2s, kappa) zero padded, and
 *   path5 [batch,5,n_path] = (path_s, x, y, heading, kappa) NaN padded; n_path > max(npts). */
int emp_cycle_path_points_dev(const double *plans, int n_enr_max, const double *scenes, int kmax, int batch,
                              int max_points, double *xy, int32_t *npts, void *stream);
int emp_cycle_path_tables_dev(const double *cart, const int32_t *npts, int max_points, int batch, int n_path,
                              double *path2, double *path5, void *stream);

/* ---- stand-alone pieces of the S-L DP (K6), host-buffer entry points ------------------------------
 * emp_quintic_host:   planning_utils.cal_quintic_coefficient (:671-703).  in [n,8] = (start_l, start_dl, start_ddl,
 *                     end_l, end_dl, end_ddl, start_s, end_s); out [n,6] coefficients of the absolute-s power basis.
 * emp_edge_cost_host: path_planning.cal_start_cost (:435) / cal_neighbor_cost (:517).  edges [n,6] = (start_s,
 *                     start_l, start_dl, start_ddl, end_s, end_l); obstacles shared by all edges; uses p->sample_s
 *                     (sample spacing sample_s/10), p->w_collision, p->w_smooth, p->w_ref.  out [n].
 * emp_obs_cost_host:  path_planning.cal_obs_cost (:588).  square_d [batch,n] -> out [batch].
 * emp_enrich_host:    path_planning.enrich_DP_s_l (:378).  node_s/node_l [batch,C], start [batch,4] = (s,l,dl,ddl);
 *                     out_s/out_l [batch,n_max], n_out [batch]. */
int emp_quintic_host(EmpPlanner *pl, const double *in, int n, double *out);
int emp_edge_cost_host(EmpPlanner *pl, const double *edges, int n_edges, const double *obs_s, const double *obs_l,
                       int n_obs, const EmpDpParams *p, double *out);
/* emp_edge_cost_ex_host: the same edge cost with explicit danger / safe distances: planner/speed_planning.py
 *                     cal_start_cost (:96-169) is that function on an (s, t) quintic with cal_obs_cost's (:172) 3 / 5. */
int emp_edge_cost_ex_host(EmpPlanner *pl, const double *edges, int n_edges, const double *obs_s, const double *obs_l,
                          int n_obs, const EmpDpParams *p, double danger_dis, double safe_dis, double *out);
int emp_obs_cost_host(EmpPlanner *pl, const double *square_d, int n, int batch, double w_collision, double danger_dis,
                      double safe_dis, double *out);
int emp_enrich_host(EmpPlanner *pl, const double *node_s, const double *node_l, const double *start, int C, int batch,
                    double resolution, int n_max, double *out_s, double *out_l, int32_t *n_out);

/* ---- small S-T / speed-draft helpers (K7), host-buffer entry points -------------------------------
 * One generic call: `op` selects the function, `in`/`out` are flat float64 arrays with the layout given below,
 * `aux` carries the per-op side inputs.
 *   EMP_OP_ST_GRAPH        speed_planning_test.generate_st_graph (:38-98)
 *                          in [batch,4,n] (s,l,s_dot,l_dot), out [batch,4,n] (s_in,s_out,t_in,t_out), NaN = absent
 *   EMP_OP_START_CONDITION speed_planning_test.calc_speed_planning_start_condition (:23-35)
 *                          in [batch,5] (vx,vy,ax,ay,heading), out [batch,2] (s_dot, s_dot2)
 *   EMP_OP_DY_OBS_DERI     planning_utils.cal_dy_obs_deri (:783-808)
 *                          in [batch,5,n] (l,vx,vy,heading,kappa), out [batch,3,128] (s_dot,l_dot,dl/ds), NaN padded
 *   EMP_OP_INDEX2S         planning_utils.trajectory_index2s (:758-780)   in [batch,2,n] (x,y), out [batch,n]
 *   EMP_OP_ST_DP_COST      speed_planning_test.CalcDpCost (:191-231)      in [n,4] (row_start,col_start,row_end,col_end)
 *   EMP_OP_ST_OBS_COST     speed_planning_test.CalcObsCost (:234-271)     in [n,4] (s_start,t_start,s_end,t_end)
 *   EMP_OP_ST_COLLISION    speed_planning_test.CalcCollisionCost (:274)   in [n] min_dis, out [n]
 *       for the three cost ops: aux.obs4 [4,n_obs] (s_in,s_out,t_in,t_out), aux.sdot_tab [40,16], aux.v0, aux.st
 *   EMP_OP_FRENET2CART / EMP_OP_CALC_PROJ_POINT  planning_utils.Frenet2Cartesian (:706-733) / CalcProjPoint (:736-755)
 *                          in [4,n] (s,l,dl,ddl), aux.path [4,n_nodes] (x,y,heading,kappa), aux.s_map [n_nodes], out [4,n] */
enum { EMP_OP_ST_GRAPH = 0, EMP_OP_START_CONDITION = 1, EMP_OP_DY_OBS_DERI = 2, EMP_OP_INDEX2S = 3,
       EMP_OP_ST_DP_COST = 4, EMP_OP_ST_OBS_COST = 5, EMP_OP_ST_COLLISION = 6, EMP_OP_FRENET2CART = 7,
       EMP_OP_CALC_PROJ_POINT = 8 };
typedef struct EmpMiscAux {
    const double *obs4; int32_t n_obs; int32_t n_nodes;
    const double *sdot_tab; const double *path; const double *s_map;
    double v0; EmpStParams st;
} EmpMiscAux;
int emp_misc_host(EmpPlanner *pl, int op, const double *in, int n, int batch, const EmpMiscAux *aux, double *out);

/* ---- fused planning step + all-gather over NVLink peer memory --------------------------------
 * Same as emp_sl_plan_dev, and every finished plan record is also stored into `peer_slices[r]` (r < n_peers):
 * device pointers into the OTHER GPUs' gather buffers at this rank's slice, obtained with emp_ipc_open.
 * The stores overlap the solves still running.  This entry point carries no completion signal: the gathered batch
 * is only guaranteed visible on a peer after the caller's own cross-rank barrier that follows the call in stream
 * order.  The product path (sharding.PeerGatherBuffer, bench.py) uses emp_sl_qp_gather_dev below instead, whose
 * arrival flags make that barrier unnecessary; this one is kept for callers that already synchronise each cycle.
 * Replaces the north_star's "single NCCL all-gather of the solved paths" with P2P writes from inside the QP kernel. */
int emp_sl_plan_p2p_dev(const double *scenes, int kmax, int batch, const EmpDpParams *dp, const EmpQpParams *qp,
                        int n_enr_max, double *plans, int32_t *rows, int n_peers, double *const *peer_slices,
                        int32_t *order_ws, void *stream);

/* The corridor+QP launch of the fused step alone (for callers that time the two kernels separately). */
int emp_sl_qp_p2p_dev(const double *scenes, int kmax, int batch, const EmpQpParams *qp, int n_enr_max,
                      double *plans, int n_peers, double *const *peer_slices, const int32_t *order, void *stream);

/* The same launch with arrival flags instead of a cross-rank barrier: when the last plan record of the launch has been
 * stored (locally and into every peer slice), `epoch` (> 0, increasing per use of the buffer) is written with release
 * semantics at system scope to *self_ready and to every *peer_ready[r] - this rank's word in the flag array of its own
 * and of each peer's gather buffer.  A consumer that acquires flag >= epoch sees the whole shard.
 * emp_gather_wait_signal_dev is that consumer and the flow control in one tiny launch: it spins (ld.acquire.sys) until
 * flags[r] >= epoch for all r < n_wait, then writes `epoch` (st.release.sys) to every signal[r], r < n_signal - e.g.
 * wait for all shards of a gather buffer, then tell every producer that the buffer may be overwritten (credit).  With
 * n_signal = 0 it only waits; with n_wait = 0 it only signals.  After timeout_ms (<= 0: 2000) without progress *err is
 * set to 1 and nothing is signalled (the GPU is never left spinning on a lost peer).
 * Together they replace the north_star's NCCL all-gather AND its completion barrier with NVLink stores from inside
 * the QP kernel; NCCL (torch.distributed) is only used to exchange the IPC handles at start-up. */
int emp_sl_qp_gather_dev(const double *scenes, int kmax, int batch, const EmpQpParams *qp, int n_enr_max,
                         double *plans, int n_peers, double *const *peer_slices, uint64_t epoch,
                         uint64_t *self_ready, uint64_t *const *peer_ready, const int32_t *order, void *stream);
int emp_gather_wait_signal_dev(const uint64_t *flags, int n_wait, uint64_t epoch, uint64_t *const *signal,
                               int n_signal, int32_t *err, int timeout_ms, void *stream);

/* Device buffers that can be shared between the per-GPU processes of one box (CUDA IPC). */
void *emp_device_alloc(size_t bytes);
void emp_device_free(void *p);
int emp_ipc_export(const void *dev_ptr, unsigned char handle[64]);
int emp_ipc_open(const unsigned char handle[64], void **dev_ptr);
int emp_ipc_close(void *dev_ptr);

/* FP64 FMA micro-benchmark (roofline denominator; absent from MEASURED_PEAKS.json).
 * Returns achieved TFLOP/s through *tflops. */
int emp_fp64_peak(int iters, double *tflops, double *ms);

/* ---- host-buffer entry points ----------------------------------------------*/

/* device < 0: current device.  CUDA is initialised here, not at library load
 * (the reference calls the planner from a forked child, test_9.py:226-227). */
int emp_planner_create(EmpPlanner **out, int device, int max_batch, int kmax, int n_enr_max, int max_col);
void emp_planner_destroy(EmpPlanner *pl);

/* The whole S-L planning step with host buffers: H2D of scenes, both kernels,
 * D2H of plans (+ rows if not NULL); returns when the host buffers are valid. */
int emp_sl_plan_host(EmpPlanner *pl, const double *scenes, int batch, const EmpDpParams *dp,
                     const EmpQpParams *qp, double *plans, int32_t *rows);
/* Pipelined form of emp_sl_plan_host for a stream of batches: submit enqueues H2D -> DP -> QP -> D2H of one batch
 * on slot `slot`'s own CUDA stream and returns at once; wait blocks until that slot's results are in `plans`
 * (/`rows`).  Slots (0 .. EMP_PLAN_SLOTS-1) are independent, so the copies and kernels of consecutive batches
 * overlap.  All host buffers must be page-locked and stay valid until the slot has been waited for. */
int emp_sl_plan_submit(EmpPlanner *pl, int slot, const double *scenes, int batch, const EmpDpParams *dp,
                       const EmpQpParams *qp, double *plans, int32_t *rows);
int emp_sl_plan_wait(EmpPlanner *pl, int slot);
int emp_sl_dp_host(EmpPlanner *pl, const double *scenes, int batch, const EmpDpParams *dp,
                   double *plans, int32_t *rows);
int emp_qp_host(EmpPlanner *pl, const double *l_min, const double *l_max, const double *start_l_dl_ddl,
                int n, int batch, const EmpQpParams *p, double *out_l_dl_ddl, int32_t *status_iters);
int emp_corridor_host(EmpPlanner *pl, const double *dp_s, const double *dp_l, int n, const double *scenes,
                      int batch, const EmpQpParams *p, double *lmin_lmax, int32_t *status_iters);
/* Page-locked host buffers.  The *_host entry points DMA straight from/to buffers that are page-locked
 * (these, cudaHostRegister'ed, or torch pinned memory) and stage pageable ones through internal pinned buffers.
 * The synchronous *_host entry points run on the planner's own stream and device buffers; every submit slot has its
 * own stream and buffers, so synchronous calls and in-flight slots never share memory. */
void *emp_host_alloc(size_t bytes);
void emp_host_free(void *p);
/* Kernel-only time (ms, CUDA events on the planner's stream) of the last *_host call. */
double emp_planner_last_kernel_ms(const EmpPlanner *pl);

#ifdef __cplusplus
}
#endif
#endif /* EMPLANNER_H_ */
