// K10 - glue between the S-L and the S-T half of one EM cycle (compiled with -fmad=false).  In the reference this
// is list manipulation inside the driver scripts, not planner functions:
//   test_9.py:204-210   mid-point averaging of (dp_path_s[::2], qp_path_l) before frenet_2_x_y_theta_kappa
//   test_10.py:234-244  unpacking the Cartesian path into trajectory_x/y/heading/kappa_init + trajectory_index2s
// cycle_path_points_kernel turns a plan record into the polyline handed to smooth_reference_line (QP#3); the
// batched cycle has no per-scene reference line among its inputs (BASELINE.json configs[4]), so it plans along the
// s-axis, for which the projection of frenet_2_x_y_theta_kappa is the identity (x = s, y = l); like that function
// (path_planning.py:31-34) it puts the plan-start point in front of the path.
// cycle_path_tables_kernel lays the smoothed path out as the tables the S-T back half reads.
#include "emp_common.cuh"
#include "cycle_kernels.h"

namespace emp {
namespace {

constexpr int PLAN_HDR = 8;

__global__ void cycle_path_points_kernel(const double *__restrict__ plans, int plan_stride, int n_enr_max,
                                         const double *__restrict__ scenes, int scene_stride, int batch,
                                         int max_points, double *__restrict__ xy, int32_t *__restrict__ npts) {
    const int lane = threadIdx.x & 31;
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= batch) return;
    const double *rec = plans + (size_t)f * plan_stride;
    const int NE = n_enr_max, NQ = (NE + 1) / 2;
    const int n = (int)rec[1];
    const bool ok = (int)rec[3] == EMP_QP_OK || (int)rec[3] == EMP_QP_UNKNOWN_MAXITER || (int)rec[3] == EMP_QP_UNKNOWN_NUMERIC || (int)rec[3] == EMP_QP_UNKNOWN_DIVERGED;
    const double *es = rec + PLAN_HDR, *ql = rec + PLAN_HDR + 2 * NE;
    double *o = xy + (size_t)f * max_points * 2;
    // frenet_2_x_y_theta_kappa (path_planning.py:31-34) puts the plan-start point (plan_start_s, plan_start_l) in front of
    // the n + 1 mid-point path points, so smooth_reference_line sees n + 2 points
    const double *sc = scenes + (size_t)f * scene_stride;
    const int m = (ok && n >= 2 && n + 2 <= max_points && n <= NQ) ? n + 2 : 0;
    for (int j = lane; j < m; j += 32) {
        const int i = j - 1;
        double s, l;
        if (j == 0) { s = sc[0]; l = sc[1]; }
        else if (i == 0) { s = es[0]; l = ql[0]; }
        else if (i == n) { s = es[2 * (n - 1)]; l = ql[n - 1]; }
        else { s = (es[2 * i] + es[2 * (i - 1)]) / 2; l = (ql[i] + ql[i - 1]) / 2; }
        o[2 * j] = s;
        o[2 * j + 1] = l;
    }
    if (lane == 0) npts[f] = m;
}

__global__ void cycle_path_tables_kernel(const double *__restrict__ cart, const int32_t *__restrict__ npts,
                                         int max_points, int batch, int n_path, double *__restrict__ path2,
                                         double *__restrict__ path5) {
    // one thread per frame: trajectory_index2s (planning_utils.py:758-780) is a running sum
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= batch) return;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    const double *c = cart + (size_t)f * max_points * 4;
    double *p2 = path2 + (size_t)f * 2 * n_path, *p5 = path5 + (size_t)f * 5 * n_path;
    const int n = min(npts[f], n_path - 1);
    double s = 0.0;
    for (int i = 0; i < n_path; ++i) {
        const bool have = i < n;
        if (have && i > 0) {
            const double dx = c[4 * i] - c[4 * (i - 1)], dy = c[4 * i + 1] - c[4 * (i - 1) + 1];
            s += sqrt(dx * dx + dy * dy);
        }
        p2[i] = have ? s : 0.0;                            // path_index2s, zero padded
        p2[n_path + i] = have ? c[4 * i + 3] : 0.0;        // trajectory_kappa_init
        p5[i] = have ? s : 0.0;
        p5[n_path + i] = have ? c[4 * i] : NaN;            // trajectory_x_init, NaN padded
        p5[2 * n_path + i] = have ? c[4 * i + 1] : NaN;
        p5[3 * n_path + i] = have ? c[4 * i + 2] : NaN;
        p5[4 * n_path + i] = have ? c[4 * i + 3] : NaN;
    }
}

}  // namespace

cudaError_t launch_cycle_path_points(const double *plans, int plan_stride, int n_enr_max, const double *scenes,
                                     int scene_stride, int batch, int max_points, double *xy, int32_t *npts,
                                     cudaStream_t st) {
    cycle_path_points_kernel<<<(batch + 3) / 4, 128, 0, st>>>(plans, plan_stride, n_enr_max, scenes, scene_stride, batch,
                                                              max_points, xy, npts);
    return cudaGetLastError();
}

cudaError_t launch_cycle_path_tables(const double *cart, const int32_t *npts, int max_points, int batch, int n_path,
                                     double *path2, double *path5, cudaStream_t st) {
    cycle_path_tables_kernel<<<(batch + 127) / 128, 128, 0, st>>>(cart, npts, max_points, batch, n_path, path2, path5);
    return cudaGetLastError();
}

}  // namespace emp
