// Device functions of the S-T speed DP shared by csrc/st_dp.cu and csrc/st_misc.cu (see st_dp.cu for citations).
#pragma once
#include "emp_common.cuh"
#include "st_kernels.h"

namespace emp {
namespace {

constexpr int NROW = 40, NCOL = 16, NSAMP = 5;
constexpr double CULL = 1.6;          // > 1.5: beyond it CalcCollisionCost returns exactly 0

__device__ __forceinline__ double s_list(int idx) {     // :114
    if (idx < 10) return 0.5 * idx;
    if (idx < 20) return 5.5 + 1.0 * (idx - 10);
    if (idx < 30) return 16.0 + 1.5 * (idx - 20);
    return 32.0 + 2.5 * (idx - 30);
}
__device__ __forceinline__ double node_s(int row) { return s_list(NROW - row - 1); }   // CalcSTCoordinate :303
__device__ __forceinline__ double node_t(int col) { return 0.5 + 0.5 * col; }           // :116

struct Obs { double s_in, s_out, t_in, t_out; };

// CalcObsCost (:234-271) for one edge.
__device__ __forceinline__ double obs_cost(double s_st, double t_st, double s_en, double t_en, const Obs *obs,
                                           int n_obs, double w_obs) {
    double cost = 0.0;
    const double dt = (t_en - t_st) / (NSAMP - 1);
    const double k = (s_en - s_st) / (t_en - t_st);
    // bounding box of the 5 samples (i = 0 and i = 4)
    const double ta = t_st + (-1.0) * dt, tb = t_st + 3.0 * dt;
    const double sa = s_st + (k * -1.0) * dt, sb = s_st + (k * 3.0) * dt;
    const double t_lo = fmin(ta, tb), t_hi = fmax(ta, tb), s_lo = fmin(sa, sb), s_hi = fmax(sa, sb);
    unsigned live = 0;
    for (int j = 0; j < n_obs; ++j) {
        const Obs o = obs[j];
        if (isnan(o.s_in)) continue;                                   // :255
        const bool far = fmin(o.t_in, o.t_out) - t_hi >= CULL || t_lo - fmax(o.t_in, o.t_out) >= CULL ||
                         fmin(o.s_in, o.s_out) - s_hi >= CULL || s_lo - fmax(o.s_in, o.s_out) >= CULL;
        if (!far) live |= 1u << j;
    }
    if (!live) return 0.0;
#pragma unroll
    for (int i = 0; i < NSAMP; ++i) {
        const double t = t_st + (double)(i - 1) * dt;
        const double s = s_st + (k * (double)(i - 1)) * dt;
        for (unsigned m = live; m; m &= m - 1) {
            const Obs o = obs[__ffs(m) - 1];
            const double v1s = o.s_in - s, v1t = o.t_in - t;
            const double v2s = o.s_out - s, v2t = o.t_out - t;
            const double v3s = v2s - v1s, v3t = v2t - v1t;
            const double d13 = v1s * v3s + v1t * v3t;
            const double d23 = v2s * v3s + v2t * v3t;
            double dmin;
            if ((d13 > 0 && d23 > 0) || (d13 < 0 && d23 < 0)) {
                const double dis1 = sqrt(v1s * v1s + v1t * v1t);
                const double dis2 = sqrt(v2s * v2s + v2t * v2t);
                dmin = fmin(dis1, dis2);
            } else {
                dmin = fabs(v1s * v3t - v1t * v3s) / sqrt(v3s * v3s + v3t * v3t);
            }
            const double ad = fabs(dmin);
            double c = 0.0;
            if (ad < 0.5) c = w_obs;
            else if (0.5 < ad && ad < 1.5) c = pow(w_obs, (0.5 - dmin) + 1);
            cost = cost + c;
        }
    }
    return cost;
}

// CalcDpCost (:191-231)
__device__ __forceinline__ double edge_cost(double s_st, double t_st, double sd_st, double s_en, double t_en,
                                            const Obs *obs, int n_obs, const EmpStParams &P) {
    const double cur = (s_en - s_st) / (t_en - t_st);
    const double acc = (cur - sd_st) / (t_en - t_st);
    const double dv = cur - P.reference_speed;
    const double c_ref = P.w_ref_speed * (dv * dv);
    double c_acc;
    if (4 > acc && acc > -6) c_acc = P.w_accel * (acc * acc);
    else c_acc = (100000 * P.w_accel) * (acc * acc);
    const double c_obs = obs_cost(s_st, t_st, s_en, t_en, obs, n_obs, P.w_obs);
    return (c_obs + c_acc) + c_ref;
}


}  // namespace
}  // namespace emp
