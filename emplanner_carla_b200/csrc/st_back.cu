// K8 - S-T back half around the speed QP, one warp per frame (compiled with -fmad=false: bit-identical to numpy).
//   speed_planning_test.generate_convex_space :308-400  -> convex_space_kernel   (lane = S-T station)
//   speed_planning_test.increase_points       :514-562  -> increase_points_kernel (lanes stride the 401 samples)
//   speed_planning_test.path_speed_merge      :565-611  -> path_speed_merge_kernel
// Exceptions the reference raises are reported as per-frame status bits (EMP_ST_VALUE_ERROR: scipy interp1d out
// of range / empty table; EMP_ST_INDEX_ERROR: index past the 16-entry S-T arrays or a profile without its NaN
// terminator); the drop-in wrappers re-raise them.
#include "emp_common.cuh"
#include "st_back_kernels.h"

namespace emp {
namespace {

constexpr int NST = EMP_ST_COLS;      // 16
constexpr int NQP = EMP_ST_QP_POINTS; // 17
constexpr int NDENSE = EMP_ST_DENSE;  // 401

// scipy interp1d(kind='linear') on a sorted table: hi = searchsorted(x, xn) clipped to [1, n-1]; caller checks range
__device__ __forceinline__ double lerp_table(const double *x, const double *y, int n, double xn) {
    int lo = 0, hi = n;                                    // first index with x[idx] >= xn
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (x[mid] < xn) lo = mid + 1; else hi = mid;
    }
    hi = min(max(lo, 1), n - 1);
    lo = hi - 1;
    const double slope = (y[hi] - y[lo]) / (x[hi] - x[lo]);
    return slope * (xn - x[lo]) + y[lo];
}

// :360-370 - interval of dp_speed_t holding t (0 when none does)
__device__ __forceinline__ int time_index(const double *dp_t, double t) {
    if (dp_t[0] > t) return 0;
    for (int j = 0; j < NST - 1; ++j)
        if (dp_t[j] <= t && t < dp_t[j + 1]) return j;
    return 0;
}

__global__ void __launch_bounds__(128)
convex_space_kernel(const double *__restrict__ st_out, const double *__restrict__ scenes, int kmax,
                    const double *__restrict__ path, int n_path, double max_lat, int repair_full, int batch,
                    double *__restrict__ bounds, int32_t *__restrict__ status) {
    __shared__ double tab[4][2][NST + 1];                  // per warp: [0]+dp_t, [0]+dp_s
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * 4 + warp;
    if (f >= batch) return;
    const double *rec = st_out + (size_t)f * EMP_ST_OUT_STRIDE;
    const double *sc = scenes + (size_t)f * (2 + 4 * kmax);
    const double *idx2s = path + (size_t)f * 2 * n_path, *kappa = idx2s + n_path;
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double *tt = tab[warp][0], *ts = tab[warp][1];
    const double my_s = lane < NST ? rec[8 + lane] : 0.0;
    if (lane == 0) { tt[0] = 0.0; ts[0] = 0.0; }
    if (lane < NST) { tt[1 + lane] = rec[24 + lane]; ts[1 + lane] = my_s; }
    __syncwarp();
    const unsigned nanmask = __ballot_sync(EMP_FULL_MASK, lane < NST && isnan(my_s));
    const int n_valid = nanmask ? __ffs(nanmask) - 1 : NST;          // stations before the first NaN
    // :330-333; repair R3 (opt-in): a full profile ends at its last index like every profile with a terminator
    const int dp_end = nanmask ? n_valid - 1 : (repair_full ? NST - 1 : NST);
    // :323-327 - end of the valid part of path_index2s (zero padded)
    int first_pad = n_path;
    for (int k = 1 + lane; k < n_path; k += 32)
        if (idx2s[k] == 0.0 && idx2s[k - 1] != 0.0) { first_pad = k; break; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first_pad = min(first_pad, __shfl_xor_sync(EMP_FULL_MASK, first_pad, o));
    const int idx_end = first_pad < n_path ? first_pad - 1 : (n_path >= 2 ? n_path - 1 : n_path);
    int st = 0;
    double s_lb = -INF, s_ub = INF, sd_lb = -INF, sd_ub = INF;
    // :336-346 - lateral-acceleration speed limit from the path curvature at dp_speed_s[i]
    bool bad = false;
    if (lane < n_valid) {
        if (idx_end < 2 || my_s < idx2s[0] || my_s > idx2s[idx_end - 1]) {
            bad = true;
        } else {
            const double cur_kappa = lerp_table(idx2s, kappa, idx_end, my_s);
            sd_lb = 0.0;
            sd_ub = sqrt(max_lat / (fabs(cur_kappa) + 1e-10));
        }
    }
    if (__any_sync(EMP_FULL_MASK, bad)) st = EMP_ST_VALUE_ERROR;
    if (dp_end < 0) st |= EMP_ST_BAD_PROFILE;
    // :348-398 - every obstacle cuts s from above (yield) or from below (overtake) on its time window
    if (st == 0) {
        const int ntab = 1 + dp_end;                                   // [0] + dp[0:dp_end]
        const double *s_in = sc + 2, *s_out = s_in + kmax, *t_in = s_out + kmax, *t_out = t_in + kmax;
        for (int j = 0; j < kmax; ++j) {
            if (isnan(s_in[j])) continue;
            const double obs_t = (t_in[j] + t_out[j]) / 2, obs_s = (s_in[j] + s_out[j]) / 2;
            const double obs_speed = (s_out[j] - s_in[j]) / (t_out[j] - t_in[j]);
            if (ntab < 2 || obs_t < tt[0] || obs_t > tt[ntab - 1]) { st = EMP_ST_VALUE_ERROR; break; }
            const double dp_at = lerp_table(tt, ts, ntab, obs_t);
            const int lb_i = max(time_index(tt + 1, t_in[j]) - 2, 3);
            const int ub_i = min(time_index(tt + 1, t_out[j]) + 2, dp_end);
            if (ub_i >= NST && lb_i <= ub_i) { st = EMP_ST_INDEX_ERROR; break; }
            if (lane >= lb_i && lane <= ub_i) {
                const double line = s_in[j] + obs_speed * (tt[1 + lane] - t_in[j]);
                if (obs_s > dp_at) { if (line < s_ub) s_ub = line; }     // min(s_ub[m], line)
                else { if (line > s_lb) s_lb = line; }                   // max(s_lb[m], line)
            }
        }
    }
    if (lane < NST) {
        double *o = bounds + (size_t)f * 4 * NST;
        o[lane] = s_lb; o[NST + lane] = s_ub; o[2 * NST + lane] = sd_lb; o[3 * NST + lane] = sd_ub;
    }
    if (lane == 0) status[f] = st;
}

__global__ void __launch_bounds__(128)
increase_points_kernel(const double *__restrict__ qp, int batch, double *__restrict__ dense, int32_t *__restrict__ status) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * 4 + warp;
    if (f >= batch) return;
    const double *s0 = qp + (size_t)f * 4 * NQP, *v0 = s0 + NQP, *a0 = v0 + NQP, *t0 = a0 + NQP;
    const unsigned nanmask = __ballot_sync(EMP_FULL_MASK, lane < NQP && isnan(t0[lane]));
    if (!nanmask || __ffs(nanmask) - 1 < 1) {              // relative_time_init[17] (:538) / [-1]
        if (lane == 0) status[f] = nanmask ? EMP_ST_BAD_PROFILE : EMP_ST_INDEX_ERROR;
        return;
    }
    const int t_end = __ffs(nanmask) - 2;
    const double T = t0[t_end];
    const double dt = T / (NDENSE - 1);
    double *o = dense + (size_t)f * 4 * NDENSE;
    for (int i = lane; i < NDENSE; i += 32) {
        const double t = (double)(i - 1) * dt;             // sic: the first sample lies at -dt (:550)
        int tmp = 0;                                       // :551-554 - last interval start <= t among 0..t_end-2
        for (int j = 0; j < t_end - 1; ++j)
            if (t0[j] <= t) tmp = j;
        const double x = t - t0[tmp];
        const double x2 = x * x;
        o[i] = ((s0[tmp] + v0[tmp] * x) + ((1.0 / 3) * a0[tmp]) * x2) + ((1.0 / 6) * a0[tmp + 1]) * x2;
        o[NDENSE + i] = (v0[tmp] + (0.5 * a0[tmp]) * x) + (0.5 * a0[tmp + 1]) * x;
        o[2 * NDENSE + i] = a0[tmp] + ((a0[tmp + 1] - a0[tmp]) * x) / (t0[tmp + 1] - t0[tmp]);
        o[3 * NDENSE + i] = t;
    }
    if (lane == 0) status[f] = 0;
}

// numpy.interp (compiled arr_interp) for one query
__device__ __forceinline__ double np_interp(double x, const double *xp, const double *fp, int n) {
    if (isnan(x)) return x;
    if (x > xp[n - 1]) return fp[n - 1];
    if (x < xp[0]) return fp[0];
    int lo = 0, hi = n;                                    // first index with xp[idx] > x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (xp[mid] <= x) lo = mid + 1; else hi = mid;
    }
    const int j = lo - 1;
    if (j == n - 1 || xp[j] == x) return fp[j];
    const double slope = (fp[j + 1] - fp[j]) / (xp[j + 1] - xp[j]);
    double r = slope * (x - xp[j]) + fp[j];
    if (isnan(r)) {
        r = slope * (x - xp[j + 1]) + fp[j + 1];
        if (isnan(r) && fp[j] == fp[j + 1]) r = fp[j];
    }
    return r;
}

__global__ void __launch_bounds__(128)
path_speed_merge_kernel(const double *__restrict__ dense, const double *__restrict__ current_time,
                        const double *__restrict__ path, int n_path, int batch, double *__restrict__ traj,
                        int32_t *__restrict__ status) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * 4 + warp;
    if (f >= batch) return;
    const double *ps = path + (size_t)f * 5 * n_path;      // path_s, x, y, heading, kappa
    const double *d = dense + (size_t)f * 4 * NDENSE;
    int first_nan = n_path;                                // :581-583
    for (int k = lane; k < n_path; k += 32)
        if (isnan(ps[n_path + k])) { first_nan = k; break; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) first_nan = min(first_nan, __shfl_xor_sync(EMP_FULL_MASK, first_nan, o));
    const int n_tab = first_nan - 1;                       // `[:index]` drops the last valid point (:594)
    if (first_nan >= n_path || n_tab < 1) {
        if (lane == 0) status[f] = first_nan >= n_path ? EMP_ST_INDEX_ERROR : EMP_ST_VALUE_ERROR;
        return;
    }
    const double now = current_time[f];
    double *o = traj + (size_t)f * 7 * NDENSE;
    for (int i = lane; i < NDENSE; i += 32) {
        const bool last = i == NDENSE - 1;                 // :603-609 - the end point copies the LAST table entry
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k * NDENSE + i] = last ? ps[(k + 1) * n_path + n_path - 1] : np_interp(d[i], ps, ps + (k + 1) * n_path, n_tab);
        o[4 * NDENSE + i] = d[NDENSE + i];
        o[5 * NDENSE + i] = d[2 * NDENSE + i];
        o[6 * NDENSE + i] = d[3 * NDENSE + i] + now;
    }
    if (lane == 0) status[f] = 0;
}

}  // namespace

cudaError_t launch_convex_space(const double *st_out, const double *scenes, int kmax, const double *path, int n_path,
                                double max_lateral_accel, int repair_full, int batch, double *bounds, int32_t *status,
                                cudaStream_t st) {
    convex_space_kernel<<<(batch + 3) / 4, 128, 0, st>>>(st_out, scenes, kmax, path, n_path, max_lateral_accel, repair_full,
                                                         batch, bounds, status);
    return cudaGetLastError();
}

cudaError_t launch_increase_points(const double *qp, int batch, double *dense, int32_t *status, cudaStream_t st) {
    increase_points_kernel<<<(batch + 3) / 4, 128, 0, st>>>(qp, batch, dense, status);
    return cudaGetLastError();
}

cudaError_t launch_path_speed_merge(const double *dense, const double *current_time, const double *path, int n_path,
                                    int batch, double *traj, int32_t *status, cudaStream_t st) {
    path_speed_merge_kernel<<<(batch + 3) / 4, 128, 0, st>>>(dense, current_time, path, n_path, batch, traj, status);
    return cudaGetLastError();
}

}  // namespace emp
