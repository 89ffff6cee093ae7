// K7 - small S-T / speed-draft helpers, one item per thread (they are microseconds each in the reference; on the
// GPU they exist so the drop-in functions of the same names need no CPU arithmetic).
//   speed_planning_test.calc_speed_planning_start_condition :23-35, generate_st_graph :38-98,
//   CalcDpCost :191-231, CalcObsCost :234-271, CalcCollisionCost :274-284
//   planning_utils.trajectory_index2s :758-780, cal_dy_obs_deri :783-808, CalcProjPoint :736-755,
//   Frenet2Cartesian :706-733
#include "st_device.cuh"
#include "st_misc_kernels.h"

namespace emp {
namespace {

__global__ void st_graph_kernel(const double *__restrict__ in, int n, int batch, double *__restrict__ out) {
    // in: [batch, 4, n] = (s, l, s_dot, l_dot); out: [batch, 4, n] = (s_in, s_out, t_in, t_out), NaN = absent.
    // one thread per frame: the reference stops at the first NaN s (:51-52)
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= batch) return;
    const double *s = in + (size_t)f * 4 * n, *l = s + n, *sd = l + n, *ld = sd + n;
    double *o = out + (size_t)f * 4 * n;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = 0; i < 4 * n; ++i) o[i] = NaN;
    for (int i = 0; i < n; ++i) {
        if (isnan(s[i])) break;
        if (fabs(ld[i]) < 0.3) continue;                               // :53-66 (both branches `continue`)
        const double t_zero = -l[i] / ld[i];
        const double b1 = 2 / ld[i] + t_zero, b2 = -2 / ld[i] + t_zero;
        const double t_max = b1 > b2 ? b1 : b2, t_min = b1 > b2 ? b2 : b1;
        if (t_max < 1 || t_min > 8) continue;
        if (t_min < 0 && t_max > 0) {
            o[i] = s[i];
            o[n + i] = s[i] + sd[i] * t_max;
            o[2 * n + i] = 0;
            o[3 * n + i] = t_max;
        } else {
            o[i] = s[i] + sd[i] * t_min;
            o[n + i] = s[i] + sd[i] * t_max;
            o[2 * n + i] = t_min;
            o[3 * n + i] = t_max;
        }
    }
}

__global__ void start_condition_kernel(const double *__restrict__ in, int batch, double *__restrict__ out) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= batch) return;
    const double *v = in + (size_t)f * 5;                              // vx, vy, ax, ay, heading
    const double c = cos(v[4]), s = sin(v[4]);
    out[2 * f] = c * v[0] + s * v[1];
    out[2 * f + 1] = c * v[2] + s * v[3];
}

// op 0: CalcDpCost(row_start, col_start, row_end, col_end)   item = (rs, cs, re, ce) as doubles
// op 1: CalcObsCost(s_start, t_start, s_end, t_end)
// op 2: CalcCollisionCost(min_dis)
__global__ void st_cost_kernel(int op, const double *__restrict__ items, int n_items, const double *__restrict__ obs4,
                               int n_obs, const double *__restrict__ sdot_tab, double v0, EmpStParams P,
                               double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    Obs obs[EMP_ST_KMAX];
    for (int j = 0; j < n_obs; ++j) {
        obs[j].s_in = obs4[j]; obs[j].s_out = obs4[n_obs + j]; obs[j].t_in = obs4[2 * n_obs + j]; obs[j].t_out = obs4[3 * n_obs + j];
    }
    if (op == 2) {
        const double d = items[i], ad = fabs(d);
        out[i] = ad < 0.5 ? P.w_obs : ((0.5 < ad && ad < 1.5) ? pow(P.w_obs, (0.5 - d) + 1) : 0.0);
    } else if (op == 1) {
        const double *e = items + (size_t)i * 4;
        out[i] = obs_cost(e[0], e[1], e[2], e[3], obs, n_obs, P.w_obs);
    } else {
        const double *e = items + (size_t)i * 4;
        const int rs = (int)e[0], cs = (int)e[1], re = (int)e[2], ce = (int)e[3];
        double s_st = 0.0, t_st = 0.0, sd = v0;
        if (rs != 0) { s_st = node_s(rs); t_st = node_t(cs); sd = sdot_tab[rs * NCOL + cs]; }     // :208-216
        out[i] = edge_cost(s_st, t_st, sd, node_s(re), node_t(ce), obs, n_obs, P);
    }
}

__global__ void dy_obs_deri_kernel(const double *__restrict__ in, int n, int batch, double *__restrict__ out) {
    // in [batch, 5, n] = (l, vx, vy, heading, kappa); out [batch, 3, 128] = (s_dot, l_dot, dl/ds), NaN padded
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= batch) return;
    const double *l = in + (size_t)f * 5 * n, *vx = l + n, *vy = vx + n, *hd = vy + n, *kp = hd + n;
    double *o = out + (size_t)f * 3 * 128;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    for (int i = 0; i < 3 * 128; ++i) o[i] = NaN;
    for (int i = 0; i < n && i < 128; ++i) {
        if (isnan(l[i])) break;
        const double sn = sin(hd[i]), cs = cos(hd[i]);
        const double l_dot = vx[i] * (-sn) + vy[i] * cs;
        const double s_dot = (vx[i] * cs + vy[i] * sn) / (1 - kp[i] * l[i]);
        o[i] = s_dot;
        o[128 + i] = l_dot;
        o[256 + i] = fabs(s_dot) < 1e-6 ? 0.0 : l_dot / s_dot;
    }
}

__global__ void index2s_kernel(const double *__restrict__ xy, int n, int batch, double *__restrict__ out) {
    // xy [batch, 2, n]; out [batch, n]: cumulative polyline length, zeros from the first NaN x on (:767-772)
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= batch) return;
    const double *x = xy + (size_t)f * 2 * n, *y = x + n;
    double *o = out + (size_t)f * n;
    for (int i = 0; i < n; ++i) o[i] = 0.0;
    double s = 0.0;
    for (int i = 1; i < n; ++i) {
        if (isnan(x[i])) break;
        const double dx = x[i] - x[i - 1], dy = y[i] - y[i - 1];
        s += sqrt(dx * dx + dy * dy);
        o[i] = s;
    }
}

__global__ void frenet2cart_kernel(const double *__restrict__ sl, int n_pts, const double *__restrict__ path,
                                   const double *__restrict__ s_map, int n_nodes, int proj_only,
                                   double *__restrict__ out) {
    // sl [4, n_pts] = (s, l, dl, ddl); path [4, n_nodes] = (x, y, heading, kappa); out [4, n_pts]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pts) return;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    // Frenet2Cartesian stops at the first NaN s (:718-719)
    bool dead = false;
    for (int j = 0; j <= i; ++j) dead = dead || isnan(sl[j]);
    if (dead) {
        for (int c = 0; c < 4; ++c) out[(size_t)c * n_pts + i] = NaN;
        return;
    }
    const double s = sl[i], l = sl[n_pts + i], dl = sl[2 * n_pts + i], ddl = sl[3 * n_pts + i];
    int m = 1;                                                          // CalcProjPoint :742-744
    while (m < n_nodes - 1 && s_map[m] < s) ++m;
    const double hd = path[2 * n_nodes + m], kp = path[3 * n_nodes + m];
    const double ds = s - s_map[m];
    const double px = path[m] + ds * cos(hd), py = path[n_nodes + m] + ds * sin(hd);
    const double ph = hd + ds * kp;
    if (proj_only) {
        out[i] = px; out[n_pts + i] = py; out[2 * n_pts + i] = ph; out[3 * n_pts + i] = kp;
        return;
    }
    const double heading = ph + atan(dl / (1 - kp * l));
    const double d = heading - ph;
    out[i] = px + l * (-sin(ph));
    out[n_pts + i] = py + l * cos(ph);
    out[2 * n_pts + i] = heading;
    out[3 * n_pts + i] = ((ddl + kp * dl * tan(d)) * (cos(d) * cos(d)) / (1 - kp * l) + kp) * cos(d) / (1 - kp * l);
}

}  // namespace

#define EMP_GRID(n) ((n) + 127) / 128, 128, 0, st
cudaError_t launch_st_graph(const double *in, int n, int batch, double *out, cudaStream_t st) {
    st_graph_kernel<<<EMP_GRID(batch)>>>(in, n, batch, out);
    return cudaGetLastError();
}
cudaError_t launch_start_condition(const double *in, int batch, double *out, cudaStream_t st) {
    start_condition_kernel<<<EMP_GRID(batch)>>>(in, batch, out);
    return cudaGetLastError();
}
cudaError_t launch_st_cost(int op, const double *items, int n_items, const double *obs4, int n_obs,
                           const double *sdot_tab, double v0, const EmpStParams &P, double *out, cudaStream_t st) {
    st_cost_kernel<<<EMP_GRID(n_items)>>>(op, items, n_items, obs4, n_obs, sdot_tab, v0, P, out);
    return cudaGetLastError();
}
cudaError_t launch_dy_obs_deri(const double *in, int n, int batch, double *out, cudaStream_t st) {
    dy_obs_deri_kernel<<<EMP_GRID(batch)>>>(in, n, batch, out);
    return cudaGetLastError();
}
cudaError_t launch_index2s(const double *xy, int n, int batch, double *out, cudaStream_t st) {
    index2s_kernel<<<EMP_GRID(batch)>>>(xy, n, batch, out);
    return cudaGetLastError();
}
cudaError_t launch_frenet2cart(const double *sl, int n_pts, const double *path, const double *s_map, int n_nodes,
                               int proj_only, double *out, cudaStream_t st) {
    frenet2cart_kernel<<<EMP_GRID(n_pts)>>>(sl, n_pts, path, s_map, n_nodes, proj_only, out);
    return cudaGetLastError();
}

}  // namespace emp
