// K6 - stand-alone pieces of the S-L DP that the reference exposes as module-level functions.
//
//   emp_quintic_dev    planning_utils.cal_quintic_coefficient (:671-703): closed form instead of a 6x6 inverse
//   emp_edge_cost_dev  path_planning.cal_start_cost (:435-514) / cal_neighbor_cost (:517-585)
//   emp_obs_cost_dev   path_planning.cal_obs_cost (:588-609)
//   emp_enrich_dev     path_planning.enrich_DP_s_l (:378-432)
// Inside the batched planner these are fused into K1 (csrc/sl_dp.cu); here they are one item per thread so the
// drop-in functions of the same names need no CPU arithmetic.  Compiled with -fmad=false; expression order
// follows oracle/sl_dp.py.
#include "emp_common.cuh"
#include "misc_kernels.h"

namespace emp {
namespace {

constexpr int NS = EMP_NSAMPLE;

struct Quintic { double a[6]; };     // l(u) = sum a_k u^k, u = s - s0

__device__ __forceinline__ Quintic quintic_local(double l0, double dl0, double ddl0, double l1, double dl1,
                                                 double ddl1, double T) {
    Quintic q;
    const double iT = 1.0 / T;
    const double iT3 = (iT * iT) * iT, iT4 = iT3 * iT, iT5 = iT4 * iT;
    const double D = l1 - l0, T2 = T * T;
    q.a[0] = l0;
    q.a[1] = dl0;
    q.a[2] = ddl0 / 2.0;
    q.a[3] = ((20.0 * D - (8.0 * dl1 + 12.0 * dl0) * T) - (3.0 * ddl0 - ddl1) * T2) * (0.5 * iT3);
    q.a[4] = ((-30.0 * D + (14.0 * dl1 + 16.0 * dl0) * T) + (3.0 * ddl0 - 2.0 * ddl1) * T2) * (0.5 * iT4);
    q.a[5] = ((12.0 * D - (6.0 * (dl1 + dl0)) * T) - (ddl0 - ddl1) * T2) * (0.5 * iT5);
    return q;
}

__global__ void quintic_kernel(const double *__restrict__ in, double *__restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *v = in + (size_t)i * 8;               // start_l, dl, ddl, end_l, dl, ddl, start_s, end_s
    const double s0 = v[6];
    const Quintic q = quintic_local(v[0], v[1], v[2], v[3], v[4], v[5], v[7] - s0);
    // absolute-s power basis: c_j = sum_{k>=j} a_k C(k,j) (-s0)^(k-j)
    const double m = -s0, m2 = m * m, m3 = m2 * m, m4 = m3 * m, m5 = m4 * m;
    double *c = out + (size_t)i * 6;
    c[0] = q.a[0] + q.a[1] * m + q.a[2] * m2 + q.a[3] * m3 + q.a[4] * m4 + q.a[5] * m5;
    c[1] = q.a[1] + 2.0 * q.a[2] * m + 3.0 * q.a[3] * m2 + 4.0 * q.a[4] * m3 + 5.0 * q.a[5] * m4;
    c[2] = q.a[2] + 3.0 * q.a[3] * m + 6.0 * q.a[4] * m2 + 10.0 * q.a[5] * m3;
    c[3] = q.a[3] + 4.0 * q.a[4] * m + 10.0 * q.a[5] * m2;
    c[4] = q.a[4] + 5.0 * q.a[5] * m;
    c[5] = q.a[5];
}

__device__ __forceinline__ double obs_cost_seq(const double *d2, int n, int stride, double w_col, double danger2,
                                               double safe2) {
    double cost = 0.0;
    for (int m = 0; m < n; ++m) {
        const double v = d2[(size_t)m * stride];
        if (v <= danger2) {
            cost = cost + w_col;
            break;
        } else if (danger2 < v && v < safe2) {
            cost = cost + 5000.0 / v;
        }
    }
    return cost;
}

__global__ void obs_cost_kernel(const double *__restrict__ d2, int n, int batch, double w_col, double danger2,
                                double safe2, double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    out[i] = obs_cost_seq(d2 + (size_t)i * n, n, 1, w_col, danger2, safe2);
}

// one edge per thread: quintic (s0,l0,dl0,ddl0) -> (s1,l1,0,0), 10 samples spaced sample_s/10 from s0
__global__ void edge_cost_kernel(const double *__restrict__ edges, const double *__restrict__ obs_s,
                                 const double *__restrict__ obs_l, int n_obs, int n_edges, EdgeParams P,
                                 double *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_edges) return;
    const double *e = edges + (size_t)i * 6;
    const double s0 = e[0], l0 = e[1], dl0 = e[2], ddl0 = e[3], s1 = e[4], l1 = e[5];
    const Quintic q = quintic_local(l0, dl0, ddl0, l1, 0.0, 0.0, s1 - s0);
    const double a1 = q.a[1], a2 = q.a[2], a3 = q.a[3], a4 = q.a[4], a5 = q.a[5];
    const double c4 = a4 - (5.0 * a5) * s0;
    const double c3 = (a3 - (4.0 * a4) * s0) + (10.0 * a5) * (s0 * s0);
    const double alpha = 6.0 * c3, beta = 24.0 * c4 + 120.0 * a5;      // dddl = 6c3 + 24c4 s + 60c5 (s*2)  (Q1)
    double sdl = 0.0, sddl = 0.0, sj = 0.0, sl2 = 0.0, sm[NS], lm[NS];
#pragma unroll
    for (int m = 0; m < NS; ++m) {
        const double u = ((double)m * P.sample_s) / 10.0;
        sm[m] = s0 + u;
        const double l = l0 + u * (a1 + u * (a2 + u * (a3 + u * (a4 + u * a5))));
        const double dl = a1 + u * (2.0 * a2 + u * (3.0 * a3 + u * (4.0 * a4 + u * (5.0 * a5))));
        const double ddl = 2.0 * a2 + u * (6.0 * a3 + u * (12.0 * a4 + u * (20.0 * a5)));
        const double jm = alpha + beta * sm[m];
        sdl = sdl + dl * dl;
        sddl = sddl + ddl * ddl;
        sj = sj + jm * jm;
        sl2 = sl2 + l * l;
        lm[m] = l;
    }
    const double smooth = (P.w0 * sdl + P.w1 * sddl) + P.w2 * sj;
    double coll = 0.0;
    for (int o = 0; o < n_obs; ++o) {
        double d2[NS];
#pragma unroll
        for (int m = 0; m < NS; ++m) {
            const double dlon = obs_s[o] - sm[m], dlat = obs_l[o] - lm[m];
            d2[m] = dlon * dlon + dlat * dlat;
        }
        coll = coll + obs_cost_seq(d2, NS, 1, P.w_col, P.danger2, P.safe2);
    }
    out[i] = (smooth + coll) + P.w_ref * sl2;
}

// enrich_DP_s_l for one path per thread: nodes (s,l)[C] + plan start -> densified (s,l), count
__global__ void enrich_kernel(const double *__restrict__ node_s, const double *__restrict__ node_l,
                              const double *__restrict__ start, int C, int batch, double res, int n_max,
                              double *__restrict__ out_s, double *__restrict__ out_l, int32_t *__restrict__ n_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const double *S = node_s + (size_t)b * C, *L = node_l + (size_t)b * C, *st = start + (size_t)b * 4;
    double *os = out_s + (size_t)b * n_max, *ol = out_l + (size_t)b * n_max;
    int cnt_out = 0;
    for (int g = 0; g < C; ++g) {
        const double ss = g == 0 ? st[0] : S[g - 1], es = S[g];
        const double sl_ = g == 0 ? st[1] : L[g - 1], el = L[g];
        const double sdl = g == 0 ? st[2] : 0.0, sddl = g == 0 ? st[3] : 0.0;
        const int cnt = (int)(es - ss);                             // :405,:423
        const int len = cnt > 0 ? (int)ceil((double)cnt / res) : 0;
        const Quintic q = quintic_local(sl_, sdl, sddl, el, 0.0, 0.0, es - ss);
        for (int k = 0; k < len; ++k) {
            if (cnt_out >= n_max) break;
            const double u = (double)k * res;
            os[cnt_out] = ss + u;
            ol[cnt_out] = sl_ + u * (sdl + u * (q.a[2] + u * (q.a[3] + u * (q.a[4] + u * q.a[5]))));
            ++cnt_out;
        }
    }
    if (cnt_out < n_max && C > 0) {
        os[cnt_out] = S[C - 1];
        ol[cnt_out] = L[C - 1];
        ++cnt_out;
    }
    n_out[b] = cnt_out;
}

}  // namespace

cudaError_t launch_quintic(const double *in, double *out, int n, cudaStream_t st) {
    quintic_kernel<<<(n + 127) / 128, 128, 0, st>>>(in, out, n);
    return cudaGetLastError();
}
cudaError_t launch_obs_cost(const double *d2, int n, int batch, double w, double danger2, double safe2, double *out,
                            cudaStream_t st) {
    obs_cost_kernel<<<(batch + 127) / 128, 128, 0, st>>>(d2, n, batch, w, danger2, safe2, out);
    return cudaGetLastError();
}
cudaError_t launch_edge_cost(const double *edges, const double *obs_s, const double *obs_l, int n_obs, int n_edges,
                             const EdgeParams &P, double *out, cudaStream_t st) {
    edge_cost_kernel<<<(n_edges + 127) / 128, 128, 0, st>>>(edges, obs_s, obs_l, n_obs, n_edges, P, out);
    return cudaGetLastError();
}
cudaError_t launch_enrich(const double *node_s, const double *node_l, const double *start, int C, int batch, double res,
                          int n_max, double *out_s, double *out_l, int32_t *n_out, cudaStream_t st) {
    enrich_kernel<<<(batch + 63) / 64, 64, 0, st>>>(node_s, node_l, start, C, batch, res, n_max, out_s, out_l, n_out);
    return cudaGetLastError();
}

}  // namespace emp
