// K9 - speed QP (S-T smoothing), one warp per frame.
//
// Replaces reference planner/speed_planning_test.py speed_QP (:403-511, cvxopt.solvers.qp at :502) with the
// documented repairs of oracle/st_back.py (R1: the equality constraint is Aeq' x = beq; R2: lb/ub are two arrays
// whose finite entries become inequality rows).  Variables x = (s, s_dot, s_dot2) per station, q <= 16 stations:
//
//   min  w_a |s_dot2|^2 + w_j sum (s_dot2[i+1] - s_dot2[i])^2 + w_v |s_dot - v_ref|^2
//   s.t. s[i+1] = s[i] + dt s_dot[i] + dt^2/3 s_dot2[i] + dt^2/6 s_dot2[i+1]
//        s_dot[i+1] = s_dot[i] + dt/2 (s_dot2[i] + s_dot2[i+1])
//        s[i] <= s[i+1];  lb <= x <= ub (finite entries);  x[0:3] pinned by lb = ub
//
// The iteration is cvxopt coneqp's (oracle/coneqp.py: initial point, Mehrotra predictor/corrector, step 0.99,
// stopping test), with s and z carried unscaled.  The linear algebra follows cvxopt's `chol2` route - Cholesky of
// S = P + G' diag(z/s) G, then of K = A S^-1 A' - but uses the structure: every row of G touches one variable
// or (s[i], s[i+1]), the jerk term couples (s_dot2[i], s_dot2[i+1]), so S splits into three tridiagonal chains
// (one per component) whose bidiagonal factors cost q steps; L^-1 A' (3q x 2(q-1)) and the dense K and its factor
// live in shared memory (21.6 KB per frame), column/row per lane.
#include "emp_common.cuh"
#include "st_back_kernels.h"

namespace emp {
namespace {

#define EMP_UNROLL _Pragma("unroll")
constexpr int NQP = EMP_ST_QP_POINTS;
constexpr int MAXQ = SQP_MAXQ, NV = SQP_NV, NE = SQP_NE, YS = SQP_YS, KS = SQP_KS;

// reciprocal / reciprocal square root: hardware seed + two Newton steps (full double precision for the well-scaled
// arguments met here; the QP's parity bar is 1e-4 m on the solution, the iterates agree to ~1e-9)
__device__ __forceinline__ double frcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double frsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-hx * y, y, 0.5);
    return fma(y, e, y);
}

struct WarpSmem {
    double Y[NV * YS];        // L^-1 A'
    double K[NE * KS];        // A S^-1 A', overwritten by its Cholesky factor (strict lower part)
    double kdi[32];           // inverse diagonal of that factor
    double x[NV];             // iterate
    double v[NV];             // x-space right-hand side -> solution
    double y[32];             // y-space scratch
    double ldi[3][MAXQ];      // chain factors: inverse diagonal
    double le[3][MAXQ];       // chain factors: sub-diagonal
    double dg[3][MAXQ];       // chain diagonals of S
    double wm[MAXQ];          // per no-reversing row: weight z/s, or a row-space value for G' products
    double ac[12];            // A_sub columns (:444-449)
};

#ifndef EMP_SQP_MINBLOCKS
#define EMP_SQP_MINBLOCKS (SQP_WARPS == 1 ? 8 : 2)
#endif
__global__ void __launch_bounds__(SQP_WARPS * 32, EMP_SQP_MINBLOCKS)
speed_qp_kernel(const double *__restrict__ start, const double *__restrict__ st_out, const double *__restrict__ bounds,
                double *__restrict__ out, int32_t *__restrict__ status_iters, const SpeedQpLaunch p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = blockIdx.x * SQP_WARPS + warp;
    if (f >= p.batch) return;
    WarpSmem &W = reinterpret_cast<WarpSmem *>(smem_raw)[warp];
    const EmpSpeedQpParams &Q = p.prm;
    const double NaN = __longlong_as_double(0x7ff8000000000000LL);
    const double *rec = st_out + (size_t)f * EMP_ST_OUT_STRIDE;
    double *o = out + (size_t)f * 4 * NQP;

    // :415-442 - dp_speed_end = (first NaN of dp_speed_s) - 1; without a terminator dp_speed_s[16] raises IndexError
    const unsigned nanmask = __ballot_sync(EMP_FULL_MASK, lane < EMP_ST_COLS && isnan(rec[8 + lane]));
    // (repair R3, opt-in: a full profile ends at its last column, index 15, like every profile with a terminator)
    const bool full_ok = !nanmask && Q.repair_full_profile == 1;
    const int end = nanmask ? __ffs(nanmask) - 2 : (full_ok ? EMP_ST_COLS - 1 : EMP_ST_COLS);
    if ((!nanmask && !full_ok) || end < 1) {
        for (int i = lane; i < 4 * NQP; i += 32) o[i] = NaN;
        if (lane == 0) {
            status_iters[2 * f] = (nanmask || full_ok) ? EMP_ST_BAD_PROFILE : EMP_ST_INDEX_ERROR;
            status_iters[2 * f + 1] = 0;
        }
        return;
    }
    const int q = end + 1, nv = 3 * q, ne = 2 * (q - 1), nm = q - 1;
    const double dt = rec[24 + end] / (double)end;
    if (lane == 0) {
        const double c[12] = {1.0, dt, (1.0 / 3) * (dt * dt), -1.0, 0.0, (1.0 / 6) * (dt * dt),
                              0.0, 1.0, (1.0 / 2) * dt, 0.0, -1.0, dt / 2};
        EMP_UNROLL
        for (int k = 0; k < 12; ++k) W.ac[k] = c[k];
    }
    // per-lane variables v = lane, lane + 32 and their rows: [2t] lower bound, [2t+1] upper bound, [4] no-reversing row
    bool vo[2], on[5];
    int vk[2], va[2];
    double hr[5], pd[2], qv[2];
    const double *bnd = bounds + (size_t)f * 4 * EMP_ST_COLS;
    EMP_UNROLL
    for (int t = 0; t < 2; ++t) {
        const int v = lane + 32 * t;
        vo[t] = v < nv;
        vk[t] = v / 3;
        va[t] = v - 3 * vk[t];
        double lb = 0.0, ub = 0.0;
        if (vo[t]) {
            if (vk[t] == 0) {
                lb = va[t] == 0 ? 0.0 : start[2 * f + va[t] - 1];           // :473-478
                ub = lb;
            } else if (va[t] == 2) {
                lb = -6.0; ub = 4.0;                                         // :467,:470
            } else {
                lb = bnd[(2 * va[t]) * EMP_ST_COLS + vk[t] - 1];             // s_lb / s_dot_lb
                ub = bnd[(2 * va[t] + 1) * EMP_ST_COLS + vk[t] - 1];         // s_ub / s_dot_ub
            }
        }
        on[2 * t] = vo[t] && isfinite(lb);
        on[2 * t + 1] = vo[t] && isfinite(ub);
        hr[2 * t] = on[2 * t] ? -lb : 0.0;
        hr[2 * t + 1] = on[2 * t + 1] ? ub : 0.0;
        const int cnt = (vk[t] > 0 ? 1 : 0) + (vk[t] < q - 1 ? 1 : 0);
        pd[t] = !vo[t] ? 0.0 : va[t] == 1 ? 2 * Q.w_v_ref : va[t] == 2 ? 2 * (Q.w_s_dot2 + Q.w_jerk * cnt) : 0.0;
        qv[t] = vo[t] && va[t] == 1 ? -2 * Q.w_v_ref * Q.reference_speed : 0.0;
    }
    on[4] = lane < nm;
    hr[4] = 0.0;
    int m_rows = 0;
    EMP_UNROLL
    for (int r = 0; r < 5; ++r) m_rows += __popc(__ballot_sync(EMP_FULL_MASK, on[r]));
    const double wj2 = 2 * Q.w_jerk;
    __syncwarp();

    // ---- products with the constraint matrices -------------------------------------------------
    auto Gmul = [&](const double *xv, double g[5]) {
        EMP_UNROLL
        for (int t = 0; t < 2; ++t) {
            const double xi = vo[t] ? xv[lane + 32 * t] : 0.0;
            g[2 * t] = -xi;
            g[2 * t + 1] = xi;
        }
        g[4] = on[4] ? xv[3 * lane] - xv[3 * lane + 3] : 0.0;
    };
    // (G' r)_v for the lane's variable t; the no-reversing rows' values must be in W.wm
    auto GTvar = [&](const double r[5], int t) -> double {
        double acc = (on[2 * t + 1] ? r[2 * t + 1] : 0.0) - (on[2 * t] ? r[2 * t] : 0.0);
        if (va[t] == 0) acc += (vk[t] < nm ? W.wm[vk[t]] : 0.0) - (vk[t] >= 1 ? W.wm[vk[t] - 1] : 0.0);
        return acc;
    };
    auto Pmul = [&](const double *xv, int t) -> double {
        const int v = lane + 32 * t;
        double r = pd[t] * xv[v];
        if (va[t] == 2) r -= wj2 * ((vk[t] > 0 ? xv[v - 3] : 0.0) + (vk[t] < q - 1 ? xv[v + 3] : 0.0));
        return r;
    };
    auto ATy = [&](const double *yv, int t) -> double {
        const int k = vk[t], a = va[t];
        double r = 0.0;
        if (k <= q - 2) r += W.ac[a] * yv[2 * k] + W.ac[6 + a] * yv[2 * k + 1];
        if (k >= 1) r += W.ac[a + 3] * yv[2 * k - 2] + W.ac[6 + a + 3] * yv[2 * k - 1];
        return r;
    };
    auto Ax = [&](const double *xv) -> double {
        if (lane >= ne) return 0.0;
        const int i = lane >> 1, par = (lane & 1) * 6;
        double r = 0.0;
        EMP_UNROLL
        for (int j = 0; j < 6; ++j) r += W.ac[par + j] * xv[3 * i + j];
        return r;
    };

    // ---- factor: chains of S, Y = L^-1 A', K = Y'Y and its Cholesky factor ------------------------
    auto factor = [&](const double w[5]) -> bool {
        bool bad = false;
        __syncwarp();
        if (lane < MAXQ) W.wm[lane] = on[4] ? w[4] : 0.0;
        __syncwarp();
        EMP_UNROLL
        for (int t = 0; t < 2; ++t)
            if (vo[t]) {
                double d = pd[t] + ((on[2 * t] ? w[2 * t] : 0.0) + (on[2 * t + 1] ? w[2 * t + 1] : 0.0));
                if (va[t] == 0) d += (vk[t] < nm ? W.wm[vk[t]] : 0.0) + (vk[t] >= 1 ? W.wm[vk[t] - 1] : 0.0);
                W.dg[va[t]][vk[t]] = d;
            }
        __syncwarp();
        if (lane < 3) {
            double prev = 0.0;
            for (int k = 0; k < q; ++k) {
                const double dd = W.dg[lane][k] - prev * prev;
                bad = bad || !(dd > 0.0);
                const double inv = frsqrt(dd);
                const double e = k < q - 1 ? (lane == 0 ? -W.wm[k] : lane == 2 ? -wj2 : 0.0) * inv : 0.0;
                W.ldi[lane][k] = inv;
                W.le[lane][k] = e;
                prev = e;
            }
        }
        if (__any_sync(EMP_FULL_MASK, bad)) return true;
        __syncwarp();
        if (lane < ne) {
            const int i = lane >> 1, par = (lane & 1) * 6;
            EMP_UNROLL
            for (int a = 0; a < 3; ++a) {
                double yprev = 0.0;
                for (int k = 0; k < q; ++k) {
                    const double b = k == i ? W.ac[par + a] : (k == i + 1 ? W.ac[par + a + 3] : 0.0);
                    const double yv = (b - (k > 0 ? W.le[a][k - 1] : 0.0) * yprev) * W.ldi[a][k];
                    W.Y[(3 * k + a) * YS + lane] = yv;
                    yprev = yv;
                }
            }
        }
        __syncwarp();
        if (lane < ne) {
            double acc[NE];
            EMP_UNROLL
            for (int c = 0; c < NE; ++c) acc[c] = 0.0;
            for (int r = 0; r < nv; ++r) {
                const double a1 = W.Y[r * YS + lane];
                const int lim = 2 * (r / 3 + 1);                     // Y[r][c] = 0 for columns of later stations
                EMP_UNROLL
                for (int c = 0; c < NE; ++c) {
                    if (c >= lim) break;                             // uniform
                    acc[c] = fma(a1, W.Y[r * YS + c], acc[c]);
                }
            }
            EMP_UNROLL
            for (int c = 0; c < NE; ++c)
                if (c <= lane) W.K[lane * KS + c] = acc[c];
        }
        __syncwarp();
        for (int j = 0; j < ne; ++j) {
            const double djj = W.K[j * KS + j];
            if (!(djj > 0.0)) return true;                            // uniform
            const double inv = frsqrt(djj);
            double lij = 0.0;
            if (lane > j && lane < ne) {
                lij = W.K[lane * KS + j] * inv;
                W.K[lane * KS + j] = lij;
            }
            if (lane == j) W.kdi[j] = inv;
            __syncwarp();
            if (lane > j && lane < ne)
                for (int k = j + 1; k <= lane; ++k) W.K[lane * KS + k] = fma(-lij, W.K[k * KS + j], W.K[lane * KS + k]);
            __syncwarp();
        }
        return false;
    };

    // ---- solve: W.v = rhs of the x block -> dx;  returns dy of this lane (by = rhs of the equality block) -------
    auto solve = [&](double by) -> double {
        __syncwarp();
        if (lane < 3) {
            double prev = 0.0;
            for (int k = 0; k < q; ++k) {
                const double val = (W.v[3 * k + lane] - (k > 0 ? W.le[lane][k - 1] : 0.0) * prev) * W.ldi[lane][k];
                W.v[3 * k + lane] = val;
                prev = val;
            }
        }
        __syncwarp();
        double yv = 0.0;
        if (lane < ne) {
            for (int r = 3 * (lane >> 1); r < nv; ++r) yv = fma(W.Y[r * YS + lane], W.v[r], yv);
            yv -= by;
        }
        for (int j = 0; j < ne; ++j) {
            if (lane == j) yv *= W.kdi[j];
            const double yj = shfl(yv, j);
            if (lane > j && lane < ne) yv = fma(-W.K[lane * KS + j], yj, yv);
        }
        for (int j = ne - 1; j >= 0; --j) {
            if (lane == j) yv *= W.kdi[j];
            const double yj = shfl(yv, j);
            if (lane < j) yv = fma(-W.K[j * KS + lane], yj, yv);
        }
        W.y[lane] = lane < ne ? yv : 0.0;
        __syncwarp();
        EMP_UNROLL
        for (int t = 0; t < 2; ++t)
            if (vo[t]) {
                const int v = lane + 32 * t;
                double acc = W.v[v];
                const int lim = min(ne, 2 * (vk[t] + 1));
                for (int c = 0; c < lim; ++c) acc = fma(-W.Y[v * YS + c], W.y[c], acc);
                W.v[v] = acc;
            }
        __syncwarp();
        if (lane < 3) {
            double next = 0.0;
            for (int k = q - 1; k >= 0; --k) {
                const double val = (W.v[3 * k + lane] - W.le[lane][k] * next) * W.ldi[lane][k];
                W.v[3 * k + lane] = val;
                next = val;
            }
        }
        __syncwarp();
        return yv;
    };

    // ---- initial point: W = I -------------------------------------------------------------------
    int status = EMP_QP_OK, it = 0;
    bool have_x = false;
    double s[5], z[5], w[5], g[5];
    EMP_UNROLL
    for (int r = 0; r < 5; ++r) { w[r] = 1.0; s[r] = 1.0; z[r] = 1.0; }
    double yv = 0.0, gap = 0.0, resx0 = 1.0, resz0 = 1.0, gap_min = 0.0;
    if (factor(w)) {
        status = EMP_QP_UNKNOWN_NUMERIC;                               // cvxopt: ValueError("Rank(A) < p or ...")
    } else {
        if (lane < MAXQ) W.wm[lane] = 0.0;                             // h of the no-reversing rows is 0
        __syncwarp();
        EMP_UNROLL
        for (int t = 0; t < 2; ++t)
            if (vo[t]) W.v[lane + 32 * t] = -qv[t] + GTvar(hr, t);
        yv = solve(0.0);
        EMP_UNROLL
        for (int t = 0; t < 2; ++t)
            if (vo[t]) W.x[lane + 32 * t] = W.v[lane + 32 * t];
        have_x = true;
        __syncwarp();
        Gmul(W.x, g);
        double red[3] = {0.0, 0.0, 0.0}, mx_s = -1e300, mx_z = -1e300;   // |s|^2, |q|^2, |h|^2
        EMP_UNROLL
        for (int r = 0; r < 5; ++r)
            if (on[r]) {
                z[r] = g[r] - hr[r];
                s[r] = -z[r];
                red[0] += s[r] * s[r];
                red[2] += hr[r] * hr[r];
                mx_s = fmax(mx_s, -s[r]);
                mx_z = fmax(mx_z, -z[r]);
            }
        red[1] = qv[0] * qv[0] + qv[1] * qv[1];
        EMP_UNROLL
        for (int k = 0; k < 3; ++k) red[k] = warp_sum(red[k]);
        const double nrm = sqrt(red[0]);
        resx0 = fmax(1.0, sqrt(red[1]));
        resz0 = fmax(1.0, sqrt(red[2]));
        const double ts = warp_max(mx_s), tz = warp_max(mx_z);
        const double sh_s = ts >= -1e-8 * fmax(nrm, 1.0) ? 1.0 + ts : 0.0;
        const double sh_z = tz >= -1e-8 * fmax(nrm, 1.0) ? 1.0 + tz : 0.0;
        double gs = 0.0;
        EMP_UNROLL
        for (int r = 0; r < 5; ++r)
            if (on[r]) {
                if (sh_s != 0.0) s[r] += sh_s;
                if (sh_z != 0.0) z[r] += sh_z;
                gs += s[r] * z[r];
            }
        gap = warp_sum(gs);
        gap_min = gap;

        // ---- interior-point iterations --------------------------------------------------------------
        double rxv[2], rz[5], ws3[5], tq[5], dsv[5], dz[5];
        double xp[2] = {0.0, 0.0};                                     // iterate before the last step (divergence exit)
        for (it = 0;; ++it) {
            __syncwarp();
            W.y[lane] = lane < ne ? yv : 0.0;
            if (lane < MAXQ) W.wm[lane] = on[4] ? z[4] : 0.0;
            __syncwarp();
            double red7[7] = {0, 0, 0, 0, 0, 0, 0};       // |rx|^2 |rz|^2 x.(Px+q) x.q z.rz |ry|^2 y.ry
            EMP_UNROLL
            for (int t = 0; t < 2; ++t) {
                rxv[t] = 0.0;
                if (vo[t]) {
                    const double xi = W.x[lane + 32 * t];
                    const double pxq = Pmul(W.x, t) + qv[t];
                    rxv[t] = (pxq + ATy(W.y, t)) + GTvar(z, t);
                    red7[0] += rxv[t] * rxv[t];
                    red7[2] += xi * pxq;
                    red7[3] += xi * qv[t];
                }
            }
            const double ry = Ax(W.x);
            red7[5] = ry * ry;
            red7[6] = yv * ry;
            Gmul(W.x, g);
            EMP_UNROLL
            for (int r = 0; r < 5; ++r) {
                rz[r] = on[r] ? (s[r] + g[r]) - hr[r] : 0.0;
                red7[1] += rz[r] * rz[r];
                red7[4] += on[r] ? z[r] * rz[r] : 0.0;
            }
            EMP_UNROLL
            for (int off = 16; off > 0; off >>= 1)
                EMP_UNROLL
                for (int k = 0; k < 7; ++k) red7[k] += shfl_xor(red7[k], off);
            const double resx = sqrt(red7[0]), resz = sqrt(red7[1]), resy = sqrt(red7[5]);
            const double pcost = 0.5 * (red7[2] + red7[3]);
            const double dcost = ((pcost + red7[6]) + red7[4]) - gap;
            bool have_rel = true;
            double relgap = 0.0;
            if (pcost < 0.0) relgap = gap / -pcost;
            else if (dcost > 0.0) relgap = gap / dcost;
            else have_rel = false;
            const double pres = fmax(resy, resz / resz0), dres = resx / resx0;       // resy0 = max(1, |beq|) = 1
            if (pres <= Q.feastol && dres <= Q.feastol && (gap <= Q.abstol || (have_rel && relgap <= Q.reltol))) break;
            if (it == Q.maxiters) { status = EMP_QP_UNKNOWN_MAXITER; break; }
            if (!isfinite((gap + resx) + resz)) { status = EMP_QP_UNKNOWN_NUMERIC; break; }
            if (it > 0 && Q.diverge_factor > 0.0 && gap > Q.diverge_factor * gap_min) {
                status = EMP_QP_UNKNOWN_DIVERGED;                      // the step that blew the gap up is discarded
                __syncwarp();                                          // every lane is done reading W.x (residuals above)
                EMP_UNROLL
                for (int t = 0; t < 2; ++t)
                    if (vo[t]) W.x[lane + 32 * t] = xp[t];
                break;
            }
            gap_min = fmin(gap_min, gap);
            EMP_UNROLL
            for (int t = 0; t < 2; ++t) xp[t] = vo[t] ? W.x[lane + 32 * t] : 0.0;

            double is[5];
            EMP_UNROLL
            for (int r = 0; r < 5; ++r) { is[r] = frcp(s[r]); w[r] = on[r] ? z[r] * is[r] : 0.0; }
            if (factor(w)) { status = EMP_QP_UNKNOWN_NUMERIC; break; }
            const double mu = gap / (double)m_rows;
            double sigma = 0.0, step = 1.0, dyv = 0.0;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                EMP_UNROLL
                for (int r = 0; r < 5; ++r) {
                    const double bs = (sigma * mu - s[r] * z[r]) - (pass == 1 ? ws3[r] : 0.0);
                    tq[r] = on[r] ? (bs + z[r] * rz[r]) * is[r] : 0.0;
                }
                __syncwarp();
                if (lane < MAXQ) W.wm[lane] = on[4] ? tq[4] : 0.0;
                __syncwarp();
                EMP_UNROLL
                for (int t = 0; t < 2; ++t)
                    if (vo[t]) W.v[lane + 32 * t] = -rxv[t] - GTvar(tq, t);
                dyv = solve(-ry);
                Gmul(W.v, g);
                double dsdz = 0.0, tmax = -1e300;
                EMP_UNROLL
                for (int r = 0; r < 5; ++r) {
                    dz[r] = tq[r] + w[r] * g[r];
                    dsv[r] = -rz[r] - g[r];
                    if (on[r]) {
                        dsdz += dsv[r] * dz[r];
                        tmax = fmax(tmax, fmax(-dsv[r] * is[r], -dz[r] * frcp(z[r])));
                    }
                    if (pass == 0) ws3[r] = on[r] ? dsv[r] * dz[r] : 0.0;
                }
                dsdz = warp_sum(dsdz);
                const double tt = fmax(0.0, warp_max(tmax));
                if (tt == 0.0) step = 1.0;
                else step = pass == 0 ? fmin(1.0, 1.0 / tt) : fmin(1.0, 0.99 / tt);
                if (pass == 0) {
                    const double sg = fmin(1.0, fmax(0.0, (1.0 - step) + (dsdz / gap) * (step * step)));
                    sigma = (sg * sg) * sg;
                }
            }
            EMP_UNROLL
            for (int t = 0; t < 2; ++t)
                if (vo[t]) W.x[lane + 32 * t] += step * W.v[lane + 32 * t];
            yv += step * dyv;
            gs = 0.0;
            EMP_UNROLL
            for (int r = 0; r < 5; ++r)
                if (on[r]) {
                    s[r] += step * dsv[r];
                    z[r] += step * dz[r];
                    gs += s[r] * z[r];
                }
            gap = warp_sum(gs);
        }
    }
    // ---- outputs (:505-509): 17 entries each, NaN beyond qp_size --------------------------------------
    __syncwarp();
    if (lane < NQP) {
        const bool have = lane < q && have_x;
        o[lane] = have ? W.x[3 * lane] : NaN;
        o[NQP + lane] = have ? W.x[3 * lane + 1] : NaN;
        o[2 * NQP + lane] = have ? W.x[3 * lane + 2] : NaN;
        o[3 * NQP + lane] = lane < q ? (double)lane * dt : NaN;
    }
    if (lane == 0) {
        status_iters[2 * f] = status;
        status_iters[2 * f + 1] = it;
    }
}

}  // namespace

size_t speed_qp_smem_bytes() { return SQP_WARPS * sizeof(WarpSmem); }

cudaError_t launch_speed_qp(const double *start, const double *st_out, const double *bounds, double *out,
                            int32_t *status_iters, const SpeedQpLaunch &p, cudaStream_t st) {
    const size_t smem = speed_qp_smem_bytes();
    cudaFuncSetAttribute(speed_qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int grid = (p.batch + SQP_WARPS - 1) / SQP_WARPS;
    speed_qp_kernel<<<grid, SQP_WARPS * 32, smem, st>>>(start, st_out, bounds, out, status_iters, p);
    return cudaGetLastError();
}

}  // namespace emp
