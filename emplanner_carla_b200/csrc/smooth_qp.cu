// K5 - reference-line smoother (box-constrained QP) + heading/curvature, one warp per polyline.
//
// Replaces reference planner/planning_utils.py smooth_reference_line (:262-361, cvxopt.solvers.qp at :353) and
// cal_heading_kappa (:185-228).  It is QP#1 (51-node local reference line, test_9.py:110) and QP#3 (the
// Cartesian path at the end of frenet_2_x_y_theta_kappa, path_planning.py:47) of every planning frame.
//
//   min  ws |A1 p|^2 + wl |A2 p|^2 + wr |p - ref|^2      A1 = second differences, A2 = first differences
//   s.t. ref - thr <= p <= ref + thr                      (4n inequality rows, no equalities)
//
// x and y decouple; for each coordinate H is pentadiagonal.  Two consecutive points per lane turn it into a
// block-tridiagonal system with 2x2 blocks (n <= 64), factored by the same lane-parallel block cyclic reduction
// as the path QP (csrc/sl_qp.cu), for the x and the y system side by side.  The iteration is cvxopt coneqp's
// (oracle/coneqp.py): same initial point, step rule and stopping test; s, z carried unscaled.
#include "emp_common.cuh"
#include "smooth_kernels.h"

namespace emp {
namespace {

#define EMP_UNROLL _Pragma("unroll")

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

// 2x2-block cyclic reduction factor of one coordinate's system (per lane)
struct Fac2 {
    double l10, i00, i11;       // Cholesky of the pivot block: L = [[1/i00, 0], [l10, 1/i11]]
    double xl[4], xr[4];        // L^-1 K(i, i-d), L^-1 K(i, i+d), row-major
};

template <int NL>
__device__ __forceinline__ bool cr_factor2(double d00, double d10, double d11, const double cl_in[4], const double cr_in[4],
                                           bool stage, int lane, int levels, Lanes<NL> &L, Fac2 &F) {
    double CL[4], CR[4];
    EMP_UNROLL
    for (int k = 0; k < 4; ++k) { CL[k] = cl_in[k]; CR[k] = cr_in[k]; F.xl[k] = 0.0; F.xr[k] = 0.0; }
    F.l10 = 0.0; F.i00 = 1.0; F.i11 = 1.0;
    bool bad = false;
#pragma unroll 1
    for (int lev = 0; lev < levels; ++lev) {
        const int d = 1 << lev, msk = 2 * d - 1;
        if ((lane & msk) == d - 1) {
            const double j00 = frsqrt(d00);
            const double c10 = d10 * j00;
            const double t11 = d11 - c10 * c10;
            const double j11 = frsqrt(t11);
            bad = bad || (stage && (!(d00 > 0.0) || !(t11 > 0.0) || !isfinite(j00 + j11)));
            F.l10 = c10; F.i00 = j00; F.i11 = j11;
            EMP_UNROLL
            for (int c = 0; c < 2; ++c) {
                const double a0 = CL[c] * j00;
                F.xl[c] = a0; F.xl[2 + c] = (CL[2 + c] - c10 * a0) * j11;
                const double b0 = CR[c] * j00;
                F.xr[c] = b0; F.xr[2 + c] = (CR[2 + c] - c10 * b0) * j11;
            }
        }
        // see csrc/sl_qp.cu: no lane predicates needed for the updates
        EMP_UNROLL
        for (int r = 0; r < 2; ++r) {
            EMP_UNROLL
            for (int c = 0; c < 2; ++c) {
                const double plr = F.xl[r] * F.xr[c] + F.xl[2 + r] * F.xr[2 + c];
                CL[2 * c + r] = -L.up(plr, d);
                CR[2 * r + c] = -L.down(plr, d);
                if (c <= r) {
                    const double prr = F.xr[r] * F.xr[c] + F.xr[2 + r] * F.xr[2 + c];
                    const double pll = F.xl[r] * F.xl[c] + F.xl[2 + r] * F.xl[2 + c];
                    const double upd = L.up(prr, d) + L.down(pll, d);
                    if (r == 0) d00 -= upd;
                    else if (c == 0) d10 -= upd;
                    else d11 -= upd;
                }
            }
        }
    }
    return bad;
}

template <int NL>
__device__ __forceinline__ void cr_solve2(const Fac2 &F, double r0, double r1, int lane, int levels, Lanes<NL> &L, double &y0,
                                          double &y1) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll 1
    for (int lev = 0; lev < levels; ++lev) {
        const int d = 1 << lev, msk = 2 * d - 1;
        double ul0 = 0.0, ul1 = 0.0, ur0 = 0.0, ur1 = 0.0;
        if ((lane & msk) == d - 1) {
            t0 = r0 * F.i00;
            t1 = (r1 - F.l10 * t0) * F.i11;
            ul0 = F.xl[0] * t0 + F.xl[2] * t1; ul1 = F.xl[1] * t0 + F.xl[3] * t1;
            ur0 = F.xr[0] * t0 + F.xr[2] * t1; ur1 = F.xr[1] * t0 + F.xr[3] * t1;
        }
        r0 -= L.up(ur0, d) + L.down(ul0, d);
        r1 -= L.up(ur1, d) + L.down(ul1, d);
    }
    y0 = 0.0; y1 = 0.0;
#pragma unroll 1
    for (int lev = levels - 1; lev >= 0; --lev) {
        const int d = 1 << lev, msk = 2 * d - 1;
        const double a0 = L.up(y0, d), a1 = L.up(y1, d);
        const double b0 = L.down(y0, d), b1 = L.down(y1, d);
        if ((lane & msk) == d - 1) {
            const double v0 = t0 - ((F.xl[0] * a0 + F.xl[1] * a1) + (F.xr[0] * b0 + F.xr[1] * b1));
            const double v1 = t1 - ((F.xl[2] * a0 + F.xl[3] * a1) + (F.xr[2] * b0 + F.xr[3] * b1));
            y1 = v1 * F.i11;
            y0 = (v0 - F.l10 * y1) * F.i00;
        }
    }
}

// One polyline on NL lanes, two consecutive points per lane: NL = 32 one warp (n <= 64), NL = 64 one CTA (n <= 128).
// `scr`: this polyline's [3][2 NL] doubles for the heading / curvature stage; `xch`: Lanes<64>'s exchange lines.
template <int NL>
__device__ __forceinline__ void
smooth_one_line(const int line, const int lane, double *__restrict__ scr_base, double *__restrict__ xch,
                const double *__restrict__ xy_in, const int32_t *__restrict__ npts, double *__restrict__ out,
                int32_t *__restrict__ status_iters, const SmoothLaunch &p) {
    Lanes<NL> L(lane, xch);
    const int n = npts ? min(npts[line], p.max_points) : p.max_points;
    const EmpSmoothParams &Q = p.prm;
    int status = 0, it = -1;
    double v[2][2] = {{0.0, 0.0}, {0.0, 0.0}};            // [coordinate][point within lane]
    if (n < 2 || n > 2 * NL) {
        status = EMP_QP_TOO_LARGE;
    } else if (p.heading_only) {
        EMP_UNROLL
        for (int q = 0; q < 2; ++q) {                      // cal_heading_kappa alone: the input is the polyline
            const int i = min(2 * lane + q, n - 1);
            v[0][q] = xy_in[((size_t)line * p.max_points + i) * 2];
            v[1][q] = xy_in[((size_t)line * p.max_points + i) * 2 + 1];
        }
    } else {
        const int NSTG = (n + 1) >> 1;                     // stages (lanes in use)
        const bool stage = lane < NSTG;
        // cyclic-reduction levels: log2(NL) eliminate lanes 0 .. NL-2; the last lane is only a stage when the polyline
        // fills every lane (n = 2 NL - 1, 2 NL) and then takes one more level of its own
        const int levels = (NL == 32 ? 5 : 6) + (NSTG == NL ? 1 : 0);
        bool on[2];
        double ref[2][2], thr[2] = {Q.x_thre, Q.y_thre};
        double hd[2], ho1[2], ho2[2];                      // pentadiagonal H: diag, (i,i+1), (i,i+2) per point
        EMP_UNROLL
        for (int q = 0; q < 2; ++q) {
            const int i = 2 * lane + q;
            on[q] = i < n;
            const int ii = on[q] ? i : 0;
            ref[0][q] = xy_in[((size_t)line * p.max_points + ii) * 2];
            ref[1][q] = xy_in[((size_t)line * p.max_points + ii) * 2 + 1];
            // A1'A1 and A2'A2 entries (planning_utils.py:313-344)
            const double d1 = (i <= n - 3 ? 1.0 : 0.0) + (i >= 1 && i <= n - 2 ? 4.0 : 0.0) + (i >= 2 ? 1.0 : 0.0);
            const double o1 = -2.0 * ((i <= n - 3 ? 1.0 : 0.0) + (i >= 1 && i <= n - 2 ? 1.0 : 0.0));
            const double o2 = i <= n - 3 ? 1.0 : 0.0;
            const double d2 = (i <= n - 2 ? 1.0 : 0.0) + (i >= 1 ? 1.0 : 0.0);
            const double p1 = i <= n - 2 ? -1.0 : 0.0;
            hd[q] = on[q] ? 2 * ((Q.w_smooth * d1 + Q.w_length * d2) + Q.w_ref) : 1.0;
            ho1[q] = on[q] ? 2 * (Q.w_smooth * o1 + Q.w_length * p1) : 0.0;
            ho2[q] = on[q] ? 2 * (Q.w_smooth * o2) : 0.0;
        }
        const int m_rows = 4 * n;
        // H x for one coordinate: needs the neighbours at distance 1 and 2 (previous / next lane)
        auto hmul = [&](const double x[2], double o[2]) {
            const double pm2 = L.up(x[0], 1), pm1 = L.up(x[1], 1);       // points 2l-2, 2l-1
            const double pp2 = L.down(x[0], 1), pp3 = L.down(x[1], 1);   // points 2l+2, 2l+3
            const double c_m2 = L.up(ho2[0], 1), c_m1a = L.up(ho1[1], 1), c_m1b = L.up(ho2[1], 1);
            const bool has_prev = lane > 0;
            // point 0 of the lane (index 2l): couples 2l-2 (ho2 of 2l-2), 2l-1 (ho1 of 2l-1), 2l+1 (ho1), 2l+2 (ho2)
            o[0] = hd[0] * x[0] + ho1[0] * x[1] + ho2[0] * pp2 + (has_prev ? c_m2 * pm2 + c_m1a * pm1 : 0.0);
            // point 1 (index 2l+1): couples 2l-1 (ho2 of 2l-1), 2l (ho1 of 2l), 2l+2 (ho1), 2l+3 (ho2)
            o[1] = hd[1] * x[1] + ho1[0] * x[0] + ho1[1] * pp2 + ho2[1] * pp3 + (has_prev ? c_m1b * pm1 : 0.0);
        };
        // rows: [c][q][0] = upper bound, [c][q][1] = lower bound
        double s[2][2][2], z[2][2][2];
        EMP_UNROLL
        for (int c = 0; c < 2; ++c)
            EMP_UNROLL
            for (int q = 0; q < 2; ++q) { s[c][q][0] = s[c][q][1] = 1.0; z[c][q][0] = z[c][q][1] = 1.0; }
        auto hrow = [&](int c, int q, int b) -> double { return b == 0 ? ref[c][q] + thr[c] : -(ref[c][q] - thr[c]); };
        double gap = 0.0, resx0 = 1.0, resz0 = 1.0, gap_min = 0.0;
        double vp[2][2] = {{0.0, 0.0}, {0.0, 0.0}};        // iterate before the last step (divergence exit)
        double rx[2][2] = {{0, 0}, {0, 0}}, rz[2][2][2], ws3[2][2][2];
        EMP_UNROLL
        for (int c = 0; c < 2; ++c)
            EMP_UNROLL
            for (int q = 0; q < 2; ++q) { rz[c][q][0] = rz[c][q][1] = 0.0; ws3[c][q][0] = ws3[c][q][1] = 0.0; }
        while (true) {
            if (it >= 0) {
                double red[5] = {0, 0, 0, 0, 0};           // |rx|^2, |rz|^2, x.(Hx+f), x.f, z.rz
                EMP_UNROLL
                for (int c = 0; c < 2; ++c) {
                    double hx[2];
                    hmul(v[c], hx);
                    EMP_UNROLL
                    for (int q = 0; q < 2; ++q) {
                        const double f = -2 * Q.w_ref * ref[c][q];
                        const double hxf = hx[q] + f;
                        rx[c][q] = on[q] ? hxf + (z[c][q][0] - z[c][q][1]) : 0.0;
                        rz[c][q][0] = on[q] ? (s[c][q][0] + v[c][q]) - hrow(c, q, 0) : 0.0;
                        rz[c][q][1] = on[q] ? (s[c][q][1] - v[c][q]) - hrow(c, q, 1) : 0.0;
                        if (on[q]) {
                            red[0] += rx[c][q] * rx[c][q];
                            red[1] += rz[c][q][0] * rz[c][q][0] + rz[c][q][1] * rz[c][q][1];
                            red[2] += v[c][q] * hxf;
                            red[3] += v[c][q] * f;
                            red[4] += z[c][q][0] * rz[c][q][0] + z[c][q][1] * rz[c][q][1];
                        }
                    }
                }
                L.sum_n(red);
                const double resx = sqrt(red[0]), resz = sqrt(red[1]);
                const double pcost = 0.5 * (red[2] + red[3]);
                const double dcost = (pcost + red[4]) - gap;
                bool have_rel = true;
                double relgap = 0.0;
                if (pcost < 0.0) relgap = gap / -pcost;
                else if (dcost > 0.0) relgap = gap / dcost;
                else have_rel = false;
                const double pres = resz / resz0, dres = resx / resx0;
                if (pres <= Q.feastol && dres <= Q.feastol && (gap <= Q.abstol || (have_rel && relgap <= Q.reltol))) break;
                if (it == Q.maxiters) { status = EMP_QP_UNKNOWN_MAXITER; break; }
                if (!isfinite((gap + resx) + resz)) { status = EMP_QP_UNKNOWN_NUMERIC; break; }
                if (it > 0 && Q.diverge_factor > 0.0 && gap > Q.diverge_factor * gap_min) {
                    status = EMP_QP_UNKNOWN_DIVERGED;              // the step that blew the gap up is discarded
                    v[0][0] = vp[0][0]; v[0][1] = vp[0][1]; v[1][0] = vp[1][0]; v[1][1] = vp[1][1];
                    break;
                }
                gap_min = fmin(gap_min, gap);
                vp[0][0] = v[0][0]; vp[0][1] = v[0][1]; vp[1][0] = v[1][0]; vp[1][1] = v[1][1];
            }
            // factor S = H + diag(z/s summed over a variable's two rows), both coordinates
            Fac2 F[2];
            double is[2][2][2], w[2][2][2];
            bool bad = false;
            EMP_UNROLL
            for (int c = 0; c < 2; ++c) {
                double wv[2];
                EMP_UNROLL
                for (int q = 0; q < 2; ++q) {
                    EMP_UNROLL
                    for (int b = 0; b < 2; ++b) {
                        is[c][q][b] = frcp(s[c][q][b]);
                        w[c][q][b] = z[c][q][b] * is[c][q][b];
                    }
                    wv[q] = on[q] ? w[c][q][0] + w[c][q][1] : 0.0;
                }
                // couplings: CR = K(lane, lane+1) = [[ho2[0], 0], [ho1[1], ho2[1]]]; CL = K(lane, lane-1) = CR(lane-1)'
                double CR[4] = {ho2[0], 0.0, ho1[1], ho2[1]};
                double CL[4];
                const double p0 = L.up(CR[0], 1), p2 = L.up(CR[2], 1), p3 = L.up(CR[3], 1);
                CL[0] = lane > 0 ? p0 : 0.0; CL[1] = lane > 0 ? p2 : 0.0; CL[2] = 0.0; CL[3] = lane > 0 ? p3 : 0.0;
                if (!stage) { EMP_UNROLL for (int k = 0; k < 4; ++k) { CR[k] = 0.0; CL[k] = 0.0; } }
                bad = cr_factor2(stage ? hd[0] + wv[0] : 1.0, stage ? ho1[0] : 0.0, stage ? hd[1] + wv[1] : 1.0, CL, CR,
                                 stage, lane, levels, L, F[c]) || bad;
            }
            if (L.any(bad)) { status = EMP_QP_UNKNOWN_NUMERIC; break; }

            const double mu = gap / (double)m_rows;
            const int npass = it < 0 ? 1 : 2;
            double sigma = 0.0, step = 1.0;
            double dx[2][2], dsv[2][2][2], dz[2][2][2];
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass) {
                double tq[2][2][2];
                EMP_UNROLL
                for (int c = 0; c < 2; ++c) {
                    double g[2];
                    EMP_UNROLL
                    for (int q = 0; q < 2; ++q) {
                        if (it < 0) {
                            // initial point: (H + G'G) x = -f + G'h
                            tq[c][q][0] = tq[c][q][1] = 0.0;
                            g[q] = on[q] ? 2 * Q.w_ref * ref[c][q] + (hrow(c, q, 0) - hrow(c, q, 1)) : 0.0;
                        } else {
                            EMP_UNROLL
                            for (int b = 0; b < 2; ++b) {
                                const double bs = (sigma * mu - s[c][q][b] * z[c][q][b]) - (pass == 1 ? ws3[c][q][b] : 0.0);
                                tq[c][q][b] = on[q] ? (bs + z[c][q][b] * rz[c][q][b]) * is[c][q][b] : 0.0;
                            }
                            g[q] = -rx[c][q] - (tq[c][q][0] - tq[c][q][1]);
                        }
                    }
                    cr_solve2(F[c], stage ? g[0] : 0.0, stage ? g[1] : 0.0, lane, levels, L, dx[c][0], dx[c][1]);
                }
                if (it < 0) break;
                double dsdz = 0.0, tmax = -1e300;
                EMP_UNROLL
                for (int c = 0; c < 2; ++c)
                    EMP_UNROLL
                    for (int q = 0; q < 2; ++q)
                        EMP_UNROLL
                        for (int b = 0; b < 2; ++b) {
                            const double gdx = b == 0 ? dx[c][q] : -dx[c][q];
                            dz[c][q][b] = tq[c][q][b] + w[c][q][b] * gdx;
                            dsv[c][q][b] = -rz[c][q][b] - gdx;
                            if (on[q]) {
                                dsdz += dsv[c][q][b] * dz[c][q][b];
                                tmax = fmax(tmax, fmax(-dsv[c][q][b] * is[c][q][b], -dz[c][q][b] / z[c][q][b]));
                            }
                            if (pass == 0) ws3[c][q][b] = dsv[c][q][b] * dz[c][q][b];
                        }
                dsdz = L.sum(dsdz);
                const double t = fmax(0.0, L.max(tmax));
                if (t == 0.0) step = 1.0;
                else step = pass == 0 ? fmin(1.0, 1.0 / t) : fmin(1.0, 0.99 / t);
                if (pass == 0) {
                    const double sg = fmin(1.0, fmax(0.0, (1.0 - step) + (dsdz / gap) * (step * step)));
                    sigma = (sg * sg) * sg;
                }
            }
            if (it < 0) {
                double red[3] = {0, 0, 0};                 // |s|^2, |f|^2, |h|^2
                double mn_s = -1e300, mn_z = -1e300;
                EMP_UNROLL
                for (int c = 0; c < 2; ++c)
                    EMP_UNROLL
                    for (int q = 0; q < 2; ++q) {
                        v[c][q] = on[q] ? dx[c][q] : 0.0;
                        const double f = -2 * Q.w_ref * ref[c][q];
                        EMP_UNROLL
                        for (int b = 0; b < 2; ++b) {
                            const double hr = hrow(c, q, b);
                            const double gx = b == 0 ? v[c][q] : -v[c][q];
                            z[c][q][b] = on[q] ? gx - hr : 1.0;
                            s[c][q][b] = on[q] ? -z[c][q][b] : 1.0;
                            if (on[q]) {
                                red[0] += s[c][q][b] * s[c][q][b];
                                red[2] += hr * hr;
                                mn_s = fmax(mn_s, -s[c][q][b]);
                                mn_z = fmax(mn_z, -z[c][q][b]);
                            }
                        }
                        if (on[q]) red[1] += f * f;
                    }
                L.sum_n(red);
                const double nrm = sqrt(red[0]);
                resx0 = fmax(1.0, sqrt(red[1]));
                resz0 = fmax(1.0, sqrt(red[2]));
                const double ts = L.max(mn_s), tz = L.max(mn_z);
                const double sh_s = ts >= -1e-8 * fmax(nrm, 1.0) ? 1.0 + ts : 0.0;
                const double sh_z = tz >= -1e-8 * fmax(nrm, 1.0) ? 1.0 + tz : 0.0;
                double gs = 0.0;
                EMP_UNROLL
                for (int c = 0; c < 2; ++c)
                    EMP_UNROLL
                    for (int q = 0; q < 2; ++q)
                        EMP_UNROLL
                        for (int b = 0; b < 2; ++b)
                            if (on[q]) {
                                if (sh_s != 0.0) s[c][q][b] += sh_s;
                                if (sh_z != 0.0) z[c][q][b] += sh_z;
                                gs += s[c][q][b] * z[c][q][b];
                            }
                gap = L.sum(gs);
                gap_min = gap;
            } else {
                double gs = 0.0;
                EMP_UNROLL
                for (int c = 0; c < 2; ++c)
                    EMP_UNROLL
                    for (int q = 0; q < 2; ++q) {
                        if (on[q]) v[c][q] += step * dx[c][q];
                        EMP_UNROLL
                        for (int b = 0; b < 2; ++b)
                            if (on[q]) {
                                s[c][q][b] += step * dsv[c][q][b];
                                z[c][q][b] += step * dz[c][q][b];
                                gs += s[c][q][b] * z[c][q][b];
                            }
                    }
                gap = L.sum(gs);
            }
            ++it;
        }
    }
    // ---- outputs: (x, y, theta, kappa) with cal_heading_kappa (:185-228) -------------------------
    double *o = out + (size_t)line * p.max_points * 4;
    if (status != EMP_QP_TOO_LARGE) {
        double *sx = scr_base, *sy = scr_base + 2 * NL, *sth = scr_base + 4 * NL;
        EMP_UNROLL
        for (int q = 0; q < 2; ++q) {
            const int i = 2 * lane + q;
            if (i < n) { sx[i] = v[0][q]; sy[i] = v[1][q]; }
        }
        L.sync();
        double ddx[2] = {0, 0}, ddy[2] = {0, 0};
        EMP_UNROLL
        for (int q = 0; q < 2; ++q) {
            const int i = 2 * lane + q;
            if (i < n) {
                const int a = max(i - 1, 0), b = min(i, n - 2);          // dx_pre = dx_[i-1], dx_aft = dx_[i] (clamped)
                ddx[q] = ((sx[a + 1] - sx[a]) + (sx[b + 1] - sx[b])) / 2;
                ddy[q] = ((sy[a + 1] - sy[a]) + (sy[b + 1] - sy[b])) / 2;
                sth[i] = atan2(ddy[q], ddx[q]);
            }
        }
        L.sync();
        EMP_UNROLL
        for (int q = 0; q < 2; ++q) {
            const int i = 2 * lane + q;
            if (i < n) {
                const int a = max(i - 1, 0), b = min(i, n - 2);
                const double dth = ((sth[a + 1] - sth[a]) + (sth[b + 1] - sth[b])) / 2;
                o[4 * i] = sx[i];
                o[4 * i + 1] = sy[i];
                o[4 * i + 2] = sth[i];
                o[4 * i + 3] = sin(dth) / sqrt(ddx[q] * ddx[q] + ddy[q] * ddy[q]);
            }
        }
    }
    if (lane == 0) {
        status_iters[2 * line] = status;
        status_iters[2 * line + 1] = it < 0 ? 0 : it;
    }
}

__global__ void __launch_bounds__(SMO_WARPS * 32)
smooth_kernel(const double *__restrict__ xy_in, const int32_t *__restrict__ npts, double *__restrict__ out,
              int32_t *__restrict__ status_iters, const SmoothLaunch p) {
    __shared__ double scr[SMO_WARPS][3 * 64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int line = blockIdx.x * SMO_WARPS + warp;
    if (line >= p.batch) return;
    smooth_one_line<32>(line, lane, scr[warp], nullptr, xy_in, npts, out, status_iters, p);
}

// polylines of 65 .. 128 points: one CTA of 64 threads per polyline
__global__ void __launch_bounds__(64)
smooth_block_kernel(const double *__restrict__ xy_in, const int32_t *__restrict__ npts, double *__restrict__ out,
                    int32_t *__restrict__ status_iters, const SmoothLaunch p) {
    __shared__ double scr[3 * 128];
    __shared__ double xch[2 * 64];
    smooth_one_line<64>((int)blockIdx.x, (int)threadIdx.x, scr, xch, xy_in, npts, out, status_iters, p);
}

// cal_heading_kappa (planning_utils.py:185-228) for polylines of any length: one thread per point.  theta_i needs the
// points i-1..i+1, kappa_i the thetas i-1..i+1, so a thread reads at most 5 points and evaluates up to 3 arctangents;
// same expressions as the tail of smooth_kernel, so both give the same bits.
__global__ void heading_kappa_kernel(const double *__restrict__ xy_in, const int32_t *__restrict__ npts,
                                     double *__restrict__ out, int32_t *__restrict__ status_iters, int batch,
                                     int max_points) {
    const int line = blockIdx.y;
    const int n = npts ? min(npts[line], max_points) : max_points;
    const double *xy = xy_in + (size_t)line * max_points * 2;
    double *o = out + (size_t)line * max_points * 4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < max_points; i += gridDim.x * blockDim.x) {
        if (i == 0) {
            status_iters[2 * line] = n < 2 ? EMP_QP_TOO_LARGE : 0;
            status_iters[2 * line + 1] = 0;
        }
        if (i >= n || n < 2) continue;
        auto theta_at = [&](int j, double &ddx, double &ddy) -> double {
            const int a = max(j - 1, 0), b = min(j, n - 2);              // dx_pre = dx_[j-1], dx_aft = dx_[j] (clamped)
            ddx = ((xy[2 * (a + 1)] - xy[2 * a]) + (xy[2 * (b + 1)] - xy[2 * b])) / 2;
            ddy = ((xy[2 * (a + 1) + 1] - xy[2 * a + 1]) + (xy[2 * (b + 1) + 1] - xy[2 * b + 1])) / 2;
            return atan2(ddy, ddx);
        };
        double dx, dy, tx, ty;
        const double th = theta_at(i, dx, dy);
        const int a = max(i - 1, 0), b = min(i, n - 2);
        const double th_a = a == i ? th : theta_at(a, tx, ty);
        const double th_a1 = a + 1 == i ? th : theta_at(a + 1, tx, ty);
        const double th_b = b == i ? th : (b == a ? th_a : theta_at(b, tx, ty));
        const double th_b1 = b + 1 == i ? th : (b + 1 == a + 1 ? th_a1 : theta_at(b + 1, tx, ty));
        const double dth = ((th_a1 - th_a) + (th_b1 - th_b)) / 2;
        o[4 * i] = xy[2 * i];
        o[4 * i + 1] = xy[2 * i + 1];
        o[4 * i + 2] = th;
        o[4 * i + 3] = sin(dth) / sqrt(dx * dx + dy * dy);
    }
}

}  // namespace

cudaError_t launch_heading_kappa(const double *xy, const int32_t *npts, double *out, int32_t *status_iters, int batch,
                                 int max_points, cudaStream_t st) {
    const int threads = 128;
    int gx = (max_points + threads - 1) / threads;
    if (gx > 1024) gx = 1024;
    for (int b0 = 0; b0 < batch; b0 += 65535) {                        // gridDim.y limit
        const int nb = batch - b0 < 65535 ? batch - b0 : 65535;
        heading_kappa_kernel<<<dim3(gx, nb), threads, 0, st>>>(xy + (size_t)b0 * max_points * 2, npts ? npts + b0 : nullptr,
                                                              out + (size_t)b0 * max_points * 4, status_iters + 2 * (size_t)b0,
                                                              nb, max_points);
    }
    return cudaGetLastError();
}

cudaError_t launch_smooth(const double *xy, const int32_t *npts, double *out, int32_t *status_iters,
                          const SmoothLaunch &p, cudaStream_t st) {
    if (p.max_points > 64) {                               // general-size path (the kernel reports > 128 as TOO_LARGE)
        smooth_block_kernel<<<p.batch, 64, 0, st>>>(xy, npts, out, status_iters, p);
        return cudaGetLastError();
    }
    const int grid = (p.batch + SMO_WARPS - 1) / SMO_WARPS;
    smooth_kernel<<<grid, SMO_WARPS * 32, 0, st>>>(xy, npts, out, status_iters, p);
    return cudaGetLastError();
}

}  // namespace emp
