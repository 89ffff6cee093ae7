// K2 - corridor + QP path smoother, one warp per ego frame, one lane per station.
//
// Replaces reference planner/path_planning.py cal_lmin_lmax (:222-273) and
// Quadratic_planning (:78-219) including its cvxopt.solvers.qp call (:211-214).
//
// The QP   min 1/2 x'Px + q'x  s.t. Gx <= h, Ax = b   (x = (l,dl,ddl) per station)
// is solved with the same Mehrotra predictor-corrector iteration as cvxopt's
// coneqp (initial point, step rule STEP=0.99, sigma = (1-a+..)^3, stopping test
// with abstol/reltol/feastol) so the iterates and the stopping iteration match
// the CPU oracle (oracle/coneqp.py); only the linear algebra is different:
//   * all 14 inequality rows of a station touch only that station's 3 variables,
//     so S = P + G' diag(z/s) G is formed per lane; its (l,dl) 2x2 block is
//     Cholesky-factored from the closed form of the rank-one sum (no cancellation
//     when a corridor row or a start/end pin becomes stiff, z/s ~ 1e18);
//   * the jerk coupling w_dddl*(ddl[i+1]-ddl[i])^2 is split off with an auxiliary
//     equality row, which makes K = A S^-1 A' block tridiagonal (3x3 blocks):
//     cvxopt's dense `chol2` becomes a block-Cholesky chain passed lane to lane
//     with warp shuffles (forward/backward substitution likewise).
// s, z are carried unscaled (mathematically identical to cvxopt's Nesterov-Todd
// scaled iterates for the non-negative orthant).
#include <atomic>
#include <cstdlib>
#include "emp_common.cuh"
#include "sl_kernels.h"

namespace emp {

namespace {

struct FlagPtrs { unsigned long long *p[EMP_MAX_PEERS + 1]; };

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long globaltimer_ns() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

constexpr int NR = 14;          // inequality rows per station: 8 vehicle-corner rows + 3 ub + 3 lb
constexpr int QP_SMEM_ROWS = 4 * NR + 3;   // per warp: 1/s, 1/z, rz, ds o dz per row, and the previous iterate
#define EMP_UNROLL _Pragma("unroll")

struct Geo {                    // per-launch constants of the dynamics / constraint geometry
    double d1, d2;              // host_d1, host_d2 (levers of the corner rows, :117-124)
    double a, b, e, g0;         // ds, ds^2/3, ds/2, ds^2/6   (:109-110)
};

// G x for one station (row order: 8 corner rows :117-124, then x <= ub, then -x <= -lb :163-166)
__device__ __forceinline__ void gmul(const double x[3], const Geo &G, double o[NR]) {
    const double pa = x[0] + G.d1 * x[1];
    const double pb = x[0] - G.d2 * x[1];
    o[0] = pa; o[1] = pa; o[2] = pb; o[3] = pb;
    o[4] = -pa; o[5] = -pa; o[6] = -pb; o[7] = -pb;
    o[8] = x[0]; o[9] = x[1]; o[10] = x[2];
    o[11] = -x[0]; o[12] = -x[1]; o[13] = -x[2];
}
// G' v for one station
__device__ __forceinline__ void gtmul(const double v[NR], const Geo &G, double o[3]) {
    const double pa = (v[0] + v[1]) - (v[4] + v[5]);
    const double pb = (v[2] + v[3]) - (v[6] + v[7]);
    o[0] = (pa + pb) + (v[8] - v[11]);
    o[1] = (G.d1 * pa - G.d2 * pb) + (v[9] - v[12]);
    o[2] = v[10] - v[13];
}

// 1/x and 1/sqrt(x) for positive, normal-range x: hardware seed + two Newton steps (no special-case paths).
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

struct Factor {                 // per-lane pieces of the KKT factorisation
    double q00, q01, q11, q22;  // Q upper triangular, S_i^-1 = Q Q'
    double l10, l20, l21;       // strict lower part of the Cholesky factor L of this stage's pivot block
    double i00, i11, i22;       // 1 / diag(L)
    double xl[9], xr[9];        // XL = L^-1 K(i, i-d), XR = L^-1 K(i, i+d) at the level the stage was eliminated
};

// Block cyclic reduction (nested-dissection Cholesky) of the block-tridiagonal K: at level d = 1,2,4,8,16 the
// stages with (lane mod 2d) == d-1 are eliminated against their current neighbours lane-d and lane+d, which
// stay with (lane mod 2d) == 2d-1.  Five levels instead of n-1 sequential steps, every level lane-parallel.
#define EMP_CR_LEVELS 5
#ifdef EMP_QP_UNROLL_LEVELS
#define EMP_LEVEL_UNROLL _Pragma("unroll")
#else
#define EMP_LEVEL_UNROLL _Pragma("unroll 1")
#endif

#ifndef EMP_QP_MINBLOCKS
#define EMP_QP_MINBLOCKS 2
#endif

// One frame, one warp.  Everything below is warp-local (registers, shuffles, this warp's slice of shared memory),
// so the warps of a CTA are independent workers.
template <int NL>
__device__ __forceinline__ void
qp_one_frame(const int frame, const int lane, const double *__restrict__ scenes, double *__restrict__ plans,
             double *__restrict__ lmin_lmax, const double *__restrict__ lmin_in, const double *__restrict__ lmax_in,
             const double *__restrict__ start_in, double *__restrict__ out_x, int32_t *__restrict__ status_iters,
             const QpLaunch &p) {
    constexpr int LEVELS = NL == 32 ? 5 : (NL == 64 ? 6 : 7);   // cyclic-reduction levels = log2(lanes)
    // shared memory: per frame QP_SMEM_ROWS rows of NL doubles ([row][lane]); the warp path has QP_WARPS frames per CTA,
    // the block path one frame plus the two exchange lines of Lanes<NL>
    extern __shared__ double qp_smem_all[];
    double *qp_smem = qp_smem_all + (NL == 32 ? (size_t)(threadIdx.x >> 5) * (QP_SMEM_ROWS * 32) : (size_t)0);
    Lanes<NL> L(lane, qp_smem_all + QP_SMEM_ROWS * NL);
    const EmpQpParams &Q = p.q;

    // ---------------- inputs ---------------------------------------------------
    int n;
    double lmin = -10.0, lmax = 10.0;             // path_planning.py:233-234
    double st[3] = {0.0, 0.0, 0.0};
    int qp_status = EMP_QP_OK;
    double *plan = nullptr;
    if (p.mode == QP_MODE_DIRECT) {
        n = p.n_direct;
        if (n > NL || n < 2) qp_status = EMP_QP_TOO_LARGE;
        if (lane < n && qp_status == EMP_QP_OK) {
            lmin = lmin_in[(size_t)frame * n + lane];
            lmax = lmax_in[(size_t)frame * n + lane];
        }
        st[0] = start_in[(size_t)frame * 3 + 0];
        st[1] = start_in[(size_t)frame * 3 + 1];
        st[2] = start_in[(size_t)frame * 3 + 2];
    } else {
        const double *sc = scenes + (size_t)frame * p.scene_stride;
        const int NE = p.n_enr_max;
        double dps = 0.0, dpl = 0.0;
        int nq_out;
        if (p.mode == QP_MODE_FUSED) {
            plan = plans + (size_t)frame * p.plan_stride;
            const int n_enr = (int)plan[0];
            n = (n_enr + 1) >> 1;                  // dp_path[::2]  (test_9.py:187-188)
            nq_out = (NE + 1) / 2;
            st[0] = sc[1];
            st[1] = sc[2];
            st[2] = sc[3];
        } else {                                   // QP_MODE_CORRIDOR: path given by the caller
            n = p.n_direct;
            nq_out = n;
        }
        if (n > NL || n < (p.mode == QP_MODE_FUSED ? 2 : 1)) qp_status = EMP_QP_TOO_LARGE;
        const bool valid0 = lane < n && qp_status == EMP_QP_OK;
        if (valid0) {
            if (p.mode == QP_MODE_FUSED) {
                dps = plan[EMP_PLAN_HDR + 2 * lane];
                dpl = plan[EMP_PLAN_HDR + NE + 2 * lane];
            } else {
                dps = lmin_in[(size_t)frame * n + lane];
                dpl = lmax_in[(size_t)frame * n + lane];
            }
        }
        // ---- cal_lmin_lmax (:222-273)
        int K = (int)sc[4];
        K = K < 0 ? 0 : (K > p.kmax ? p.kmax : K);
        const double INF = __longlong_as_double(0x7ff0000000000000LL);
        if (qp_status == EMP_QP_OK) {
            for (int o = 0; o < K; ++o) {
                const double os = sc[EMP_SCENE_HDR + o];
                const double ol = sc[EMP_SCENE_HDR + p.kmax + o];
                const double smin = os - Q.obs_length / 2;
                const double smax = os + Q.obs_length / 2;
                // np.argmin of an all-NaN array (NaN obstacle) is index 0
                const bool os_nan = isnan(os);
                const int lo = (os_nan ? 0 : L.argmin(valid0 ? fabs(dps - smin) : INF)) + 2;   // offset = 2 (:235,:240)
                const int hi = (os_nan ? 0 : L.argmin(valid0 ? fabs(dps - smax) : INF)) + 2;
                const int ctr = os_nan ? 0 : L.argmin(valid0 ? fabs(dps - os) : INF);
                const double path_l = L.get(dpl, ctr);
                if (lo <= hi && hi >= n) {         // the reference raises IndexError here (:265-272)
                    qp_status = EMP_QP_CORRIDOR_INDEX;
                    break;
                }
                if (lane >= lo && lane <= hi) {
                    if (path_l < ol) lmax = fmin(lmax, ol - Q.obs_width / 2);
                    else lmin = fmax(lmin, ol + Q.obs_width / 2);
                }
            }
        }
        if (lmin_lmax != nullptr && lane < nq_out) {
            lmin_lmax[((size_t)frame * 2 + 0) * nq_out + lane] = lane < n ? lmin : 0.0;
            lmin_lmax[((size_t)frame * 2 + 1) * nq_out + lane] = lane < n ? lmax : 0.0;
        }
        if (p.mode == QP_MODE_CORRIDOR) {
            if (lane == 0) {
                status_iters[2 * frame] = qp_status;
                status_iters[2 * frame + 1] = 0;
            }
            return;
        }
    }

    int iters = 0;
    double gap = 0.0;
    double x[3] = {0.0, 0.0, 0.0};

    if (qp_status == EMP_QP_OK) {
        const bool valid = lane < n;
        const bool stage = lane < n - 1;           // lane i owns equality pair (i, i+1)
        const int m_rows = NR * n;
        // ---------------- problem data (Quadratic_planning :103-205) -----------
        Geo G;
        G.d1 = Q.host_d1;
        G.d2 = Q.host_d2;
        const double ds = Q.dp_sampling_res;
        G.a = ds;
        G.b = ds * ds / 3;
        G.e = ds / 2;
        G.g0 = ds * ds / 6;
        const int fi = (int)ceil(Q.host_d1 / ds), bi = (int)ceil(Q.host_d2 / ds);
        const double hU = L.get(lmax, min(lane + fi, n - 1));     // l_max[index1] (:130)
        const double hL = L.get(lmin, max(lane - bi, 0));         // l_min[index2] (:131)
        const double hw = Q.host_w / 2;
        // h is recomputed from 5 registers where it is needed (:133-166): rows 0-7 from the corridor, rows 8-13
        // from the +-1e5 boxes, pinned to the plan start on lane 0 (:148-153) and to 0 on lane n-1 (:155-160)
        const double hq[4] = {hU - hw, hU + hw, -hL + hw, -hL - hw};
        double ubv[3];
        EMP_UNROLL
        for (int c = 0; c < 3; ++c) ubv[c] = lane == 0 ? st[c] : (lane == n - 1 ? 0.0 : 100000.0);
        const bool interior = lane != 0 && lane != n - 1;
        auto hrow = [&](int r) -> double {
            if (r < 8) return hq[((r >> 2) << 1) | (r & 1)];
            if (r < 11) return ubv[r - 8];
            return interior ? 100000.0 : -ubv[r - 11];
        };
        // P (diagonal part) with and without the jerk terms; cc = coupling ddl_i <-> ddl_{i+1}
        double Pd[3], Pn[3];
        Pd[0] = 2 * (Q.w_l + Q.w_centre);
        Pd[1] = 0.0;                                             // w_cost_dl*(H_DL.T@H_L) == 0 (:193)
        Pn[2] = 2 * Q.w_ddl;
        if (lane == n - 1) {
            Pd[0] += 2 * Q.w_end_l;
            Pd[1] += 2 * Q.w_end_dl;
            Pn[2] += 2 * Q.w_end_ddl;
        }
        Pn[0] = Pd[0];
        Pn[1] = Pd[1];
        Pd[2] = Pn[2] + 2 * Q.w_dddl * ((lane == 0 || lane == n - 1) ? 1.0 : 2.0);
        const double cc = -2 * Q.w_dddl;
        const bool nojerk = !(Q.w_dddl > 0.0);
        const double jinv = nojerk ? 1.0 : 1.0 / (2 * Q.w_dddl);
        const double q0 = valid ? Q.w_centre * (-2 * ((lmin + lmax) / 2)) : 0.0;   // :199-205

        // ---------------- helpers ------------------------------------------------
        Factor F;
        bool broke = false;

        // factor(w): per-lane S^-1 = QQ', then the block-Cholesky chain of K
        auto factor = [&](const double w[NR]) {
            const double wa = (w[0] + w[1]) + (w[4] + w[5]);
            const double wb = (w[2] + w[3]) + (w[6] + w[7]);
            const double D0 = Pn[0] + (w[8] + w[11]);
            const double D1 = Pn[1] + (w[9] + w[12]);
            const double s22 = Pn[2] + (w[10] + w[13]);
            const double s01 = G.d1 * wa - G.d2 * wb;
            const double r00sq = D0 + (wa + wb);
            const double dsum = G.d1 + G.d2;
            const double r11sq = D1 + (D0 * (G.d1 * G.d1 * wa + G.d2 * G.d2 * wb) + (wa * wb) * (dsum * dsum)) / r00sq;
            const double ir00 = frsqrt(r00sq);
            const double ir11 = frsqrt(r11sq);
            F.q00 = ir00;
            F.q11 = ir11;
            F.q01 = -(s01 * ir00) * (ir00 * ir11);
            F.q22 = frsqrt(s22);
            if (!valid) { F.q00 = 1.0; F.q11 = 1.0; F.q01 = 0.0; F.q22 = 1.0; }
            // U = E Q, V = F Q
            const double u01 = F.q01 + G.a * F.q11, u02 = G.b * F.q22, u12 = G.e * F.q22, v02 = G.g0 * F.q22;
            double e00 = F.q00 * F.q00 + u01 * u01 + u02 * u02;
            double e10 = u01 * F.q11 + u02 * u12;
            double e11 = F.q11 * F.q11 + u12 * u12;
            double e20 = u02 * F.q22, e21 = u12 * F.q22, e22 = F.q22 * F.q22;
            double f00 = F.q00 * F.q00 + F.q01 * F.q01 + v02 * v02;
            double f10 = F.q01 * F.q11 + v02 * u12;
            double f11 = e11;
            double f20 = -v02 * F.q22, f21 = -u12 * F.q22, f22 = e22;
            double B[9];                                           // F X_i E' = V U'
            B[0] = (v02 * u02 - F.q00 * F.q00) - F.q01 * u01;
            B[1] = v02 * u12 - F.q01 * F.q11;
            B[2] = v02 * F.q22;
            B[3] = u12 * u02 - F.q11 * u01;
            B[4] = u12 * u12 - F.q11 * F.q11;
            B[5] = u12 * F.q22;
            B[6] = -F.q22 * u02;
            B[7] = -F.q22 * u12;
            B[8] = -e22;
            if (nojerk) {
                e20 = e21 = 0.0; f20 = f21 = 0.0; e22 = 0.0; f22 = 0.0;
                B[2] = B[5] = B[6] = B[7] = B[8] = 0.0;
            }
            // K_ii = E X_i E' + F X_{i+1} F' + diag(0,0,1/(2 w_dddl))
            const double k00 = e00 + L.down(f00, 1);
            const double k10 = e10 + L.down(f10, 1);
            const double k11 = e11 + L.down(f11, 1);
            const double k20 = e20 + L.down(f20, 1);
            const double k21 = e21 + L.down(f21, 1);
            const double k22 = (e22 + L.down(f22, 1)) + jinv;
            // pivot block D, couplings CL = K(i, i-1) = B', CR = K(i, i+1) = B of the next lane
            double D00 = k00, D10 = k10, D11 = k11, D20 = k20, D21 = k21, D22 = k22;
            double CL[9], CR[9];
            EMP_UNROLL
            for (int r = 0; r < 3; ++r) {
                EMP_UNROLL
                for (int c = 0; c < 3; ++c) {
                    CL[3 * r + c] = (stage && lane > 0) ? B[3 * c + r] : 0.0;
                    const double nb = L.down(B[3 * r + c], 1);
                    CR[3 * r + c] = (lane < n - 2) ? nb : 0.0;
                }
            }
            if (!stage) { D00 = 1.0; D10 = 0.0; D11 = 1.0; D20 = 0.0; D21 = 0.0; D22 = 1.0; }
            bool bad = false;
            EMP_UNROLL
            for (int k = 0; k < 9; ++k) { F.xl[k] = 0.0; F.xr[k] = 0.0; }
EMP_LEVEL_UNROLL
            for (int lev = 0; lev < LEVELS; ++lev) {
                const int d = 1 << lev, msk = 2 * d - 1;
                const bool el = (lane & msk) == d - 1;
                if (el) {
                    const double j00 = frsqrt(D00);
                    const double c10 = D10 * j00, c20 = D20 * j00;
                    const double t11 = D11 - c10 * c10;
                    const double j11 = frsqrt(t11);
                    const double c21 = (D21 - c20 * c10) * j11;
                    const double t22 = (D22 - c20 * c20) - c21 * c21;
                    const double j22 = frsqrt(t22);
                    bad = bad || (stage && (!(D00 > 0.0) || !(t11 > 0.0) || !(t22 > 0.0) || !isfinite(j00 + j11 + j22)));
                    F.l10 = c10; F.l20 = c20; F.l21 = c21;
                    F.i00 = j00; F.i11 = j11; F.i22 = j22;
                    EMP_UNROLL
                    for (int c = 0; c < 3; ++c) {                 // X = L^-1 C, column c
                        const double a0 = CL[c] * j00;
                        const double a1 = (CL[3 + c] - c10 * a0) * j11;
                        F.xl[c] = a0; F.xl[3 + c] = a1; F.xl[6 + c] = ((CL[6 + c] - c20 * a0) - c21 * a1) * j22;
                        const double b0 = CR[c] * j00;
                        const double b1 = (CR[3 + c] - c10 * b0) * j11;
                        F.xr[c] = b0; F.xr[3 + c] = b1; F.xr[6 + c] = ((CR[6 + c] - c20 * b0) - c21 * b1) * j22;
                    }
                }
                // Products of the eliminated stages, handed to the remaining neighbours element by element.
                // No lane predicates are needed: a lane with t trailing one-bits is "remaining" at levels < t and
                // eliminated at level t, so its D/CL/CR are dead afterwards and may be overwritten with garbage;
                // remaining lanes only read from lanes eliminated at this level, and a missing left/right
                // neighbour shows up as an exactly-zero coupling block (lane 31 is never a stage).
                // The last level eliminates lane 15 alone: its neighbours (lane -1, lane 31) do not exist, so its
                // coupling blocks are exactly zero and there is nobody to hand products to.
                if (lev == LEVELS - 1) break;
                EMP_UNROLL
                for (int r = 0; r < 3; ++r) {
                    EMP_UNROLL
                    for (int c = 0; c < 3; ++c) {
                        const double plr = (F.xl[r] * F.xr[c] + F.xl[3 + r] * F.xr[3 + c]) + F.xl[6 + r] * F.xr[6 + c];
                        CL[3 * c + r] = -L.up(plr, d);         // K(j, j-2d) = -(XL'XR of lane j-d)'
                        CR[3 * r + c] = -L.down(plr, d);       // K(j, j+2d) = -(XL'XR of lane j+d)
                        if (c <= r) {
                            const double prr = (F.xr[r] * F.xr[c] + F.xr[3 + r] * F.xr[3 + c]) + F.xr[6 + r] * F.xr[6 + c];
                            const double pll = (F.xl[r] * F.xl[c] + F.xl[3 + r] * F.xl[3 + c]) + F.xl[6 + r] * F.xl[6 + c];
                            const double upd = L.up(prr, d) + L.down(pll, d);   // (XR'XR)(j-d) + (XL'XL)(j+d)
                            if (r == 0) D00 -= upd;
                            else if (r == 1 && c == 0) D10 -= upd;
                            else if (r == 1) D11 -= upd;
                            else if (c == 0) D20 -= upd;
                            else if (c == 1) D21 -= upd;
                            else D22 -= upd;
                        }
                    }
                }
            }
            broke = L.any(bad);
        };

        // solve [S A'; A 0][dx; dy] = [g; by] with the current factorisation
        auto kkt_solve = [&](const double g[3], const double by[2], double dx[3], double dy[2]) {
            // X g = Q (Q' g)
            auto applyX = [&](const double v[3], double o[3]) {
                const double t0 = F.q00 * v[0];
                const double t1 = F.q01 * v[0] + F.q11 * v[1];
                const double t2 = F.q22 * v[2];
                o[0] = F.q00 * t0 + F.q01 * t1;
                o[1] = F.q11 * t1;
                o[2] = F.q22 * t2;
            };
            double xg[3];
            applyX(g, xg);
            // r_i = E Xg_i + F Xg_{i+1} - (by, 0)
            const double fx0 = -xg[0] + G.g0 * xg[2];
            const double fx1 = -xg[1] + G.e * xg[2];
            const double fx2 = -xg[2];
            double r0 = ((xg[0] + G.a * xg[1]) + G.b * xg[2]) + L.down(fx0, 1) - by[0];
            double r1 = (xg[1] + G.e * xg[2]) + L.down(fx1, 1) - by[1];
            double r2 = nojerk ? 0.0 : xg[2] + L.down(fx2, 1);
            // forward elimination, level by level: t_i = L_i^-1 r_i, then r_{i-d} -= XL' t_i, r_{i+d} -= XR' t_i
            if (!stage) { r0 = 0.0; r1 = 0.0; r2 = 0.0; }
            double t0 = 0.0, t1 = 0.0, t2 = 0.0;
EMP_LEVEL_UNROLL
            for (int lev = 0; lev < LEVELS; ++lev) {
                const int d = 1 << lev, msk = 2 * d - 1;
                const bool el = (lane & msk) == d - 1;
                double ul0 = 0.0, ul1 = 0.0, ul2 = 0.0, ur0 = 0.0, ur1 = 0.0, ur2 = 0.0;
                if (el) {
                    t0 = r0 * F.i00;
                    t1 = (r1 - F.l10 * t0) * F.i11;
                    t2 = ((r2 - F.l20 * t0) - F.l21 * t1) * F.i22;
                    ul0 = (F.xl[0] * t0 + F.xl[3] * t1) + F.xl[6] * t2;
                    ul1 = (F.xl[1] * t0 + F.xl[4] * t1) + F.xl[7] * t2;
                    ul2 = (F.xl[2] * t0 + F.xl[5] * t1) + F.xl[8] * t2;
                    ur0 = (F.xr[0] * t0 + F.xr[3] * t1) + F.xr[6] * t2;
                    ur1 = (F.xr[1] * t0 + F.xr[4] * t1) + F.xr[7] * t2;
                    ur2 = (F.xr[2] * t0 + F.xr[5] * t1) + F.xr[8] * t2;
                }
                // remaining lanes: r -= XR'(j-d) t(j-d) + XL'(j+d) t(j+d)   (dead values elsewhere, see factor);
                // nobody remains after the last level (lane 15 alone, zero coupling blocks)
                if (lev == LEVELS - 1) break;
                r0 -= L.up(ur0, d) + L.down(ul0, d);
                r1 -= L.up(ur1, d) + L.down(ul1, d);
                r2 -= L.up(ur2, d) + L.down(ul2, d);
            }
            // back substitution, top level first: y_i = L_i^-T (t_i - XL y_{i-d} - XR y_{i+d})
            double y0 = 0.0, y1 = 0.0, y2 = 0.0;
EMP_LEVEL_UNROLL
            for (int lev = LEVELS - 1; lev >= 0; --lev) {
                const int d = 1 << lev, msk = 2 * d - 1;
                const bool el = (lane & msk) == d - 1;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
                if (lev < LEVELS - 1) {                   // the top level has no eliminated neighbours yet (y = 0)
                    a0 = L.up(y0, d); a1 = L.up(y1, d); a2 = L.up(y2, d);
                    b0 = L.down(y0, d); b1 = L.down(y1, d); b2 = L.down(y2, d);
                }
                if (el) {
                    const double v0 = t0 - (((F.xl[0] * a0 + F.xl[1] * a1) + F.xl[2] * a2) + ((F.xr[0] * b0 + F.xr[1] * b1) + F.xr[2] * b2));
                    const double v1 = t1 - (((F.xl[3] * a0 + F.xl[4] * a1) + F.xl[5] * a2) + ((F.xr[3] * b0 + F.xr[4] * b1) + F.xr[5] * b2));
                    const double v2 = t2 - (((F.xl[6] * a0 + F.xl[7] * a1) + F.xl[8] * a2) + ((F.xr[6] * b0 + F.xr[7] * b1) + F.xr[8] * b2));
                    y2 = v2 * F.i22;
                    y1 = (v1 - F.l21 * y2) * F.i11;
                    y0 = ((v0 - F.l10 * y1) - F.l20 * y2) * F.i00;
                }
            }
            if (!stage) { y0 = 0.0; y1 = 0.0; y2 = 0.0; }
            // x = X (g - E' y_i - F' y_{i-1})
            double pu0 = L.up(y0, 1), pu1 = L.up(y1, 1), pu2 = L.up(y2, 1);
            if (lane == 0) { pu0 = 0.0; pu1 = 0.0; pu2 = 0.0; }
            double v[3];
            v[0] = g[0] - (y0 - pu0);
            v[1] = g[1] - ((G.a * y0 + y1) - pu1);
            v[2] = g[2] - (((G.b * y0 + G.e * y1) + y2) + ((G.g0 * pu0 + G.e * pu1) - pu2));
            applyX(v, dx);
            dy[0] = y0;
            dy[1] = y1;
        };

        // P x (with the jerk coupling), A x, A' y
        auto pmul = [&](const double xx[3], double o[3]) {
            double up = L.up(xx[2], 1), dn = L.down(xx[2], 1);
            if (lane == 0) up = 0.0;
            if (lane >= n - 1) dn = 0.0;
            o[0] = Pd[0] * xx[0];
            o[1] = Pd[1] * xx[1];
            o[2] = Pd[2] * xx[2] + cc * (up + dn);
        };
        auto amul = [&](const double xx[3], double o[2]) {
            const double n0 = L.down(xx[0], 1), n1 = L.down(xx[1], 1), n2 = L.down(xx[2], 1);
            o[0] = stage ? (((xx[0] + G.a * xx[1]) + G.b * xx[2]) - n0) + G.g0 * n2 : 0.0;
            o[1] = stage ? ((xx[1] + G.e * xx[2]) - n1) + G.e * n2 : 0.0;
        };
        auto atmul = [&](const double yy[2], double o[3]) {
            double p0 = L.up(yy[0], 1), p1 = L.up(yy[1], 1);
            if (lane == 0) { p0 = 0.0; p1 = 0.0; }
            o[0] = yy[0] - p0;
            o[1] = (G.a * yy[0] + yy[1]) - p1;
            o[2] = (G.b * yy[0] + G.e * yy[1]) + (G.g0 * p0 + G.e * p1);
        };

        // ---------------- coneqp iteration -----------------------------------------
        // One loop body serves the initial point (it == -1: W = I, one solve) and the Mehrotra iterations
        // (two solves), so factor() and kkt_solve() are instantiated once (code size: instruction cache).
        // Per-row state: s, z live in registers; 1/s, 1/z, rz and the predictor's ds o dz live in shared
        // memory ([row][lane], conflict-free); everything else per row is recomputed where it is used.
        double *sm_is = qp_smem + lane;                            // this frame's rows, [row][lane]
        double *sm_iz = sm_is + NR * NL, *sm_rz = sm_iz + NR * NL, *sm_ws = sm_rz + NR * NL;
        double s[NR], z[NR], y[2] = {0.0, 0.0};
        double *sm_xp = sm_ws + NR * NL;                           // iterate before the last step (divergence exit)
        double rx[3] = {0.0, 0.0, 0.0}, ry[2] = {0.0, 0.0};
        EMP_UNROLL
        for (int r = 0; r < NR; ++r) { s[r] = 1.0; z[r] = 1.0; sm_ws[r * NL] = 0.0; sm_rz[r * NL] = 0.0; }
        double resx0 = 1.0, resz0 = 1.0, gap_min = 0.0;
        const double resy0 = 1.0;                                  // b = 0
        int it = -1;
        while (true) {
            if (it >= 0) {
                double pxq[3], aty[3], gtz[3], gx[NR];
                pmul(x, pxq);
                pxq[0] += q0;
                atmul(y, aty);
                {
                    double zt[NR];
                    EMP_UNROLL
                    for (int r = 0; r < NR; ++r) zt[r] = z[r];
                    gtmul(zt, G, gtz);
                }
                amul(x, ry);
                gmul(x, G, gx);
                double red[7] = {0, 0, 0, 0, 0, 0, 0};
                EMP_UNROLL
                for (int c = 0; c < 3; ++c) {
                    rx[c] = valid ? (pxq[c] + aty[c]) + gtz[c] : 0.0;
                    red[0] += rx[c] * rx[c];
                    red[3] += valid ? x[c] * pxq[c] : 0.0;
                }
                red[4] = valid ? x[0] * q0 : 0.0;
                red[1] = ry[0] * ry[0] + ry[1] * ry[1];
                red[5] = y[0] * ry[0] + y[1] * ry[1];
                EMP_UNROLL
                for (int r = 0; r < NR; ++r) {
                    const double rzr = valid ? (s[r] + gx[r]) - hrow(r) : 0.0;
                    sm_rz[r * NL] = rzr;
                    red[2] += rzr * rzr;
                    red[6] += valid ? z[r] * rzr : 0.0;
                }
                L.sum_n(red);
                const double resx = sqrt(red[0]), resy = sqrt(red[1]), resz = sqrt(red[2]);
                const double pcost = 0.5 * (red[3] + red[4]);
                const double dcost = ((pcost + red[5]) + red[6]) - gap;
                bool have_rel = true;
                double relgap = 0.0;
                if (pcost < 0.0) relgap = gap / -pcost;
                else if (dcost > 0.0) relgap = gap / dcost;
                else have_rel = false;
                const double pres = fmax(resy / resy0, resz / resz0);
                const double dres = resx / resx0;
                if (pres <= Q.feastol && dres <= Q.feastol && (gap <= Q.abstol || (have_rel && relgap <= Q.reltol))) break;
                if (it == Q.maxiters) { qp_status = EMP_QP_UNKNOWN_MAXITER; break; }
                if (!isfinite((gap + resx) + resz)) { qp_status = EMP_QP_UNKNOWN_NUMERIC; break; }
                // Divergence exit (include/emplanner.h, EmpQpParams.diverge_factor): on an infeasible corridor coneqp
                // has no certificate; cvxopt iterates until its Cholesky breaks down or maxiters and reports 'unknown'.
                // The gap of a converging solve decreases monotonically, so the first clear increase ends the solve.
                if (it > 0 && Q.diverge_factor > 0.0 && gap > Q.diverge_factor * gap_min) {
                    qp_status = EMP_QP_UNKNOWN_DIVERGED;           // the step that blew the gap up is discarded
                    x[0] = sm_xp[0]; x[1] = sm_xp[NL]; x[2] = sm_xp[2 * NL];
                    break;
                }
                gap_min = fmin(gap_min, gap);
                sm_xp[0] = x[0]; sm_xp[NL] = x[1]; sm_xp[2 * NL] = x[2];
            }

            {
                double w[NR];
                EMP_UNROLL
                for (int r = 0; r < NR; ++r) {
                    const double isr = frcp(s[r]);
                    sm_is[r * NL] = isr;
                    sm_iz[r * NL] = frcp(z[r]);
                    w[r] = z[r] * isr;
                }
                factor(w);
            }
            if (broke) { qp_status = EMP_QP_UNKNOWN_NUMERIC; break; }   // cvxopt: ValueError (init) / ArithmeticError

            const double mu = gap / (double)m_rows;
            const int npass = it < 0 ? 1 : 2;
            double sigma = 0.0, step = 1.0;
            double dx[3], dy[2];
#pragma unroll 1
            for (int pass = 0; pass < npass; ++pass) {
                double g[3], by[2];
                if (it < 0) {
                    double gth[3], hh[NR];
                    EMP_UNROLL
                    for (int r = 0; r < NR; ++r) hh[r] = hrow(r);
                    gtmul(hh, G, gth);                             // rhs of the initial point: -q + G'h, b = 0
                    g[0] = valid ? -q0 + gth[0] : 0.0;
                    g[1] = valid ? gth[1] : 0.0;
                    g[2] = valid ? gth[2] : 0.0;
                    by[0] = 0.0;
                    by[1] = 0.0;
                } else {
                    double tq[NR], gtq[3];
                    const double sm_ = sigma * mu;
                    EMP_UNROLL
                    for (int r = 0; r < NR; ++r) {
                        const double sr_ = s[r], zr_ = z[r];
                        const double bs = (sm_ - sr_ * zr_) - (pass == 1 ? sm_ws[r * NL] : 0.0);
                        tq[r] = valid ? (bs + zr_ * sm_rz[r * NL]) * sm_is[r * NL] : 0.0;
                    }
                    gtmul(tq, G, gtq);
                    EMP_UNROLL
                    for (int c = 0; c < 3; ++c) g[c] = -rx[c] - gtq[c];
                    by[0] = -ry[0];
                    by[1] = -ry[1];
                }
                kkt_solve(g, by, dx, dy);
                if (it < 0) break;
                double gdx[NR];
                gmul(dx, G, gdx);
                double dsdz = 0.0, ts = -1e300, tz = -1e300;
                const double sm_ = sigma * mu;
                EMP_UNROLL
                for (int r = 0; r < NR; ++r) {
                    const double sr_ = s[r], zr_ = z[r];
                    const double isr = sm_is[r * NL], rzr = sm_rz[r * NL];
                    const double bs = (sm_ - sr_ * zr_) - (pass == 1 ? sm_ws[r * NL] : 0.0);
                    const double tqr = (bs + zr_ * rzr) * isr;
                    const double dzr = tqr + (zr_ * isr) * gdx[r];
                    const double dsr = -rzr - gdx[r];
                    if (valid) {
                        dsdz += dsr * dzr;
                        ts = fmax(ts, -dsr * isr);
                        tz = fmax(tz, -dzr * sm_iz[r * NL]);
                    }
                    if (pass == 0) sm_ws[r * NL] = dsr * dzr;
                }
                dsdz = L.sum(dsdz);
                const double t = fmax(0.0, L.max(fmax(ts, tz)));
                if (t == 0.0) step = 1.0;
                else step = pass == 0 ? fmin(1.0, 1.0 / t) : fmin(1.0, 0.99 / t);
                if (pass == 0) {
                    const double sg = fmin(1.0, fmax(0.0, (1.0 - step) + (dsdz / gap) * (step * step)));
                    sigma = (sg * sg) * sg;
                }
            }
            if (it < 0) {
                // initial point (coneqp, W = I): x, y from the solve; z = Gx - h, s = -z, shifted into the cone
                x[0] = dx[0]; x[1] = dx[1]; x[2] = dx[2];
                y[0] = dy[0]; y[1] = dy[1];
                double gx[NR];
                gmul(x, G, gx);
                double red[3] = {0.0, valid ? q0 * q0 : 0.0, 0.0};
                double mn_s = -1e300, mn_z = -1e300;               // max(-s), max(-z)
                EMP_UNROLL
                for (int r = 0; r < NR; ++r) {
                    const double hr = hrow(r);
                    const double z0 = valid ? gx[r] - hr : 1.0;
                    const double s0 = valid ? -z0 : 1.0;
                    z[r] = z0;
                    s[r] = s0;
                    if (valid) {
                        red[0] += s0 * s0;
                        red[2] += hr * hr;
                        mn_s = fmax(mn_s, -s0);
                        mn_z = fmax(mn_z, -z0);
                    }
                }
                L.sum_n(red);
                const double nrm = sqrt(red[0]);                   // ||s|| == ||z||
                resx0 = fmax(1.0, sqrt(red[1]));
                resz0 = fmax(1.0, sqrt(red[2]));
                const double ts = L.max(mn_s), tz = L.max(mn_z);
                const double sh_s = ts >= -1e-8 * fmax(nrm, 1.0) ? 1.0 + ts : 0.0;
                const double sh_z = tz >= -1e-8 * fmax(nrm, 1.0) ? 1.0 + tz : 0.0;
                double gs = 0.0;
                EMP_UNROLL
                for (int r = 0; r < NR; ++r) {
                    if (valid) {
                        double sr_ = s[r], zr_ = z[r];
                        if (sh_s != 0.0) sr_ += sh_s;
                        if (sh_z != 0.0) zr_ += sh_z;
                        s[r] = sr_;
                        z[r] = zr_;
                        gs += sr_ * zr_;
                    }
                }
                gap = L.sum(gs);
                gap_min = gap;
            } else {
                // step: recompute ds, dz of the corrector pass row by row (cheaper than keeping 28 registers)
                double gdx[NR];
                gmul(dx, G, gdx);
                const double sm_ = sigma * mu;
                double gs = 0.0;
                EMP_UNROLL
                for (int r = 0; r < NR; ++r) {
                    double sr_ = s[r], zr_ = z[r];
                    const double isr = sm_is[r * NL], rzr = sm_rz[r * NL];
                    const double bs = (sm_ - sr_ * zr_) - sm_ws[r * NL];
                    const double tqr = (bs + zr_ * rzr) * isr;
                    const double dzr = tqr + (zr_ * isr) * gdx[r];
                    const double dsr = -rzr - gdx[r];
                    if (valid) {
                        sr_ += step * dsr;
                        zr_ += step * dzr;
                        s[r] = sr_;
                        z[r] = zr_;
                        gs += sr_ * zr_;
                    }
                }
                EMP_UNROLL
                for (int c = 0; c < 3; ++c) x[c] += step * dx[c];
                y[0] += step * dy[0];
                y[1] += step * dy[1];
                gap = L.sum(gs);
            }
            ++it;
        }
        iters = it < 0 ? 0 : it;
    }

    // ---------------- outputs ----------------------------------------------------
    if (plan != nullptr) {
        const int NE = p.n_enr_max, NQ = (NE + 1) / 2;
        double *o = plan + EMP_PLAN_HDR + 2 * NE;
        if (lane < n && lane < NQ) {
            o[lane] = x[0];
            o[NQ + lane] = x[1];
            o[2 * NQ + lane] = x[2];
        }
        if (lane == 0) {
            plan[1] = (double)n;
            plan[3] = (double)qp_status;
            plan[4] = (double)iters;
            plan[6] = gap;
        }
        if (p.n_peers > 0) {
            // Fused all-gather: push the finished plan record into every peer's gather buffer with 16-byte
            // stores over NVLink (peer pointers opened through CUDA IPC); the transfers of finished frames overlap
            // the solves still running on this GPU.  A stream-ordered barrier after the launch makes them visible.
            L.sync();
            const double2 *src = reinterpret_cast<const double2 *>(plan);
            const int words2 = p.plan_stride >> 1;
            for (int w = lane; w < words2; w += NL) {
                const double2 vv = src[w];
#pragma unroll 1
                for (int r = 0; r < p.n_peers; ++r) {
                    double2 *dst = reinterpret_cast<double2 *>(p.peer[r] + (size_t)frame * p.plan_stride);
                    dst[w] = vv;
                }
            }
        }
    } else {
        if (lane < n && n <= NL) {
            double *o = out_x + (size_t)frame * 3 * n;
            o[lane] = x[0];
            o[n + lane] = x[1];
            o[2 * n + lane] = x[2];
        }
        if (lane == 0) {
            status_iters[2 * frame] = qp_status;
            status_iters[2 * frame + 1] = iters;
        }
    }
}

// Frame -> warp mapping.  Default: static, warp w of the grid solves frame order[w] (heaviest frames first when the
// emp_sl_order_dev schedule is given), one frame per warp; CTAs retire one by one, so the next step's DP kernel (another
// stream) moves in while this launch drains - worth 2 % of the pipelined step over the queue below.
// Queue (`queue` != 0, EMP_QP_QUEUE=1): grid = resident CTAs (EMP_QP_MINBLOCKS per SM), every warp pulls the next frame
// from a ticket counter until the batch is exhausted.  Solve times differ by 2x between frames (6 .. 13 interior-point
// iterations) and 4096 frames are 3.46 waves of the 1184 resident warps; the queue ends the launch within one solve of
// the average (1-3 % faster launched alone), but its CTAs hold every register of the SM until the very end.
// ticket[0] = next frame, ticket[1] = warps that are done; the last one re-arms both for the next launch that is handed
// this slot (launch_sl_qp rotates slots, so concurrent launches never share one) and, in the fused all-gather,
// publishes the launch's epoch.
__global__ void __launch_bounds__(QP_WARPS * 32, EMP_QP_MINBLOCKS)
sl_qp_kernel(const double *__restrict__ scenes, double *__restrict__ plans, double *__restrict__ lmin_lmax,
             const double *__restrict__ lmin_in, const double *__restrict__ lmax_in,
             const double *__restrict__ start_in, double *__restrict__ out_x,
             int32_t *__restrict__ status_iters, unsigned *__restrict__ ticket, const int queue,
             const QpLaunch p) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (;;) {
        int frame = blockIdx.x * QP_WARPS + (threadIdx.x >> 5);    // static mapping
        if (queue) {
            unsigned t = 0;
            if (lane == 0) t = atomicAdd(ticket, 1u);
            frame = (int)__shfl_sync(EMP_FULL_MASK, t, 0);
        }
        if (frame >= p.batch) break;
        if (p.order != nullptr) frame = p.order[frame];           // heaviest frames first
        qp_one_frame<32>(frame, lane, scenes, plans, lmin_lmax, lmin_in, lmax_in, start_in, out_x, status_iters, p);
        __syncwarp();
        if (!queue) break;
    }
    if (ticket == nullptr) return;
    if (lane == 0) {
        const unsigned total = gridDim.x * QP_WARPS;
        // This warp's peer stores before its arrival below.  A device-scope fence is enough here: the arrival counter
        // synchronises the warps of this GPU, and the last warp's system-scope fence + release below is cumulative over
        // everything it has observed (PTX memory model: causality order composes across scopes) - the same shape as
        // "stores, __syncthreads(), one thread fences at system scope and signals".
        if (p.epoch != 0ull) __threadfence();
        if (atomicAdd(ticket + 1, 1u) == total - 1) {             // every other warp is past its last ticket read
            ticket[0] = 0u;
            ticket[1] = 0u;
            if (p.epoch != 0ull) {
                // Last warp of the launch: every plan record has been stored (here and in the peers' buffers) and
                // fenced at system scope before its warp arrived, so publishing the epoch after one more fence makes
                // the whole shard visible to whoever acquires the flag - no collective, no host round trip.
                __threadfence_system();
                for (int r = 0; r < p.n_peers; ++r) st_release_sys(p.peer_ready[r], p.epoch);
                st_release_sys(p.self_ready, p.epoch);
            }
        }
    }
}

// General-size path: one CTA of NL threads per frame (33 .. NL stations), static frame -> CTA mapping.
template <int NL>
__global__ void __launch_bounds__(NL, 1)
sl_qp_block_kernel(const double *__restrict__ scenes, double *__restrict__ plans, double *__restrict__ lmin_lmax,
                   const double *__restrict__ lmin_in, const double *__restrict__ lmax_in,
                   const double *__restrict__ start_in, double *__restrict__ out_x,
                   int32_t *__restrict__ status_iters, const QpLaunch p) {
    const int frame = p.order != nullptr ? p.order[blockIdx.x] : (int)blockIdx.x;
    qp_one_frame<NL>(frame, (int)threadIdx.x, scenes, plans, lmin_lmax, lmin_in, lmax_in, start_in, out_x, status_iters, p);
}

// Consumer / flow control of the fused all-gather: one thread per flag spins until it reaches `epoch`, then thread 0
// publishes `epoch` to the words in `signal` (the credit that lets the producers overwrite the buffer again).
__global__ void flag_wait_signal_kernel(const unsigned long long *flags, int n_wait, unsigned long long epoch,
                                        FlagPtrs sig, int n_signal, int *err, long long timeout_ns) {
    __shared__ int bad;
    if (threadIdx.x == 0) bad = 0;
    __syncthreads();
    if ((int)threadIdx.x < n_wait) {
        const long long t0 = globaltimer_ns();
        if (err && *reinterpret_cast<volatile int *>(err)) bad = 1;          // an earlier wait already gave up: do not wait again
        while (!bad && ld_acquire_sys(flags + threadIdx.x) < epoch) {
            __nanosleep(64);
            if (globaltimer_ns() - t0 > timeout_ns) {            // never hang the GPU on a lost peer
                bad = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (bad) {
            if (err) *err = 1;
        } else {
            __threadfence_system();
            for (int r = 0; r < n_signal; ++r) st_release_sys(sig.p[r], epoch);
        }
    }
}

// Ticket slots (zero-initialised, self re-arming).  One array per device (a __device__ variable is per device).
__device__ unsigned g_qp_tickets[2 * 256];

}  // namespace

template <int NL>
static cudaError_t launch_block(const double *scenes, double *plans, double *lmin_lmax, const double *lmin_in,
                                const double *lmax_in, const double *start_in, double *out_x, int32_t *status_iters,
                                const QpLaunch &p, cudaStream_t st) {
    const size_t smem = (size_t)(QP_SMEM_ROWS + 2) * NL * sizeof(double);
    cudaFuncSetAttribute(sl_qp_block_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sl_qp_block_kernel<NL><<<p.batch, NL, smem, st>>>(scenes, plans, lmin_lmax, lmin_in, lmax_in, start_in, out_x,
                                                     status_iters, p);
    return cudaGetLastError();
}

cudaError_t launch_sl_qp(const double *scenes, double *plans, double *lmin_lmax, const double *lmin_in,
                         const double *lmax_in, const double *start_in, double *out_x, int32_t *status_iters,
                         const QpLaunch &p, cudaStream_t st) {
    static std::atomic<unsigned> next_slot{0};
    static unsigned *ticket_base[64] = {nullptr};
    static int sm_count[64] = {0};
    static const bool use_queue = getenv("EMP_QP_QUEUE") != nullptr;         // tuning switch: persistent CTAs + ticket queue
    // stations per frame decide the path: <= 32 one warp per frame; <= 64 / <= 128 one CTA per frame
    const int n_max = p.mode == QP_MODE_FUSED ? p.n_fused_max : p.n_direct;
    if (n_max > 32) {
        if (p.n_peers > 0 || p.epoch != 0ull) return cudaErrorNotSupported;   // the fused gather is a warp-path feature
        if (n_max <= 64) return launch_block<64>(scenes, plans, lmin_lmax, lmin_in, lmax_in, start_in, out_x, status_iters, p, st);
        return launch_block<128>(scenes, plans, lmin_lmax, lmin_in, lmax_in, start_in, out_x, status_iters, p, st);
    }
    const size_t smem = (size_t)QP_WARPS * QP_SMEM_ROWS * 32 * sizeof(double);
    cudaFuncSetAttribute(sl_qp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int all = (p.batch + QP_WARPS - 1) / QP_WARPS;
    int dev = 0;
    cudaGetDevice(&dev);
    unsigned *ticket = nullptr;
    int grid = all, queue = 0;
    if ((use_queue || p.epoch != 0ull) && dev >= 0 && dev < 64) {
        if (ticket_base[dev] == nullptr) {
            void *sym = nullptr;
            if (cudaGetSymbolAddress(&sym, g_qp_tickets) != cudaSuccess) return cudaGetLastError();
            cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
            ticket_base[dev] = static_cast<unsigned *>(sym);
        }
        const int resident = sm_count[dev] * EMP_QP_MINBLOCKS;
        if (use_queue && all > resident) {                         // more frames than resident warps: queue them
            grid = resident;
            queue = 1;
        }
        if (queue || p.epoch != 0ull)                              // the arrival flags need the "last warp" count
            ticket = ticket_base[dev] + 2 * (next_slot.fetch_add(1) % 256u);
    }
    if (p.epoch != 0ull && ticket == nullptr) return cudaErrorNotSupported;
    sl_qp_kernel<<<grid, QP_WARPS * 32, smem, st>>>(scenes, plans, lmin_lmax, lmin_in, lmax_in, start_in, out_x,
                                                 status_iters, ticket, queue, p);
    return cudaGetLastError();
}

cudaError_t launch_flag_wait_signal(const unsigned long long *flags, int n_wait, unsigned long long epoch,
                                    unsigned long long *const *signal, int n_signal, int *err, int timeout_ms,
                                    cudaStream_t st) {
    FlagPtrs sig;
    for (int r = 0; r <= EMP_MAX_PEERS; ++r) sig.p[r] = r < n_signal ? signal[r] : nullptr;
    flag_wait_signal_kernel<<<1, 32, 0, st>>>(flags, n_wait, epoch, sig, n_signal, err, (long long)timeout_ms * 1000000ll);
    return cudaGetLastError();
}

}  // namespace emp
