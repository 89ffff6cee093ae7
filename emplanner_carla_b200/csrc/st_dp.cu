// K3 - S-T speed DP, one CTA (128 threads) per frame, one launch per batch.
//
// Replaces reference planner/speed_planning_test.py speed_DP (:101-188) with CalcDpCost (:191-231),
// CalcObsCost (:234-271), CalcCollisionCost (:274-284) and CalcSTCoordinate (:287-305): 40 non-uniform s rows
// (row 0 = largest s, :114,:303) x 16 t columns of 0.5 s.
//
// The reference draft cannot finish (its back-track indexes with a float array and aliases dp_speed_t to
// dp_speed_s, :122,:155-156,:182-184); the forward tables it builds before raising are the parity target
// (tests/golden/st_dp.npz) and the back-track here is the documented two-line repair (oracle/st_dp.py).
// Bug-compatible details kept on purpose:
//   * an edge whose start ROW is 0 is treated as starting from the DP origin (s=0, t=0, plan_start_s_dot) (:208-212);
//   * the 5 obstacle samples of an edge sit at t_start + (i-1)*dt, i = 0..4 (:251-252);
//   * the end-node scans use '<=' so the last minimum wins (:163,:169).
// Arithmetic follows oracle/st_dp.py's association order; compiled with -fmad=false.  w^y is evaluated as
// exp(y ln w) with ln w in double-double (a few ulp, like pow()), so costs are compared at 1e-12 relative,
// parents exactly.
//
// Work layout.  An edge costs 5 samples x K obstacles point-to-segment distances, but only ~10 % of those
// (sample, obstacle) pairs lie within 1.5 of the obstacle's S-T segment and cost anything.  One lane per edge
// with the distance code inline ran at 13 of 32 lanes active (ncu, profiles/r1_st_v1_summary.txt).  Now:
//   * per frame, for every (column, sample, obstacle) the s-interval in which a sample at that column's sample
//     time can be within 1.6 of the segment is tabulated once (the sample times of a column are shared by all
//     its edges) - the per-pair test is two comparisons;
//   * the pairs that pass are compacted into a per-warp queue (lane-major, positions from a warp scan of the
//     per-lane counts); the warp evaluates them 32 at a time with all lanes busy (reference arithmetic,
//     operands fetched from the owning lane by shuffle) and each owner adds its results in (sample, obstacle)
//     order, so the sums stay bit-identical.
#include <type_traits>
#include "emp_common.cuh"
#include "st_kernels.h"
#include "st_device.cuh"

namespace emp {
namespace {

constexpr int ST_WARPS = ST_THREADS / 32;
#ifndef EMP_ST_PRUNE_M
#define EMP_ST_PRUNE_M 1
#endif
#ifndef EMP_ST_PRUNE_ROUNDS
#define EMP_ST_PRUNE_ROUNDS 3
#endif
constexpr int PRUNE_M = EMP_ST_PRUNE_M;            // candidates per target row evaluated per bounding round
constexpr int PRUNE_ROUNDS = EMP_ST_PRUNE_ROUNDS;  // bounding rounds (the last one takes every candidate left under the bound)
constexpr double NEAR_R = 1.6;        // > 1.5: beyond it CalcCollisionCost returns exactly 0 (margin covers rounding)

struct Interval { double lo, hi; };

// s-range of the horizontal line t = T inside the capsule of radius NEAR_R around the segment (s1,t1)-(s2,t2)
__device__ Interval capsule_slice(double s1, double t1, double s2, double t2, double T) {
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    double lo = INF, hi = -INF;
    const double d1 = T - t1, d2 = T - t2, R2 = NEAR_R * NEAR_R;
    if (fabs(d1) < NEAR_R) { const double w = sqrt(R2 - d1 * d1); lo = fmin(lo, s1 - w); hi = fmax(hi, s1 + w); }
    if (fabs(d2) < NEAR_R) { const double w = sqrt(R2 - d2 * d2); lo = fmin(lo, s2 - w); hi = fmax(hi, s2 + w); }
    const double ds = s2 - s1, dtt = t2 - t1, len2 = ds * ds + dtt * dtt;
    if (dtt != 0.0) {
        const double sc = s1 + d1 * ds / dtt, hw = NEAR_R * sqrt(len2) / fabs(dtt);
        double a = sc - hw, b = sc + hw;
        bool ok = true;
        if (ds != 0.0) {
            const double u0 = s1 - d1 * dtt / ds, u1 = s1 + (len2 - d1 * dtt) / ds;
            a = fmax(a, fmin(u0, u1));
            b = fmin(b, fmax(u0, u1));
        } else {
            const double u = d1 / dtt;
            ok = u >= 0.0 && u <= 1.0;
        }
        if (ok && a <= b) { lo = fmin(lo, a); hi = fmax(hi, b); }
    } else if (fabs(d1) < NEAR_R) {
        lo = fmin(lo, fmin(s1, s2) - NEAR_R);
        hi = fmax(hi, fmax(s1, s2) + NEAR_R);
    }
    if (lo <= hi) { lo -= 1e-9 * (1.0 + fabs(lo)); hi += 1e-9 * (1.0 + fabs(hi)); }
    return Interval{lo, hi};
}

// CalcObsCost's per-sample, per-obstacle term (:253-269) in the reference's arithmetic order
__device__ __forceinline__ double point_cost(double s, double t, const Obs o, double w_obs, double lnw_hi, double lnw_lo) {
    const double v1s = o.s_in - s, v1t = o.t_in - t;
    const double v2s = o.s_out - s, v2t = o.t_out - t;
    const double v3s = v2s - v1s, v3t = v2t - v1t;
    const double d13 = v1s * v3s + v1t * v3t;
    const double d23 = v2s * v3s + v2t * v3t;
    const bool ends = (d13 > 0 && d23 > 0) || (d13 < 0 && d23 < 0);
    // sqrt is monotone and correctly rounded: min(sqrt a, sqrt b) = sqrt(min(a, b))
    const double num = ends ? fmin(v1s * v1s + v1t * v1t, v2s * v2s + v2t * v2t) : v3s * v3s + v3t * v3t;
    const double root = sqrt(num);
    const double dmin = ends ? root : fabs(v1s * v3t - v1t * v3s) / root;
    const double ad = fabs(dmin);
    if (ad < 0.5) return w_obs;
    if (0.5 < ad && ad < 1.5) {
        const double y = (0.5 - dmin) + 1;                 // w_obs ** y
        const double ph = y * lnw_hi;
        const double pl = fma(y, lnw_hi, -ph) + y * lnw_lo;
        const double e = exp(ph);
        return fma(e, pl, e);
    }
    return 0.0;
}

// Per-warp compaction queue: descriptors of every near (sample, obstacle) pair of the warp's 32 edges, lane-major,
// and the 32 values of the batch in flight.
struct WarpQueue {
    uint16_t *desc;            // [32 * NSAMP * kmax]: owner lane | sample << 5 | obstacle << 8
    double *val;               // [32]
};

// Obstacle cost (CalcObsCost :234-271) of one edge per lane, for a warp's 32 edges at once.  `tab` is warp-uniform.
// HALF: every edge of the warp spans exactly one lattice column (t_en - t_st == 0.5): the divisions by 0.5 and
// 0.125 become exact multiplications.
template <bool HALF>
__device__ __forceinline__ double warp_obs_cost(bool valid, double s_st, double t_st, double s_en, double t_en,
                                                const Interval *tab /* [NSAMP][kmax] */, int kmax, const Obs *obs,
                                                uint32_t col_obs /* obstacles with a non-empty interval in this column */,
                                                const WarpQueue &Q, double w_obs, double lnw_hi, double lnw_lo, int lane) {
    const double dt = HALF ? 0.125 : (t_en - t_st) / (NSAMP - 1);
    const double k = HALF ? (s_en - s_st) * 2.0 : (s_en - s_st) / (t_en - t_st);
    // 1. which pairs can cost anything: two comparisons against the column's interval table
    uint32_t near[NSAMP];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < NSAMP; ++i) {
        const double s = s_st + (k * (double)(i - 1)) * dt;
        uint32_t m = 0;
        for (uint32_t rest = col_obs; rest; rest &= rest - 1) {
            const int j = __ffs(rest) - 1;
            const Interval iv = tab[i * kmax + j];
            m |= (iv.lo <= s && s <= iv.hi ? 1u : 0u) << j;
        }
        near[i] = valid ? m : 0u;
        cnt += __popc(near[i]);
    }
    // 2. lane-major queue positions
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(EMP_FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(EMP_FULL_MASK, incl, 31);
    if (total == 0) return 0.0;
    const int start = incl - cnt;
    {
        int pos = start;
#pragma unroll
        for (int i = 0; i < NSAMP; ++i)
            for (uint32_t m = near[i]; m; m &= m - 1)
                Q.desc[pos++] = (uint16_t)(lane | (i << 5) | ((__ffs(m) - 1) << 8));
    }
    // 3. evaluate 32 pairs at a time with all lanes busy; owners add their results in (sample, obstacle) order
    double cost = 0.0;
    for (int b0 = 0; b0 < total; b0 += 32) {
        __syncwarp();
        const bool have = b0 + lane < total;
        const uint32_t d = have ? Q.desc[b0 + lane] : (uint32_t)lane;
        const int ol = (int)(d & 31u), i = (int)((d >> 5) & 7u), j = (int)(d >> 8);
        const double os = shfl(s_st, ol), ot = shfl(t_st, ol), oe = shfl(s_en, ol);
        if (have) {
            const double odt = HALF ? 0.125 : (t_en - ot) / (NSAMP - 1);
            const double ok = HALF ? (oe - os) * 2.0 : (oe - os) / (t_en - ot);
            const double t = ot + (double)(i - 1) * odt;               // sic: samples at (i-1)*dt (:251-252)
            const double s = os + (ok * (double)(i - 1)) * odt;
            Q.val[lane] = point_cost(s, t, obs[j], w_obs, lnw_hi, lnw_lo);
        }
        __syncwarp();
        const int lo = max(start, b0), hi = min(start + cnt, b0 + 32);
        for (int q = lo; q < hi; ++q) cost = cost + Q.val[q - b0];
    }
    return cost;
}

#ifndef EMP_ST_MINBLOCKS
#define EMP_ST_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(ST_THREADS, EMP_ST_MINBLOCKS)
st_dp_kernel(const double *__restrict__ scenes, double *__restrict__ out, double *__restrict__ cost_tab,
             double *__restrict__ sdot_tab, int32_t *__restrict__ node_tab, const StLaunch p) {
    __shared__ double cand[NROW * NROW];              // [k][j]: source-major, conflict-free both ways
    __shared__ double cost_sm[2][NROW], sdot_sm[2][NROW], s_of[NROW];
    __shared__ unsigned char node_sm[NCOL][NROW];
    __shared__ Obs obs[EMP_ST_KMAX];
    __shared__ double top_cost[NCOL];
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    // [NCOL][NSAMP][kmax] each; tab_a: edges between lattice columns (t_start = node_t(c-1)); tab_b: edges from the
    // DP origin (column 0, and source row 0)
    Interval *tab_a = reinterpret_cast<Interval *>(dyn_smem);
    Interval *tab_b = tab_a + NCOL * NSAMP * p.kmax;
    uint16_t *qdesc = reinterpret_cast<uint16_t *>(tab_b + NCOL * NSAMP * p.kmax);
    __shared__ double qval[ST_WARPS][32];
    __shared__ double ub_sm[NROW];
    __shared__ unsigned short elist[NROW * NROW];      // edges to evaluate exactly in this column
    __shared__ unsigned char estate[NROW * NROW];      // 1: already listed
    __shared__ int n_list;
    __shared__ uint32_t col_obs_a[NCOL], col_obs_b[NCOL];  // per column: obstacles any sample of the column can be near
    __shared__ int n_act_sm;
    const int frame = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *sc = scenes + (size_t)frame * p.scene_stride;
    const double v0 = sc[0];
    const int K = p.kmax;
    if (tid == 0) {                                    // present obstacles, order kept (NaN s_in = absent, :255)
        int n = 0;
        for (int j = 0; j < K; ++j)
            if (!isnan(sc[2 + j])) {
                obs[n].s_in = sc[2 + j];
                obs[n].s_out = sc[2 + K + j];
                obs[n].t_in = sc[2 + 2 * K + j];
                obs[n].t_out = sc[2 + 3 * K + j];
                ++n;
            }
        n_act_sm = n;
    }
    if (tid < NROW) s_of[tid] = node_s(tid);
    __syncthreads();
    const int n_act = n_act_sm;
    for (int q = tid; q < NCOL * NSAMP * n_act; q += ST_THREADS) {
        const int j = q % n_act, i = (q / n_act) % NSAMP, c = q / (n_act * NSAMP);
        const Obs o = obs[j];
        const double t_en = node_t(c);
        if (c >= 1) {
            const double t_st = node_t(c - 1);
            tab_a[(c * NSAMP + i) * K + j] = capsule_slice(o.s_in, o.t_in, o.s_out, o.t_out, t_st + (double)(i - 1) * ((t_en - t_st) / (NSAMP - 1)));
        }
        tab_b[(c * NSAMP + i) * K + j] = capsule_slice(o.s_in, o.t_in, o.s_out, o.t_out, 0.0 + (double)(i - 1) * ((t_en - 0.0) / (NSAMP - 1)));
    }
    __syncthreads();
    if (tid < 2 * NCOL) {
        const int c = tid >> 1;
        const Interval *t = ((tid & 1) ? tab_b : tab_a) + c * NSAMP * K;
        uint32_t m = 0;
        if ((tid & 1) || c >= 1)
            for (int i = 0; i < NSAMP; ++i)
                for (int j = 0; j < n_act; ++j)
                    if (t[i * K + j].lo <= t[i * K + j].hi) m |= 1u << j;
        ((tid & 1) ? col_obs_b : col_obs_a)[c] = m;
    }
    __syncthreads();
    const double INF = __longlong_as_double(0x7ff0000000000000LL);
    const EmpStParams &P = p.prm;
    const bool prune = P.w_obs >= 0.0;                 // the bound needs a non-negative obstacle term
    const WarpQueue Q{qdesc + (size_t)warp * 32 * NSAMP * p.kmax, qval[warp]};

    // CalcDpCost (:191-231) around the warp-cooperative obstacle term
    auto edge = [&](auto half, bool valid, double s_st, double t_st, double sd_st, double s_en, double t_en,
                    const Interval *tab, uint32_t col_obs) {
        constexpr bool HALF = decltype(half)::value;
        const double cur = HALF ? (s_en - s_st) * 2.0 : (s_en - s_st) / (t_en - t_st);
        const double acc = HALF ? (cur - sd_st) * 2.0 : (cur - sd_st) / (t_en - t_st);
        const double dv = cur - P.reference_speed;
        const double c_ref = P.w_ref_speed * (dv * dv);
        double c_acc;
        if (4 > acc && acc > -6) c_acc = P.w_accel * (acc * acc);
        else c_acc = (100000 * P.w_accel) * (acc * acc);
        const double c_obs = warp_obs_cost<HALF>(valid, s_st, t_st, s_en, t_en, tab, K, obs, col_obs, Q, P.w_obs,
                                           p.lnw_hi, p.lnw_lo, lane);
        return (c_obs + c_acc) + c_ref;
    };

    // column 0: edges from the DP origin (:125-131)
    for (int base = warp * 32; base < NROW; base += ST_THREADS) {
        const int r = base + lane;
        const bool valid = r < NROW;
        const double c = edge(std::true_type{}, valid, 0.0, 0.0, v0, s_of[valid ? r : 0], node_t(0), tab_b, col_obs_b[0]);
        if (valid) {
            cost_sm[0][r] = c;
            sdot_sm[0][r] = s_of[r] / node_t(0);
            node_sm[0][r] = 0;
            if (cost_tab) cost_tab[((size_t)frame * NROW + r) * NCOL] = c;
            if (sdot_tab) sdot_tab[((size_t)frame * NROW + r) * NCOL] = sdot_sm[0][r];
            if (node_tab) node_tab[((size_t)frame * NROW + r) * NCOL] = 0;
            if (r == 0) top_cost[0] = c;
        }
    }
    __syncthreads();

    for (int c = 1; c < NCOL; ++c) {
        const int prev = (c - 1) & 1, cur = c & 1;
        const double t_en = node_t(c), t_prev = node_t(c - 1);
        // ---- exact branch and bound over the 40 x 40 candidates of the column --------------------------------
        // The obstacle term is >= 0 and rounding is monotone, so a candidate's value without it,
        //   lb = (c_acc + c_ref) + dp_st_cost[k],   is a lower bound of   ((c_obs + c_acc) + c_ref) + dp_st_cost[k].
        // Per target row and round: the PRUNE_M not yet evaluated candidates with the smallest lb (<= ub) are
        // evaluated exactly; the best exact value so far, ub, bounds the row's result from above; the last round
        // evaluates every candidate left with lb <= ub; the rest (lb > ub) cannot be the minimum nor tie with it,
        // and keep lb in `cand`.  The (cost, k) arg-min below therefore sees
        // the same winner, bit for bit, as the reference's full sweep (:134-152) - for ~1/5 of the obstacle work.
        constexpr int NE1 = (NROW - 1) * NROW;
        const Interval *ta = tab_a + c * NSAMP * K;
        auto eval = [&](bool valid, int e) {                       // exact candidate value of edge e (k >= 1)
            const int k = valid ? e / NROW : 1, j = valid ? e - k * NROW : 0;
            const double ec = edge(std::true_type{}, valid, s_of[k], t_prev, sdot_sm[prev][k], s_of[j], t_en, ta, col_obs_a[c]);
            if (valid) cand[e] = ec + cost_sm[prev][k];          // cost_temp + dp_st_cost[pre] (:145)
        };
        for (int e = NROW + tid; e < NROW * NROW; e += ST_THREADS) {
            const int k = e / NROW, j = e - k * NROW;
            const double cur_v = (s_of[j] - s_of[k]) * 2.0;        // t_en - t_prev == 0.5 exactly
            const double acc = (cur_v - sdot_sm[prev][k]) * 2.0;
            const double dv = cur_v - P.reference_speed;
            const double c_ref = P.w_ref_speed * (dv * dv);
            const double c_acc = (4 > acc && acc > -6) ? P.w_accel * (acc * acc) : (100000 * P.w_accel) * (acc * acc);
            cand[e] = (c_acc + c_ref) + cost_sm[prev][k];
            estate[e] = 0;
        }
        if (tid == 0) n_list = 0;
        __syncthreads();
#pragma unroll 1
        for (int round = 0, lo = 0; round < PRUNE_ROUNDS; ++round) {
            const bool last = round == PRUNE_ROUNDS - 1;
            if (!prune) {                                          // no bound available: every edge is listed
                if (round == 0) {
                    for (int e = NROW + tid; e < NROW * NROW; e += ST_THREADS) elist[e - NROW] = (unsigned short)e;
                    if (tid == 0) n_list = NE1;
                }
            } else {
                if (tid < NROW) {
                    // upper bound of the row's result: the best exact value so far (source row 0 is exact after round 0)
                    double ub = round == 0 ? INF : cand[tid];
                    if (!(ub == ub)) ub = INF;
                    if (round > 0)
                        for (int k = 1; k < NROW; ++k) {
                            const double v = cand[k * NROW + tid];
                            if (estate[k * NROW + tid] && v < ub) ub = v;
                        }
                    ub_sm[tid] = ub;
                    if (!last) {                                   // next: the PRUNE_M smallest lower bounds still <= ub
                        for (int r = 0; r < PRUNE_M; ++r) {
                            double best = INF;
                            int kb = -1;
                            for (int k = 1; k < NROW; ++k) {
                                const double v = cand[k * NROW + tid];
                                if (!estate[k * NROW + tid] && v < best) { best = v; kb = k; }
                            }
                            if (kb < 0 || !(best <= ub)) break;
                            estate[kb * NROW + tid] = 1;
                            elist[atomicAdd(&n_list, 1)] = (unsigned short)(kb * NROW + tid);
                        }
                    }
                }
                if (last) {                                        // finally: every candidate that can still win or tie
                    __syncthreads();
                    for (int e = NROW + tid; e < NROW * NROW; e += ST_THREADS)
                        if (!estate[e] && cand[e] <= ub_sm[e % NROW]) elist[atomicAdd(&n_list, 1)] = (unsigned short)e;
                }
            }
            __syncthreads();
            const int hi = n_list;
            const int n_grp = (hi - lo + 31) >> 5;                 // groups of 32 listed edges ...
            const int n_all = n_grp + (round == 0 ? (NROW + 31) / 32 : 0);   // ... + in round 0 the source-row-0 edges
            for (int g = warp; g < n_all; g += ST_WARPS) {
                if (g < n_grp) {
                    const int base = lo + 32 * g;
                    const bool valid = base + lane < hi;
                    eval(valid, valid ? elist[base + lane] : NROW);
                } else {
                    // the 40 edges from source row 0, which the reference treats as starting at the DP origin
                    // (:208-212): always evaluated exactly
                    const int j = 32 * (g - n_grp) + lane;
                    const bool valid = j < NROW;
                    const double ec = edge(std::false_type{}, valid, 0.0, 0.0, v0, s_of[valid ? j : 0], t_en, tab_b + c * NSAMP * K, col_obs_b[c]);
                    if (valid) cand[j] = ec + cost_sm[prev][0];
                }
            }
            lo = hi;
            __syncthreads();
        }
        if (tid < NROW) {
            double best = INF;                                    // dp_st_cost initialised to inf (:119)
            int kb = 0;
            for (int k = 0; k < NROW; ++k) {
                const double v = cand[k * NROW + tid];
                if (v < best) {                                   // strict '<', ascending k
                    best = v;
                    kb = k;
                }
            }
            const bool upd = best < INF;
            const double sd = upd ? (s_of[tid] - s_of[kb]) / (t_en - t_prev) : 0.0;   // true coordinates (:148-150)
            cost_sm[cur][tid] = best;
            sdot_sm[cur][tid] = sd;
            node_sm[c][tid] = (unsigned char)(upd ? kb : 0);
            if (cost_tab) cost_tab[((size_t)frame * NROW + tid) * NCOL + c] = best;
            if (sdot_tab) sdot_tab[((size_t)frame * NROW + tid) * NCOL + c] = sd;
            if (node_tab) node_tab[((size_t)frame * NROW + tid) * NCOL + c] = upd ? kb : 0;
            if (tid == 0) top_cost[c] = best;
        }
        __syncthreads();
    }

    if (tid == 0) {
        // end node: right boundary, then top boundary, '<=' (last minimum wins) (:158-172)
        const int last = (NCOL - 1) & 1;
        double min_cost = INF;
        int min_row = -1, min_col = -1;
        for (int i = 0; i < NROW; ++i)
            if (cost_sm[last][i] <= min_cost) { min_cost = cost_sm[last][i]; min_row = i; min_col = NCOL - 1; }
        for (int j = 0; j < NCOL; ++j)
            if (top_cost[j] <= min_cost) { min_cost = top_cost[j]; min_row = 0; min_col = j; }
        double *o = out + (size_t)frame * EMP_ST_OUT_STRIDE;
        const double NaN = __longlong_as_double(0x7ff8000000000000LL);
        o[0] = (double)min_row; o[1] = (double)min_col; o[2] = min_cost;
        o[3] = 0.0; o[4] = 0.0; o[5] = 0.0; o[6] = 0.0; o[7] = 0.0;
        for (int j = 0; j < NCOL; ++j) { o[8 + j] = NaN; o[8 + NCOL + j] = NaN; }
        if (min_row >= 0) {                                       // repaired back-track (oracle/st_dp.py)
            int r = min_row, c = min_col;
            o[8 + c] = s_of[r]; o[8 + NCOL + c] = node_t(c);
            while (c != 0) {
                r = node_sm[c][r];
                --c;
                o[8 + c] = s_of[r]; o[8 + NCOL + c] = node_t(c);
            }
        }
    }
}

}  // namespace

cudaError_t launch_st_dp(const double *scenes, double *out, double *cost_tab, double *sdot_tab, int32_t *node_tab,
                         const StLaunch &p, cudaStream_t st) {
    StLaunch q = p;
    const long double lw = logl((long double)p.prm.w_obs);          // ln w_obs as a double-double
    q.lnw_hi = (double)lw;
    q.lnw_lo = (double)(lw - (long double)q.lnw_hi);
    const size_t km = p.kmax > 0 ? p.kmax : 1;
    const size_t smem = 2 * NCOL * NSAMP * km * sizeof(Interval) + (size_t)ST_WARPS * 32 * NSAMP * km * sizeof(uint16_t);
    cudaFuncSetAttribute(st_dp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    st_dp_kernel<<<p.batch, ST_THREADS, smem, st>>>(scenes, out, cost_tab, sdot_tab, node_tab, q);
    return cudaGetLastError();
}

}  // namespace emp
