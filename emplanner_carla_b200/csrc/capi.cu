// C ABI of libemplanner.so (see include/emplanner.h).  Host-side glue only:
// argument checking, shared-memory carve-up, launches, the host-buffer planner
// handle, and the FP64 FMA micro-benchmark used as roofline denominator.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include "emp_common.cuh"
#include "sl_kernels.h"
#include "frenet_kernels.h"
#include "st_kernels.h"
#include "smooth_kernels.h"
#include "misc_kernels.h"
#include "st_misc_kernels.h"
#include "st_back_kernels.h"
#include "cycle_kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
int fail_cuda(cudaError_t e, const char *where) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return -100 - (int)e;
}

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

int make_dp_launch(emp::DpLaunch &L, const EmpDpParams *p, int kmax, int batch, int n_enr_max) {
    if (!p) return fail(-1, "EmpDpParams is NULL");
    if (p->row < 1 || p->row > EMP_MAX_ROW) return fail(-2, "row must be in [1, 16]");
    if (p->col < 1 || p->col > 4096) return fail(-2, "col must be in [1, 4096]");
    if (kmax < 0 || kmax > EMP_MAX_KMAX || (kmax & 1)) return fail(-2, "kmax must be even and in [0, 32]");
    if (!(p->sample_s > 0) || !(p->sampling_res > 0)) return fail(-2, "sample_s and sampling_res must be > 0");
    // the branch and bound of K1 uses "cost without the collision term" as a lower bound: needs w_collision >= 0
    if (!(p->w_collision >= 0)) return fail(-2, "w_collision_cost must be >= 0");
    if (batch < 0 || n_enr_max < 2) return fail(-2, "bad batch / n_enr_max");
    L.res = p->sampling_res;
    L.w_col = p->w_collision;
    L.w0 = p->w_smooth[0];
    L.w1 = p->w_smooth[1];
    L.w2 = p->w_smooth[2];
    L.w_ref = p->w_ref;
    L.sample_s = p->sample_s;
    L.sample_l = p->sample_l;
    L.R = p->row;
    L.C = p->col;
    L.kmax = kmax;
    L.n_enr_max = n_enr_max;
    L.scene_stride = emp_scene_stride(kmax);
    L.plan_stride = emp_plan_stride(n_enr_max);
    L.batch = batch;
    L.order = nullptr;
    const int RR = L.R * L.R;
    int off = 16;                                        // mbarrier
    L.off_slab = off;  off += align_up(emp::DP_WARPS * L.scene_stride * 8, 16);
    L.off_f0 = off;    off += align_up(2 * EMP_NSAMPLE * 8, 16);       // f0[10] + um[10]
    L.off_nl = off;    off += EMP_MAX_ROW * 8;
    L.off_refc = off;  off += align_up(RR * 8, 16);
    L.off_ltab = off;  off += align_up(EMP_NSAMPLE * RR * 8, 16);     // l of every (sample, edge): lattice constant
    L.off_warp = off;
    int w = 0;
    L.w_off_cand = w;   w += align_up(RR * 8, 16);
    L.w_off_cost = w;   w += 2 * EMP_MAX_ROW * 8;
    L.w_off_rows = w;   w += align_up(L.C * 4, 16);
    L.w_off_offs = w;   w += align_up((L.C + 1) * 4, 16);
    L.w_off_parent = w; w += align_up(L.C * L.R, 16);
    L.w_off_sj = w;     w += align_up(L.C * 8, 16);
    L.w_off_dlon = w;   w += align_up((kmax > 0 ? kmax : 1) * (EMP_NSAMPLE + 1) * 8, 16);
    L.w_off_k0 = w;     w += EMP_MAX_ROW * 4;
    L.w_off_wl = w;     w += align_up(RR * 2, 16);
    L.w_off_pv = w;     w += align_up(emp::DP_PAIR_CAND * EMP_NSAMPLE * 8 + emp::DP_PAIR_CAND * 2, 16);
    L.warp_bytes = w;
    L.smem_bytes = (unsigned)(off + emp::DP_WARPS * w);
    if (L.smem_bytes > 227u * 1024u) return fail(-3, "lattice too large for shared memory (row*col)");
    return 0;
}

int make_qp_launch(emp::QpLaunch &L, const EmpQpParams *p, int kmax, int batch, int n_enr_max, int n_direct, int mode) {
    if (!p) return fail(-1, "EmpQpParams is NULL");
    if (!(p->dp_sampling_res > 0)) return fail(-2, "dp_sampling_res must be > 0");
    if (kmax < 0 || kmax > EMP_MAX_KMAX || (kmax & 1)) return fail(-2, "kmax must be even and in [0, 32]");
    L.q = *p;
    L.batch = batch;
    L.kmax = kmax;
    L.scene_stride = emp_scene_stride(kmax);
    L.plan_stride = emp_plan_stride(n_enr_max);
    L.n_enr_max = n_enr_max;
    L.n_direct = n_direct;
    L.n_fused_max = (n_enr_max + 1) / 2;
    L.mode = mode;
    L.order = nullptr;
    L.n_peers = 0;
    L.epoch = 0;
    L.self_ready = nullptr;
    for (int r = 0; r < EMP_MAX_PEERS; ++r) { L.peer[r] = nullptr; L.peer_ready[r] = nullptr; }
    return 0;
}

}  // namespace

namespace {
// Temporary device buffers for the host-buffer Frenet entry points (stream-ordered allocation).
struct Tmp {
    cudaStream_t st;
    void *ptrs[24];
    int n = 0;
    cudaError_t err = cudaSuccess;
    explicit Tmp(cudaStream_t s) : st(s) {}
    ~Tmp() { for (int i = 0; i < n; ++i) cudaFreeAsync(ptrs[i], st); }
    template <typename T> T *in(const T *host, size_t count) {
        if (!host || count == 0) return nullptr;
        T *d = out<T>(count);
        if (d && err == cudaSuccess) err = cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, st);
        return d;
    }
    template <typename T> T *out(size_t count) {
        void *d = nullptr;
        if (err == cudaSuccess) err = cudaMallocAsync(&d, count * sizeof(T), st);
        if (d) ptrs[n++] = d;
        return (T *)d;
    }
    template <typename T> void back(T *host, const T *dev, size_t count) {
        if (host && dev && err == cudaSuccess) err = cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, st);
    }
};
}  // namespace

struct EmpPlanner {
    int device = 0, max_batch = 0, kmax = 0, n_enr_max = 0, max_col = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double *d_scenes = nullptr, *d_plans = nullptr, *d_qp_in = nullptr, *d_qp_out = nullptr;
    int32_t *d_rows = nullptr, *d_status = nullptr, *d_order = nullptr;
    double *h_scenes = nullptr, *h_plans = nullptr;       // pinned staging
    int32_t *h_rows = nullptr;
    float last_ms = 0.f;
    // pipelined planning (emp_sl_plan_submit / emp_sl_plan_wait): every slot has its own stream and device buffers,
    // none shared with the synchronous entry points above (which run on `stream`)
    struct Slot {
        cudaStream_t stream = nullptr;
        double *d_scenes = nullptr, *d_plans = nullptr;
        int32_t *d_rows = nullptr, *d_order = nullptr;
        bool busy = false;
    } slot[EMP_PLAN_SLOTS];
};

extern "C" {

void emp_smooth_params_default(EmpSmoothParams *p) {
    p->w_smooth = 0.4; p->w_length = 0.3; p->w_ref = 0.3; p->x_thre = 0.2; p->y_thre = 0.2;
    p->abstol = 1e-7; p->reltol = 1e-6; p->feastol = 1e-7; p->maxiters = 100; p->heading_only = 0;
    p->diverge_factor = 2.0;
}

void emp_speed_qp_params_default(EmpSpeedQpParams *p) {
    p->w_s_dot2 = 10; p->w_v_ref = 50; p->w_jerk = 500; p->reference_speed = 50; p->max_lateral_accel = 0.2 * 9.8;
    p->abstol = 1e-7; p->reltol = 1e-6; p->feastol = 1e-7; p->maxiters = 100; p->repair_full_profile = 0;
    p->diverge_factor = 100.0;
}

void emp_st_params_default(EmpStParams *p) {
    p->reference_speed = 50; p->w_ref_speed = 4000; p->w_accel = 100; p->w_obs = 10000000;
}

void emp_dp_params_default(EmpDpParams *p) {
    p->sampling_res = 2; p->w_collision = 1e12;
    p->w_smooth[0] = 300; p->w_smooth[1] = 1000; p->w_smooth[2] = 5000;
    p->w_ref = 20; p->sample_s = 15; p->sample_l = 1.5; p->row = 12; p->col = 6;
}

void emp_qp_params_default(EmpQpParams *p) {
    p->dp_sampling_res = 2;
    p->w_l = 1000; p->w_dl = 10000; p->w_ddl = 3000; p->w_dddl = 150; p->w_centre = 250;
    p->w_end_l = 40; p->w_end_dl = 40; p->w_end_ddl = 40;
    p->host_d1 = 3; p->host_d2 = 3; p->host_w = 3;
    p->obs_length = 5; p->obs_width = 5;
    p->abstol = 1e-7; p->reltol = 1e-6; p->feastol = 1e-7;
    p->maxiters = 100; p->reserved = 0;
    p->diverge_factor = 2.0;
}

int emp_scene_stride(int kmax) { return EMP_SCENE_HDR + 2 * kmax; }

int emp_plan_stride(int n_enr_max) {
    const int nq = (n_enr_max + 1) / 2;
    return (EMP_PLAN_HDR + 2 * n_enr_max + 3 * nq + 1) & ~1;
}

int emp_max_enriched(const EmpDpParams *p) {
    const int per = (int)std::floor(p->sample_s / p->sampling_res) + 1;
    return p->col * per + 1;
}

int emp_version(void) { return EMP_VERSION; }
const char *emp_last_error(void) { return g_err; }

int emp_sl_order_dev(const double *scenes, int kmax, int batch, int32_t *order, void *stream) {
    if (kmax < 0 || kmax > EMP_MAX_KMAX || (kmax & 1)) return fail(-2, "kmax must be even and in [0, 32]");
    if (batch < 0) return fail(-2, "bad batch");
    if (batch == 0) return 0;
    if (!scenes || !order) return fail(-1, "scenes/order is NULL");
    cudaError_t e = emp::launch_sl_order(scenes, emp_scene_stride(kmax), kmax, batch, order, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "sl_order launch");
}

int emp_sl_dp_dev(const double *scenes, int kmax, int batch, const EmpDpParams *p, int n_enr_max,
                  double *plans, int32_t *rows, const int32_t *order, void *stream) {
    emp::DpLaunch L;
    if (int rc = make_dp_launch(L, p, kmax, batch, n_enr_max)) return rc;
    if (batch == 0) return 0;
    if (!scenes || !plans) return fail(-1, "scenes/plans is NULL");
    L.order = order;
    cudaError_t e = emp::launch_sl_dp(scenes, plans, rows, L, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "sl_dp launch");
}

namespace {
// FUSED-mode launches issued right after a DP launch know the lattice: the densified path has at most
// emp_max_enriched(dp) points, which bounds the QP stations tighter than the record capacity n_enr_max does.
thread_local int g_fused_station_bound = 0;
struct FusedBound {
    explicit FusedBound(const EmpDpParams *dp) { g_fused_station_bound = dp ? (emp_max_enriched(dp) + 1) / 2 : 0; }
    ~FusedBound() { g_fused_station_bound = 0; }
};
void apply_fused_bound(emp::QpLaunch &L) {
    if (g_fused_station_bound > 0 && g_fused_station_bound < L.n_fused_max) L.n_fused_max = g_fused_station_bound;
}
}  // namespace

int emp_sl_qp_dev(const double *scenes, int kmax, int batch, const EmpQpParams *p, int n_enr_max,
                  double *plans, double *lmin_lmax, const int32_t *order, void *stream) {
    emp::QpLaunch L;
    if (int rc = make_qp_launch(L, p, kmax, batch, n_enr_max, 0, emp::QP_MODE_FUSED)) return rc;
    if (batch == 0) return 0;
    if (!scenes || !plans) return fail(-1, "scenes/plans is NULL");
    apply_fused_bound(L);
    L.order = order;
    cudaError_t e = emp::launch_sl_qp(scenes, plans, lmin_lmax, nullptr, nullptr, nullptr, nullptr, nullptr, L,
                                      (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "sl_qp launch");
}

int emp_qp_dev(const double *l_min, const double *l_max, const double *start, int n, int batch,
               const EmpQpParams *p, double *out, int32_t *status_iters, void *stream) {
    emp::QpLaunch L;
    if (n < 1) return fail(-2, "n must be >= 1");
    if (int rc = make_qp_launch(L, p, 0, batch, 2, n, emp::QP_MODE_DIRECT)) return rc;
    if (batch == 0) return 0;
    if (!l_min || !l_max || !start || !out || !status_iters) return fail(-1, "NULL argument");
    cudaError_t e = emp::launch_sl_qp(nullptr, nullptr, nullptr, l_min, l_max, start, out, status_iters, L,
                                      (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "qp launch");
}

int emp_corridor_dev(const double *dp_s, const double *dp_l, int n, const double *scenes, int kmax, int batch,
                     const EmpQpParams *p, double *lmin_lmax, int32_t *status_iters, void *stream) {
    emp::QpLaunch L;
    if (n < 1) return fail(-2, "n must be >= 1");
    if (int rc = make_qp_launch(L, p, kmax, batch, 2, n, emp::QP_MODE_CORRIDOR)) return rc;
    if (batch == 0) return 0;
    if (!dp_s || !dp_l || !scenes || !lmin_lmax || !status_iters) return fail(-1, "NULL argument");
    cudaError_t e = emp::launch_sl_qp(scenes, nullptr, lmin_lmax, dp_s, dp_l, nullptr, nullptr, status_iters, L,
                                      (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "corridor launch");
}

int emp_sl_plan_dev(const double *scenes, int kmax, int batch, const EmpDpParams *dp, const EmpQpParams *qp,
                    int n_enr_max, double *plans, int32_t *rows, int32_t *order_ws, void *stream) {
    if (order_ws)
        if (int rc = emp_sl_order_dev(scenes, kmax, batch, order_ws, stream)) return rc;
    if (int rc = emp_sl_dp_dev(scenes, kmax, batch, dp, n_enr_max, plans, rows, order_ws, stream)) return rc;
    FusedBound bound(dp);
    return emp_sl_qp_dev(scenes, kmax, batch, qp, n_enr_max, plans, nullptr, order_ws, stream);
}

int emp_frenet_project_dev(const double *lines, int n_nodes, const double *origin, const double *points,
                           const int32_t *npts, int max_points, const double *vel, const double *acc,
                           const double *s_map_in, const int32_t *pre_match, int batch,
                           double *s_map, int32_t *match, double *proj, double *sl, double *deriv, void *stream) {
    if (batch < 0 || n_nodes < 2 || n_nodes > 1024) return fail(-2, "bad batch / n_nodes (2..1024)");
    if (batch == 0) return 0;
    if (!lines) return fail(-1, "lines is NULL");
    if (deriv && (!vel || !acc || !points || !origin)) return fail(-1, "deriv needs points, vel, acc and origin");
    if (sl && !origin && !s_map_in) return fail(-1, "sl needs origin or s_map_in");
    if (points && max_points < 1) return fail(-2, "max_points must be >= 1");
    emp::FrenetLaunch L{batch, n_nodes, max_points};
    if (16 + (size_t)emp::FR_WARPS * 5 * n_nodes * 8 > 227u * 1024u) return fail(-3, "reference line too long for shared memory");
    cudaError_t e = emp::launch_frenet_project(lines, origin, points, npts, vel, acc, s_map_in, pre_match, proj, s_map, match, sl, deriv, L,
                                               (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "frenet_project launch");
}

int emp_frenet_to_cart_dev(const double *lines, const double *s_map, int n_nodes, const double *start_sl,
                           const double *path_s, const double *path_l, const int32_t *npath, int max_points,
                           int batch, double *cart, int32_t *ncart, void *stream) {
    if (batch < 0 || n_nodes < 2 || max_points < 1) return fail(-2, "bad batch / n_nodes / max_points");
    if (batch == 0) return 0;
    if (!lines || !s_map || !start_sl || !path_s || !path_l || !cart || !ncart) return fail(-1, "NULL argument");
    emp::FrenetLaunch L{batch, n_nodes, max_points};
    cudaError_t e = emp::launch_frenet_to_cart(lines, s_map, start_sl, path_s, path_l, npath, cart, ncart, L,
                                               (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "frenet_to_cart launch");
}

int emp_smooth_line_dev(const double *xy, const int32_t *npts, int max_points, int batch, const EmpSmoothParams *p,
                        double *out, int32_t *status_iters, void *stream) {
    if (!p) return fail(-1, "EmpSmoothParams is NULL");
    if (batch < 0 || max_points < 2) return fail(-2, "bad batch / max_points");
    if (batch == 0) return 0;
    if (!xy || !out || !status_iters) return fail(-1, "NULL argument");
    if (p->heading_only == 1) {                                        // cal_heading_kappa alone: any length
        cudaError_t e = emp::launch_heading_kappa(xy, npts, out, status_iters, batch, max_points, (cudaStream_t)stream);
        return e == cudaSuccess ? 0 : fail_cuda(e, "heading_kappa launch");
    }
    emp::SmoothLaunch L{*p, batch, max_points, 0};
    cudaError_t e = emp::launch_smooth(xy, npts, out, status_iters, L, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "smooth launch");
}

int emp_st_dp_dev(const double *scenes, int kmax, int batch, const EmpStParams *p, double *out, double *cost_tab,
                  double *sdot_tab, int32_t *node_tab, void *stream) {
    if (!p) return fail(-1, "EmpStParams is NULL");
    if (kmax < 0 || kmax > EMP_ST_KMAX || (kmax & 1)) return fail(-2, "kmax must be even and in [0, 16]");
    if (batch < 0) return fail(-2, "bad batch");
    // w^(1.5-d) is evaluated as exp((1.5-d) ln w), and the pruning bound needs a non-negative obstacle term
    if (!(p->w_obs > 0)) return fail(-2, "w_cost_obs must be > 0");
    if (batch == 0) return 0;
    if (!scenes || !out) return fail(-1, "scenes/out is NULL");
    emp::StLaunch L{*p, batch, kmax, 2 + 4 * kmax};
    cudaError_t e = emp::launch_st_dp(scenes, out, cost_tab, sdot_tab, node_tab, L, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "st_dp launch");
}

int emp_sl_qp_p2p_dev(const double *scenes, int kmax, int batch, const EmpQpParams *qp, int n_enr_max,
                      double *plans, int n_peers, double *const *peer_slices, const int32_t *order, void *stream) {
    if (n_peers < 0 || n_peers > EMP_MAX_PEERS) return fail(-2, "n_peers must be in [0, 8]");
    if (n_peers > 0 && !peer_slices) return fail(-1, "peer_slices is NULL");
    emp::QpLaunch L;
    if (int rc = make_qp_launch(L, qp, kmax, batch, n_enr_max, 0, emp::QP_MODE_FUSED)) return rc;
    if (batch == 0) return 0;
    if (!scenes || !plans) return fail(-1, "scenes/plans is NULL");
    apply_fused_bound(L);
    L.n_peers = n_peers;
    L.order = order;
    for (int r = 0; r < n_peers; ++r) {
        if (!peer_slices[r]) return fail(-1, "peer slice is NULL");
        L.peer[r] = peer_slices[r];
    }
    cudaError_t e = emp::launch_sl_qp(scenes, plans, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, L,
                                      (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "sl_qp (p2p) launch");
}

int emp_sl_qp_gather_dev(const double *scenes, int kmax, int batch, const EmpQpParams *qp, int n_enr_max,
                         double *plans, int n_peers, double *const *peer_slices, uint64_t epoch,
                         uint64_t *self_ready, uint64_t *const *peer_ready, const int32_t *order, void *stream) {
    if (n_peers < 0 || n_peers > EMP_MAX_PEERS) return fail(-2, "n_peers must be in [0, 8]");
    if (n_peers > 0 && (!peer_slices || !peer_ready)) return fail(-1, "peer_slices / peer_ready is NULL");
    if (epoch == 0 || !self_ready) return fail(-2, "epoch must be > 0 and self_ready not NULL");
    emp::QpLaunch L;
    if (int rc = make_qp_launch(L, qp, kmax, batch, n_enr_max, 0, emp::QP_MODE_FUSED)) return rc;
    if (batch <= 0) return fail(-2, "batch must be > 0 (the arrival flag is written by the launch's last warp)");
    if (!scenes || !plans) return fail(-1, "scenes/plans is NULL");
    L.n_peers = n_peers;
    L.order = order;
    L.epoch = epoch;
    L.self_ready = reinterpret_cast<unsigned long long *>(self_ready);
    for (int r = 0; r < n_peers; ++r) {
        if (!peer_slices[r] || !peer_ready[r]) return fail(-1, "peer slice / flag is NULL");
        L.peer[r] = peer_slices[r];
        L.peer_ready[r] = reinterpret_cast<unsigned long long *>(peer_ready[r]);
    }
    cudaError_t e = emp::launch_sl_qp(scenes, plans, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, L,
                                      (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "sl_qp (gather) launch");
}

int emp_gather_wait_signal_dev(const uint64_t *flags, int n_wait, uint64_t epoch, uint64_t *const *signal,
                               int n_signal, int32_t *err, int timeout_ms, void *stream) {
    if (n_wait < 0 || n_wait > 32 || n_signal < 0 || n_signal > EMP_MAX_PEERS + 1) return fail(-2, "bad n_wait / n_signal");
    if ((n_wait > 0 && !flags) || (n_signal > 0 && !signal)) return fail(-1, "NULL argument");
    if (timeout_ms <= 0) timeout_ms = 2000;
    cudaError_t e = emp::launch_flag_wait_signal(reinterpret_cast<const unsigned long long *>(flags), n_wait, epoch,
                                                 reinterpret_cast<unsigned long long *const *>(signal), n_signal, err,
                                                 timeout_ms, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "gather wait/signal launch");
}

int emp_sl_plan_p2p_dev(const double *scenes, int kmax, int batch, const EmpDpParams *dp, const EmpQpParams *qp,
                        int n_enr_max, double *plans, int32_t *rows, int n_peers, double *const *peer_slices,
                        int32_t *order_ws, void *stream) {
    if (order_ws)
        if (int rc = emp_sl_order_dev(scenes, kmax, batch, order_ws, stream)) return rc;
    if (int rc = emp_sl_dp_dev(scenes, kmax, batch, dp, n_enr_max, plans, rows, order_ws, stream)) return rc;
    FusedBound bound(dp);
    return emp_sl_qp_p2p_dev(scenes, kmax, batch, qp, n_enr_max, plans, n_peers, peer_slices, order_ws, stream);
}

void *emp_device_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        fail_cuda(cudaGetLastError(), "emp_device_alloc");
        return nullptr;
    }
    cudaMemset(p, 0, bytes);
    return p;
}

void emp_device_free(void *p) {
    if (p) cudaFree(p);
}

int emp_ipc_export(const void *dev_ptr, unsigned char handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!dev_ptr || !handle) return fail(-1, "NULL argument");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr));
    if (e != cudaSuccess) return fail_cuda(e, "cudaIpcGetMemHandle");
    memcpy(handle, &h, 64);
    return 0;
}

int emp_ipc_open(const unsigned char handle[64], void **dev_ptr) {
    if (!handle || !dev_ptr) return fail(-1, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cudaIpcOpenMemHandle");
}

int emp_ipc_close(void *dev_ptr) {
    if (!dev_ptr) return 0;
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cudaIpcCloseMemHandle");
}

int emp_fp64_peak(int iters, double *tflops, double *ms) {
    cudaError_t e = emp::launch_fp64_peak(iters, tflops, ms);
    return e == cudaSuccess ? 0 : fail_cuda(e, "fp64 peak");
}

// ---- host-buffer planner --------------------------------------------------------
#define EMP_CK(call, what)                                   \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return fail_cuda(e_, what);   \
    } while (0)

int emp_planner_create(EmpPlanner **out, int device, int max_batch, int kmax, int n_enr_max, int max_col) {
    if (!out) return fail(-1, "out is NULL");
    *out = nullptr;
    if (max_batch < 1 || n_enr_max < 2 || max_col < 1) return fail(-2, "bad planner sizes");
    if (kmax < 0 || kmax > EMP_MAX_KMAX || (kmax & 1)) return fail(-2, "kmax must be even and in [0, 32]");
    if (device < 0) EMP_CK(cudaGetDevice(&device), "cudaGetDevice");
    EMP_CK(cudaSetDevice(device), "cudaSetDevice");
    EmpPlanner *pl = new (std::nothrow) EmpPlanner();
    if (!pl) return fail(-4, "out of host memory");
    pl->device = device; pl->max_batch = max_batch; pl->kmax = kmax; pl->n_enr_max = n_enr_max; pl->max_col = max_col;
    const size_t sb = (size_t)max_batch * emp_scene_stride(kmax) * 8;
    const size_t pb = (size_t)max_batch * emp_plan_stride(n_enr_max) * 8;
    const size_t rb = (size_t)max_batch * max_col * 4;
    const size_t qb = (size_t)max_batch * 3 * EMP_MAX_QP_POINTS * 8;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    ok(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
    ok(cudaEventCreate(&pl->ev0));
    ok(cudaEventCreate(&pl->ev1));
    ok(cudaMalloc(&pl->d_scenes, sb));
    ok(cudaMalloc(&pl->d_plans, pb));
    ok(cudaMalloc(&pl->d_rows, rb));
    ok(cudaMalloc(&pl->d_qp_in, qb));
    ok(cudaMalloc(&pl->d_qp_out, qb));
    ok(cudaMalloc(&pl->d_status, (size_t)max_batch * 2 * 4));
    ok(cudaMalloc(&pl->d_order, (size_t)max_batch * 4));
    ok(cudaMallocHost(&pl->h_scenes, sb));
    ok(cudaMallocHost(&pl->h_plans, pb));
    ok(cudaMallocHost(&pl->h_rows, rb));
    if (e != cudaSuccess) {
        emp_planner_destroy(pl);
        return fail_cuda(e, "planner allocation");
    }
    *out = pl;
    return 0;
}

void emp_planner_destroy(EmpPlanner *pl) {
    if (!pl) return;
    cudaSetDevice(pl->device);
    if (pl->stream) cudaStreamSynchronize(pl->stream);
    cudaFree(pl->d_scenes); cudaFree(pl->d_plans); cudaFree(pl->d_rows);
    cudaFree(pl->d_qp_in); cudaFree(pl->d_qp_out); cudaFree(pl->d_status); cudaFree(pl->d_order);
    cudaFreeHost(pl->h_scenes); cudaFreeHost(pl->h_plans); cudaFreeHost(pl->h_rows);
    for (int k = 0; k < EMP_PLAN_SLOTS; ++k) {
        if (pl->slot[k].stream) { cudaStreamSynchronize(pl->slot[k].stream); cudaStreamDestroy(pl->slot[k].stream); }
        cudaFree(pl->slot[k].d_scenes); cudaFree(pl->slot[k].d_plans); cudaFree(pl->slot[k].d_rows); cudaFree(pl->slot[k].d_order);
    }
    if (pl->ev0) cudaEventDestroy(pl->ev0);
    if (pl->ev1) cudaEventDestroy(pl->ev1);
    if (pl->stream) cudaStreamDestroy(pl->stream);
    delete pl;
}

static int plan_host_impl(EmpPlanner *pl, const double *scenes, int batch, const EmpDpParams *dp,
                          const EmpQpParams *qp, double *plans, int32_t *rows) {
    if (!pl || !scenes || !plans || !dp) return fail(-1, "NULL argument");
    if (batch < 0 || batch > pl->max_batch) return fail(-2, "batch exceeds planner capacity");
    if (dp->col > pl->max_col) return fail(-2, "col exceeds planner capacity");
    if (emp_max_enriched(dp) > pl->n_enr_max) return fail(-2, "lattice needs a larger n_enr_max than the planner was created with");
    if (batch == 0) return 0;
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    const size_t sb = (size_t)batch * emp_scene_stride(pl->kmax) * 8;
    const size_t pb = (size_t)batch * emp_plan_stride(pl->n_enr_max) * 8;
    const size_t rb = (size_t)batch * dp->col * 4;
    // Pageable caller buffers are staged through the planner's pinned buffers so the copies are real DMA;
    // buffers that are already page-locked (emp_host_alloc, cudaHostRegister, torch pin_memory) are used directly.
    const bool in_pinned = is_pinned(scenes), out_pinned = is_pinned(plans), rows_pinned = rows && is_pinned(rows);
    const double *h_in = scenes;
    if (!in_pinned) { memcpy(pl->h_scenes, scenes, sb); h_in = pl->h_scenes; }
    EMP_CK(cudaMemcpyAsync(pl->d_scenes, h_in, sb, cudaMemcpyHostToDevice, pl->stream), "H2D scenes");
    EMP_CK(cudaEventRecord(pl->ev0, pl->stream), "event");
    int rc = emp_sl_order_dev(pl->d_scenes, pl->kmax, batch, pl->d_order, pl->stream);
    if (rc) return rc;
    rc = emp_sl_dp_dev(pl->d_scenes, pl->kmax, batch, dp, pl->n_enr_max, pl->d_plans, rows ? pl->d_rows : nullptr, pl->d_order,
                       pl->stream);
    if (rc) return rc;
    if (qp) {
        FusedBound bound(dp);
        rc = emp_sl_qp_dev(pl->d_scenes, pl->kmax, batch, qp, pl->n_enr_max, pl->d_plans, nullptr, pl->d_order, pl->stream);
        if (rc) return rc;
    }
    EMP_CK(cudaEventRecord(pl->ev1, pl->stream), "event");
    EMP_CK(cudaMemcpyAsync(out_pinned ? plans : pl->h_plans, pl->d_plans, pb, cudaMemcpyDeviceToHost, pl->stream), "D2H plans");
    if (rows) EMP_CK(cudaMemcpyAsync(rows_pinned ? rows : pl->h_rows, pl->d_rows, rb, cudaMemcpyDeviceToHost, pl->stream), "D2H rows");
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    if (!out_pinned) memcpy(plans, pl->h_plans, pb);
    if (rows && !rows_pinned) memcpy(rows, pl->h_rows, rb);
    cudaEventElapsedTime(&pl->last_ms, pl->ev0, pl->ev1);
    return 0;
}

int emp_sl_plan_submit(EmpPlanner *pl, int slot, const double *scenes, int batch, const EmpDpParams *dp,
                       const EmpQpParams *qp, double *plans, int32_t *rows) {
    if (!pl || !scenes || !plans || !dp || !qp) return fail(-1, "NULL argument");
    if (slot < 0 || slot >= EMP_PLAN_SLOTS) return fail(-2, "slot out of range");
    if (batch < 1 || batch > pl->max_batch) return fail(-2, "batch exceeds planner capacity");
    if (dp->col > pl->max_col) return fail(-2, "col exceeds planner capacity");
    if (emp_max_enriched(dp) > pl->n_enr_max) return fail(-2, "lattice needs a larger n_enr_max than the planner was created with");
    EmpPlanner::Slot &S = pl->slot[slot];
    if (S.busy) return fail(-5, "slot still in flight: call emp_sl_plan_wait first");
    if (!is_pinned(scenes) || !is_pinned(plans) || (rows && !is_pinned(rows)))
        return fail(-2, "emp_sl_plan_submit needs page-locked host buffers (emp_host_alloc / cudaHostRegister)");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    const size_t sb_max = (size_t)pl->max_batch * emp_scene_stride(pl->kmax) * 8;
    const size_t pb_max = (size_t)pl->max_batch * emp_plan_stride(pl->n_enr_max) * 8;
    if (!S.stream) {                                   // first use of the slot: its buffers, then (only then) its stream
        cudaError_t e = cudaSuccess;
        auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
        ok(cudaMalloc(&S.d_scenes, sb_max));
        ok(cudaMalloc(&S.d_plans, pb_max));
        ok(cudaMalloc(&S.d_rows, (size_t)pl->max_batch * pl->max_col * 4));
        ok(cudaMalloc(&S.d_order, (size_t)pl->max_batch * 4));
        if (e == cudaSuccess) ok(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
        if (e != cudaSuccess) {
            cudaFree(S.d_scenes); cudaFree(S.d_plans); cudaFree(S.d_rows); cudaFree(S.d_order);
            S.d_scenes = S.d_plans = nullptr;
            S.d_rows = S.d_order = nullptr;
            S.stream = nullptr;
            return fail_cuda(e, "slot allocation");
        }
    }
    const size_t sb = (size_t)batch * emp_scene_stride(pl->kmax) * 8;
    const size_t pb = (size_t)batch * emp_plan_stride(pl->n_enr_max) * 8;
    EMP_CK(cudaMemcpyAsync(S.d_scenes, scenes, sb, cudaMemcpyHostToDevice, S.stream), "H2D scenes");
    if (int rc = emp_sl_plan_dev(S.d_scenes, pl->kmax, batch, dp, qp, pl->n_enr_max, S.d_plans, rows ? S.d_rows : nullptr,
                                 S.d_order, S.stream)) return rc;
    EMP_CK(cudaMemcpyAsync(plans, S.d_plans, pb, cudaMemcpyDeviceToHost, S.stream), "D2H plans");
    if (rows) EMP_CK(cudaMemcpyAsync(rows, S.d_rows, (size_t)batch * dp->col * 4, cudaMemcpyDeviceToHost, S.stream), "D2H rows");
    S.busy = true;
    return 0;
}

int emp_sl_plan_wait(EmpPlanner *pl, int slot) {
    if (!pl) return fail(-1, "NULL argument");
    if (slot < 0 || slot >= EMP_PLAN_SLOTS) return fail(-2, "slot out of range");
    EmpPlanner::Slot &S = pl->slot[slot];
    if (!S.busy) return 0;
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    EMP_CK(cudaStreamSynchronize(S.stream), "slot stream");
    S.busy = false;
    return 0;
}

int emp_sl_plan_host(EmpPlanner *pl, const double *scenes, int batch, const EmpDpParams *dp,
                     const EmpQpParams *qp, double *plans, int32_t *rows) {
    if (!qp) return fail(-1, "EmpQpParams is NULL");
    return plan_host_impl(pl, scenes, batch, dp, qp, plans, rows);
}

int emp_sl_dp_host(EmpPlanner *pl, const double *scenes, int batch, const EmpDpParams *dp, double *plans,
                   int32_t *rows) {
    return plan_host_impl(pl, scenes, batch, dp, nullptr, plans, rows);
}

int emp_qp_host(EmpPlanner *pl, const double *l_min, const double *l_max, const double *start, int n, int batch,
                const EmpQpParams *p, double *out, int32_t *status_iters) {
    if (!pl || !l_min || !l_max || !start || !out || !status_iters) return fail(-1, "NULL argument");
    if (batch < 0 || batch > pl->max_batch) return fail(-2, "batch exceeds planner capacity");
    if (n < 1 || n > EMP_MAX_QP_POINTS) {
        // report like the kernel would, without launching
        for (int b = 0; b < batch; ++b) { status_iters[2 * b] = EMP_QP_TOO_LARGE; status_iters[2 * b + 1] = 0; }
        return 0;
    }
    if (batch == 0) return 0;
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    const size_t nb = (size_t)batch * n * 8;
    double *d_lmin = pl->d_qp_in, *d_lmax = pl->d_qp_in + (size_t)batch * n, *d_start = pl->d_qp_in + 2 * (size_t)batch * n;
    EMP_CK(cudaMemcpyAsync(d_lmin, l_min, nb, cudaMemcpyHostToDevice, pl->stream), "H2D lmin");
    EMP_CK(cudaMemcpyAsync(d_lmax, l_max, nb, cudaMemcpyHostToDevice, pl->stream), "H2D lmax");
    EMP_CK(cudaMemcpyAsync(d_start, start, (size_t)batch * 3 * 8, cudaMemcpyHostToDevice, pl->stream), "H2D start");
    EMP_CK(cudaEventRecord(pl->ev0, pl->stream), "event");
    int rc = emp_qp_dev(d_lmin, d_lmax, d_start, n, batch, p, pl->d_qp_out, pl->d_status, pl->stream);
    if (rc) return rc;
    EMP_CK(cudaEventRecord(pl->ev1, pl->stream), "event");
    EMP_CK(cudaMemcpyAsync(out, pl->d_qp_out, (size_t)batch * 3 * n * 8, cudaMemcpyDeviceToHost, pl->stream), "D2H qp");
    EMP_CK(cudaMemcpyAsync(status_iters, pl->d_status, (size_t)batch * 2 * 4, cudaMemcpyDeviceToHost, pl->stream), "D2H status");
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    cudaEventElapsedTime(&pl->last_ms, pl->ev0, pl->ev1);
    return 0;
}

int emp_corridor_host(EmpPlanner *pl, const double *dp_s, const double *dp_l, int n, const double *scenes, int batch,
                      const EmpQpParams *p, double *lmin_lmax, int32_t *status_iters) {
    if (!pl || !dp_s || !dp_l || !scenes || !lmin_lmax || !status_iters) return fail(-1, "NULL argument");
    if (batch < 0 || batch > pl->max_batch) return fail(-2, "batch exceeds planner capacity");
    if (n < 1 || n > EMP_MAX_QP_POINTS) {
        for (int b = 0; b < batch; ++b) { status_iters[2 * b] = EMP_QP_TOO_LARGE; status_iters[2 * b + 1] = 0; }
        return 0;
    }
    if (batch == 0) return 0;
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    const size_t nb = (size_t)batch * n * 8;
    const size_t sb = (size_t)batch * emp_scene_stride(pl->kmax) * 8;
    double *d_s = pl->d_qp_in, *d_l = pl->d_qp_in + (size_t)batch * n;
    EMP_CK(cudaMemcpyAsync(d_s, dp_s, nb, cudaMemcpyHostToDevice, pl->stream), "H2D dp_s");
    EMP_CK(cudaMemcpyAsync(d_l, dp_l, nb, cudaMemcpyHostToDevice, pl->stream), "H2D dp_l");
    EMP_CK(cudaMemcpyAsync(pl->d_scenes, scenes, sb, cudaMemcpyHostToDevice, pl->stream), "H2D scenes");
    int rc = emp_corridor_dev(d_s, d_l, n, pl->d_scenes, pl->kmax, batch, p, pl->d_qp_out, pl->d_status, pl->stream);
    if (rc) return rc;
    EMP_CK(cudaMemcpyAsync(lmin_lmax, pl->d_qp_out, 2 * nb, cudaMemcpyDeviceToHost, pl->stream), "D2H corridor");
    EMP_CK(cudaMemcpyAsync(status_iters, pl->d_status, (size_t)batch * 2 * 4, cudaMemcpyDeviceToHost, pl->stream), "D2H status");
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    return 0;
}

int emp_frenet_project_host(EmpPlanner *pl, const double *lines, int n_nodes, const double *origin,
                            const double *points, const int32_t *npts, int max_points, const double *vel,
                            const double *acc, const double *s_map_in, const int32_t *pre_match, int batch,
                            double *s_map, int32_t *match, double *proj, double *sl, double *deriv) {
    if (!pl) return fail(-1, "planner is NULL");
    if (batch <= 0) return batch == 0 ? 0 : fail(-2, "bad batch");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    const size_t B = batch, P = points ? max_points : 0;
    int rc;
    {
        Tmp t(pl->stream);
        const double *d_lines = t.in(lines, B * n_nodes * 4), *d_org = t.in(origin, B * 2);
        const double *d_pts = t.in(points, B * P * 2), *d_vel = t.in(vel, B * P * 2), *d_acc = t.in(acc, B * P * 2);
        const int32_t *d_np = t.in(npts, B), *d_pre = t.in(pre_match, B);
        const double *d_smin = t.in(s_map_in, B * n_nodes);
        double *d_smap = s_map ? t.out<double>(B * n_nodes) : nullptr;
        double *d_proj = (proj && P) ? t.out<double>(B * P * 4) : nullptr;
        int32_t *d_match = (match && P) ? t.out<int32_t>(B * P) : nullptr;
        double *d_sl = (sl && P) ? t.out<double>(B * P * 2) : nullptr;
        double *d_der = (deriv && P) ? t.out<double>(B * P * 7) : nullptr;
        if (t.err != cudaSuccess) return fail_cuda(t.err, "frenet_project_host staging");
        rc = emp_frenet_project_dev(d_lines, n_nodes, d_org, d_pts, d_np, max_points, d_vel, d_acc, d_smin, d_pre,
                                    batch, d_smap, d_match, d_proj, d_sl, d_der, pl->stream);
        if (rc == 0) {
            t.back(s_map, d_smap, B * n_nodes);
            t.back(match, d_match, B * P);
            t.back(proj, d_proj, B * P * 4);
            t.back(sl, d_sl, B * P * 2);
            t.back(deriv, d_der, B * P * 7);
            if (t.err != cudaSuccess) return fail_cuda(t.err, "frenet_project_host copy-back");
        }
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    return rc;
}

int emp_frenet_to_cart_host(EmpPlanner *pl, const double *lines, const double *s_map, int n_nodes,
                            const double *start_sl, const double *path_s, const double *path_l,
                            const int32_t *npath, int max_points, int batch, double *cart, int32_t *ncart) {
    if (!pl || !cart || !ncart) return fail(-1, "NULL argument");
    if (batch <= 0) return batch == 0 ? 0 : fail(-2, "bad batch");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    const size_t B = batch, M = max_points;
    int rc;
    {
        Tmp t(pl->stream);
        const double *d_lines = t.in(lines, B * n_nodes * 4), *d_smap = t.in(s_map, B * n_nodes);
        const double *d_start = t.in(start_sl, B * 2), *d_ps = t.in(path_s, B * M), *d_pl = t.in(path_l, B * M);
        const int32_t *d_np = t.in(npath, B);
        double *d_cart = t.out<double>(B * (M + 1) * 4);
        int32_t *d_nc = t.out<int32_t>(B);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "frenet_to_cart_host staging");
        cudaMemsetAsync(d_cart, 0xff, B * (M + 1) * 4 * sizeof(double), pl->stream);     // unwritten rows read as NaN
        rc = emp_frenet_to_cart_dev(d_lines, d_smap, n_nodes, d_start, d_ps, d_pl, d_np, max_points, batch, d_cart,
                                    d_nc, pl->stream);
        if (rc == 0) {
            t.back(cart, d_cart, B * (M + 1) * 4);
            t.back(ncart, d_nc, B);
            if (t.err != cudaSuccess) return fail_cuda(t.err, "frenet_to_cart_host copy-back");
        }
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    return rc;
}

int emp_smooth_line_host(EmpPlanner *pl, const double *xy, const int32_t *npts, int max_points, int batch,
                         const EmpSmoothParams *p, double *out, int32_t *status_iters) {
    if (!pl || !xy || !out || !status_iters) return fail(-1, "NULL argument");
    if (batch <= 0) return batch == 0 ? 0 : fail(-2, "bad batch");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    const size_t B = batch, M = max_points;
    int rc;
    {
        Tmp t(pl->stream);
        const double *d_xy = t.in(xy, B * M * 2);
        const int32_t *d_np = t.in(npts, B);
        double *d_out = t.out<double>(B * M * 4);
        int32_t *d_si = t.out<int32_t>(B * 2);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "smooth_line_host staging");
        cudaMemsetAsync(d_out, 0, B * M * 4 * sizeof(double), pl->stream);
        rc = emp_smooth_line_dev(d_xy, d_np, max_points, batch, p, d_out, d_si, pl->stream);
        if (rc == 0) {
            t.back(out, d_out, B * M * 4);
            t.back(status_iters, d_si, B * 2);
            if (t.err != cudaSuccess) return fail_cuda(t.err, "smooth_line_host copy-back");
        }
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    return rc;
}

int emp_st_dp_host(EmpPlanner *pl, const double *scenes, int kmax, int batch, const EmpStParams *p, double *out,
                   double *cost_tab, double *sdot_tab, int32_t *node_tab) {
    if (!pl || !scenes || !out) return fail(-1, "NULL argument");
    if (batch <= 0) return batch == 0 ? 0 : fail(-2, "bad batch");
    if (kmax < 0 || kmax > EMP_ST_KMAX || (kmax & 1)) return fail(-2, "kmax must be even and in [0, 16]");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    const size_t B = batch, T = (size_t)EMP_ST_ROWS * EMP_ST_COLS;
    int rc;
    {
        Tmp t(pl->stream);
        const double *d_sc = t.in(scenes, B * (2 + 4 * kmax));
        double *d_out = t.out<double>(B * EMP_ST_OUT_STRIDE);
        double *d_cost = cost_tab ? t.out<double>(B * T) : nullptr;
        double *d_sdot = sdot_tab ? t.out<double>(B * T) : nullptr;
        int32_t *d_node = node_tab ? t.out<int32_t>(B * T) : nullptr;
        if (t.err != cudaSuccess) return fail_cuda(t.err, "st_dp_host staging");
        EMP_CK(cudaEventRecord(pl->ev0, pl->stream), "event");
        rc = emp_st_dp_dev(d_sc, kmax, batch, p, d_out, d_cost, d_sdot, d_node, pl->stream);
        EMP_CK(cudaEventRecord(pl->ev1, pl->stream), "event");
        if (rc == 0) {
            t.back(out, d_out, B * EMP_ST_OUT_STRIDE);
            t.back(cost_tab, d_cost, B * T);
            t.back(sdot_tab, d_sdot, B * T);
            t.back(node_tab, d_node, B * T);
            if (t.err != cudaSuccess) return fail_cuda(t.err, "st_dp_host copy-back");
        }
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    cudaEventElapsedTime(&pl->last_ms, pl->ev0, pl->ev1);
    return rc;
}

int emp_quintic_host(EmpPlanner *pl, const double *in, int n, double *out) {
    if (!pl || !in || !out) return fail(-1, "NULL argument");
    if (n <= 0) return n == 0 ? 0 : fail(-2, "bad n");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    {
        Tmp t(pl->stream);
        const double *d_in = t.in(in, (size_t)n * 8);
        double *d_out = t.out<double>((size_t)n * 6);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "quintic staging");
        EMP_CK(emp::launch_quintic(d_in, d_out, n, pl->stream), "quintic launch");
        t.back(out, d_out, (size_t)n * 6);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "quintic copy-back");
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    return 0;
}

int emp_edge_cost_host(EmpPlanner *pl, const double *edges, int n_edges, const double *obs_s, const double *obs_l,
                       int n_obs, const EmpDpParams *p, double *out) {
    return emp_edge_cost_ex_host(pl, edges, n_edges, obs_s, obs_l, n_obs, p, 4.0, 6.0, out);   // path_planning.py:588
}

int emp_edge_cost_ex_host(EmpPlanner *pl, const double *edges, int n_edges, const double *obs_s, const double *obs_l,
                          int n_obs, const EmpDpParams *p, double danger_dis, double safe_dis, double *out) {
    if (!pl || !edges || !out || !p) return fail(-1, "NULL argument");
    if (n_edges <= 0) return n_edges == 0 ? 0 : fail(-2, "bad n_edges");
    if (n_obs < 0 || (n_obs > 0 && (!obs_s || !obs_l))) return fail(-2, "bad obstacle list");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    {
        Tmp t(pl->stream);
        const double *d_e = t.in(edges, (size_t)n_edges * 6);
        const double *d_os = t.in(obs_s, (size_t)n_obs), *d_ol = t.in(obs_l, (size_t)n_obs);
        double *d_out = t.out<double>((size_t)n_edges);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "edge_cost staging");
        emp::EdgeParams P{p->sample_s, p->w_collision, p->w_smooth[0], p->w_smooth[1], p->w_smooth[2], p->w_ref,
                           danger_dis * danger_dis, safe_dis * safe_dis};
        EMP_CK(emp::launch_edge_cost(d_e, d_os, d_ol, n_obs, n_edges, P, d_out, pl->stream), "edge_cost launch");
        t.back(out, d_out, (size_t)n_edges);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "edge_cost copy-back");
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    return 0;
}

int emp_obs_cost_host(EmpPlanner *pl, const double *square_d, int n, int batch, double w_collision, double danger_dis,
                      double safe_dis, double *out) {
    if (!pl || !square_d || !out) return fail(-1, "NULL argument");
    if (batch <= 0 || n < 0) return batch == 0 ? 0 : fail(-2, "bad sizes");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    {
        Tmp t(pl->stream);
        const double *d_in = t.in(square_d, (size_t)batch * (n > 0 ? n : 1));
        double *d_out = t.out<double>((size_t)batch);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "obs_cost staging");
        EMP_CK(emp::launch_obs_cost(d_in, n, batch, w_collision, danger_dis * danger_dis, safe_dis * safe_dis, d_out,
                                    pl->stream), "obs_cost launch");
        t.back(out, d_out, (size_t)batch);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "obs_cost copy-back");
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    return 0;
}

int emp_enrich_host(EmpPlanner *pl, const double *node_s, const double *node_l, const double *start, int C, int batch,
                    double resolution, int n_max, double *out_s, double *out_l, int32_t *n_out) {
    if (!pl || !node_s || !node_l || !start || !out_s || !out_l || !n_out) return fail(-1, "NULL argument");
    if (batch <= 0 || C < 1 || n_max < 1 || !(resolution > 0)) return batch == 0 ? 0 : fail(-2, "bad sizes");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    {
        Tmp t(pl->stream);
        const double *d_s = t.in(node_s, (size_t)batch * C), *d_l = t.in(node_l, (size_t)batch * C);
        const double *d_st = t.in(start, (size_t)batch * 4);
        double *d_os = t.out<double>((size_t)batch * n_max), *d_ol = t.out<double>((size_t)batch * n_max);
        int32_t *d_n = t.out<int32_t>((size_t)batch);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "enrich staging");
        EMP_CK(emp::launch_enrich(d_s, d_l, d_st, C, batch, resolution, n_max, d_os, d_ol, d_n, pl->stream), "enrich launch");
        t.back(out_s, d_os, (size_t)batch * n_max);
        t.back(out_l, d_ol, (size_t)batch * n_max);
        t.back(n_out, d_n, (size_t)batch);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "enrich copy-back");
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    return 0;
}

int emp_misc_host(EmpPlanner *pl, int op, const double *in, int n, int batch, const EmpMiscAux *aux, double *out) {
    if (!pl || !in || !out) return fail(-1, "NULL argument");
    if (n < 0 || batch < 0) return fail(-2, "bad sizes");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    size_t n_in = 0, n_out = 0;
    switch (op) {
        case EMP_OP_ST_GRAPH: n_in = (size_t)batch * 4 * n; n_out = n_in; break;
        case EMP_OP_START_CONDITION: n_in = (size_t)batch * 5; n_out = (size_t)batch * 2; break;
        case EMP_OP_DY_OBS_DERI: n_in = (size_t)batch * 5 * n; n_out = (size_t)batch * 3 * 128; break;
        case EMP_OP_INDEX2S: n_in = (size_t)batch * 2 * n; n_out = (size_t)batch * n; break;
        case EMP_OP_ST_DP_COST: case EMP_OP_ST_OBS_COST: n_in = (size_t)n * 4; n_out = n; break;
        case EMP_OP_ST_COLLISION: n_in = n; n_out = n; break;
        case EMP_OP_FRENET2CART: case EMP_OP_CALC_PROJ_POINT: n_in = (size_t)n * 4; n_out = (size_t)n * 4; break;
        default: return fail(-2, "unknown op");
    }
    if (n_in == 0 || n_out == 0) return 0;
    const bool cost_op = op == EMP_OP_ST_DP_COST || op == EMP_OP_ST_OBS_COST || op == EMP_OP_ST_COLLISION;
    const bool cart_op = op == EMP_OP_FRENET2CART || op == EMP_OP_CALC_PROJ_POINT;
    if ((cost_op || cart_op) && !aux) return fail(-1, "aux is NULL");
    if (cost_op && (aux->n_obs < 0 || aux->n_obs > EMP_ST_KMAX)) return fail(-2, "n_obs must be in [0, 16]");
    if (cart_op && (!aux->path || !aux->s_map || aux->n_nodes < 2)) return fail(-2, "bad path / s_map");
    {
        Tmp t(pl->stream);
        const double *d_in = t.in(in, n_in);
        double *d_out = t.out<double>(n_out);
        const double *d_obs = nullptr, *d_tab = nullptr, *d_path = nullptr, *d_smap = nullptr;
        if (cost_op) {
            d_obs = t.in(aux->obs4, (size_t)4 * aux->n_obs);
            d_tab = t.in(aux->sdot_tab, (size_t)EMP_ST_ROWS * EMP_ST_COLS);
        }
        if (cart_op) {
            d_path = t.in(aux->path, (size_t)4 * aux->n_nodes);
            d_smap = t.in(aux->s_map, (size_t)aux->n_nodes);
        }
        if (t.err != cudaSuccess) return fail_cuda(t.err, "misc staging");
        cudaError_t e = cudaSuccess;
        switch (op) {
            case EMP_OP_ST_GRAPH: e = emp::launch_st_graph(d_in, n, batch, d_out, pl->stream); break;
            case EMP_OP_START_CONDITION: e = emp::launch_start_condition(d_in, batch, d_out, pl->stream); break;
            case EMP_OP_DY_OBS_DERI: e = emp::launch_dy_obs_deri(d_in, n, batch, d_out, pl->stream); break;
            case EMP_OP_INDEX2S: e = emp::launch_index2s(d_in, n, batch, d_out, pl->stream); break;
            case EMP_OP_ST_DP_COST: case EMP_OP_ST_OBS_COST: case EMP_OP_ST_COLLISION:
                if (op == EMP_OP_ST_DP_COST && !d_tab) return fail(-1, "sdot_tab is NULL");
                e = emp::launch_st_cost(op == EMP_OP_ST_DP_COST ? 0 : (op == EMP_OP_ST_OBS_COST ? 1 : 2), d_in, n, d_obs,
                                        aux->n_obs, d_tab, aux->v0, aux->st, d_out, pl->stream);
                break;
            default:
                e = emp::launch_frenet2cart(d_in, n, d_path, d_smap, aux->n_nodes, op == EMP_OP_CALC_PROJ_POINT, d_out,
                                            pl->stream);
        }
        if (e != cudaSuccess) return fail_cuda(e, "misc launch");
        t.back(out, d_out, n_out);
        if (t.err != cudaSuccess) return fail_cuda(t.err, "misc copy-back");
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    return 0;
}

int emp_cycle_path_points_dev(const double *plans, int n_enr_max, const double *scenes, int kmax, int batch,
                              int max_points, double *xy, int32_t *npts, void *stream) {
    if (batch < 0 || n_enr_max < 2 || max_points < 4) return fail(-2, "bad batch / n_enr_max / max_points");
    if (kmax < 0 || kmax > EMP_MAX_KMAX || (kmax & 1)) return fail(-2, "kmax must be even and in [0, 32]");
    if (batch == 0) return 0;
    if (!plans || !scenes || !xy || !npts) return fail(-1, "NULL argument");
    cudaError_t e = emp::launch_cycle_path_points(plans, emp_plan_stride(n_enr_max), n_enr_max, scenes,
                                                  emp_scene_stride(kmax), batch, max_points, xy, npts,
                                                  (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cycle_path_points launch");
}

int emp_cycle_path_tables_dev(const double *cart, const int32_t *npts, int max_points, int batch, int n_path,
                              double *path2, double *path5, void *stream) {
    if (batch < 0 || max_points < 1 || n_path < 2) return fail(-2, "bad batch / max_points / n_path");
    if (batch == 0) return 0;
    if (!cart || !npts || !path2 || !path5) return fail(-1, "NULL argument");
    cudaError_t e = emp::launch_cycle_path_tables(cart, npts, max_points, batch, n_path, path2, path5,
                                                  (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "cycle_path_tables launch");
}

int emp_st_convex_space_dev(const double *st_out, const double *scenes, int kmax, const double *path, int n_path,
                            const EmpSpeedQpParams *p, int batch, double *bounds, int32_t *status, void *stream) {
    if (!p) return fail(-1, "EmpSpeedQpParams is NULL");
    if (kmax < 0 || kmax > EMP_ST_KMAX || (kmax & 1)) return fail(-2, "kmax must be even and in [0, 16]");
    if (batch < 0 || n_path < 1) return fail(-2, "bad batch / n_path");
    if (batch == 0) return 0;
    if (!st_out || !scenes || !path || !bounds || !status) return fail(-1, "NULL argument");
    cudaError_t e = emp::launch_convex_space(st_out, scenes, kmax, path, n_path, p->max_lateral_accel,
                                             p->repair_full_profile == 1 ? 1 : 0, batch, bounds, status,
                                             (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "convex_space launch");
}

int emp_st_speed_qp_dev(const double *start, const double *st_out, const double *bounds, int batch,
                        const EmpSpeedQpParams *p, double *qp, int32_t *status_iters, void *stream) {
    if (!p) return fail(-1, "EmpSpeedQpParams is NULL");
    if (batch < 0) return fail(-2, "bad batch");
    if (batch == 0) return 0;
    if (!start || !st_out || !bounds || !qp || !status_iters) return fail(-1, "NULL argument");
    emp::SpeedQpLaunch L{*p, batch};
    cudaError_t e = emp::launch_speed_qp(start, st_out, bounds, qp, status_iters, L, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "speed_qp launch");
}

int emp_st_increase_points_dev(const double *qp, int batch, double *dense, int32_t *status, void *stream) {
    if (batch < 0) return fail(-2, "bad batch");
    if (batch == 0) return 0;
    if (!qp || !dense || !status) return fail(-1, "NULL argument");
    cudaError_t e = emp::launch_increase_points(qp, batch, dense, status, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "increase_points launch");
}

int emp_st_path_speed_merge_dev(const double *dense, const double *current_time, const double *path5, int n_path,
                                int batch, double *traj, int32_t *status, void *stream) {
    if (batch < 0 || n_path < 1) return fail(-2, "bad batch / n_path");
    if (batch == 0) return 0;
    if (!dense || !current_time || !path5 || !traj || !status) return fail(-1, "NULL argument");
    cudaError_t e = emp::launch_path_speed_merge(dense, current_time, path5, n_path, batch, traj, status,
                                                 (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "path_speed_merge launch");
}

int emp_st_back_host(EmpPlanner *pl, int batch, const EmpSpeedQpParams *p,
                     const double *st_out, const double *scenes, int kmax, const double *path, int n_path,
                     const double *start, const double *bounds_in, const double *qp_in, const double *dense_in,
                     const double *current_time, const double *path5, int n_path5,
                     double *bounds, int32_t *convex_status, double *qp, int32_t *qp_status_iters,
                     double *dense, int32_t *dense_status, double *traj, int32_t *merge_status) {
    if (!pl || !p) return fail(-1, "NULL argument");
    if (batch <= 0) return batch == 0 ? 0 : fail(-2, "bad batch");
    // which stages run: a stage runs when its output is wanted or a later stage needs it and no *_in replaces it
    const bool want_merge = traj != nullptr;
    const bool run_dense = (dense != nullptr || (want_merge && !dense_in));
    const bool run_qp = (qp != nullptr || (run_dense && !qp_in)) && !(qp_in && !qp);
    const bool run_convex = (bounds != nullptr || (run_qp && !bounds_in)) && !(bounds_in && !bounds);
    if (run_convex && (!st_out || !scenes || !path)) return fail(-1, "convex space needs st_out, scenes and path");
    if (run_qp && (!st_out || !start)) return fail(-1, "speed QP needs st_out and start");
    if (run_dense && !run_qp && !qp_in) return fail(-1, "increase_points needs a QP result");
    if (want_merge && (!current_time || !path5)) return fail(-1, "path_speed_merge needs current_time and path5");
    EMP_CK(cudaSetDevice(pl->device), "cudaSetDevice");
    const size_t B = batch;
    int rc = 0;
    {
        Tmp t(pl->stream);
        const double *d_st = (run_convex || run_qp) ? t.in(st_out, B * EMP_ST_OUT_STRIDE) : nullptr;
        const double *d_bnd = bounds_in ? t.in(bounds_in, B * 4 * EMP_ST_COLS) : nullptr;
        const double *d_qp = qp_in ? t.in(qp_in, B * 4 * EMP_ST_QP_POINTS) : nullptr;
        const double *d_dense = dense_in ? t.in(dense_in, B * 4 * EMP_ST_DENSE) : nullptr;
        int32_t *d_cs = nullptr, *d_qs = nullptr, *d_ds = nullptr, *d_ms = nullptr;
        double *d_traj = nullptr;
        if (run_convex) {
            const double *d_sc = t.in(scenes, B * (2 + 4 * kmax));
            const double *d_path = t.in(path, B * 2 * n_path);
            double *o = t.out<double>(B * 4 * EMP_ST_COLS);
            d_cs = t.out<int32_t>(B);
            if (t.err != cudaSuccess) return fail_cuda(t.err, "st_back staging");
            rc = emp_st_convex_space_dev(d_st, d_sc, kmax, d_path, n_path, p, batch, o, d_cs, pl->stream);
            if (rc) return rc;
            d_bnd = o;
        }
        if (run_qp) {
            const double *d_start = t.in(start, B * 2);
            double *o = t.out<double>(B * 4 * EMP_ST_QP_POINTS);
            d_qs = t.out<int32_t>(B * 2);
            if (t.err != cudaSuccess) return fail_cuda(t.err, "st_back staging");
            EMP_CK(cudaEventRecord(pl->ev0, pl->stream), "event");
            rc = emp_st_speed_qp_dev(d_start, d_st, d_bnd, batch, p, o, d_qs, pl->stream);
            EMP_CK(cudaEventRecord(pl->ev1, pl->stream), "event");
            if (rc) return rc;
            d_qp = o;
        }
        if (run_dense) {
            double *o = t.out<double>(B * 4 * EMP_ST_DENSE);
            d_ds = t.out<int32_t>(B);
            if (t.err != cudaSuccess) return fail_cuda(t.err, "st_back staging");
            rc = emp_st_increase_points_dev(d_qp, batch, o, d_ds, pl->stream);
            if (rc) return rc;
            d_dense = o;
        }
        if (want_merge) {
            if (!d_dense) return fail(-1, "path_speed_merge needs a densified profile");
            const double *d_now = t.in(current_time, B);
            const double *d_p5 = t.in(path5, B * 5 * n_path5);
            d_traj = t.out<double>(B * 7 * EMP_ST_DENSE);
            d_ms = t.out<int32_t>(B);
            if (t.err != cudaSuccess) return fail_cuda(t.err, "st_back staging");
            rc = emp_st_path_speed_merge_dev(d_dense, d_now, d_p5, n_path5, batch, d_traj, d_ms, pl->stream);
            if (rc) return rc;
        }
        if (run_convex) { t.back(bounds, d_bnd, B * 4 * EMP_ST_COLS); t.back(convex_status, d_cs, B); }
        if (run_qp) { t.back(qp, d_qp, B * 4 * EMP_ST_QP_POINTS); t.back(qp_status_iters, d_qs, B * 2); }
        if (run_dense) { t.back(dense, d_dense, B * 4 * EMP_ST_DENSE); t.back(dense_status, d_ds, B); }
        if (want_merge) { t.back(traj, d_traj, B * 7 * EMP_ST_DENSE); t.back(merge_status, d_ms, B); }
        if (t.err != cudaSuccess) return fail_cuda(t.err, "st_back copy-back");
    }
    EMP_CK(cudaStreamSynchronize(pl->stream), "planner stream");
    if (run_qp) cudaEventElapsedTime(&pl->last_ms, pl->ev0, pl->ev1);
    return 0;
}

void *emp_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        fail_cuda(cudaGetLastError(), "emp_host_alloc");
        return nullptr;
    }
    return p;
}

void emp_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

double emp_planner_last_kernel_ms(const EmpPlanner *pl) { return pl ? (double)pl->last_ms : 0.0; }

}  // extern "C"
