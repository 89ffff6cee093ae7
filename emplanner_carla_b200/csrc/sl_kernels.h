// Internal launch descriptors shared by the kernels and the C-ABI glue.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/emplanner.h"

#define EMP_PLAN_HDR 8        // header words of a plan record (include/emplanner.h)
#define EMP_SCENE_HDR 6       // header words of a scene record

namespace emp {

#ifndef EMP_DP_WARPS
#define EMP_DP_WARPS 4
#endif
#ifndef EMP_QP_WARPS
#define EMP_QP_WARPS 4
#endif
constexpr int DP_WARPS = EMP_DP_WARPS;   // frames (warps) per CTA in the DP kernel
constexpr int QP_WARPS = EMP_QP_WARPS;   // frames (warps) per CTA in the QP kernel

constexpr int DP_PAIR_CAND = 16;   // candidates per chunk of the DP kernel's (candidate, sample) pair evaluation (>= EMP_MAX_ROW)

struct DpLaunch {
    double res, w_col, w0, w1, w2, w_ref, sample_s, sample_l;
    int R, C, kmax, n_enr_max, scene_stride, plan_stride, batch;
    // shared-memory carve-up (bytes)
    int off_slab, off_f0, off_nl, off_refc, off_ltab, off_warp, warp_bytes;
    int w_off_cand, w_off_cost, w_off_rows, w_off_offs, w_off_parent, w_off_sj, w_off_dlon, w_off_k0, w_off_wl, w_off_pv;
    unsigned smem_bytes;
    const int32_t *order;     // optional [batch]: work slot -> frame (launch_sl_order); NULL = natural order
};

enum { QP_MODE_FUSED = 0,      // path from the plan record ([::2]), corridor + QP
       QP_MODE_DIRECT = 1,     // corridor given by the caller, QP only
       QP_MODE_CORRIDOR = 2 }; // path given by the caller, corridor only

struct QpLaunch {
    EmpQpParams q;
    int batch, kmax, scene_stride, plan_stride, n_enr_max;
    int n_direct;             // stations per frame in the DIRECT / CORRIDOR modes
    int n_fused_max;          // FUSED mode: upper bound on the stations of any frame (selects the warp / block path)
    int mode;                 // QP_MODE_*
    const int32_t *order;     // optional [batch]: ticket -> frame, heaviest frames first (launch_sl_order)
    int n_peers;              // fused all-gather: number of peer slices to push each plan record to
    double *peer[EMP_MAX_PEERS];
    // arrival flags of the fused all-gather: when the last plan record of the launch has been stored, `epoch` is
    // written (release, system scope) to this rank's word in every peer's flag array and in its own (0 = no flags)
    unsigned long long epoch;
    unsigned long long *self_ready;
    unsigned long long *peer_ready[EMP_MAX_PEERS];
};

// Consumer side of the fused all-gather (csrc/sl_qp.cu): spin (acquire, system scope) until flags[r] >= epoch for all
// r < n_wait, then write `epoch` to every signal[r] (release, system scope).  *err is set to 1 after ~timeout_ms.
cudaError_t launch_flag_wait_signal(const unsigned long long *flags, int n_wait, unsigned long long epoch,
                                    unsigned long long *const *signal, int n_signal, int *err, int timeout_ms,
                                    cudaStream_t st);

// order[0..batch) = frame indices sorted by obstacle count, descending (counting sort, one CTA).  Both kernels' cost
// grows with the obstacle count (DP: exact evaluations per column; QP: tighter corridors, more iterations), and the
// slowest frame alone takes ~2/3 of a whole 4096-frame launch: starting the heavy frames first and together is worth
// 15 % (DP) and 10 % (QP) of the launch (profiles/r2_sched_probe.txt).
cudaError_t launch_sl_order(const double *scenes, int scene_stride, int kmax, int batch, int32_t *order, cudaStream_t st);
cudaError_t launch_sl_dp(const double *scenes, double *plans, int32_t *rows, const DpLaunch &p, cudaStream_t st);
cudaError_t launch_sl_qp(const double *scenes, double *plans, double *lmin_lmax, const double *lmin_in,
                         const double *lmax_in, const double *start_in, double *out_x, int32_t *status_iters,
                         const QpLaunch &p, cudaStream_t st);
cudaError_t launch_fp64_peak(int iters, double *tflops, double *ms);

}  // namespace emp
