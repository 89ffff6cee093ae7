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

struct DpLaunch {
    double res, w_col, w0, w1, w2, w_ref, sample_s, sample_l;
    int R, C, kmax, n_enr_max, scene_stride, plan_stride, batch;
    // shared-memory carve-up (bytes)
    int off_slab, off_f0, off_nl, off_refc, off_ltab, off_warp, warp_bytes;
    int w_off_cand, w_off_cost, w_off_rows, w_off_offs, w_off_parent, w_off_sj, w_off_dlon, w_off_k0, w_off_wl;
    unsigned smem_bytes;
};

enum { QP_MODE_FUSED = 0,      // path from the plan record ([::2]), corridor + QP
       QP_MODE_DIRECT = 1,     // corridor given by the caller, QP only
       QP_MODE_CORRIDOR = 2 }; // path given by the caller, corridor only

struct QpLaunch {
    EmpQpParams q;
    int batch, kmax, scene_stride, plan_stride, n_enr_max;
    int n_direct;             // stations per frame in the DIRECT / CORRIDOR modes
    int mode;                 // QP_MODE_*
    int n_peers;              // fused all-gather: number of peer slices to push each plan record to
    double *peer[EMP_MAX_PEERS];
};

cudaError_t launch_sl_dp(const double *scenes, double *plans, int32_t *rows, const DpLaunch &p, cudaStream_t st);
cudaError_t launch_sl_qp(const double *scenes, double *plans, double *lmin_lmax, const double *lmin_in,
                         const double *lmax_in, const double *start_in, double *out_x, int32_t *status_iters,
                         const QpLaunch &p, cudaStream_t st);
cudaError_t launch_fp64_peak(int iters, double *tflops, double *ms);

}  // namespace emp
