#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/emplanner.h"

namespace emp {

#ifndef EMP_SQP_WARPS
#define EMP_SQP_WARPS 1
#endif
constexpr int SQP_WARPS = EMP_SQP_WARPS;   // frames per CTA of the speed-QP kernel (1: every shared-memory address is an immediate)
constexpr int SQP_MAXQ = 16;          // stations (qp_size <= 16: the DP profile holds 16 entries, one is the NaN terminator)
constexpr int SQP_NV = 3 * SQP_MAXQ;  // variables
constexpr int SQP_NE = 2 * (SQP_MAXQ - 1);
constexpr int SQP_YS = SQP_NE + 1;    // row stride of L^-1 A' (odd: conflict-free 64-bit rows and columns)
constexpr int SQP_KS = SQP_NE + 1;

struct SpeedQpLaunch {
    EmpSpeedQpParams prm;
    int batch;
};

size_t speed_qp_smem_bytes();
cudaError_t launch_speed_qp(const double *start, const double *st_out, const double *bounds, double *out,
                            int32_t *status_iters, const SpeedQpLaunch &p, cudaStream_t st);
cudaError_t launch_convex_space(const double *st_out, const double *scenes, int kmax, const double *path, int n_path,
                                double max_lateral_accel, int repair_full, int batch, double *bounds, int32_t *status,
                                cudaStream_t st);
cudaError_t launch_increase_points(const double *qp, int batch, double *dense, int32_t *status, cudaStream_t st);
cudaError_t launch_path_speed_merge(const double *dense, const double *current_time, const double *path, int n_path,
                                    int batch, double *traj, int32_t *status, cudaStream_t st);

}  // namespace emp
