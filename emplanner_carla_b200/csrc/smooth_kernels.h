#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/emplanner.h"

namespace emp {

constexpr int SMO_WARPS = 4;

struct SmoothLaunch {
    EmpSmoothParams prm;
    int batch, max_points, heading_only;
};

cudaError_t launch_heading_kappa(const double *xy, const int32_t *npts, double *out, int32_t *status_iters, int batch,
                                 int max_points, cudaStream_t st);
cudaError_t launch_smooth(const double *xy, const int32_t *npts, double *out, int32_t *status_iters,
                          const SmoothLaunch &p, cudaStream_t st);

}  // namespace emp
