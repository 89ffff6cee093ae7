#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/emplanner.h"

namespace emp {
cudaError_t launch_cycle_path_points(const double *plans, int plan_stride, int n_enr_max, const double *scenes,
                                     int scene_stride, int batch, int max_points, double *xy, int32_t *npts,
                                     cudaStream_t st);
cudaError_t launch_cycle_path_tables(const double *cart, const int32_t *npts, int max_points, int batch, int n_path,
                                     double *path2, double *path5, cudaStream_t st);
}  // namespace emp
