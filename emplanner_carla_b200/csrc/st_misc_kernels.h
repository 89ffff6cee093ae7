#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/emplanner.h"

namespace emp {
cudaError_t launch_st_graph(const double *in, int n, int batch, double *out, cudaStream_t st);
cudaError_t launch_start_condition(const double *in, int batch, double *out, cudaStream_t st);
cudaError_t launch_st_cost(int op, const double *items, int n_items, const double *obs4, int n_obs,
                           const double *sdot_tab, double v0, const EmpStParams &P, double *out, cudaStream_t st);
cudaError_t launch_dy_obs_deri(const double *in, int n, int batch, double *out, cudaStream_t st);
cudaError_t launch_index2s(const double *xy, int n, int batch, double *out, cudaStream_t st);
cudaError_t launch_frenet2cart(const double *sl, int n_pts, const double *path, const double *s_map, int n_nodes,
                               int proj_only, double *out, cudaStream_t st);
}  // namespace emp
