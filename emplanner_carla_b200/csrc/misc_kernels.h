#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/emplanner.h"

namespace emp {

struct EdgeParams { double sample_s, w_col, w0, w1, w2, w_ref, danger2, safe2; };

cudaError_t launch_quintic(const double *in, double *out, int n, cudaStream_t st);
cudaError_t launch_obs_cost(const double *d2, int n, int batch, double w, double danger2, double safe2, double *out,
                            cudaStream_t st);
cudaError_t launch_edge_cost(const double *edges, const double *obs_s, const double *obs_l, int n_obs, int n_edges,
                             const EdgeParams &P, double *out, cudaStream_t st);
cudaError_t launch_enrich(const double *node_s, const double *node_l, const double *start, int C, int batch, double res,
                          int n_max, double *out_s, double *out_l, int32_t *n_out, cudaStream_t st);

}  // namespace emp
