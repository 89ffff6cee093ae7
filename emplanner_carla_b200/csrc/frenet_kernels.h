#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace emp {

constexpr int FR_WARPS = 4;   // scenes (warps) per CTA in the Frenet kernels

struct FrenetLaunch {
    int batch, n_nodes, max_points;
};

cudaError_t launch_frenet_project(const double *lines, const double *origin, const double *points,
                                  const int32_t *npts, const double *vel, const double *acc, const double *s_map_in,
                                  const int32_t *pre_match, double *proj, double *s_map, int32_t *match, double *sl,
                                  double *deriv, const FrenetLaunch &p, cudaStream_t st);
cudaError_t launch_frenet_to_cart(const double *lines, const double *s_maps, const double *start_sl,
                                  const double *path_s, const double *path_l, const int32_t *npath, double *cart,
                                  int32_t *ncart, const FrenetLaunch &p, cudaStream_t st);

}  // namespace emp
