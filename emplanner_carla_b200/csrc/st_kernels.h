#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/emplanner.h"

namespace emp {

#ifndef EMP_ST_THREADS
#define EMP_ST_THREADS 256
#endif
constexpr int ST_THREADS = EMP_ST_THREADS;

struct StLaunch {
    EmpStParams prm;
    int batch, kmax, scene_stride;
    double lnw_hi, lnw_lo;            // ln(w_obs), filled by launch_st_dp
};

cudaError_t launch_st_dp(const double *scenes, double *out, double *cost_tab, double *sdot_tab, int32_t *node_tab,
                         const StLaunch &p, cudaStream_t st);

}  // namespace emp
