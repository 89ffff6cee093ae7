// FP64 FMA micro-benchmark: the compute roofline denominator for the S-L DP/QP
// kernels (MEASURED_PEAKS.json only carries HBM and bf16 numbers).
#include "emp_common.cuh"
#include "sl_kernels.h"

namespace emp {
namespace {

__global__ void __launch_bounds__(256) fp64_fma_kernel(double *out, int iters, double a, double b) {
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = fma(v[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += v[k];
    if (s == 123.456) out[0] = s;      // never true; keeps the chain alive
}

}  // namespace

cudaError_t launch_fp64_peak(int iters, double *tflops, double *ms_out) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = sms * 8, block = 256;
    fp64_fma_kernel<<<grid, block>>>(d, 64, 0.999, 1e-3);           // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        fp64_fma_kernel<<<grid, block>>>(d, iters, 0.999, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    e = cudaGetLastError();
    const double flops = 2.0 * 8.0 * (double)iters * (double)grid * (double)block;
    if (tflops) *tflops = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return e;
}

}  // namespace emp
