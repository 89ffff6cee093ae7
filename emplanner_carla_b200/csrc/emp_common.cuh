// Shared device helpers for libemplanner (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define EMP_FULL_MASK 0xffffffffu

namespace emp {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk -> SASS UBLKCP) -------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(phase)
        : "memory");
}

// ---- 64-bit warp shuffles / reductions --------------------------------------
__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(EMP_FULL_MASK, v, src); }
__device__ __forceinline__ double shfl_down(double v, int d) { return __shfl_down_sync(EMP_FULL_MASK, v, d); }
__device__ __forceinline__ double shfl_up(double v, int d) { return __shfl_up_sync(EMP_FULL_MASK, v, d); }
__device__ __forceinline__ double shfl_xor(double v, int m) { return __shfl_xor_sync(EMP_FULL_MASK, v, m); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor(v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, shfl_xor(v, o));
    return v;
}

// ---- lane-to-lane communication policy of the banded QP kernels (sl_qp.cu, smooth_qp.cu) ---------------------------
// How the lanes (stations) of one frame talk to each other.  Lanes<32>: one warp per frame, shuffles.  Lanes<64> /
// Lanes<128>: one CTA per frame, one thread per lane, every exchange goes through a double-buffered shared-memory
// line and one __syncthreads() - the general-size path (Quadratic_planning with more than 32 stations, e.g.
// sampling_res = 1 -> 46); ~10x the latency of the warp path, same arithmetic.
template <int NL>
struct Lanes {
    int lane;
    double *scr;                  // [2][NL] doubles: double-buffered exchange line
    int flip;
    __device__ __forceinline__ Lanes(int lane_, double *scr_) : lane(lane_), scr(scr_), flip(0) {}
    __device__ __forceinline__ double *next() { flip ^= 1; return scr + flip * NL; }
    __device__ __forceinline__ double get(double v, int src) {
        double *b = next();
        b[lane] = v;
        __syncthreads();
        return b[src];
    }
    __device__ __forceinline__ double up(double v, int d) { return get(v, lane >= d ? lane - d : lane); }
    __device__ __forceinline__ double down(double v, int d) { return get(v, lane + d < NL ? lane + d : lane); }
    __device__ __forceinline__ double sum(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(EMP_FULL_MASK, v, o);
        double *b = next();
        if ((lane & 31) == 0) b[lane >> 5] = v;
        __syncthreads();
        double t = b[0];
#pragma unroll
        for (int w = 1; w < NL / 32; ++w) t += b[w];
        return t;
    }
    __device__ __forceinline__ double max(double v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(EMP_FULL_MASK, v, o));
        double *b = next();
        if ((lane & 31) == 0) b[lane >> 5] = v;
        __syncthreads();
        double t = b[0];
#pragma unroll
        for (int w = 1; w < NL / 32; ++w) t = fmax(t, b[w]);
        return t;
    }
    template <int N>
    __device__ __forceinline__ void sum_n(double (&v)[N]) {
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = sum(v[k]);
    }
    __device__ __forceinline__ int argmin(double v) {      // first minimum (np.argmin)
        int idx = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(EMP_FULL_MASK, v, o);
            const int oi = __shfl_xor_sync(EMP_FULL_MASK, idx, o);
            if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        double *b = next();
        if ((lane & 31) == 0) { b[2 * (lane >> 5)] = v; b[2 * (lane >> 5) + 1] = (double)idx; }
        __syncthreads();
        double bv = b[0];
        int bi = (int)b[1];
#pragma unroll
        for (int w = 1; w < NL / 32; ++w)
            if (b[2 * w] < bv) { bv = b[2 * w]; bi = (int)b[2 * w + 1]; }      // ascending warps: ties keep the lower index
        return bi;
    }
    __device__ __forceinline__ bool any(bool p) { return __syncthreads_or(p ? 1 : 0) != 0; }
    __device__ __forceinline__ void sync() { __syncthreads(); }
};

template <>
struct Lanes<32> {                 // stateless: nothing of it lives in registers or local memory
    __device__ __forceinline__ Lanes(int, double *) {}
    __device__ __forceinline__ double get(double v, int src) { return __shfl_sync(EMP_FULL_MASK, v, src); }
    __device__ __forceinline__ double up(double v, int d) { return __shfl_up_sync(EMP_FULL_MASK, v, d); }
    __device__ __forceinline__ double down(double v, int d) { return __shfl_down_sync(EMP_FULL_MASK, v, d); }
    __device__ __forceinline__ double sum(double v) { return warp_sum(v); }
    __device__ __forceinline__ double max(double v) { return warp_max(v); }
    template <int N>
    __device__ __forceinline__ void sum_n(double (&v)[N]) {
        _Pragma("unroll")
        for (int o = 16; o > 0; o >>= 1) {
            _Pragma("unroll")
            for (int k = 0; k < N; ++k) v[k] += shfl_xor(v[k], o);
        }
    }
    __device__ __forceinline__ int argmin(double v) {      // first minimum over lanes (np.argmin semantics)
        int idx = threadIdx.x & 31;
        _Pragma("unroll")
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = shfl_xor(v, o);
            const int oi = __shfl_xor_sync(EMP_FULL_MASK, idx, o);
            if (ov < v || (ov == v && oi < idx)) {
                v = ov;
                idx = oi;
            }
        }
        return idx;
    }
    __device__ __forceinline__ bool any(bool p) { return __any_sync(EMP_FULL_MASK, p) != 0; }
    __device__ __forceinline__ void sync() { __syncwarp(); }
};

}  // namespace emp
