// Shared host/device helpers of libse_b200.so: error reporting across the C ABI,
// CUDA call checking, warp/block reductions.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../include/se_b200.h"

namespace secommon {

inline char* last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define SE_CUDA_CHECK(expr)                                                                         \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            return secommon::fail(SE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define SE_REQUIRE(cond, ...)                                              \
    do {                                                                   \
        if (!(cond)) return secommon::fail(SE_ERR_BAD_ARG, __VA_ARGS__);   \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SE_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return SE_OK;
}

#if defined(__CUDACC__)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum NS per-thread values over the CTA (as doubles) and add the totals to dst[0..NS).
// All threads of the CTA must call it.
template <int NS, typename T>
__device__ __forceinline__ void block_accumulate_to(const T* acc, double* dst) {
    __shared__ double scratch[NS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        double v = warp_sum((double)acc[i]);
        if (lane == 0) scratch[i][warp] = v;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            double v = lane < nwarps ? scratch[i][lane] : 0.0;
            v = warp_sum(v);
            if (lane == 0 && v != 0.0) atomicAdd(dst + i, v);
        }
    }
    __syncthreads();
}
#endif

}  // namespace secommon
