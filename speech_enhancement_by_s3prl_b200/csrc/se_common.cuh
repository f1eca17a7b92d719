// Shared host/device helpers of libse_b200.so: error reporting across the C ABI,
// CUDA call checking, warp/block reductions.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
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

// Programmatic-dependent-launch mask (experiments): SE_B200_PDL bit 0 = head, bit 1 = mask->iSTFT, bit 2 = finalize; default all.
inline int pdl_mask() {
    static int m = -1;
    if (m < 0) { const char* e = getenv("SE_B200_PDL"); m = e ? atoi(e) : 7; }
    return m;
}

// Per-device one-time setup (cudaFuncSetAttribute is per device): true the first time it is called on the current device
// for the given flag word (one static word per kernel family).
inline bool first_use_on_device(unsigned long long& seen) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ULL << (dev & 63);
    if (seen & bit) return false;
    seen |= bit;
    return true;
}

// SM count of the current device (cached per device)
inline int device_sms() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& v = sms[dev & 63];
    if (v == 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
    }
    return v;
}

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(SE_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return SE_OK;
}

// Optional CTA timeline (debugging aid, se_set_trace): when a trace buffer is set, every CTA of the fused-step kernels
// appends {kernel id << 32 | blockIdx.x, SM id, globaltimer at start, globaltimer at end} to it.
inline unsigned long long*& trace_ptr() {
    static unsigned long long* p = nullptr;
    return p;
}

#if defined(__CUDACC__)
struct TraceScope {
    unsigned long long* buf;
    unsigned long long t0;
    int kernel_id;
    __device__ __forceinline__ static unsigned long long now() {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        return t;
    }
    __device__ __forceinline__ TraceScope(unsigned long long* b, int id) : buf(b), t0(0), kernel_id(id) {
        if (buf && threadIdx.x == 0) t0 = now();
    }
    __device__ __forceinline__ void mark(int id) {        // intermediate timestamp (calling thread), record id = `id`
        if (buf) {
            const unsigned long long i = atomicAdd(buf, 1ULL);
            unsigned long long* r = buf + 1 + 4 * i;
            r[0] = ((unsigned long long)id << 32) | blockIdx.x;
            r[1] = 0;
            r[2] = t0;
            r[3] = now();
        }
    }
    __device__ __forceinline__ void finish() {           // call from thread 0 when the CTA's work is done
        if (buf && threadIdx.x == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            const unsigned long long i = atomicAdd(buf, 1ULL);
            unsigned long long* r = buf + 1 + 4 * i;
            r[0] = ((unsigned long long)kernel_id << 32) | blockIdx.x;
            r[1] = smid;
            r[2] = t0;
            r[3] = now();
        }
    }
};

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
