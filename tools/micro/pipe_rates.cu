// Issue rates of the fp32 instructions the FFT kernels are made of (sm_100a): scalar FFMA vs packed FFMA2 / FADD2 / FMUL2,
// alone and mixed with shared-memory loads and shuffles.  nvcc -arch=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float fadd(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

constexpr int ITERS = 2048, ILP = 8;
template <int MODE> __global__ void k(float* out, float seed) {
    __shared__ float sm[1024];
    sm[threadIdx.x] = seed;
    __syncthreads();
    u64 a[ILP]; float f[ILP];
    u64 b = ((u64)__float_as_uint(seed) << 32) | __float_as_uint(1.0001f);
#pragma unroll
    for (int i = 0; i < ILP; ++i) { a[i] = b + i; f[i] = seed + i; }
    float fs = seed;
    int idx = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) f[i] = ffma(f[i], fs, fs);
            if (MODE == 1) a[i] = fma2(a[i], b, b);
            if (MODE == 2) a[i] = add2(a[i], b);
            if (MODE == 3) a[i] = mul2(a[i], b);
            if (MODE == 4) f[i] = fadd(f[i], fs);
            if (MODE == 5) { a[i] = fma2(a[i], b, b); f[i] = ffma(f[i], fs, fs); }                  // packed + scalar
            if (MODE == 6) { a[i] = fma2(a[i], b, b); if (i % 4 == 0) f[i] += sm[(idx + i) & 1023]; }   // + 1 LDS per 4
            if (MODE == 7) { a[i] = fma2(a[i], b, b); if (i % 4 == 0) f[i] += __shfl_xor_sync(0xffffffffu, f[i], 1); }
            if (MODE == 8) { a[i] = fma2(a[i], b, b); idx = idx * 3 + i; }                            // + integer IMAD
            if (MODE == 9) { a[i] = fma2(a[i], b, b); idx = (idx + i) ^ it; }                          // + integer ALU
        }
    }
    float s = 0; 
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += f[i] + __uint_as_float((unsigned)a[i]) + __uint_as_float((unsigned)(a[i] >> 32));
    if (s == 123.456f) out[0] = s + idx;
}
template <int MODE> void run(const char* name, double inst_per_iter, int warps_per_smsp) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 4);
    const int threads = 128 * warps_per_smsp;    // one block per SM, warps spread over the 4 SMSPs
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, threads>>>(out, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<sms, threads>>>(out, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double inst = (double)ITERS * ILP * inst_per_iter * warps_per_smsp;     // warp-instructions per SMSP
    printf("%-34s warps/SMSP %2d: %7.3f ms  %6.3f cycles per warp-instruction per SMSP\n", name, warps_per_smsp, ms, cycles / inst);
}
int main() {
    for (int w : {1, 2, 4, 8}) {
        run<0>("FFMA (scalar)", 1, w);
        run<1>("FFMA2 (packed)", 1, w);
        run<2>("FADD2", 1, w);
        run<3>("FMUL2", 1, w);
        run<4>("FADD (scalar)", 1, w);
        run<5>("FFMA2 + FFMA", 2, w);
        run<6>("FFMA2 + LDS/4", 1.25, w);
        run<7>("FFMA2 + SHFL/4", 1.25, w);
        run<8>("FFMA2 + IMAD", 2, w);
        run<9>("FFMA2 + IADD/LOP", 3, w);
    }
    return 0;
}
