// Do the operand selectors of the packed fp32 instructions (half swap, broadcast, per-half negation) cost pipe cycles?
// nvcc -arch=sm_100a -O3 -o packed_swizzle packed_swizzle.cu && ./packed_swizzle
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
constexpr int ITERS = 2048, ILP = 8;
template <int MODE> __global__ void k(float* out, float seed) {
    u64 a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = pk2(seed + i, seed - i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const u64 o = a[(i + 3) % ILP];
            float lo, hi; upk(o, lo, hi);
            if (MODE == 0) a[i] = add2(a[i], o);                                 // plain
            if (MODE == 1) a[i] = add2(a[i], pk2(hi, lo));                       // halves swapped
            if (MODE == 2) a[i] = add2(a[i], pk2(lo, lo));                       // broadcast
            if (MODE == 3) a[i] = add2(a[i], pk2(-hi, lo));                      // swap + negate one half (multiply by i)
            if (MODE == 4) a[i] = fma2(pk2(hi, hi), pk2(-lo, lo), a[i]);         // the second half of a complex multiply
            if (MODE == 5) a[i] = fma2(o, o, a[i]);                              // plain fma
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { float lo, hi; upk(a[i], lo, hi); s += lo + hi; }
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char* name, int warps_per_smsp) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, 4);
    const int threads = 128 * warps_per_smsp;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, threads>>>(out, 1.0f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<sms, threads>>>(out, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-44s warps/SMSP %d: %7.3f ms  %6.3f cycles per packed op per SMSP\n", name, warps_per_smsp, ms,
           cycles / ((double)ITERS * ILP * warps_per_smsp));
}
int main() {
    for (int w : {2, 8}) {
        run<0>("FADD2 plain", w);
        run<1>("FADD2 swapped halves", w);
        run<2>("FADD2 broadcast", w);
        run<3>("FADD2 swap + negate (x i)", w);
        run<4>("FFMA2 (a.y,a.y)*(-b.y,b.x)+c", w);
        run<5>("FFMA2 plain", w);
    }
    return 0;
}
