// libse_b200.so -- tensor-core mask head (precision = 1):  offset = act(cmvn(x) W^T + b)  [* linears]
//
// tcgen05 (5th-gen tensor core) TF32 GEMM with the accumulator in tensor memory:
//   * CTA tile: 128 rows (frames) x BN columns (BN <= 512, all of Dout when it fits) x 32-deep k-blocks
//   * warps 0-3  producers: read x / W rows with coalesced 128-byte loads, apply the per-utterance CMVN
//                (model.py:30) on the fly, round to TF32 (cvt.rna) and store into the K-major
//                SWIZZLE_128B shared-memory layout the MMA descriptors expect; later the same four
//                warps run the epilogue (tcgen05.ld -> +bias -> activation -> coalesced stores)
//   * warp 4     allocates TMEM and issues tcgen05.mma (one elected lane), committing each stage to
//                its "empty" mbarrier and the finished accumulator to the epilogue barrier
// Operands go through registers instead of TMA on purpose: the rows of the (B, F, K) feature tensor are
// K*4 = 1028 bytes apart (not 16-byte aligned, so no tensor map can describe them), and the CMVN has to
// be applied between the load and the MMA anyway.
#include "se_common.cuh"

using secommon::fail;

namespace {

constexpr int BM = 128, BK = 32, kStagesMax = 4;
constexpr int kProducerWarps = 8, kProducerThreads = kProducerWarps * 32, kThreads = kProducerThreads + 32;   // + 1 MMA warp
constexpr int kMaxWRows = 8;                                      // W rows per producer thread: bn <= 256 -> 8
constexpr unsigned kSpinLimit = 1u << 28;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (unsigned spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (spin > kSpinLimit) __trap();                          // never hang the GPU on a protocol bug
    }
}
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// rows are 128 B apart, 8-row groups 1024 B apart (SBO), 16-byte chunks XOR-swizzled with (row & 7).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                                       // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                             // SBO
    d |= (uint64_t)1 << 46;                                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                       // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {           // kind::tf32, fp32 accumulate, A/B K-major, M = 128
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float activate(float z, int act) {
    if (act == SE_ACT_RELU) return z > 0.f ? z : 0.f;
    if (act == SE_ACT_SIGMOID) return __fdividef(1.0f, 1.0f + __expf(-z));
    return z;
}

struct HeadArgs {
    const float* x; const float* mean; const float* stdv; float cmvn_eps;
    const float* W; const float* bias;
    long long R; int n_frames, Din, Dout, act;
    const float* linears; float* offset_out; float* pred_out;
    long long ldx, ld_stats, ldw, ld_out;   // row strides (floats)
    int vec;         // 1: every stride and base pointer allows 128-bit loads
    int bn;          // columns per CTA (multiple of 16, <= 256)
    int tmem_cols;   // power of two >= bn
    int stages, kblocks;
};

// element (row, kk) of a [rows][32 fp32] K-major SWIZZLE_128B tile
__device__ __forceinline__ int sw128(int row, int kk) { return row * 32 + ((((kk >> 2) ^ (row & 7)) << 2) | (kk & 3)); }

__global__ void __launch_bounds__(kThreads, 2) linear_head_tc_kernel(HeadArgs a) {
    // [stages][A 128x32 | B bn x 32] fp32 (1024-byte aligned: SWIZZLE_128B atoms), epilogue staging, barriers, TMEM slot
    extern __shared__ __align__(1024) float tiles[];
    const int stage_floats = (BM + a.bn) * BK;
    float* stage_out = tiles;                                 // epilogue staging (8 warps x 32 x 33 floats) reuses the operand
                                                              // ring: every MMA has finished reading it by then
    const int ring_floats = max(a.stages * stage_floats, kProducerWarps * 32 * 33);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + ring_floats);
    uint64_t* empty = full + kStagesMax;
    uint64_t* accum_full = empty + kStagesMax;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long r0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * a.bn;

    if (threadIdx.x == 0) {
        if (smem_u32(tiles) & 1023) __trap();                                    // swizzle atoms need the alignment
        for (int s = 0; s < a.stages; ++s) { mbar_init(&full[s], kProducerWarps); mbar_init(&empty[s], 1); }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kProducerWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kProducerWarps) {
        // ===================== producers =====================
        const int t = threadIdx.x;
        if (a.vec) {
            // 128-bit path: thread -> 16-byte chunk c of the k-block, rows rbase + 32 i.  Each quarter-warp reads one
            // row's 128 B; every load of the k-block is issued before the first one is consumed.
            const int c = t & 7, rbase = t >> 3;
            // per-row constants, independent of the k-block: source pointers and the utterance of each A row
            const float* xrow[4];
            int urow[4];
            {
                const long long u0 = r0 / a.n_frames;
                const int rem0 = (int)(r0 - u0 * a.n_frames);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = rbase + 32 * i;
                    xrow[i] = (r0 + row < a.R) ? a.x + (r0 + row) * a.ldx : nullptr;
                    urow[i] = (int)u0 + (rem0 + row) / a.n_frames;
                }
            }
            const float* wrow[kMaxWRows];
#pragma unroll
            for (int i = 0; i < kMaxWRows; ++i) {
                const int row = rbase + 32 * i;
                wrow[i] = (row < a.bn && n0 + row < a.Dout) ? a.W + (long long)(n0 + row) * a.ldw : nullptr;
            }
            for (int kb = 0; kb < a.kblocks; ++kb) {
                const int s = kb % a.stages;
                if (kb >= a.stages) mbar_wait(&empty[s], ((kb / a.stages) - 1) & 1);
                float* As = tiles + s * stage_floats;
                float* Bs = As + BM * BK;
                const int k = kb * BK + 4 * c;
                const bool kin = k < a.Din;
                const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 xa[4], wb[kMaxWRows];
#pragma unroll
                for (int i = 0; i < 4; ++i) xa[i] = (kin && xrow[i]) ? __ldg(reinterpret_cast<const float4*>(xrow[i] + k)) : zero4;
#pragma unroll
                for (int i = 0; i < kMaxWRows; ++i) wb[i] = (kin && wrow[i]) ? __ldg(reinterpret_cast<const float4*>(wrow[i] + k)) : zero4;
                const bool k1 = k + 1 < a.Din, k2 = k + 2 < a.Din, k3 = k + 3 < a.Din;
                int u_cur = -1;
                float4 mu = zero4, inv = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = rbase + 32 * i;
                    float4 v = xa[i];
                    if (a.mean && kin && xrow[i]) {
                        if (urow[i] != u_cur) {
                            u_cur = urow[i];
                            mu = __ldg(reinterpret_cast<const float4*>(a.mean + (long long)u_cur * a.ld_stats + k));
                            const float4 sd = __ldg(reinterpret_cast<const float4*>(a.stdv + (long long)u_cur * a.ld_stats + k));
                            inv = make_float4(__fdividef(1.0f, sd.x + a.cmvn_eps), __fdividef(1.0f, sd.y + a.cmvn_eps),
                                              __fdividef(1.0f, sd.z + a.cmvn_eps), __fdividef(1.0f, sd.w + a.cmvn_eps));
                        }
                        v = make_float4((v.x - mu.x) * inv.x, (v.y - mu.y) * inv.y, (v.z - mu.z) * inv.z, (v.w - mu.w) * inv.w);
                    }
                    v = make_float4(to_tf32(v.x), k1 ? to_tf32(v.y) : 0.f, k2 ? to_tf32(v.z) : 0.f, k3 ? to_tf32(v.w) : 0.f);
                    *reinterpret_cast<float4*>(As + row * 32 + ((c ^ (row & 7)) << 2)) = v;
                }
#pragma unroll
                for (int i = 0; i < kMaxWRows; ++i) {
                    const int row = rbase + 32 * i;
                    if (row < a.bn) {
                        float4 v = wb[i];
                        v = make_float4(to_tf32(v.x), k1 ? to_tf32(v.y) : 0.f, k2 ? to_tf32(v.z) : 0.f, k3 ? to_tf32(v.w) : 0.f);
                        *reinterpret_cast<float4*>(Bs + row * 32 + ((c ^ (row & 7)) << 2)) = v;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the MMA (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        } else {
            // scalar path (dense rows that are only 4-byte aligned): warp -> rows, lane = k within the block
            for (int kb = 0; kb < a.kblocks; ++kb) {
                const int s = kb % a.stages;
                if (kb >= a.stages) mbar_wait(&empty[s], ((kb / a.stages) - 1) & 1);
                float* As = tiles + s * stage_floats;
                float* Bs = As + BM * BK;
                const int k = kb * BK + lane;
                const bool kin = k < a.Din;
                const long long u0 = r0 / a.n_frames;
                const int rem0 = (int)(r0 - u0 * a.n_frames);
                int u_cur = -1;
                float mu = 0.f, inv = 1.f;
#pragma unroll 4
                for (int i = 0; i < BM / kProducerWarps; ++i) {
                    const int row = warp * (BM / kProducerWarps) + i;
                    const long long r = r0 + row;
                    float v = 0.f;
                    if (kin && r < a.R) {
                        v = __ldg(a.x + r * a.ldx + k);
                        if (a.mean) {
                            const int u = (int)u0 + (rem0 + row) / a.n_frames;
                            if (u != u_cur) { u_cur = u; mu = __ldg(a.mean + (long long)u * a.ld_stats + k); inv = __fdividef(1.0f, __ldg(a.stdv + (long long)u * a.ld_stats + k) + a.cmvn_eps); }
                            v = (v - mu) * inv;
                        }
                    }
                    As[sw128(row, lane)] = to_tf32(v);
                }
#pragma unroll 4
                for (int row = warp; row < a.bn; row += kProducerWarps) {
                    const int n = n0 + row;
                    const float w = (kin && n < a.Dout) ? __ldg(a.W + (long long)n * a.ldw + k) : 0.f;
                    Bs[sw128(row, lane)] = to_tf32(w);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
            }
        }
        // ===================== epilogue =====================
        // warp w reads TMEM lanes 32 (w & 3) .. +31 (its quadrant) and every second 32-column chunk
        mbar_wait(accum_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* st = stage_out + warp * 32 * 33;
        const int quad = warp & 3;
        for (int c0 = 32 * (warp >> 2); c0 < a.bn; c0 += 64) {
            if (n0 + c0 >= a.Dout) break;
            uint32_t acc[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, acc);
#pragma unroll
            for (int c = 0; c < 32; ++c) st[lane * 33 + c] = __uint_as_float(acc[c]);
            __syncwarp();
            const int n = n0 + c0 + lane;
            const bool nin = n < a.Dout && (c0 + lane) < a.bn;
            const float bz = (nin && a.bias) ? __ldg(a.bias + n) : 0.f;
            const long long rq = r0 + quad * 32;
            const int rows = (int)(a.R - rq < 32 ? a.R - rq : 32);
            if (nin) {
                float* optr = a.offset_out ? a.offset_out + rq * a.ld_out + n : nullptr;
                float* pptr = a.pred_out ? a.pred_out + rq * a.ld_out + n : nullptr;
                const float* lptr = a.pred_out ? a.linears + rq * a.ld_out + n : nullptr;
#pragma unroll 4
                for (int rr = 0; rr < rows; ++rr) {
                    const float o = activate(st[rr * 33 + lane] + bz, a.act);
                    if (optr) optr[(long long)rr * a.ld_out] = o;
                    if (pptr) pptr[(long long)rr * a.ld_out] = __ldg(lptr + (long long)rr * a.ld_out) * o;
                }
            }
            __syncwarp();
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===================== MMA issuer =====================
        const uint32_t idesc = make_idesc(a.bn);
        for (int kb = 0; kb < a.kblocks; ++kb) {
            const int s = kb % a.stages;
            mbar_wait(&full[s], (kb / a.stages) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t a_addr = smem_u32(tiles + s * stage_floats);
                const uint32_t b_addr = a_addr + BM * BK * 4;
#pragma unroll
                for (int kk = 0; kk < BK / 8; ++kk)
                    umma_tf32(tmem_base, make_desc(a_addr + kk * 32), make_desc(b_addr + kk * 32), idesc, (kb | kk) ? 1u : 0u);
                umma_commit(&empty[s]);                                   // stage reusable once these MMAs have read it
                if (kb == a.kblocks - 1) umma_commit(accum_full);         // accumulator complete
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == kProducerWarps) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
    }
}

}  // namespace

namespace sehead {

int launch_linear_head_tc(const float* x, long long ldx, const float* mean, const float* stdv, long long ld_stats, float cmvn_eps,
                          const float* W, long long ldw, const float* b, long long R, int n_frames, int Din, int Dout, int act,
                          const float* linears, float* offset_out, float* pred_out, long long ld_out, cudaStream_t st) {
    HeadArgs a{};
    a.ldx = ldx; a.ld_stats = ld_stats; a.ldw = ldw; a.ld_out = ld_out;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    a.vec = (ldx % 4 == 0 && ldw % 4 == 0 && al16(x) && al16(W) && (!mean || (ld_stats % 4 == 0 && al16(mean) && al16(stdv)))) ? 1 : 0;
    a.x = x; a.mean = mean; a.stdv = stdv; a.cmvn_eps = cmvn_eps; a.W = W; a.bias = b;
    a.R = R; a.n_frames = n_frames; a.Din = Din; a.Dout = Dout; a.act = act;
    a.linears = linears; a.offset_out = offset_out; a.pred_out = pred_out;
    const int chunks = (Dout + 255) / 256;                       // <= 256 accumulator columns per CTA: two CTAs share an SM's TMEM
    a.bn = ((((Dout + chunks - 1) / chunks) + 15) / 16) * 16;
    a.tmem_cols = 32;
    while (a.tmem_cols < a.bn) a.tmem_cols *= 2;
    a.kblocks = (Din + BK - 1) / BK;
    const size_t stage_bytes = (size_t)(BM + a.bn) * BK * 4;
    const size_t fixed = 256, staging = (size_t)kProducerWarps * 32 * 33 * 4;
    int stages = (int)((110 * 1024 - fixed) / stage_bytes);        // <= ~110 KB so that two CTAs fit per SM
    a.stages = stages > kStagesMax ? kStagesMax : stages;
    if (a.stages > a.kblocks) a.stages = a.kblocks;
    if (a.stages < 1) return fail(SE_ERR_UNSUPPORTED, "head tile does not fit in shared memory (Dout=%d)", Dout);
    const size_t ring = (size_t)a.stages * stage_bytes;
    const size_t smem = (ring > staging ? ring : staging) + fixed;
    static unsigned long long opted = 0;                       // per device: cudaFuncSetAttribute is not process-wide
    if (secommon::first_use_on_device(opted)) {
        SE_CUDA_CHECK(cudaFuncSetAttribute(linear_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    }
    dim3 grid((unsigned)((R + BM - 1) / BM), (unsigned)chunks);
    linear_head_tc_kernel<<<grid, kThreads, smem, st>>>(a);
    return secommon::check_launch("linear_head_tc_kernel");
}

}  // namespace sehead
