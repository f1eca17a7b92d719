// libse_b200.so -- n_fft = 512 fast paths: one frame per half-warp, FFT values in registers
// (fft256_warp.cuh), frames gathered straight from global memory with coalesced 8-byte loads
// (each half-warp reads the 2 KB of its frame as 16 x 128 B), spectra written straight from
// registers.  No block-level barrier in the main loops: half-warps run independently.
//
//   stft512_kernel        K1: frame -> window -> rFFT -> power / phase / log-power
//   mask_istft512_kernel  K3: frame -> rFFT -> x sqrt(mask) -> irFFT -> window -> overlap-add in
//                             registers along a run of consecutive frames -> / envelope -> wav,
//                             plus the per-utterance metric sums
//
// Both kernels are bound by instruction issue, not by HBM, so the code is organised
// to keep the instruction count and the instruction footprint down:
//   * the analysis window is pre-scaled by 1/2, which removes the 1/2 of the real-FFT split;
//     X[k] = E + T and X[M-k] = conj(E - T) share one complex multiply (T = W^k * (-i)(Z[k] - conj Z[M-k]))
//   * the inverse transform reuses the forward FFT code: IFFT(Z) = conj(FFT(conj Z)), with the
//     conjugations, the 1/M, the synthesis window and the 1/envelope folded into one table
//   * in K3 the three transforms of a frame (clean, noisy, inverse) run through ONE copy of the FFT
//     code inside a non-unrolled pass loop, so the kernel fits the instruction cache
//   * rarely taken paths (reflect padding at utterance edges, unaligned rows) go through a small
//     rolled loop into the half-warp's shared-memory buffer instead of being unrolled in registers
#include <cstdlib>
#include "se_common.cuh"
#include "fft256_warp.cuh"
#include "fast_common.cuh"
#include "tile_kernels.cuh"

using namespace fft256w;
using namespace fastc;
using sekern::StftArgs;
using sekern::MaskIstftArgs;

namespace {

constexpr int N = 512, H = 256;

// stage `count` floats (multiple of 4) of a waveform row starting at original coordinate t0 into dst (16-byte
// aligned shared memory): 16-byte cp.async when the span is interior and aligned, else a rolled reflect loop
__device__ __forceinline__ void stage_wave(float* __restrict__ dst, const float* __restrict__ row, int T, int t0, int count, int j) {
    const float* src = row + t0;
    if ((t0 >= 0) && (t0 + count <= T) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        for (int c = j; c < count / 4; c += 16) cp_async16(dst + 4 * c, src + 4 * c);
    } else {
        // utterance edges (reflection) and rows that are not 16-byte aligned: element-wise, but still asynchronous -- a
        // synchronous loop here serialises ~32 L2 round trips and stretches the whole CTA (seen as a bimodal CTA timeline)
#pragma unroll 4
        for (int i = j; i < count; i += 16) {
            int t = t0 + i;
            t = t < 0 ? -t : t;
            t = t >= T ? 2 * (T - 1) - t : t;
            cp_async4(dst + i, row + t);
        }
    }
}
// stage a row of `count` floats (any alignment) that needs no reflection
__device__ __forceinline__ void stage_row(float* __restrict__ dst, const float* __restrict__ src, int count, int j, bool padded) {
    if (padded && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {                              // padded: row stride >= round4(count)
        const int n16 = (count + 3) / 4;
#pragma unroll
        for (int c = 0; c < 5; ++c)                                                              // count <= 320 here (K = 257)
            if (j + 16 * c < n16) cp_async16(dst + 4 * (j + 16 * c), src + 4 * (j + 16 * c));
    } else {
#pragma unroll 4
        for (int i = j; i < count; i += 16) cp_async4(dst + i, src + i);
    }
}

// v[r] = (x[2m], x[2m+1]) * win2[m], m = j + 16 r, from a staged frame
__device__ __forceinline__ void frame_from_stage(const float* __restrict__ st, int j, const float2* __restrict__ win2, float2 (&v)[16]) {
    const float2* s2 = reinterpret_cast<const float2*>(st);
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        v[r] = pmul(s2[j + 16 * r], win2[j + 16 * r]);
    }
}

// stage H floats of a waveform row starting at original coordinate t0 (reflect outside [0, T))
__device__ __forceinline__ void stage_half(float* __restrict__ dst, const float* __restrict__ row, int T, int t0, int j) {
    const float* src = row + t0;
    if ((t0 >= 0) && (t0 + H <= T) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
        for (int c = 0; c < H / 64; ++c) cp_async16(dst + 4 * (j + 16 * c), src + 4 * (j + 16 * c));
    } else {
#pragma unroll 4
        for (int i = j; i < H; i += 16) {
            int t = t0 + i;
            t = t < 0 ? -t : t;
            t = t >= T ? 2 * (T - 1) - t : t;
            cp_async4(dst + i, row + t);
        }
    }
}
// v[r] = (x[2m], x[2m+1]) * win2[m], m = j + 16 r: r < 8 from the first-half slot, r >= 8 from the second-half slot
__device__ __forceinline__ void frame_from_slots(const float* __restrict__ first, const float* __restrict__ second, int j,
                                                 const float2* __restrict__ win2, float2 (&v)[16]) {
    const float2* f2 = reinterpret_cast<const float2*>(first);
    const float2* s2 = reinterpret_cast<const float2*>(second);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        v[r] = pmul(f2[j + 16 * r], win2[j + 16 * r]);
        v[r + 8] = pmul(s2[j + 16 * r], win2[j + 16 * r + 128]);
    }
}

// ------------------------------------------------------------------ K1
constexpr int kWarps1 = 8, kThreads1 = kWarps1 * 32;

// dynamic shared memory of K1: transpose buffers [16][256] float2, frame staging [16][2][512] float, window pairs
constexpr size_t kSmem1 = (size_t)(kWarps1 * 2) * (M * 8 + 2 * N * 4) + M * 8;


// write the requested outputs of one frame (lane j of a half-warp holds Z[j + 16 q] and the mirrored bins);
// `acc` (STATS): the half-warp's private (sum x, sum x^2) accumulators in shared memory, updated for the feature written
// STATS: sacc = the half-warp's (sum x, sum x^2) accumulators in REGISTERS: slot q <-> bin j + 16 q, slot 8 + q <-> bin
// 256 - (j + 16 q), slot 16 <-> bin 128 (lane 0)
template <bool POWER, bool PHASE, bool LOGP, bool STATS>
__device__ __forceinline__ void emit_frame(const float2 (&v)[16], const float2 (&zm)[8], const float2 (&twn)[8], int j,
                                           const StftArgs& a, long long o, float2 (&sacc)[17]) {
    float* pw = POWER ? a.power + o : nullptr;
    float* lg = LOGP ? a.logp + o : nullptr;
    float* ph = PHASE ? a.phase + o : nullptr;
    auto stat = [&](int slot, float x) {
        if (STATS) sacc[slot] = pfma(make_float2(x, x), make_float2(1.0f, x), sacc[slot]);
    };
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int k = j + 16 * q;
        float2 xa, xb;
        split_pair(v[q], zm[q], twn[q], xa, xb);
        const float pa = xa.x * xa.x + xa.y * xa.y, pb = xb.x * xb.x + xb.y * xb.y;
        if (POWER) { pw[k] = pa; pw[M - k] = pb; }
        float la = 0.f, lb = 0.f;
        if (LOGP) { la = fast_log(pa + a.log_eps); lb = fast_log(pb + a.log_eps); lg[k] = la; lg[M - k] = lb; }
        if (PHASE) { ph[k] = atan2f(k == 0 ? 0.0f : xa.y, xa.x); ph[M - k] = atan2f(k == 0 ? 0.0f : xb.y, xb.x); }
        stat(q, LOGP ? la : pa);
        stat(8 + q, LOGP ? lb : pb);
    }
    if (j == 0) {                                               // k = 128 pairs with itself: X = 2 conj(Z[128])
        const float2 x = make_float2(2.0f * v[8].x, -2.0f * v[8].y);
        const float p = x.x * x.x + x.y * x.y;
        float l = 0.f;
        if (POWER) pw[128] = p;
        if (LOGP) { l = fast_log(p + a.log_eps); lg[128] = l; }
        if (PHASE) ph[128] = atan2f(x.y, x.x);
        stat(16, LOGP ? l : p);
    }
}

// General hop: frames are dealt to CTAs in contiguous ranges [blockIdx.x * per_cta, ...); iteration `it` of a CTA transforms
// the 16 consecutive frames first + 16 it + hw (one per half-warp), whole frames double-buffered through cp.async.
template <bool POWER, bool PHASE, bool LOGP>
__global__ void __launch_bounds__(kThreads1, 2) stft512_kernel(StftArgs a, long long total_frames, int per_cta) {
    extern __shared__ __align__(16) unsigned char smem1[];
    const int lane = threadIdx.x & 31, j = lane & 15;
    const int hw = (threadIdx.x >> 4);
    float2* xbuf = reinterpret_cast<float2*>(smem1) + hw * M;
    float* stage = reinterpret_cast<float*>(smem1 + (size_t)(kWarps1 * 2) * M * 8) + hw * 2 * N;
    float2* s_win2 = reinterpret_cast<float2*>(smem1 + (size_t)(kWarps1 * 2) * (M * 8 + 2 * N * 4));
    for (int i = threadIdx.x; i < M; i += kThreads1) s_win2[i] = make_float2(0.5f * a.tab.window[2 * i], 0.5f * a.tab.window[2 * i + 1]);
    float2 tw[15], twn[8];
    load_lane_constants(j, a.tab.twM, a.tab.twN, tw, twn);
    __syncthreads();
    griddep_launch();
    const int total = (int)total_frames;                              // < 2^31 (checked by the launcher): 32-bit index math
    const int cta_lo = blockIdx.x * per_cta;
    const int cta_hi = cta_lo + per_cta < total ? cta_lo + per_cta : total;
    const unsigned hmask = half_mask(lane);
    auto prefetch = [&](int gg, int buf) {
        const int u = gg / a.n_frames, f = gg - u * a.n_frames;
        stage_wave(stage + buf * N, a.wav + (long long)u * a.utt_stride, a.T, f * a.hop - N / 2, N, j);
        cp_async_commit();
    };
    if (cta_lo + hw < cta_hi) prefetch(cta_lo + hw, 0);
    int buf = 0;
#pragma unroll 1
    for (int gg = cta_lo + hw; gg < cta_hi; gg += kThreads1 / 16, buf ^= 1) {
        if (gg + kThreads1 / 16 < cta_hi) { prefetch(gg + kThreads1 / 16, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp(hmask);
        float2 v[16];
        frame_from_stage(stage + buf * N, j, s_win2, v);
        fft256<-1>(v, xbuf, j, tw, hmask);
        float2 zm[8];
        fetch_mirror(v, lane, zm);
        float2 no_stats[17];
        emit_frame<POWER, PHASE, LOGP, false>(v, zm, twn, j, a, (long long)gg * a.spec_stride, no_stats);
        __syncwarp(hmask);                                          // stage[buf] is free for the prefetch after next
    }
}

// hop = 256 (= N/2): every half-warp owns a RUN of consecutive frames of one utterance (F / rpu frames, the first F mod rpu
// runs one more).  The kernel is bound by shared-memory wavefronts and the fp32 pipe in equal parts (tools/micro/fft_loop.cu:
// the transform alone is 62 SM-cycles, its transpose 34 of them), so everything except the transpose stays out of shared
// memory: the frame's raw samples live in REGISTERS -- lane j holds the pairs m = j + 16 r -- as two half-frame buffers;
// frame f reads (first, second) = (A, B), and as soon as it is windowed the dead first buffer receives the second half of
// frame f+1 straight from global memory (coalesced 8-byte loads, a whole transform to land behind); the next frame reads
// (B, A).  The window (pre-scaled by 1/2) and the CMVN accumulators are per-lane registers too.  No block-level barrier in
// the frame loop.
// STATS: the sum and the sum of squares per (utterance, bin) of the feature written (log-power if LOGP, else power) -- the
// CMVN statistics of the mask head (model.py:30) -- are accumulated in registers (fp32 over the few frames of a run), parked
// in the half-warp's shared-memory row at the end, combined over the CTA's half-warps in double precision after ONE barrier,
// and added to stat_sums with one double atomicAdd pair per (CTA, utterance, bin).
#ifndef SE_K1_WARPS
#define SE_K1_WARPS 2
#endif
// small CTAs (2 warps, four per SM at 255 registers per thread): with two steps in flight the other step's kernels get SMs
// back in finer grains (8-warp CTAs: +1.4 us per step; the kernel alone is the same 13.9 us either way)
constexpr int kWarpsRun = SE_K1_WARPS, kThreadsRun = kWarpsRun * 32;
constexpr int kAccFloat2 = M + 2;                                         // bins 0..256, padded to a 16-byte multiple
constexpr int kHwBytes1 = M * 8 + kAccFloat2 * 8;                         // transpose buffer | accumulators
constexpr size_t kSmem1Run = (size_t)(kWarpsRun * 2) * kHwBytes1 + (kWarpsRun * 2) * 4;

struct StftRunPlan { int runs_per_utt; long long total_runs; };

// buf[r] = samples (2m, 2m+1), m = j + 16 r, r < 8, of the H samples starting at original coordinate t0 (reflect outside [0, T))
__device__ __forceinline__ void load_half_regs(float2 (&buf)[8], const float* __restrict__ row, int T, int t0, int j) {
    const float* src = row + t0;
    if ((t0 >= 0) && (t0 + H <= T) && ((reinterpret_cast<uintptr_t>(src) & 7) == 0)) {
        const float2* s2 = reinterpret_cast<const float2*>(src) + j;
#pragma unroll
        for (int r = 0; r < 8; ++r) buf[r] = __ldg(s2 + 16 * r);
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int ta = t0 + 2 * (j + 16 * r), tb = ta + 1;
            ta = ta < 0 ? -ta : ta; ta = ta >= T ? 2 * (T - 1) - ta : ta;
            tb = tb < 0 ? -tb : tb; tb = tb >= T ? 2 * (T - 1) - tb : tb;
            buf[r] = make_float2(__ldg(row + ta), __ldg(row + tb));
        }
    }
}

template <bool POWER, bool PHASE, bool LOGP, bool STATS>
__global__ void __launch_bounds__(kThreadsRun, 256 / kThreadsRun) stft512_run_kernel(StftArgs a, StftRunPlan plan) {
    extern __shared__ __align__(16) unsigned char smem1[];
    secommon::TraceScope trace(a.trace, 1);
    const int lane = threadIdx.x & 31, j = lane & 15;
    const int hw = (threadIdx.x >> 4);
    unsigned char* mine = smem1 + (size_t)hw * kHwBytes1;
    float2* xbuf = reinterpret_cast<float2*>(mine);
    float2* acc = reinterpret_cast<float2*>(mine + M * 8);
    int* s_utt = reinterpret_cast<int*>(smem1 + (size_t)(kWarpsRun * 2) * kHwBytes1);   // utterance of every half-warp's run (-1: none)
    const unsigned hmask = half_mask(lane);
    if (a.zero_ptr)                                                   // SE_FLAG_WS_SELF_CLEAN: K3's sums of this step
        for (long long i = (long long)blockIdx.x * kThreadsRun + threadIdx.x; i < a.zero_count; i += (long long)gridDim.x * kThreadsRun)
            a.zero_ptr[i] = 0.0;
    const long long unit = (long long)blockIdx.x * (kThreadsRun / 16) + hw;
    const bool active = unit < plan.total_runs;
    const int u = active ? (int)(unit / plan.runs_per_utt) : -1;
    // the first frame's samples come from HBM: put those loads in flight before anything else (tables, accumulators)
    int fa = 0, fb = 0;
    const float* row = a.wav;
    float2 A[8], B[8];
    if (active) {
        const int ri = (int)(unit - (long long)u * plan.runs_per_utt);
        // the first (F mod rpu) runs are one frame longer; rpu is even, so the two half-warps of a warp (runs 2k, 2k+1 of
        // one utterance) have equal lengths except for one pair per utterance
        const int base_len = a.n_frames / plan.runs_per_utt, rem_runs = a.n_frames - base_len * plan.runs_per_utt;
        fa = ri * base_len + min(ri, rem_runs);
        fb = fa + base_len + (ri < rem_runs ? 1 : 0);
        row = a.wav + (long long)u * a.utt_stride;
        load_half_regs(A, row, a.T, (fa - 1) * H, j);
        load_half_regs(B, row, a.T, fa * H, j);
    }
    float2 wlo[8], whi[8];                                            // window pairs of this lane, pre-scaled by 1/2
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const float2 w0 = __ldg(reinterpret_cast<const float2*>(a.tab.window) + j + 16 * r);
        const float2 w1 = __ldg(reinterpret_cast<const float2*>(a.tab.window) + j + 16 * r + 128);
        wlo[r] = make_float2(0.5f * w0.x, 0.5f * w0.y);
        whi[r] = make_float2(0.5f * w1.x, 0.5f * w1.y);
    }
    float2 tw[15], twn[8];
    load_lane_constants(j, a.tab.twM, a.tab.twN, tw, twn);
    float2 sacc[17];
#pragma unroll
    for (int i = 0; i < 17; ++i) sacc[i] = make_float2(0.0f, 0.0f);
    if (STATS && j == 0) s_utt[hw] = u;
    griddep_launch();                                   // a dependent kernel may start its prologue (it waits for our completion)
    if (threadIdx.x == 0) trace.mark(17);
    if (active) {
        const int F = a.n_frames;
        // one frame: window (first, second) into v, refill `first` with the second half of frame f + 1, transform, write
        auto frame = [&](float2 (&first)[8], float2 (&second)[8], int f) {
            float2 v[16];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                v[r] = pmul(first[r], wlo[r]);
                v[r + 8] = pmul(second[r], whi[r]);
            }
            if (f + 1 < fb) load_half_regs(first, row, a.T, (f + 1) * H, j);
            fft256<-1>(v, xbuf, j, tw, hmask);
            float2 zm[8];
            fetch_mirror(v, lane, zm);
            emit_frame<POWER, PHASE, LOGP, STATS>(v, zm, twn, j, a, ((long long)u * F + f) * a.spec_stride, sacc);
        };
#pragma unroll 1
        for (int f = fa; f < fb; f += 2) {
            frame(A, B, f);
            if (f + 1 < fb) frame(B, A, f + 1);
        }
    }
    if (threadIdx.x == 0) trace.mark(18);
    if (STATS) {
        // park the register accumulators in the half-warp's row: bin k at acc[k]
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            acc[j + 16 * q] = sacc[q];
            acc[M - j - 16 * q] = sacc[8 + q];
        }
        if (j == 0) acc[128] = sacc[16];
        __syncthreads();                                            // every half-warp's accumulators are final
        if (threadIdx.x == 0) trace.mark(19);
        // thread t owns bin t (thread 0 also bin 256); the CTA's runs are consecutive, so utterances are non-decreasing
        const float2* base = reinterpret_cast<const float2*>(smem1 + M * 8);
        constexpr int kStride = kHwBytes1 / 8;
        for (int bin = threadIdx.x; bin <= M; bin += kThreadsRun) {
            double s1 = 0.0, s2 = 0.0;
            int cur = -1;
            for (int h = 0; h < kThreadsRun / 16; ++h) {
                const int uh = s_utt[h];
                if (uh < 0) break;
                if (uh != cur) {
                    if (cur >= 0) {
                        double* pdst = a.stat_sums + ((long long)cur * a.ld_stats + bin) * 2;
                        atomicAdd(pdst, s1);
                        atomicAdd(pdst + 1, s2);
                    }
                    cur = uh; s1 = 0.0; s2 = 0.0;
                }
                const float2 t = base[h * kStride + bin];
                s1 += (double)t.x;
                s2 += (double)t.y;
            }
            if (cur >= 0) {
                double* pdst = a.stat_sums + ((long long)cur * a.ld_stats + bin) * 2;
                atomicAdd(pdst, s1);
                atomicAdd(pdst + 1, s2);
            }
        }
    }
    if (a.trace) { __syncthreads(); trace.finish(); }
}

// ------------------------------------------------------------------ K3: fused mask -> iSTFT, hop = 256
// One half-warp owns a RUN of consecutive output blocks b = b0 .. b1 of one utterance (block b = output
// samples [(b-1)*256, b*256), the sum of the second half of frame b-1 and the first half of frame b).
// It walks frames b0-1 .. b1; the second half of each inverse transform stays in registers ("carry")
// and is added to the first half of the next one, so the overlap-add needs neither shared memory nor
// atomics.  Frame b0-1 is a halo frame (recomputed by the neighbouring run).
// Synthesis table: bw[n] = s(n) * w[n] / (2 M (w[n mod H]^2 + w[n mod H + H]^2)) with s = +1 for even n
// and -1 for odd n (the conjugation of the forward-as-inverse FFT); the 2 undoes merge_pair_conj's factor.
//
// Staging (cp.async) is sized for THREE resident CTAs per SM -- the kernel is bound by dependent-issue
// latency, so resident warps are what buys throughput.  Consecutive frames share half of their samples,
// so a half-warp keeps two 256-sample slots per waveform: frame f reads (first, second) = (slot p, slot p^1),
// and as soon as the frame is in registers the dead first slot receives the second half of frame f+1.  Every
// prefetch therefore has a full frame of work to land behind.  Pass order per frame:
//   0: noisy frame -> rFFT -> x sqrt(mask) (mask row slot, refilled right after) -> merged spectrum
//   1: inverse FFT -> window -> overlap-add with the carry -> store, waveform sums (clean first slot)
//   2: clean frame -> rFFT -> spectral SI-SDR sums against the masked noisy power kept from pass 0
// The three transforms run through ONE copy of the FFT code inside a non-unrolled pass loop.
// cp.async groups are committed in the fixed order N(oisy) M(ask) C(lean) once per frame (empty groups at the
// end of a run), so "all but the two most recent groups" is exactly the data the next consumer needs.
#ifndef SE_K3_WARPS
#define SE_K3_WARPS 4
#endif
constexpr int kWarps3 = SE_K3_WARPS, kThreads3 = kWarps3 * 32;

// Run partition of one utterance's `bpu` output blocks into `rpu` runs: base = bpu / rpu blocks each, `rem` = bpu mod rpu
// runs get one more.  Runs come in pairs (2k, 2k+1) -- the two half-warps of a warp -- and both runs of a pair have the
// same length (so no warp idles half its lanes) except for one mixed pair when rem is odd; long and short pairs are
// interleaved evenly (Bresenham), so every CTA -- and every SM -- gets the same mix of long and short runs.
// Returns the first block (1-based) of run ri and its length.
__host__ __device__ __forceinline__ void run_bounds(int ri, int bpu, int rpu, int& b0, int& len) {
    const int base = bpu / rpu, rem = bpu - base * rpu;
    if (rpu & 1) {                                                  // odd run count (tiny problems): longer runs first
        b0 = 1 + ri * base + (ri < rem ? ri : rem);
        len = base + (ri < rem ? 1 : 0);
        return;
    }
    const int P = rpu >> 1, L = rem >> 1, odd = rem & 1;            // pairs, long pairs, one extra long run
    const int k = ri >> 1, h = ri & 1;
    const int nl = (int)(((long long)k * L) / P), nl1 = (int)(((long long)(k + 1) * L) / P);
    const int is_long = nl1 - nl;                                   // 0 or 1; pair 0 is always short (L < P)
    // odd rem: the extra block makes run 0 (first run of the short pair 0) long -- the one mixed pair
    b0 = 1 + 2 * (k * base + nl) + h * (base + is_long) + (ri > 0 ? odd : 0);
    len = base + is_long + (ri == 0 ? odd : 0);
}

struct RunPlan { int blocks_per_utt; int runs_per_utt; long long total_runs; };   // run ri: bpu/rpu blocks, the first bpu%rpu runs one more

// resident warps per SM = 4 x SE_K3_MIN_BLOCKS.  2 (8 warps, 255-register cap: no spills, runs of 6.8 blocks at 64 x 4 s so the
// halo frame is 15 % instead of 22 %) beats 3 (12 warps, 168 registers): K3 alone 38.2 vs 41.1 us, two steps in flight 67.2 vs
// 68.1 us per step (round 2, tools/sweep_lib.sh)
#ifndef SE_K3_MIN_BLOCKS
#define SE_K3_MIN_BLOCKS 2
#endif
constexpr int kMaskFloats3 = 272;
// per half-warp: transpose buffer | noisy slots 2 x 256 | clean slots 2 x 256 | mask row
constexpr int kNoisyBytes3 = 2 * H * 4;
constexpr int kHwBytes3 = M * 8 + kNoisyBytes3 + 2 * H * 4 + kMaskFloats3 * 4;
constexpr size_t kSmem3 = (size_t)(kWarps3 * 2) * kHwBytes3 + 2 * M * 8;
static_assert(kHwBytes3 % 16 == 0 && kNoisyBytes3 % 16 == 0, "16-byte aligned cp.async destinations");

template <bool PM>
__global__ void __launch_bounds__(kThreads3, SE_K3_MIN_BLOCKS * 4 / kWarps3) mask_istft512_kernel(MaskIstftArgs a, RunPlan plan) {
    extern __shared__ __align__(16) unsigned char smem3[];
    secommon::TraceScope trace(a.trace, 3);
    float2* s_win2 = reinterpret_cast<float2*>(smem3 + (size_t)(kWarps3 * 2) * kHwBytes3);
    float2* s_bw2 = s_win2 + M;
    {   // the first frame's noisy samples come from HBM and do not depend on the upstream kernel: loads in flight first
        const int hw0 = threadIdx.x >> 4, j0 = threadIdx.x & 15;
        const long long unit0 = (long long)blockIdx.x * (kThreads3 / 16) + hw0;
        if (unit0 < plan.total_runs) {
            const int u0 = (int)(unit0 / plan.runs_per_utt), ri0 = (int)(unit0 - (long long)u0 * plan.runs_per_utt);
            int b00, len00;
            run_bounds(ri0, plan.blocks_per_utt, plan.runs_per_utt, b00, len00);
            const int f00 = b00 - 1;
            float* nb0 = reinterpret_cast<float*>(smem3 + (size_t)hw0 * kHwBytes3 + M * 8);
            const float* nrow0 = a.noisy + (long long)u0 * a.utt_stride;
            stage_half(nb0, nrow0, a.T, (f00 - 1) * H, j0);
            stage_half(nb0 + H, nrow0, a.T, f00 * H, j0);
        }
        cp_async_commit();
    }
    for (int i = threadIdx.x; i < M; i += kThreads3) {
        const float w0 = a.tab.window[2 * i], w1 = a.tab.window[2 * i + 1];
        s_win2[i] = make_float2(0.5f * w0, 0.5f * w1);
        const int n0 = (2 * i) & (H - 1), n1 = (2 * i + 1) & (H - 1);
        const float e0 = a.tab.window[n0] * a.tab.window[n0] + a.tab.window[n0 + H] * a.tab.window[n0 + H];
        const float e1 = a.tab.window[n1] * a.tab.window[n1] + a.tab.window[n1 + H] * a.tab.window[n1 + H];
        s_bw2[i] = make_float2(w0 / (2.0f * M * e0), -w1 / (2.0f * M * e1));
    }
    const int lane = threadIdx.x & 31, j = lane & 15, hw = threadIdx.x >> 4;
    const unsigned hmask = half_mask(lane);
    float2 tw[15], twn[8];
    load_lane_constants(j, a.tab.twM, a.tab.twN, tw, twn);
    __syncthreads();
    griddep_launch();
    unsigned char* mine = smem3 + (size_t)hw * kHwBytes3;
    float2* xbuf = reinterpret_cast<float2*>(mine);
    float* nb = reinterpret_cast<float*>(mine + M * 8);            // noisy slots
    float* cb = reinterpret_cast<float*>(mine + M * 8 + kNoisyBytes3);   // clean slots
    float* mb = cb + 2 * H;                                        // mask row
    const long long unit = (long long)blockIdx.x * (kThreads3 / 16) + hw;
    auto zero_ws = [&]() {                                        // SE_FLAG_WS_SELF_CLEAN: the CMVN sums the head has consumed
        if (a.zero_ptr)
            for (long long i = (long long)blockIdx.x * kThreads3 + threadIdx.x; i < a.zero_count; i += (long long)gridDim.x * kThreads3)
                a.zero_ptr[i] = 0.0;
    };
    if (unit >= plan.total_runs) { griddep_wait(); zero_ws(); return; }   // no block-level barrier below (tracing: approximate for ragged CTAs)
    const int u = (int)(unit / plan.runs_per_utt), ri = (int)(unit - (long long)u * plan.runs_per_utt);
    const int F = a.n_frames;
    int b0, run_len;
    run_bounds(ri, plan.blocks_per_utt, plan.runs_per_utt, b0, run_len);
    const int b1 = b0 + run_len - 1;
    const float* nrow = a.noisy + (long long)u * a.utt_stride;
    const float* crow = a.clean ? a.clean + (long long)u * a.utt_stride : nullptr;
    float* orow = a.wav_out + (long long)u * a.out_stride;
    const int len = a.lengths ? (int)min((long long)a.T, max(0LL, a.lengths[u])) : a.T;   // (a length beyond the padded row would read past it)
    const int valid_frames = min(F, len / H + 1);                  // runner.py:455
    const bool spec = a.want_spec && crow && a.sums;
    const bool need_clean = crow && a.sums;
    const bool out_aligned = (reinterpret_cast<uintptr_t>(orow) & 7) == 0;
    const bool mask_padded = a.mask_stride >= 260;
    const float* mrow0 = a.mask + (long long)u * F * a.mask_stride;
    float acc[sekern::NSUMS];
#pragma unroll
    for (int i = 0; i < sekern::NSUMS; ++i) acc[i] = 0.0f;
    float2 carry[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) carry[q] = make_float2(0.0f, 0.0f);
    float2 yy2 = make_float2(0.0f, 0.0f), yc2 = yy2, cc2 = yy2;     // (even, odd) sample partial sums of the fast overlap-add path
    float2 st2 = yy2, tt2 = yy2, ss2 = yy2;                         // (bin k, bin M-k) partial spectral sums

    // prologue: both halves of the first (halo) frame, its mask row, both clean halves -- groups N, M, C.  The waveforms
    // are inputs of the step, so their first loads are issued before waiting for the upstream kernel (the mask's producer).
    const int f0 = b0 - 1;                                        // (its noisy samples were requested at the top: group N)
    griddep_wait();                                               // the mask (and the zeroed sums) come from upstream kernels
    zero_ws();
    stage_row(mb, mrow0 + (long long)f0 * a.mask_stride, M + 1, j, mask_padded);
    cp_async_commit();
    if (need_clean) {
        stage_half(cb, crow, a.T, (f0 - 1) * H, j);
        stage_half(cb + H, crow, a.T, f0 * H, j);
    }
    cp_async_commit();

    int p = 0;                                                      // slot holding the first half of the current frame
#pragma unroll 1
    for (int f = f0; f <= b1; ++f, p ^= 1) {
        const bool halo = (f == f0);
        const bool own = spec && (!halo || f == 0) && f < valid_frames;
        const bool more = f < b1;
        float ra[8], rb[8], r128 = 0.0f;                            // relu(mask * |X|^2) of this frame (objective.py:89)
#pragma unroll
        for (int q = 0; q < 8; ++q) { ra[q] = 0.0f; rb[q] = 0.0f; }
        float2 v[16];
#pragma unroll 1
        for (int pass = 0; pass < (own ? 3 : 2); ++pass) {
            if (pass == 0) {
                cp_async_wait<2>();                                 // N(f) has landed
                __syncwarp(hmask);
                frame_from_slots(nb + p * H, nb + (p ^ 1) * H, j, s_win2, v);
                __syncwarp(hmask);
                if (more) stage_half(nb + p * H, nrow, a.T, (f + 1) * H, j);       // second half of frame f+1 -> the dead first slot
                cp_async_commit();
            } else if (pass == 2) {
                frame_from_slots(cb + p * H, cb + (p ^ 1) * H, j, s_win2, v);        // C(f) landed before pass 1's overlap-add
            }
            fft256<-1>(v, xbuf, j, tw, hmask);
            if (pass == 0) {
                float2 zm[8];
                fetch_mirror(v, lane, zm);
                cp_async_wait<2>();                                 // M(f) has landed
                __syncwarp(hmask);
                float2 ca[8], cbv[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float2 xa, xb;
                    split_pair(v[q], zm[q], twn[q], xa, xb);
                    mask_merge<PM>(xa, xb, mb[j + 16 * q], mb[M - j - 16 * q], twn[q], own, ra[q], rb[q], ca[q], cbv[q]);
                }
                const float g128 = mb[128];
                const float2 x128 = make_float2(2.0f * v[8].x, -2.0f * v[8].y);    // k = 128 pairs with itself: X = 2 conj(Z[128])
                if (own) r128 = fmaxf(PM ? g128 : g128 * (x128.x * x128.x + x128.y * x128.y), 0.0f);
                __syncwarp(hmask);
                if (more) stage_row(mb, mrow0 + (long long)(f + 1) * a.mask_stride, M + 1, j, mask_padded);
                cp_async_commit();
                // Zinv[128] = 2 conj(Y[128]) (same factor 2 as merge_pair_conj); its conjugate feeds the FFT
                float2 y128;
                if (PM) { y128 = apply_gain<true>(x128, g128); y128.x *= 2.0f; y128.y *= 2.0f; }
                else { const float s128 = 2.0f * fast_sqrt(g128); y128 = make_float2(s128 * x128.x, s128 * x128.y); }
                scatter_mirror(ca, cbv, y128, lane, v);
            } else if (pass == 1) {
                // v[q] = conj(z[m]), z[m] = (x[2m], x[2m+1]) unnormalised, m = j + 16 q; signs and scales are in s_bw2
                cp_async_wait<2>();                                 // C(f) has landed
                __syncwarp(hmask);
                if (!halo) {
                    const int t0 = (f - 1) * H;
                    const float* cfirst = cb + p * H;               // first half of the clean frame = this output block
                    if (out_aligned && t0 + H <= len && need_clean) {   // whole block inside the utterance, all sums wanted:
                        float2* o2 = reinterpret_cast<float2*>(orow + t0) + j;     // straight-line code, no per-sample predicates
                        const float2* c2 = reinterpret_cast<const float2*>(cfirst) + j;
                        const float2* bw2 = s_bw2 + j;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float2 y = pfma(bw2[16 * q], v[q], carry[q]);
                            const float2 c = c2[16 * q];
                            o2[16 * q] = y;
                            yy2 = pfma(y, y, yy2);
                            yc2 = pfma(y, c, yc2);
                            cc2 = pfma(c, c, cc2);
                        }
                    } else if (out_aligned && t0 + H <= len) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int m = j + 16 * q;
                            const float2 y = pfma(s_bw2[m], v[q], carry[q]);
                            *reinterpret_cast<float2*>(orow + t0 + 2 * m) = y;
                            if (a.sums) yy2 = pfma(y, y, yy2);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const int m = j + 16 * q;
                            const float2 y = pfma(s_bw2[m], v[q], carry[q]);
                            const int t = t0 + 2 * m;
                            if (out_aligned) *reinterpret_cast<float2*>(orow + t) = y;
                            else { orow[t] = y.x; orow[t + 1] = y.y; }
                            if (a.sums) {
                                float2 c = make_float2(0.0f, 0.0f);
                                if (crow) c = *reinterpret_cast<const float2*>(cfirst + 2 * m);
                                if (t < len) { acc[sekern::SUM_YY] += y.x * y.x; acc[sekern::SUM_YC] += y.x * c.x; acc[sekern::SUM_CC] += c.x * c.x; }
                                if (t + 1 < len) { acc[sekern::SUM_YY] += y.y * y.y; acc[sekern::SUM_YC] += y.y * c.y; acc[sekern::SUM_CC] += c.y * c.y; }
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) carry[q] = pmul(s_bw2[j + 16 * q + 128], v[q + 8]);
            } else {
                float2 zm[8];
                fetch_mirror(v, lane, zm);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    float2 xa, xb;
                    split_pair(v[q], zm[q], twn[q], xa, xb);
                    const float2 pt = make_float2(xa.x * xa.x + xa.y * xa.y, xb.x * xb.x + xb.y * xb.y);
                    const float2 r = make_float2(ra[q], rb[q]);
                    const float2 rt = pmul(r, pt);
                    st2 = cadd(st2, make_float2(fast_sqrt(rt.x), fast_sqrt(rt.y)));
                    tt2 = cadd(tt2, pt);
                    ss2 = cadd(ss2, r);
                }
                if (j == 0) {
                    const float pt128 = 4.0f * (v[8].x * v[8].x + v[8].y * v[8].y);
                    acc[sekern::SUM_SPEC_ST] += fast_sqrt(r128 * pt128);
                    acc[sekern::SUM_SPEC_TT] += pt128;
                    acc[sekern::SUM_SPEC_SS] += r128;
                }
            }
        }
        __syncwarp(hmask);                                          // the clean first slot is dead now
        if (need_clean && more) stage_half(cb + p * H, crow, a.T, (f + 1) * H, j);
        cp_async_commit();
    }
    cp_async_wait<0>();
    // the last run of an utterance also zero-fills [out_len, pad_to) and finishes sum c^2 over [out_len, len)
    if (b1 == F - 1) {
        for (int t = a.out_len + j; t < max(a.pad_to, len); t += 16) {
            if (t < a.pad_to) orow[t] = 0.0f;
            if (a.sums && crow && t < len) { const float c = __ldg(crow + t); acc[sekern::SUM_CC] += c * c; }
        }
    }
    if (a.sums) {
        acc[sekern::SUM_YY] += yy2.x + yy2.y;
        acc[sekern::SUM_YC] += yc2.x + yc2.y;
        acc[sekern::SUM_CC] += cc2.x + cc2.y;
        acc[sekern::SUM_SPEC_ST] += st2.x + st2.y;
        acc[sekern::SUM_SPEC_TT] += tt2.x + tt2.y;
        acc[sekern::SUM_SPEC_SS] += ss2.x + ss2.y;
#pragma unroll
        for (int i = 0; i < sekern::NSUMS; ++i) {
            float s = acc[i];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(hmask, s, o);
            if (j == 0 && s != 0.0f) atomicAdd(a.sums + (long long)u * sekern::NSUMS + i, (double)s);
        }
    }
    if (a.trace) { __syncthreads(); trace.finish(); }
}

}  // namespace

namespace sefast {

int num_sms() { return secommon::device_sms(); }

// opt the kernels into their dynamic shared-memory sizes (called once per device from se_prepare / first use)
int prepare512() {
#define SE_OPT(K, BYTES) SE_CUDA_CHECK(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)))
    SE_OPT((stft512_kernel<true, false, false>), kSmem1);
    SE_OPT((stft512_kernel<false, true, false>), kSmem1);
    SE_OPT((stft512_kernel<true, true, false>), kSmem1);
    SE_OPT((stft512_kernel<false, false, true>), kSmem1);
    SE_OPT((stft512_kernel<true, false, true>), kSmem1);
    SE_OPT((stft512_kernel<false, true, true>), kSmem1);
    SE_OPT((stft512_kernel<true, true, true>), kSmem1);
    SE_OPT((stft512_run_kernel<true, false, false, false>), kSmem1Run);
    SE_OPT((stft512_run_kernel<false, true, false, false>), kSmem1Run);
    SE_OPT((stft512_run_kernel<true, true, false, false>), kSmem1Run);
    SE_OPT((stft512_run_kernel<false, false, true, false>), kSmem1Run);
    SE_OPT((stft512_run_kernel<true, false, true, false>), kSmem1Run);
    SE_OPT((stft512_run_kernel<false, true, true, false>), kSmem1Run);
    SE_OPT((stft512_run_kernel<true, true, true, false>), kSmem1Run);
    SE_OPT((stft512_run_kernel<true, false, false, true>), kSmem1Run);
    SE_OPT((stft512_run_kernel<false, false, true, true>), kSmem1Run);
    SE_OPT((stft512_run_kernel<true, false, true, true>), kSmem1Run);
    SE_OPT(mask_istft512_kernel<false>, kSmem3);
    SE_OPT(mask_istft512_kernel<true>, kSmem3);
#undef SE_OPT
    return SE_OK;
}

int launch_stft512(const StftArgs& a, cudaStream_t st) {
    const long long total = (long long)a.n_utt * a.n_frames;
    if (total > 0x7fffff00LL) return secommon::fail(SE_ERR_BAD_ARG, "too many frames (%lld)", total);
    const int sel = (a.power ? 1 : 0) | (a.phase ? 2 : 0) | (a.logp ? 4 : 0);
    if (sel == 0) return SE_OK;                                  // nothing requested
    if (a.hop == H) {
        const int per_it = kThreadsRun / 16;
        // runs: one balanced wave of 2 CTAs per SM when the batch is small, runs of about 32 frames otherwise
        const long long slots = (256LL / kThreadsRun) * num_sms() * per_it;   // 8 warps per SM (255 registers per thread)
        long long rpu;
        if (total <= slots * 32) {
            rpu = slots / a.n_utt;
            if (rpu > a.n_frames) rpu = a.n_frames;
        } else rpu = (a.n_frames + 31) / 32;
        if (rpu > 2) rpu &= ~1LL;
        if (rpu < 1) rpu = 1;
        StftRunPlan plan;
        plan.runs_per_utt = (int)rpu;
        plan.total_runs = (long long)a.n_utt * rpu;
        const long long grid = (plan.total_runs + per_it - 1) / per_it;
        if (grid > 0x7fffffffLL) return secommon::fail(SE_ERR_BAD_ARG, "grid too large");
#define SE_RUN(P, Q, L, S) stft512_run_kernel<P, Q, L, S><<<(unsigned)grid, kThreadsRun, kSmem1Run, st>>>(a, plan)
        if (a.stat_sums) {
            // statistics of the ONE feature written: log-power (sel 4) or power (sel 1)
            if (sel == 4) SE_RUN(false, false, true, true);
            else if (sel == 1) SE_RUN(true, false, false, true);
            else if (sel == 5) SE_RUN(true, false, true, true);                      // power + log-power, statistics of log-power
            else return secommon::fail(SE_ERR_BAD_ARG, "statistics need power and / or logpower (no phase)");
        } else {
            switch (sel) {
                case 1: SE_RUN(true, false, false, false); break;
                case 2: SE_RUN(false, true, false, false); break;
                case 3: SE_RUN(true, true, false, false); break;
                case 4: SE_RUN(false, false, true, false); break;
                case 5: SE_RUN(true, false, true, false); break;
                case 6: SE_RUN(false, true, true, false); break;
                default: SE_RUN(true, true, true, false); break;
            }
        }
#undef SE_RUN
        return secommon::check_launch("stft512_run_kernel");
    }
    if (a.stat_sums) return secommon::fail(SE_ERR_UNSUPPORTED, "fused statistics need hop = 256");
    const int per_it = kThreads1 / 16;
    const long long want = (total + per_it - 1) / per_it;
    const long long cap = 2LL * num_sms();
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    const int per_cta = (int)((total + grid - 1) / grid);
    switch (sel) {
        case 1: stft512_kernel<true, false, false><<<grid, kThreads1, kSmem1, st>>>(a, total, per_cta); break;
        case 2: stft512_kernel<false, true, false><<<grid, kThreads1, kSmem1, st>>>(a, total, per_cta); break;
        case 3: stft512_kernel<true, true, false><<<grid, kThreads1, kSmem1, st>>>(a, total, per_cta); break;
        case 4: stft512_kernel<false, false, true><<<grid, kThreads1, kSmem1, st>>>(a, total, per_cta); break;
        case 5: stft512_kernel<true, false, true><<<grid, kThreads1, kSmem1, st>>>(a, total, per_cta); break;
        case 6: stft512_kernel<false, true, true><<<grid, kThreads1, kSmem1, st>>>(a, total, per_cta); break;
        default: stft512_kernel<true, true, true><<<grid, kThreads1, kSmem1, st>>>(a, total, per_cta); break;
    }
    return secommon::check_launch("stft512_kernel");
}

int launch_mask_istft512(const MaskIstftArgs& a, cudaStream_t st) {
    // Runs: every utterance's F-1 output blocks are cut into runs_per_utt near-equal runs (lengths differ by at most one).
    // Small batches: as many runs as there are resident half-warps (SE_K3_MIN_BLOCKS CTAs per SM), so that ONE balanced
    // wave covers the GPU -- but at least 4 blocks per run (the halo frame costs 2/3 of a frame).  Large batches: runs of
    // about 32 blocks, many waves.
    const int blocks_per_utt = a.n_frames - 1;
    const long long slots = (long long)(SE_K3_MIN_BLOCKS * 4 / kWarps3) * num_sms() * (kThreads3 / 16);
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("SE_B200_RUN_LEN"); forced = e ? atoi(e) : 0; }
    RunPlan plan;
    plan.blocks_per_utt = blocks_per_utt;
    long long rpu;
    if (forced > 0) rpu = (blocks_per_utt + forced - 1) / forced;
    else if ((long long)a.n_utt * blocks_per_utt <= slots * 32) {
        rpu = slots / a.n_utt;
        const long long cap = blocks_per_utt / 4;
        if (rpu > cap) rpu = cap;
    } else rpu = (blocks_per_utt + 31) / 32;
    if (rpu > 2) rpu &= ~1LL;                                        // even: warps pair runs of the same utterance
    if (rpu < 1) rpu = 1;
    if (rpu > blocks_per_utt) rpu = blocks_per_utt;
    plan.runs_per_utt = (int)rpu;
    plan.total_runs = (long long)a.n_utt * plan.runs_per_utt;
    const long long grid = (plan.total_runs + (kThreads3 / 16) - 1) / (kThreads3 / 16);
    if (grid > 0x7fffffffLL) return secommon::fail(SE_ERR_BAD_ARG, "grid too large");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads3);
    cfg.dynamicSmemBytes = kSmem3;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // prologue overlaps the upstream kernel's tail
    attr[0].val.programmaticStreamSerializationAllowed = (secommon::pdl_mask() & 2) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (a.mask_is_power) SE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mask_istft512_kernel<true>, a, plan));
    else SE_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mask_istft512_kernel<false>, a, plan));
    return secommon::check_launch("mask_istft512_kernel");
}

}  // namespace sefast
