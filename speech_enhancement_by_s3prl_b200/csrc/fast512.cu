// libse_b200.so -- n_fft = 512 fast paths: one frame per half-warp, FFT values in registers
// (fft256_warp.cuh), frames gathered straight from global memory with coalesced 8-byte loads
// (each half-warp reads the 2 KB of its frame as 16 x 128 B), spectra written straight from
// registers.  No block-level barrier anywhere: warps run independently and the kernels are
// persistent (grid sized from the SM count).
//
//   stft512_kernel        K1: frame -> window -> rFFT -> power / phase / log-power
//   mask_istft512_kernel  K3: frame -> rFFT -> x sqrt(mask) -> irFFT -> window -> overlap-add in
//                             registers along a run of consecutive frames -> / envelope -> wav,
//                             plus the per-utterance metric sums
#include <cstdlib>
#include "se_common.cuh"
#include "fft256_warp.cuh"
#include "tile_kernels.cuh"

using namespace fft256w;
using sekern::StftArgs;
using sekern::MaskIstftArgs;

namespace {

constexpr int N = 512, K = 257, kWarps = 8, kThreads = kWarps * 32;

// gather z[m] = (x[2m], x[2m+1]) * (w[2m], w[2m+1]) for m = j + 16 r of the frame starting at original
// coordinate t0 (may run over either end of the row -> reflect, torch.stft pad_mode='reflect')
__device__ __forceinline__ void load_frame(const float* __restrict__ row, int T, int t0, int j, const float2* __restrict__ win2,
                                           float2 (&v)[16]) {
    const bool interior = (t0 >= 0) && (t0 + N <= T);
    const float* src = row + t0;
    if (interior && ((reinterpret_cast<uintptr_t>(src) & 7) == 0)) {
        const float2* s2 = reinterpret_cast<const float2*>(src);
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = __ldg(s2 + j + 16 * r);
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            int ta = t0 + 2 * (j + 16 * r), tb = ta + 1;
            ta = ta < 0 ? -ta : ta; ta = ta >= T ? 2 * (T - 1) - ta : ta;
            tb = tb < 0 ? -tb : tb; tb = tb >= T ? 2 * (T - 1) - tb : tb;
            v[r] = make_float2(__ldg(row + ta), __ldg(row + tb));
        }
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        const float2 w = win2[j + 16 * r];
        v[r].x *= w.x;
        v[r].y *= w.y;
    }
}

__global__ void __launch_bounds__(kThreads, 2) stft512_kernel(StftArgs a, long long total_frames) {
    __shared__ __align__(16) float2 s_x[kWarps * 2][M];       // transpose buffers, one per half-warp
    __shared__ __align__(16) float2 s_win2[M];                 // window as (w[2m], w[2m+1])
    for (int i = threadIdx.x; i < M; i += kThreads) s_win2[i] = make_float2(a.tab.window[2 * i], a.tab.window[2 * i + 1]);
    const int lane = threadIdx.x & 31, j = lane & 15;
    const int hw = (threadIdx.x >> 4);                          // half-warp in CTA
    float2 tw[15], twn[8];
    load_lane_constants(j, a.tab.twM, a.tab.twN, tw, twn);
    __syncthreads();
    float2* xbuf = s_x[hw];
    const long long n_hw = (long long)gridDim.x * (kThreads / 16);
    const long long first = (long long)blockIdx.x * (kThreads / 16) + hw;
    const unsigned hmask = half_mask(lane);                     // half-warps are independent of each other
    for (long long gg = first; gg < total_frames; gg += n_hw) {
        const int u = (int)(gg / a.n_frames), f = (int)(gg - (long long)u * a.n_frames);
        float2 v[16];
        load_frame(a.wav + (long long)u * a.utt_stride, a.T, f * a.hop - N / 2, j, s_win2, v);
        fft256<-1>(v, xbuf, j, tw, hmask);
        float2 zm[8];
        fetch_mirror(v, lane, zm);
        const long long o = gg * a.spec_stride;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = j + 16 * q;
            float2 xa = rfft_split(v[q], zm[q], twn[q]);
            float2 xb = rfft_split(zm[q], v[q], make_float2(-twn[q].x, twn[q].y));     // W_N^(M-k) = -conj(W_N^k)
            if (k == 0) { xa.y = 0.0f; xb.y = 0.0f; }
            const float pa = xa.x * xa.x + xa.y * xa.y, pb = xb.x * xb.x + xb.y * xb.y;
            if (a.power) { a.power[o + k] = pa; a.power[o + M - k] = pb; }
            if (a.logp) { a.logp[o + k] = logf(pa + a.log_eps); a.logp[o + M - k] = logf(pb + a.log_eps); }
            if (a.phase) { a.phase[o + k] = atan2f(xa.y, xa.x); a.phase[o + M - k] = atan2f(xb.y, xb.x); }
        }
        if (j == 0) {                                               // k = 128 pairs with itself
            const float2 x = rfft_split(v[8], v[8], make_float2(0.0f, -1.0f));
            const float p = x.x * x.x + x.y * x.y;
            if (a.power) a.power[o + 128] = p;
            if (a.logp) a.logp[o + 128] = logf(p + a.log_eps);
            if (a.phase) a.phase[o + 128] = atan2f(x.y, x.x);
        }
    }
}


// ------------------------------------------------------------------ K3: fused mask -> iSTFT, hop = 256
// One half-warp owns a RUN of consecutive output blocks b = b0 .. b1 of one utterance (block b = output
// samples [(b-1)*256, b*256), the sum of the second half of frame b-1 and the first half of frame b).
// It walks frames b0-1 .. b1; the second half of each inverse transform stays in registers ("carry")
// and is added to the first half of the next one, so the overlap-add needs neither shared memory nor
// atomics.  Frame b0-1 is a halo frame (recomputed by the neighbouring run).  The synthesis window, the
// 1/M of the inverse transform and the division by the overlap-added squared window are folded into
// one table:  bw[n] = w[n] / (M * (w[n mod H]^2 + w[n mod H + H]^2)).
constexpr int H = 256, kWarps3 = 4, kThreads3 = kWarps3 * 32;

struct RunPlan { int run_len; int runs_per_utt; long long total_runs; };

__global__ void __launch_bounds__(kThreads3, 3) mask_istft512_kernel(MaskIstftArgs a, RunPlan plan) {
    __shared__ __align__(16) float2 s_x[kWarps3 * 2][M];
    __shared__ __align__(16) float2 s_win2[M];        // analysis window pairs
    __shared__ __align__(16) float2 s_bw2[M];         // synthesis table pairs (see above)
    for (int i = threadIdx.x; i < M; i += kThreads3) {
        const float w0 = a.tab.window[2 * i], w1 = a.tab.window[2 * i + 1];
        s_win2[i] = make_float2(w0, w1);
        const int n0 = (2 * i) & (H - 1), n1 = (2 * i + 1) & (H - 1);
        const float e0 = a.tab.window[n0] * a.tab.window[n0] + a.tab.window[n0 + H] * a.tab.window[n0 + H];
        const float e1 = a.tab.window[n1] * a.tab.window[n1] + a.tab.window[n1 + H] * a.tab.window[n1 + H];
        s_bw2[i] = make_float2(w0 / (M * e0), w1 / (M * e1));
    }
    const int lane = threadIdx.x & 31, j = lane & 15, hw = threadIdx.x >> 4;
    const unsigned hmask = half_mask(lane);
    float2 tw[15], twn[8];
    load_lane_constants(j, a.tab.twM, a.tab.twN, tw, twn);
    __syncthreads();
    float2* xbuf = s_x[hw];
    const long long unit = (long long)blockIdx.x * (kThreads3 / 16) + hw;
    if (unit >= plan.total_runs) return;                          // no block-level barrier below
    const int u = (int)(unit / plan.runs_per_utt), ri = (int)(unit - (long long)u * plan.runs_per_utt);
    const int F = a.n_frames;
    const int b0 = 1 + ri * plan.run_len;
    const int b1 = min(F - 1, b0 + plan.run_len - 1);
    const float* nrow = a.noisy + (long long)u * a.utt_stride;
    const float* crow = a.clean ? a.clean + (long long)u * a.utt_stride : nullptr;
    float* orow = a.wav_out + (long long)u * a.out_stride;
    const int len = a.lengths ? (int)a.lengths[u] : a.T;
    const int valid_frames = min(F, len / H + 1);                  // runner.py:455
    const bool spec = a.want_spec && crow && a.sums;
    const bool out_aligned = (reinterpret_cast<uintptr_t>(orow) & 7) == 0;
    const bool clean_aligned = crow && (reinterpret_cast<uintptr_t>(crow) & 7) == 0;
    float acc[sekern::NSUMS];
#pragma unroll
    for (int i = 0; i < sekern::NSUMS; ++i) acc[i] = 0.0f;
    float2 carry[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) carry[q] = make_float2(0.0f, 0.0f);

    for (int f = b0 - 1; f <= b1; ++f) {
        const bool halo = (f == b0 - 1);
        const bool own = spec && (!halo || f == 0) && f < valid_frames;
        const float* mk = a.mask + ((long long)u * F + f) * a.mask_stride;
        float pta[8], ptb[8], pt128 = 0.0f;
        if (own) {                                                  // |STFT(clean)|^2 for the spectral SI-SDR sums
            float2 c[16];
            load_frame(crow, a.T, f * H - N / 2, j, s_win2, c);
            fft256<-1>(c, xbuf, j, tw, hmask);
            float2 cm[8];
            fetch_mirror(c, lane, cm);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float2 xa = rfft_split(c[q], cm[q], twn[q]);
                float2 xb = rfft_split(cm[q], c[q], make_float2(-twn[q].x, twn[q].y));
                if (j == 0 && q == 0) { xa.y = 0.0f; xb.y = 0.0f; }
                pta[q] = xa.x * xa.x + xa.y * xa.y;
                ptb[q] = xb.x * xb.x + xb.y * xb.y;
            }
            const float2 x = rfft_split(c[8], c[8], make_float2(0.0f, -1.0f));
            pt128 = x.x * x.x + x.y * x.y;
        }
        float2 v[16];
        load_frame(nrow, a.T, f * H - N / 2, j, s_win2, v);
        float ga[8], gb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { ga[q] = __ldg(mk + j + 16 * q); gb[q] = __ldg(mk + M - j - 16 * q); }
        const float g128 = __ldg(mk + 128);
        fft256<-1>(v, xbuf, j, tw, hmask);
        float2 zm[8];
        fetch_mirror(v, lane, zm);
        float2 za[8], zb[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float2 wk = twn[q], wmk = make_float2(-twn[q].x, twn[q].y);
            float2 xa = rfft_split(v[q], zm[q], wk);
            float2 xb = rfft_split(zm[q], v[q], wmk);
            if (j == 0 && q == 0) { xa.y = 0.0f; xb.y = 0.0f; }
            if (own) {
                const float ra = fmaxf(ga[q] * (xa.x * xa.x + xa.y * xa.y), 0.0f);      // relu(predicted), objective.py:89
                const float rb = fmaxf(gb[q] * (xb.x * xb.x + xb.y * xb.y), 0.0f);
                acc[sekern::SUM_SPEC_ST] += sqrtf(ra * pta[q]) + sqrtf(rb * ptb[q]);
                acc[sekern::SUM_SPEC_TT] += pta[q] + ptb[q];
                acc[sekern::SUM_SPEC_SS] += ra + rb;
            }
            const float2 ya = cscale(xa, sqrtf(ga[q])), yb = cscale(xb, sqrtf(gb[q]));
            za[q] = irfft_merge(ya, yb, wk);
            zb[q] = irfft_merge(yb, ya, wmk);
        }
        float2 z128;
        {
            const float2 x = rfft_split(v[8], v[8], make_float2(0.0f, -1.0f));
            if (own && j == 0) {
                const float r = fmaxf(g128 * (x.x * x.x + x.y * x.y), 0.0f);
                acc[sekern::SUM_SPEC_ST] += sqrtf(r * pt128);
                acc[sekern::SUM_SPEC_TT] += pt128;
                acc[sekern::SUM_SPEC_SS] += r;
            }
            const float2 y = cscale(x, sqrtf(g128));
            z128 = irfft_merge(y, y, make_float2(0.0f, -1.0f));
        }
        scatter_mirror(za, zb, z128, lane, v);
        fft256<+1>(v, xbuf, j, tw, hmask);                          // v[q] = (x[2m], x[2m+1]), m = j + 16 q
        if (!halo) {
            const int t0 = (f - 1) * H;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int m = j + 16 * q;
                const float2 bw = s_bw2[m];
                const float2 y = make_float2(carry[q].x + bw.x * v[q].x, carry[q].y + bw.y * v[q].y);
                const int t = t0 + 2 * m;
                if (out_aligned) *reinterpret_cast<float2*>(orow + t) = y;
                else { orow[t] = y.x; orow[t + 1] = y.y; }
                if (a.sums) {
                    float2 c = make_float2(0.0f, 0.0f);
                    if (crow) c = clean_aligned ? __ldg(reinterpret_cast<const float2*>(crow + t)) : make_float2(__ldg(crow + t), __ldg(crow + t + 1));
                    if (t < len) { acc[sekern::SUM_YY] += y.x * y.x; acc[sekern::SUM_YC] += y.x * c.x; acc[sekern::SUM_CC] += c.x * c.x; }
                    if (t + 1 < len) { acc[sekern::SUM_YY] += y.y * y.y; acc[sekern::SUM_YC] += y.y * c.y; acc[sekern::SUM_CC] += c.y * c.y; }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float2 bw = s_bw2[j + 16 * q + 128];
            carry[q] = make_float2(bw.x * v[q + 8].x, bw.y * v[q + 8].y);
        }
    }
    // the last run of an utterance also zero-fills [out_len, pad_to) and finishes sum c^2 over [out_len, len)
    if (b1 == F - 1) {
        for (int t = a.out_len + j; t < max(a.pad_to, len); t += 16) {
            if (t < a.pad_to) orow[t] = 0.0f;
            if (a.sums && crow && t < len) { const float c = __ldg(crow + t); acc[sekern::SUM_CC] += c * c; }
        }
    }
    if (a.sums) {
#pragma unroll
        for (int i = 0; i < sekern::NSUMS; ++i) {
            float s = acc[i];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(hmask, s, o);
            if (j == 0 && s != 0.0f) atomicAdd(a.sums + (long long)u * sekern::NSUMS + i, (double)s);
        }
    }
}

}  // namespace

namespace sefast {

int num_sms() {
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

int launch_stft512(const StftArgs& a, cudaStream_t st) {
    const long long total = (long long)a.n_utt * a.n_frames;
    const long long want = (total + (kThreads / 16) - 1) / (kThreads / 16);
    const long long cap = 2LL * num_sms();
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    stft512_kernel<<<grid, kThreads, 0, st>>>(a, total);
    return secommon::check_launch("stft512_kernel");
}

int launch_mask_istft512(const MaskIstftArgs& a, cudaStream_t st) {
    // run length: long enough that the halo frame is a small overhead, short enough to fill the GPU
    const long long blocks_total = (long long)a.n_utt * (a.n_frames - 1);
    const long long slots = 3LL * num_sms() * (kThreads3 / 16);           // resident half-warps
    long long rl = (blocks_total + slots - 1) / slots;
    static int forced = -1;
    if (forced < 0) { const char* e = getenv("SE_B200_RUN_LEN"); forced = e ? atoi(e) : 0; }
    if (forced > 0) rl = forced;
    else rl = rl < 8 ? 8 : (rl > 64 ? 64 : rl);
    RunPlan plan;
    plan.runs_per_utt = (int)((a.n_frames - 1 + rl - 1) / rl);
    plan.run_len = (int)((a.n_frames - 1 + plan.runs_per_utt - 1) / plan.runs_per_utt);     // balanced runs
    plan.total_runs = (long long)a.n_utt * plan.runs_per_utt;
    const long long grid = (plan.total_runs + (kThreads3 / 16) - 1) / (kThreads3 / 16);
    if (grid > 0x7fffffffLL) return secommon::fail(SE_ERR_BAD_ARG, "grid too large");
    mask_istft512_kernel<<<(unsigned)grid, kThreads3, 0, st>>>(a, plan);
    return secommon::check_launch("mask_istft512_kernel");
}

}  // namespace sefast
