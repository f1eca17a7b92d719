// g++ harness: runs the tile bodies of csrc/tile_kernels.cuh serially on the CPU
// (HostExec) so tests/test_fft_core.py can compare them with numpy without a GPU.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../speech_enhancement_by_s3prl_b200/csrc/tile_kernels.cuh"
#include "../speech_enhancement_by_s3prl_b200/csrc/host_plan.h"

using namespace sekern;

struct HostExec {
    template <class F> void foreach(int n, F f) { for (int w = 0; w < n; ++w) f(w); }
    void sync() {}
    template <int NS> void block_accumulate(const float* acc, double* dst) { for (int i = 0; i < NS; ++i) dst[i] += (double)acc[i]; }
};

template <int N> static int run_fft(const float* in, float* out, int nfft, int dir) {
    constexpr int M = N / 2, PAD = Padded<M>::SIZE;
    std::vector<float> twM, twN;
    seplan::make_twiddles(N, twM, twN);
    std::vector<float2> bx(nfft * PAD), by(nfft * PAD);
    HostExec ex;
    const float2* zin = reinterpret_cast<const float2*>(in);
    auto load0 = [=](int g, int i) { return zin[g * M + i]; };
    const float2* z = dir < 0 ? fft_run<Plan<M>, -1>(ex, nfft, load0, bx.data(), by.data(), reinterpret_cast<const float2*>(twM.data()))
                              : fft_run<Plan<M>, +1>(ex, nfft, load0, bx.data(), by.data(), reinterpret_cast<const float2*>(twM.data()));
    for (int g = 0; g < nfft; ++g)
        for (int k = 0; k < M; ++k) { out[2 * (g * M + k)] = z[g * PAD + phys(k)].x; out[2 * (g * M + k) + 1] = z[g * PAD + phys(k)].y; }
    return 0;
}

template <int N> static int run_stft(StftArgs a) {
    std::vector<float> twM, twN;
    seplan::make_twiddles(N, twM, twN);
    a.tab.twM = reinterpret_cast<const float2*>(twM.data());
    a.tab.twN = reinterpret_cast<const float2*>(twN.data());
    constexpr int G = Cfg<N>::G;
    std::vector<unsigned char> smem(Smem<N>::bytes((G - 1) * a.hop + N, 0));
    HostExec ex;
    const int tiles = (a.n_frames + G - 1) / G;
    for (int u = 0; u < a.n_utt; ++u)
        for (int t = 0; t < tiles; ++t) stft_tile<N>(ex, a, u, t, smem.data());
    return 0;
}

template <int N> static int run_istft(IstftArgs a) {
    std::vector<float> twM, twN;
    seplan::make_twiddles(N, twM, twN);
    a.tab.twM = reinterpret_cast<const float2*>(twM.data());
    a.tab.twN = reinterpret_cast<const float2*>(twN.data());
    constexpr int G = Cfg<N>::G;
    a.tile_len = seplan::inverse_tile_len(N, a.hop, G);
    if (a.tile_len <= 0) return -3;
    std::vector<unsigned char> smem(Smem<N>::bytes(0, 0));
    HostExec ex;
    const int tiles = seplan::inverse_num_tiles(a.out_len, a.pad_to, a.tile_len);
    for (int u = 0; u < a.n_utt; ++u)
        for (int t = 0; t < tiles; ++t) istft_tile<N>(ex, a, u, t, smem.data());
    return 0;
}

template <int N> static int run_mask_istft(MaskIstftArgs a) {
    std::vector<float> twM, twN;
    seplan::make_twiddles(N, twM, twN);
    a.tab.twM = reinterpret_cast<const float2*>(twM.data());
    a.tab.twN = reinterpret_cast<const float2*>(twN.data());
    constexpr int G = Cfg<N>::G;
    a.tile_len = seplan::inverse_tile_len(N, a.hop, G);
    if (a.tile_len <= 0) return -3;
    std::vector<unsigned char> smem(Smem<N>::bytes((G - 1) * a.hop + N, G * (N / 2 + 1)));
    HostExec ex;
    const int tiles = seplan::inverse_num_tiles(a.out_len, a.pad_to, a.tile_len);
    for (int u = 0; u < a.n_utt; ++u)
        for (int t = 0; t < tiles; ++t) mask_istft_tile<N>(ex, a, u, t, smem.data());
    return 0;
}

#define DISPATCH(n_fft, CALL)              \
    switch (n_fft) {                       \
        case 256: return CALL(256);        \
        case 400: return CALL(400);        \
        case 512: return CALL(512);        \
        case 1024: return CALL(1024);      \
        case 2048: return CALL(2048);      \
        default: return -2;                \
    }

extern "C" {

int h_fft(int n_fft, const float* in, float* out, int nfft, int dir) {
#define CALL(NN) run_fft<NN>(in, out, nfft, dir)
    DISPATCH(n_fft, CALL)
#undef CALL
}

int h_stft(int n_fft, const float* wav, int n_utt, long long utt_stride, int T, int hop, const float* window,
           float* power, float* phase, float* logp, float log_eps) {
    StftArgs a{};
    a.wav = wav; a.utt_stride = utt_stride; a.n_utt = n_utt; a.T = T; a.hop = hop; a.n_frames = T / hop + 1;
    a.tab.window = window; a.power = power; a.phase = phase; a.logp = logp; a.log_eps = log_eps; a.spec_stride = n_fft / 2 + 1;
#define CALL(NN) run_stft<NN>(a)
    DISPATCH(n_fft, CALL)
#undef CALL
}

int h_istft(int n_fft, const float* power, const float* phase, int n_utt, int n_frames, int hop, const float* window,
            float* wav_out, long long out_stride, int pad_to) {
    IstftArgs a{};
    a.power = power; a.phase = phase; a.n_utt = n_utt; a.n_frames = n_frames; a.hop = hop; a.tab.window = window;
    a.wav_out = wav_out; a.out_stride = out_stride; a.out_len = hop * (n_frames - 1); a.pad_to = pad_to;
#define CALL(NN) run_istft<NN>(a)
    DISPATCH(n_fft, CALL)
#undef CALL
}

int h_mask_istft(int n_fft, const float* noisy, const float* clean, long long utt_stride, const float* mask,
                 const long long* lengths, int n_utt, int T, int hop, const float* window, float* wav_out,
                 long long out_stride, int pad_to, double* sums, int want_spec) {
    MaskIstftArgs a{};
    a.noisy = noisy; a.clean = clean; a.utt_stride = utt_stride; a.mask = mask; a.lengths = lengths;
    a.n_utt = n_utt; a.T = T; a.hop = hop; a.n_frames = T / hop + 1; a.tab.window = window;
    a.wav_out = wav_out; a.out_stride = out_stride; a.out_len = hop * (a.n_frames - 1); a.pad_to = pad_to;
    a.sums = sums; a.want_spec = want_spec; a.mask_stride = n_fft / 2 + 1;
#define CALL(NN) run_mask_istft<NN>(a)
    DISPATCH(n_fft, CALL)
#undef CALL
}

}  // extern "C"
