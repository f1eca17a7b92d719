// Host-side planning shared by the CUDA library and the g++ test harness:
// twiddle/window tables (computed in double, rounded once to float) and the tile geometry.
#pragma once
#include <cmath>
#include <vector>

namespace seplan {

inline bool supported_nfft(int n) { return n == 256 || n == 400 || n == 512 || n == 1024 || n == 2048; }

// twM[j] = exp(-2*pi*i*j/M) (j < M), twN[k] = exp(-2*pi*i*k/N) (k <= M), interleaved re/im
inline void make_twiddles(int n_fft, std::vector<float>& twM, std::vector<float>& twN) {
    const int M = n_fft / 2;
    const double pi = 3.14159265358979323846;
    twM.resize(2 * M);
    twN.resize(2 * (M + 2));
    for (int j = 0; j < M; ++j) {
        twM[2 * j] = (float)std::cos(2.0 * pi * j / M);
        twM[2 * j + 1] = (float)(-std::sin(2.0 * pi * j / M));
    }
    for (int k = 0; k <= M; ++k) {
        twN[2 * k] = (float)std::cos(2.0 * pi * k / n_fft);
        twN[2 * k + 1] = (float)(-std::sin(2.0 * pi * k / n_fft));
    }
    twN[2 * (M + 1)] = twN[2 * (M + 1) + 1] = 0.0f;
}

// Output samples per inverse tile such that the frames overlapping a tile never exceed G.
// Always a multiple of hop (tile i owns frames [i*len/hop, (i+1)*len/hop)).
inline int inverse_tile_len(int n_fft, int hop, int G) {
    int halo = ((n_fft / 2) % hop == 0 && n_fft % hop == 0) ? n_fft / hop - 1 : (n_fft - 1 + hop - 1) / hop;
    int ft = G - halo;
    return ft > 0 ? ft * hop : 0;
}

// tiles needed so that every output sample in [0, max(out_len, pad_to)) and every frame is owned
inline int inverse_num_tiles(int out_len, int pad_to, int tile_len) {
    int cover = out_len > pad_to ? out_len : pad_to;
    int a = out_len / tile_len + 1;
    int b = (cover + tile_len - 1) / tile_len;
    return a > b ? a : b;
}

}  // namespace seplan
