"""CPU: the tile bodies of csrc/tile_kernels.cuh (the code the CUDA kernels run per CTA)
executed serially under g++ (tests/host_harness.cpp) and compared with numpy / the oracle.
This checks Stockham indexing, twiddles, padding, reflect framing, split/merge and
overlap-add without a GPU; the -m gpu tests repeat the comparisons on the device."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import stft_f64
from oracle.preprocessor import OnlinePreprocessor

HERE = os.path.dirname(os.path.abspath(__file__))
CFGS = {256: (129, 16, 8), 400: (201, 25, 10), 512: (257, 32, 16), 1024: (513, 64, 16), 2048: (1025, 128, 32)}


@pytest.fixture(scope="session")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("harness") / "host_harness.so"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", str(out),
                           os.path.join(HERE, "host_harness.cpp")])
    return ctypes.CDLL(str(out))


def fptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


@pytest.mark.parametrize("n_fft", [256, 400, 512, 1024, 2048])
@pytest.mark.parametrize("direction", [-1, 1])
def test_complex_fft_matches_numpy(harness, n_fft, direction):
    M = n_fft // 2
    rng = np.random.default_rng(n_fft)
    x = (rng.standard_normal((5, M)) + 1j * rng.standard_normal((5, M))).astype(np.complex64)
    out = np.zeros_like(x)
    assert harness.h_fft(n_fft, fptr(x), fptr(out), 5, direction) == 0
    ref = np.fft.fft(x.astype(np.complex128), axis=1) if direction < 0 else np.fft.ifft(x.astype(np.complex128), axis=1) * M
    assert np.abs(out - ref).max() / np.abs(ref).max() < 1e-6


def make_pre(n_fft):
    n_freq, win_ms, hop_ms = CFGS[n_fft]
    return OnlinePreprocessor(win_ms=win_ms, hop_ms=hop_ms, n_freq=n_freq)


@pytest.mark.parametrize("n_fft,T", [(512, 4000), (512, 4096), (512, 300), (400, 3333), (1024, 5000), (256, 777), (2048, 9000)])
def test_stft_tile_matches_oracle(harness, n_fft, T):
    pre = make_pre(n_fft)
    hop = pre._win_args["hop_length"]
    K = n_fft // 2 + 1
    g = torch.Generator().manual_seed(T)
    wavs = torch.randn(2, 3, T, generator=g) * 0.05
    F = T // hop + 1
    win = pre._window.numpy().astype(np.float32)
    for ch in (0, 1):
        x = np.ascontiguousarray(wavs.numpy())
        power = np.zeros((2, F, K), np.float32)
        phase = np.zeros_like(power)
        logp = np.zeros_like(power)
        rc = harness.h_stft(n_fft, ctypes.c_void_p(x.ctypes.data + 4 * ch * T), 2, ctypes.c_longlong(3 * T), T, hop,
                            fptr(win), fptr(power), fptr(phase), fptr(logp), ctypes.c_float(1e-10))
        assert rc == 0
        c = pre.get_feat_config
        lin, ph, lg = pre(wavs, [c("linear", ch), c("phase", ch), c("linear", ch, log=True)])
        z64 = stft_f64.stft(wavs[:, ch].numpy(), n_fft, hop)
        scale = np.abs(z64).max() ** 2
        assert np.abs(power - lin.numpy()).max() / scale < 1e-5        # fp32 vs fp32 library FFT
        assert np.abs(power - np.abs(z64) ** 2).max() / scale < 1e-5   # vs exact arithmetic
        # phase only where the bin is not numerically empty
        strong = np.abs(z64) > 1e-3 * np.abs(z64).max()
        d = np.angle(np.exp(1j * (phase - np.angle(z64))))
        assert np.abs(d[strong]).max() < 1e-3
        np.testing.assert_allclose(logp, np.log(power + 1e-10), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n_fft,T,pad", [(512, 4000, 4000), (512, 4096, 4096), (400, 3333, 3333), (1024, 5000, 5000), (512, 1200, 0)])
def test_istft_tile_matches_oracle(harness, n_fft, T, pad):
    pre = make_pre(n_fft)
    hop = pre._win_args["hop_length"]
    g = torch.Generator().manual_seed(T + 1)
    wavs = torch.randn(2, 1, T, generator=g) * 0.05
    c = pre.get_feat_config
    lin, ph = pre(wavs, [c("linear", 0), c("phase", 0)])
    lin = lin * torch.rand(lin.shape, generator=g)              # not a consistent spectrogram any more
    ref = pre.istft(lin, ph).numpy()
    F = lin.shape[1]
    width = max(pad, hop * (F - 1))
    out = np.full((2, width + 3), 7.0, np.float32)
    win = pre._window.numpy().astype(np.float32)
    rc = harness.h_istft(n_fft, fptr(np.ascontiguousarray(lin.numpy())), fptr(np.ascontiguousarray(ph.numpy())), 2, F, hop,
                         fptr(win), fptr(out), ctypes.c_longlong(width + 3), pad)
    assert rc == 0
    n = hop * (F - 1)
    assert np.abs(out[:, :n] - ref).max() < 2e-6
    assert (out[:, n:width] == 0).all() and (out[:, width:] == 7.0).all()


@pytest.mark.parametrize("n_fft,T", [(512, 4000), (512, 4096), (400, 3333), (1024, 5000), (512, 2560)])
def test_mask_istft_tile_matches_unfused_oracle(harness, n_fft, T):
    pre = make_pre(n_fft)
    hop = pre._win_args["hop_length"]
    K = n_fft // 2 + 1
    g = torch.Generator().manual_seed(T + 2)
    B = 3
    wavs = torch.randn(B, 3, T, generator=g) * 0.05
    wavs[:, 1] = wavs[:, 0] * 0.8 + 0.01 * torch.randn(B, T, generator=g)
    lengths = torch.LongTensor([T, T - 517, T // 2])
    F = T // hop + 1
    mask = torch.rand(B, F, K, generator=g)
    c = pre.get_feat_config
    lin, ph, lin_t = pre(wavs, [c("linear", 0), c("phase", 0), c("linear", 1)])
    ref = pre.istft(lin * mask, ph)
    ref = torch.cat([ref, ref.new_zeros(B, T - ref.shape[1])], 1).numpy()
    out = np.full((B, T), 7.0, np.float32)
    sums = np.zeros((B, 6), np.float64)
    x = np.ascontiguousarray(wavs.numpy())
    win = pre._window.numpy().astype(np.float32)
    rc = harness.h_mask_istft(n_fft, fptr(x), ctypes.c_void_p(x.ctypes.data + 4 * T), ctypes.c_longlong(3 * T),
                              fptr(np.ascontiguousarray(mask.numpy())), fptr(lengths.numpy()), B, T, hop, fptr(win),
                              fptr(out), ctypes.c_longlong(T), T, fptr(sums), 1)
    assert rc == 0
    assert np.abs(out - ref).max() < 2e-6
    clean = wavs[:, 1].numpy().astype(np.float64)
    for b in range(B):
        n = int(lengths[b])
        y = ref[b, :n].astype(np.float64)
        np.testing.assert_allclose(sums[b, 0], (y * clean[b, :n]).sum(), rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(sums[b, 1], (clean[b, :n] ** 2).sum(), rtol=1e-5)
        np.testing.assert_allclose(sums[b, 2], (y * y).sum(), rtol=1e-4)
        nfr = n // hop + 1
        src = np.sqrt((lin[b, :nfr] * mask[b, :nfr]).numpy().astype(np.float64))
        tar = np.sqrt(lin_t[b, :nfr].numpy().astype(np.float64))
        np.testing.assert_allclose(sums[b, 3], (src * tar).sum(), rtol=1e-4)
        np.testing.assert_allclose(sums[b, 4], (tar * tar).sum(), rtol=1e-4)
        np.testing.assert_allclose(sums[b, 5], (src * src).sum(), rtol=1e-4)
