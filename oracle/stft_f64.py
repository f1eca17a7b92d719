"""Independent float64 restatement of the STFT / iSTFT contract (NumPy, direct DFT).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  This does not call any FFT
library: frames are multiplied by an explicit DFT matrix in float64, so it
anchors the ``torch.stft``-based oracle (``oracle/preprocessor.py``) and gives
the error budget of the fp32 paths against exact arithmetic.

Contract (torch.stft, center=True, pad_mode='reflect', periodic Hann, one-sided,
normalized=False; torch.istft with the same arguments):
  x_pad[p] = x[reflect(p - N/2)],  reflect(-i) = i, reflect(T-1+i) = T-1-i
  X[f, k]  = sum_n w[n] * x_pad[f*H + n] * exp(-2j*pi*k*n/N),  f = 0..T//H, k = 0..N/2
  y_pad[p] = sum_f w[p-fH] * irfft(Y[f])[p-fH] / sum_f w[p-fH]^2,  y[t] = y_pad[t + N/2], t < H*(F-1)
"""
import numpy as np


def hann_periodic(win, n_fft=None):
    w = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(win) / win)
    if n_fft is not None and n_fft > win:           # torch centres a short window in the frame
        left = (n_fft - win) // 2
        w = np.pad(w, (left, n_fft - win - left))
    return w


def reflect_index(p, T):
    """Index into x for padded coordinate p (may be <0 or >=T), single reflection."""
    p = np.where(p < 0, -p, p)
    return np.where(p >= T, 2 * (T - 1) - p, p)


def stft(x, n_fft, hop, window=None):
    """x: (R, T) float -> complex128 (R, F, K) time-major."""
    x = np.asarray(x, dtype=np.float64)
    R, T = x.shape
    w = hann_periodic(n_fft) if window is None else np.asarray(window, dtype=np.float64)
    F = T // hop + 1
    K = n_fft // 2 + 1
    n = np.arange(n_fft)
    pos = (np.arange(F)[:, None] * hop + n[None, :]) - n_fft // 2          # (F, N) original coords
    idx = reflect_index(pos, T)
    frames = x[:, idx] * w[None, None, :]                                   # (R, F, N)
    dft = np.exp(-2j * np.pi * np.outer(n, np.arange(K)) / n_fft)           # (N, K)
    return frames @ dft


def istft(Y, n_fft, hop, window=None):
    """Y: complex (R, F, K) time-major -> (R, hop*(F-1)) float64."""
    Y = np.asarray(Y, dtype=np.complex128)
    R, F, K = Y.shape
    w = hann_periodic(n_fft) if window is None else np.asarray(window, dtype=np.float64)
    n = np.arange(n_fft)
    k = np.arange(K)
    # irfft by explicit synthesis: imaginary parts of DC and Nyquist are ignored
    scale = np.full(K, 2.0)
    scale[0] = 1.0
    if n_fft % 2 == 0:
        scale[-1] = 1.0
    basis = np.exp(2j * np.pi * np.outer(k, n) / n_fft) * scale[:, None]    # (K, N)
    Yc = Y.copy()
    Yc[..., 0] = Yc[..., 0].real
    if n_fft % 2 == 0:
        Yc[..., -1] = Yc[..., -1].real
    frames = (Yc @ basis).real / n_fft                                      # (R, F, N)
    total = hop * (F - 1) + n_fft
    y = np.zeros((R, total))
    env = np.zeros(total)
    for f in range(F):
        y[:, f * hop:f * hop + n_fft] += frames[:, f] * w[None, :]
        env[f * hop:f * hop + n_fft] += w * w
    start = n_fft // 2
    stop = start + hop * (F - 1)
    return y[:, start:stop] / env[None, start:stop]
