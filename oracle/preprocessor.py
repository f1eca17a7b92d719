"""Oracle restatement of S3PRL's ``utility.preprocessor.OnlinePreprocessor``.

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  PARITY UNPINNED: the class
restated here is not under ``/root/reference`` (un-vendored S3PRL dependency,
no version pin: reference README.md:12-27, requirements.txt has no s3prl
entry).  Every behaviour is reconstructed from the reference's call sites:

* constructor kwargs = the whole ``online:`` block        config/pretrain_sample.yaml:32-65, run_downstream.py:159
* ``n_fft=(n_freq-1)*2``, hop/win from ms                 config/pretrain_sample.yaml:39,46-48
* STFT over ``wavs.reshape(-1, T)`` with ``_window``       sampler.py:225-227
* ``linear`` = power spectrum, ``phase`` = atan2          sampler.py:228-229, objective.py:89-90
* ``center=True`` (frames = T // hop + 1)                 runner.py:455,572; sampler.py:68
* outputs time-major ``(B, F, D)``                        model.py:30, objective.py:22-23
* ``get_feat_config(feat_type, channel, log, delta, cmvn)`` run_downstream.py:153-156, runner.py:50
* no-argument call -> features of a dummy waveform        run_downstream.py:163,183; model.py:146
* ``istft(linears, phases) -> (B, hop*(F-1))``            runner.py:267-268

Choices that the call sites cannot show (kept fixed here):
periodic Hann window, ``pad_mode='reflect'``, ``normalized=False``, one-sided;
``log(x + 1e-10)``; HTK mel filterbank without normalisation on the power
spectrum; 5-tap regression deltas with replicate padding applied recursively;
CMVN over time with the unbiased std and ``+ eps``.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

FEAT_TYPES = ("complx", "linear", "phase", "mel", "mfcc")


def hz_to_mel_htk(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate):
    """HTK triangular filterbank, no area normalisation: (n_freqs, n_mels)."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(hz_to_mel_htk(f_min), hz_to_mel_htk(f_max), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)          # (n_freqs, n_mels+2)
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


def compute_deltas(x, win_length=5):
    """Regression deltas over the last axis with replicate padding.

    d[t] = sum_{i=-n..n} i * x[t+i] / (n(n+1)(2n+1)/3),  n = (win_length-1)//2
    """
    n = (win_length - 1) // 2
    denom = n * (n + 1) * (2 * n + 1) / 3.0
    shape = x.shape
    flat = x.reshape(-1, 1, shape[-1])
    flat = F.pad(flat, (n, n), mode="replicate")
    kernel = torch.arange(-n, n + 1, dtype=x.dtype, device=x.device).view(1, 1, -1)
    return (F.conv1d(flat, kernel) / denom).reshape(shape)


class OnlinePreprocessor(nn.Module):
    def __init__(self, sample_rate=16000, win_ms=25, hop_ms=10, n_freq=201, n_mels=40,
                 n_mfcc=13, feat_list=None, eps=1e-10, **kwargs):
        super().__init__()
        self._sample_rate = sample_rate
        self._n_freq = n_freq
        self._n_mels = n_mels
        win = round(win_ms * sample_rate / 1000)
        hop = round(hop_ms * sample_rate / 1000)
        n_fft = (n_freq - 1) * 2
        self._win_args = {"n_fft": n_fft, "hop_length": hop, "win_length": win}
        self.register_buffer("_window", torch.hann_window(win))
        self.register_buffer("_melfb", melscale_fbanks(n_freq, 0.0, sample_rate / 2.0, n_mels, sample_rate))
        gen = torch.Generator().manual_seed(0)
        self.register_buffer("_pseudo_wav", torch.randn(sample_rate, generator=gen))
        self.feat_list = feat_list
        self.eps = eps

    @classmethod
    def get_feat_config(cls, feat_type, channel=0, log=False, delta=0, cmvn=False):
        assert feat_type in FEAT_TYPES
        return {"feat_type": feat_type, "channel": channel, "log": log, "delta": delta, "cmvn": cmvn}

    # -- the two primitives the reference also reaches for directly (sampler.py:226-228)
    def _stft(self, x, window=None):
        window = self._window if window is None else window
        z = torch.stft(x, window=window, center=True, pad_mode="reflect", normalized=False,
                       onesided=True, return_complex=True, **self._win_args)
        return torch.view_as_real(z)                     # (rows, K, F, 2)

    @staticmethod
    def _magphase(complx):
        re, im = complx[..., 0], complx[..., 1]
        return re * re + im * im, torch.atan2(im, re)    # power, phase

    def _select(self, tensors, feat_type, channel=0, log=False, delta=0, cmvn=False):
        raw = tensors[feat_type].select(dim=-3, index=int(channel))      # (B, D, F)
        if bool(log):
            raw = (raw + self.eps).log()
        feats = [raw.contiguous()]
        for _ in range(int(delta)):
            feats.append(compute_deltas(feats[-1]))
        feats = torch.cat(feats, dim=-2)
        if bool(cmvn):
            feats = (feats - feats.mean(dim=-1, keepdim=True)) / (feats.std(dim=-1, keepdim=True) + self.eps)
        return feats

    def forward(self, wavs=None, feat_list=None):
        feat_list = self.feat_list if feat_list is None else feat_list
        assert feat_list is not None
        if wavs is None:
            n_ch = max(int(cfg.get("channel", 0)) for cfg in feat_list) + 1
            wavs = self._pseudo_wav.view(1, 1, -1).repeat(1, n_ch, 1)
        assert wavs.dim() >= 3                                            # (B, C, T)
        lead = wavs.shape[:-1]
        complx = self._stft(wavs.reshape(-1, wavs.shape[-1]))
        complx = complx.reshape(lead + complx.shape[-3:])                 # (B, C, K, F, 2)
        linear, phase = self._magphase(complx)
        tensors = {"linear": linear, "phase": phase}
        if any(cfg["feat_type"] == "mel" for cfg in feat_list):
            tensors["mel"] = torch.matmul(linear.transpose(-1, -2), self._melfb).transpose(-1, -2)
        if any(cfg["feat_type"] == "complx" for cfg in feat_list):
            tensors["complx"] = complx.transpose(-1, -2).reshape(lead + (-1, complx.shape[-2]))
        out = [self._select(tensors, **cfg) for cfg in feat_list]
        return [o.transpose(-1, -2).contiguous() for o in out]            # (B, F, D)

    def istft(self, linears, phases, linear_power=2):
        # linears, phases: (B, F, K) time-major power spectrum + phase
        mag = linears.transpose(-1, -2).pow(1.0 / linear_power)
        ph = phases.transpose(-1, -2)
        z = torch.polar(mag, ph)
        return torch.istft(z, window=self._window, center=True, normalized=False, onesided=True,
                           **self._win_args)
