"""Drop-in ``OnlinePreprocessor`` (S3PRL ``utility.preprocessor``) on the sm_100a kernels.

Same constructor kwargs (the whole ``online:`` block of config/pretrain_sample.yaml:32-65
is splatted in at run_downstream.py:159), same ``forward(wavs=None, feat_list=None)``
returning time-major ``(B, F, D)`` tensors (call sites runner.py:433, 558, 297, 51;
sampler.py:60), same ``istft(linears, phases)`` (runner.py:267), ``get_feat_config``
(run_downstream.py:153-156; runner.py:50) and the attributes the reference reaches for
(``_win_args``, ``_sample_rate``, ``_window``, ``_stft``, ``_magphase``, settable
``channel_inp`` / ``channel_tar``).  Behaviour is pinned by ``oracle/preprocessor.py``.

The kernels only run on a CUDA device.  The reference also calls a CPU deep copy of the
preprocessor on CPU data for audio logging (runner.py:50-51, 65): such inputs are moved
to the current CUDA device, processed there and the results moved back -- never
computed on the CPU.
"""
import math

import torch
import torch.nn as nn

from . import ops

FEAT_TYPES = ("complx", "linear", "phase", "mel", "mfcc")


def _hz_to_mel(f):
    return 2595.0 * math.log10(1.0 + f / 700.0)


def mel_filterbank(n_freqs, n_mels, sample_rate):
    """HTK triangular filters, no normalisation, f in [0, sr/2]: (n_freqs, n_mels) (host side, built once)."""
    freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_pts = torch.linspace(_hz_to_mel(0.0), _hz_to_mel(sample_rate / 2.0), n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    widths = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - freqs.unsqueeze(1)
    return torch.clamp(torch.min(-slopes[:, :-2] / widths[:-1], slopes[:, 2:] / widths[1:]), min=0.0).contiguous()


class OnlinePreprocessor(nn.Module):
    def __init__(self, sample_rate=16000, win_ms=25, hop_ms=10, n_freq=201, n_mels=40, n_mfcc=13,
                 feat_list=None, eps=1e-10, **kwargs):
        super().__init__()
        self._sample_rate = sample_rate
        self._n_freq = n_freq
        self._n_mels = n_mels
        win = round(win_ms * sample_rate / 1000)
        hop = round(hop_ms * sample_rate / 1000)
        n_fft = (n_freq - 1) * 2
        if n_fft not in ops.SUPPORTED_NFFT:
            raise ValueError(f"n_freq={n_freq} -> n_fft={n_fft}: supported n_fft are {ops.SUPPORTED_NFFT}")
        if win > n_fft:
            raise ValueError("win_ms longer than n_fft")
        self._win_args = {"n_fft": n_fft, "hop_length": hop, "win_length": win}
        self.register_buffer("_window", torch.hann_window(win))
        self.register_buffer("_frame_window", ops.centered_window(torch.hann_window(win), n_fft))
        self.register_buffer("_melfb", mel_filterbank(n_freq, n_mels, sample_rate))
        gen = torch.Generator().manual_seed(0)
        self.register_buffer("_pseudo_wav", torch.randn(sample_rate, generator=gen))
        self.feat_list = feat_list
        self.eps = eps

    @classmethod
    def get_feat_config(cls, feat_type, channel=0, log=False, delta=0, cmvn=False):
        assert feat_type in FEAT_TYPES
        return {"feat_type": feat_type, "channel": channel, "log": log, "delta": delta, "cmvn": cmvn}

    # ---- device handling: compute always happens on CUDA ---------------------------------
    def _compute_device(self, *tensors):
        for t in tensors:
            if t is not None and t.is_cuda:
                return t.device
        if self._window.is_cuda:
            return self._window.device
        if not torch.cuda.is_available():
            raise RuntimeError("se_b200.OnlinePreprocessor needs a CUDA device (no CPU fallback)")
        return torch.device("cuda", torch.cuda.current_device())

    def _tables(self, dev):
        return self._frame_window.to(dev), self._melfb.to(dev)

    # ---- the primitives the reference calls directly (sampler.py:226-228) ------------------
    def _stft(self, x, window=None):
        """(rows, T) -> (rows, K, F, 2) like torch.stft(...) viewed as real."""
        dev = self._compute_device(x)
        rows, T = x.shape
        n_fft, hop = self._win_args["n_fft"], self._win_args["hop_length"]
        win = self._frame_window.to(dev) if window is None else ops.centered_window(window.to(dev), n_fft)
        res = ops.stft(x.to(dev).reshape(rows, 1, T).contiguous(), 0, n_fft, hop, win, power=True, phase=True)
        mag = res["power"].sqrt()
        z = torch.stack([mag * res["phase"].cos(), mag * res["phase"].sin()], dim=-1)
        return z.transpose(1, 2).to(x.device)

    @staticmethod
    def _magphase(complx):
        re, im = complx[..., 0], complx[..., 1]
        return re * re + im * im, torch.atan2(im, re)

    # ---- forward ---------------------------------------------------------------------------
    def forward(self, wavs=None, feat_list=None):
        feat_list = self.feat_list if feat_list is None else feat_list
        assert feat_list is not None, "feat_list was given neither at construction nor at call time"
        if wavs is None:                              # run_downstream.py:163,183: learn the feature dims
            n_ch = max(int(cfg.get("channel", 0)) for cfg in feat_list) + 1
            wavs = self._pseudo_wav.view(1, 1, -1).repeat(1, n_ch, 1)
        assert wavs.dim() == 3, "wavs must be (batch, channel, samples)"
        home = wavs.device
        dev = self._compute_device(wavs)
        x = wavs.to(device=dev, dtype=torch.float32).contiguous()
        window, melfb = self._tables(dev)
        n_fft, hop = self._win_args["n_fft"], self._win_args["hop_length"]

        need = {}                                     # channel -> set of kernel outputs
        for cfg in feat_list:
            ft, ch = cfg["feat_type"], int(cfg.get("channel", 0))
            if ft in ("complx", "mfcc"):
                raise NotImplementedError(f"feat_type {ft!r} is outside the accelerated path (no config selects it)")
            want = need.setdefault(ch, set())
            if ft == "phase":
                want.add("phase")
            elif ft == "linear" and bool(cfg.get("log", False)):
                want.add("logpower")
            else:
                want.add("power")
        spectra = {ch: ops.stft(x, ch, n_fft, hop, window, power="power" in w, phase="phase" in w,
                                logpower="logpower" in w, log_eps=self.eps) for ch, w in need.items()}

        outs = []
        for cfg in feat_list:
            ft, ch = cfg["feat_type"], int(cfg.get("channel", 0))
            log, delta, cmvn = bool(cfg.get("log", False)), int(cfg.get("delta", 0)), bool(cfg.get("cmvn", False))
            sp = spectra[ch]
            if ft == "mel" and delta <= 2 and self._n_mels <= 64:
                # K1b fused: mel -> log -> deltas (-> CMVN) in one launch (two with CMVN), final layout
                outs.append(ops.mel_features(sp["power"], melfb, log, self.eps, order=delta, cmvn=cmvn).to(home))
                continue
            if ft == "mel":
                feat = ops.mel(sp["power"], melfb, log, self.eps, out_cols=(delta + 1) * self._n_mels)
                base = self._n_mels
            else:
                feat = sp["phase"] if ft == "phase" else (sp["logpower"] if log else sp["power"])
                if ft == "phase" and log:
                    feat = (feat + self.eps).log()
                base = feat.shape[-1]
                if delta > 0:
                    wide = feat.new_empty(feat.shape[0], feat.shape[1], (delta + 1) * base)
                    wide[..., :base] = feat
                    feat = wide
                elif cmvn:
                    feat = feat.clone()               # cmvn is applied in place; keep the shared spectrum intact
            if delta > 0:
                ops.delta_(feat, base, delta)
            if cmvn:
                mean, std = ops.cmvn_stats(feat)
                ops.cmvn_apply_(feat, mean, std, self.eps)
            outs.append(feat.to(home))
        return outs

    # ---- inverse ---------------------------------------------------------------------------
    def istft(self, linears, phases, linear_power=2):
        """(B, F, K) power (or magnitude**linear_power) + phase -> (B, hop*(F-1))."""
        home = linears.device
        dev = self._compute_device(linears, phases)
        power = linears.to(dev) if linear_power == 2 else linears.to(dev).pow(2.0 / linear_power)
        window, _ = self._tables(dev)
        wav = ops.istft(power.float(), phases.to(dev).float(), self._win_args["n_fft"], self._win_args["hop_length"], window)
        return wav.to(home)
