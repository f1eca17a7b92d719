"""The hot-path glue of reference runner.py as free functions: length masks
(runner.py:216-220, sampler.py:35-39), frame counts (runner.py:455) and ``_decode_wav``
(runner.py:266-270)."""
import torch

from . import ops
from .utils import masked_normalize_decibel


def stft_lengths(lengths, hop):
    return lengths // hop + 1


def get_length_masks(lengths, ascending=None):
    """(B,) int64 -> (B, max(lengths)) int64 0/1.  ``ascending`` (the reference's fixed arange table,
    runner.py:32,79) is accepted and ignored: the mask is generated from the lengths, so audio
    longer than 50 s is handled too."""
    return ops.length_masks(lengths)


def decode_wav(preprocessor, linears, phases, lengths, target_level=-25):
    """runner.py:266-270: iSTFT -> zero-pad to max(lengths) -> level normalisation."""
    n_fft, hop = preprocessor._win_args["n_fft"], preprocessor._win_args["hop_length"]
    pad_to = int(lengths.max().item()) if torch.is_tensor(lengths) else int(max(lengths))
    wav = ops.istft(linears, phases, n_fft, hop, preprocessor._frame_window.to(linears.device), pad_to=pad_to)
    lengths = torch.as_tensor(lengths, device=wav.device, dtype=torch.int64)
    return masked_normalize_decibel(wav, target_level, lengths)


def pseudo_wav(preprocessor, linear_predicted, wavs, lengths, channel=None, target_level=-25):
    """runner.py:272-305 (``_pseudo_clean`` / ``_pseudo_noise`` -> ``_decode_wav``): waveform of a predicted power spectrum
    (the upstream SpecHead's output) with the phase of ``wavs[:, channel]`` (default ``preprocessor.channel_inp``), zero-padded
    to max(lengths) and level-normalised to ``target_level`` dB -- one fused kernel (STFT of the noisy frames, polar with the
    predicted magnitude, iSTFT, overlap-add, sum of squares) plus the gain pass; no phase tensor, no complex spectrum."""
    n_fft, hop = preprocessor._win_args["n_fft"], preprocessor._win_args["hop_length"]
    ch = int(getattr(preprocessor, "channel_inp", 0)) if channel is None else int(channel)
    wavs = wavs.to(linear_predicted.device)
    lengths = torch.as_tensor(lengths, device=wavs.device, dtype=torch.int64)
    T = wavs.shape[2]
    window = preprocessor._frame_window.to(wavs.device)
    wav, sums = ops.mask_istft(wavs, ch, None, linear_predicted, lengths, n_fft, hop, window, pad_to=T, want_sums=True,
                               want_spec=False, mask_is_power=True)
    ops.finalize_metrics(sums, lengths, T, wav=wav, target_db=target_level, want_gain=False, want_sisdr=False, want_loss=False)
    return wav[:, :int(lengths.max().item())]
