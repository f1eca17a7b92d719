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
