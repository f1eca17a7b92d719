"""Drop-in waveform-level helpers (reference utils.py:26-46)."""
import torch

from . import ops


def _lengths_of(length_masks):
    """(B, T) 0/1 prefix masks (runner.py:216-220) or already a (B,) length vector."""
    if length_masks.dim() == 1:
        return length_masks.to(torch.int64)
    return length_masks.sum(dim=-1).to(torch.int64)


def masked_mean(batch, length_masks, keepdim=False, eps=1e-8):
    """utils.py:26-29 (host-side convenience; the kernels compute this internally)."""
    return (batch * length_masks).sum(dim=-1, keepdim=keepdim) / (length_masks.sum(dim=-1, keepdim=keepdim) + eps)


def masked_normalize_decibel(audio, target, length_masks, eps=1e-8):
    """utils.py:31-46 -- ``target``: dB number, (B,) tensor of dB levels or a (B, T) reference waveform.
    ``length_masks`` may be the reference's (B, T) int64 masks or simply the (B,) lengths."""
    lengths = _lengths_of(length_masks).to(audio.device)
    if isinstance(target, (int, float)):
        tdb = torch.full((audio.shape[0],), float(target), device=audio.device)
        return ops.masked_normalize_db(audio, lengths, target_db=tdb, eps=eps)
    if target.dim() > 1:
        return ops.masked_normalize_db(audio, lengths, ref=target, eps=eps)
    return ops.masked_normalize_db(audio, lengths, target_db=target.to(torch.float32), eps=eps)
