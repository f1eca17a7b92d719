"""Drop-in metric (reference evaluation.py:5-10) plus its batched device form.

The reference evaluates SI-SDR one utterance at a time on CPU slices through joblib
(runner.py:587-602).  ``sisdr_eval(src, tar)`` keeps that signature (1-D tensors, returns a
Python float) but computes on the GPU; ``sisdr_eval_batch`` does the whole batch in one
launch and is what the fused evaluation step uses.
"""
import torch

from . import ops


def sisdr_eval_batch(src, tar, lengths=None, eps=1e-10):
    """src, tar: (B, T) CUDA tensors, lengths (B,) int64 -> (B,) SI-SDR in dB."""
    return ops.sisdr_wave(src, tar, lengths, eps)


def sisdr_eval(src, tar, sr=16000, eps=1e-10):
    if not torch.cuda.is_available():
        raise RuntimeError("se_b200.sisdr_eval needs a CUDA device (no CPU fallback)")
    dev = src.device if src.is_cuda else torch.device("cuda", torch.cuda.current_device())
    s = src.reshape(1, -1).to(dev, torch.float32).contiguous()
    t = tar.reshape(1, -1).to(dev, torch.float32).contiguous()
    return ops.sisdr_wave(s, t, None, eps).item()
