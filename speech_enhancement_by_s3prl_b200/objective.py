"""Drop-in objectives (reference objective.py).  Instantiated by name with the config block
(``eval(f'{args.objective}(**cfg)')`` runner.py:83) and called with every local variable of
the training step as keyword arguments (runner.py:458, 575; sampler.py:90), so each
``forward`` names what it consumes and swallows the rest in ``**kwargs``.

Instead of the (B, F) int64 ``stft_length_masks`` the kernels take per-utterance frame
counts; they are recovered from the masks (prefix masks, runner.py:216-220) or passed
directly as ``stft_lengths`` by callers that have them.
"""
import torch
import torch.nn as nn

from . import ops


def _frames(stft_length_masks=None, stft_lengths=None):
    if stft_lengths is not None:
        return stft_lengths.to(torch.int64)
    return stft_length_masks.sum(dim=-1).to(torch.int64)


class SISDR(nn.Module):
    """objective.py:81-100 -- SI-SDR between spectral magnitudes sqrt(relu(power)); mean over the batch."""

    def __init__(self, eps=1e-10, **kwargs):
        super().__init__()
        self.eps = eps

    def forward(self, predicted, linear_tar, stft_length_masks=None, stft_lengths=None, **kwargs):
        per_utt = ops.sisdr_spec(predicted, linear_tar, _frames(stft_length_masks, stft_lengths), self.eps)
        return per_utt.mean(), {}


class L1(nn.Module):
    """objective.py:103-117 -- mean |log_predicted - log(linear_tar + eps)| over the valid elements of the batch."""

    def __init__(self, eps=1e-10, **kwargs):
        super().__init__()
        self.eps = eps

    def forward(self, log_predicted, linear_tar, stft_length_masks=None, stft_lengths=None, **kwargs):
        acc = ops.l1_logspec_sums(log_predicted, linear_tar, _frames(stft_length_masks, stft_lengths), self.eps)
        return (acc[0] / acc[1]).to(torch.float32), {}


class WSD(nn.Module):
    """objective.py:120-153 -- weighted speech distortion.  SURVEY.md 8f row 1 ("next"): kept on
    stock torch ops for now so the config key keeps working; not yet a fused kernel."""

    def __init__(self, alpha=0.5, db_interval=30, eps=1e-10, **kwargs):
        super().__init__()
        self.alpha, self.db_interval, self.eps = alpha, db_interval, eps

    def forward(self, linear_inp, offset, linear_tar, stft_length_masks, **kwargs):
        m = stft_length_masks.unsqueeze(-1)
        noise = torch.clamp(linear_inp - linear_tar, min=0.0)
        energy = linear_tar.sum(dim=-1, keepdim=True)
        thres = 10.0 * torch.log10(energy.max() + self.eps) - self.db_interval
        voiced = ((10.0 * torch.log10(energy + self.eps)) > thres).long()
        speech = ((linear_tar - offset * linear_tar) * voiced * m).pow(2).sum(-1).sum(-1).mean()
        noise_term = (offset * noise * m).pow(2).sum(-1).sum(-1).mean()
        return self.alpha * speech + (1.0 - self.alpha) * noise_term, {}
