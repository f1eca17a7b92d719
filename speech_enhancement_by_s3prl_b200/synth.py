"""Seeded synthetic noisy-speech batches in the reference's input layout (host side, CPU torch).

Recipe of SURVEY.md 8d: per utterance, seed = 1337 + global utterance index (1337 is the
reference's default seed, run_downstream.py:62); ``speech`` = five harmonically related,
amplitude-modulated sinusoids plus 1/f-shaped Gaussian noise; ``noise`` = white Gaussian; both
RMS-normalised to -25 dBFS (dataset.py:106-111); mixed at the requested SNR with the reference's
noise scaling (dataset.py:66-72); stacked ``[noisy, clean, scaled_noise]`` -> (T, 3)
(dataset.py:161) and collated to ``lengths (B,) int64, wavs (B, 3, Tmax)`` (dataset.py:169-179).
This is the data-preparation side the reference runs in DataLoader workers; it is not part of the
accelerated path.
"""
import math

import torch
from torch.nn.utils.rnn import pad_sequence

SNRS = (-8, -6, -4, -2, 0, 2, 4, 6, 8)


def normalize_db(audio, target_level=-25.0):
    rms = audio.pow(2).mean().pow(0.5)
    return audio * ((10.0 ** (target_level / 20.0)) / (rms + 1e-10))


def mix(speech, noise, snr_db, eps=1e-10):
    ratio = 10.0 ** (snr_db / 10.0)
    gain = (speech.pow(2).sum() / (ratio * noise.pow(2).sum() + eps)).pow(0.5)
    scaled = gain * noise
    return speech + scaled, scaled


def speech_like(T, gen, sample_rate=16000):
    t = torch.arange(T, dtype=torch.float32) / sample_rate
    f0 = 90.0 + 130.0 * torch.rand(1, generator=gen).item()
    sig = torch.zeros(T)
    for h in range(1, 6):
        phase = 2 * math.pi * torch.rand(1, generator=gen).item()
        sig += torch.sin(2 * math.pi * f0 * h * t + phase) / h
    rate = 2.0 + 3.0 * torch.rand(1, generator=gen).item()
    sig *= 0.55 + 0.45 * torch.sin(2 * math.pi * rate * t)
    white = torch.randn(T, generator=gen)
    spec = torch.fft.rfft(white)
    spec /= torch.arange(1, spec.numel() + 1, dtype=torch.float32).sqrt()      # ~1/f power
    pink = torch.fft.irfft(spec, n=T)
    return sig + 0.3 * pink / pink.std()


def utterance(T, index, snr_db=None):
    gen = torch.Generator().manual_seed(1337 + int(index))
    speech = normalize_db(speech_like(T, gen))
    noise = normalize_db(torch.randn(T, generator=gen))
    if snr_db is None:
        snr_db = SNRS[int(torch.randint(len(SNRS), (1,), generator=gen))]
    noisy, scaled = mix(speech, noise, float(snr_db))
    return torch.stack([noisy, speech, scaled], dim=-1)                          # (T, 3)


def collate(samples):
    lengths = torch.LongTensor([len(s) for s in samples])
    wavs = pad_sequence(samples, batch_first=True).transpose(-1, -2).contiguous()
    return lengths, wavs


def batch(n_utt, seconds, first_index=0, sample_rate=16000, snr_db=None, min_seconds=None):
    """(lengths, wavs (B, 3, Tmax)).  min_seconds: draw each length uniformly in [min_seconds, seconds]."""
    items = []
    for i in range(n_utt):
        T = int(seconds * sample_rate)
        if min_seconds is not None:
            gen = torch.Generator().manual_seed(99991 + first_index + i)
            T = int((min_seconds + (seconds - min_seconds) * torch.rand(1, generator=gen).item()) * sample_rate)
        items.append(utterance(T, first_index + i, snr_db))
    return collate(items)
