"""Oracle restatement of the reference's in-repo hot-path arithmetic (CPU torch).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Each function names the
reference lines it follows; ``oracle/make_golden.py`` pins them against the
unmodified reference modules (``tests/golden/signal_path_*.npz``).
"""
import math

import torch
import torch.nn.functional as F
from torch.nn.utils.rnn import pad_sequence


# ----------------------------------------------------------------- data (a1, a13)
def normalize_wav_decibel(audio, target_level=-25.0):
    """dataset.py:106-111 -- RMS-normalise a 1-D signal to ``target_level`` dBFS."""
    rms = audio.pow(2).mean().pow(0.5)
    return audio * ((10.0 ** (target_level / 20.0)) / (rms + 1e-10))


def add_noise(speech, noise, snrs, eps=1e-10):
    """dataset.py:54-74 -- tile/crop ``noise`` to ``speech`` and mix at ``snrs`` dB.

    speech, noise: (B, T*), snrs: (B,).  Returns (noisy, scaled_noise).  The reference
    broadcasts snrs (B,) against (B, 1) powers, which only type-checks for B == 1 (the only
    way it is called: dataset.py:158, sampler.py:51); here snrs is viewed as (B, 1)."""
    t_s, t_n = speech.size(-1), noise.size(-1)
    if t_s >= t_n:
        reps, rem = divmod(t_s, t_n)
        noise = torch.cat([noise.repeat(1, reps), noise[:, :rem]], dim=-1)
    else:
        noise = noise[:, :t_s]
    ratio = 10.0 ** (snrs.reshape(-1, 1) / 10.0)
    p_s = speech.pow(2).sum(dim=-1, keepdim=True)
    p_n = noise.pow(2).sum(dim=-1, keepdim=True)
    gain = (p_s / (ratio * p_n + eps)).pow(0.5)
    scaled = gain * noise
    return speech + scaled, scaled


def collate(samples):
    """dataset.py:169-179 -- list of (T_i, 3) -> lengths (B,) int64, wavs (B, 3, Tmax)."""
    lengths = torch.LongTensor([len(s) for s in samples])
    wavs = pad_sequence(samples, batch_first=True).transpose(-1, -2).contiguous()
    return lengths, wavs


# ------------------------------------------------------------------ masks (a4)
def length_masks(lengths):
    """runner.py:216-220 / sampler.py:35-39 -- (B,) -> (B, max(lengths)) int64 0/1."""
    steps = torch.arange(int(lengths.max().item()), device=lengths.device)
    return (steps.unsqueeze(0) < lengths.unsqueeze(-1)).long()


def stft_lengths(lengths, hop):
    """runner.py:455, 572 -- frames that belong to each utterance."""
    return lengths // hop + 1


# ------------------------------------------------------------------ heads (a5, a6)
def linear_head(features, weight, bias, activation="ReLU"):
    """model.py:14-17 -- act(W feat + b)."""
    return getattr(torch.nn, activation)()(F.linear(features, weight, bias))


def linear_residual_head(features, linears, weight, bias, activation="Sigmoid", cmvn=True, eps=1e-6):
    """model.py:28-34 -- CMVN over time (unbiased std) -> Linear -> act -> linears * offset."""
    if cmvn:
        mu = features.mean(dim=1, keepdim=True)
        sd = features.std(dim=1, keepdim=True)
        features = (features - mu) / (sd + eps)
    offset = getattr(torch.nn, activation)()(F.linear(features, weight, bias))
    return linears * offset, offset


# ----------------------------------------------------------------- losses (a8, a9, f1)
def sisdr_spectral(predicted, linear_tar, stft_length_masks, eps=1e-10):
    """objective.py:86-100 -- SI-SDR between sqrt(relu(.)) magnitudes, mean over batch."""
    m = stft_length_masks.unsqueeze(-1)
    src = (F.relu(predicted).pow(0.5) * m).flatten(start_dim=1)
    tar = (F.relu(linear_tar).pow(0.5) * m).flatten(start_dim=1)
    alpha = (src * tar).sum(dim=1) / ((tar * tar).sum(dim=1) + eps)
    proj = alpha.unsqueeze(1) * tar
    resid = ((proj - src) ** 2).sum(dim=1) + eps
    per_utt = -10.0 * torch.log10((proj * proj).sum(dim=1) / resid + eps)
    return per_utt.mean(), per_utt


def l1_logspectral(log_predicted, linear_tar, stft_length_masks, eps=1e-10):
    """objective.py:109-117 -- mean |log_pred - log(tar+eps)| over valid elements of the batch."""
    keep = stft_length_masks.unsqueeze(-1).bool()
    src = log_predicted.masked_select(keep)
    tar = linear_tar.masked_select(keep)
    return (src - (tar + eps).log()).abs().mean()


def wsd(linear_inp, offset, linear_tar, stft_length_masks, alpha=0.5, db_interval=30, eps=1e-10):
    """objective.py:127-153 -- weighted speech-distortion loss."""
    m = stft_length_masks.unsqueeze(-1)
    noise = torch.clamp(linear_inp - linear_tar, min=0.0)
    energy = linear_tar.sum(dim=-1, keepdim=True)
    thres = 10.0 * torch.log10(energy.max() + eps) - db_interval
    voiced = ((10.0 * torch.log10(energy + eps)) > thres).long()
    speech_term = ((linear_tar - offset * linear_tar) * voiced * m).pow(2).sum(-1).sum(-1).mean()
    noise_term = (offset * noise * m).pow(2).sum(-1).sum(-1).mean()
    return alpha * speech_term + (1.0 - alpha) * noise_term


# ------------------------------------------------------ waveform level (a10, a11, a12)
def masked_mean(batch, masks, keepdim=False, eps=1e-8):
    """utils.py:26-29."""
    return (batch * masks).sum(dim=-1, keepdim=keepdim) / (masks.sum(dim=-1, keepdim=keepdim) + eps)


def masked_normalize_decibel(audio, target, masks, eps=1e-8):
    """utils.py:31-46 -- rescale each row so its masked mean-square hits the target level.

    ``target``: a number (dB, same for every row), a (B,) tensor of dB levels, or a
    (B, T) reference waveform whose masked level is used."""
    if isinstance(target, (int, float)):
        target = torch.full((len(audio),), float(target), device=audio.device)
    elif target.dim() > 1:
        target = 10.0 * torch.log10(masked_mean(target.pow(2), masks))
    gain_sq = (10.0 ** (target.unsqueeze(-1) / 10.0)) / (masked_mean(audio.pow(2), masks, keepdim=True) + eps)
    return audio * gain_sq.pow(0.5)


def decode_wav(preprocessor, linears, phases, lengths, target_level=-25):
    """runner.py:266-270 -- iSTFT, zero-pad to max(lengths), normalise level."""
    wav = preprocessor.istft(linears, phases)
    pad = int(max(lengths)) - wav.size(1)
    wav = torch.cat([wav, wav.new_zeros(wav.size(0), pad)], dim=1)
    return masked_normalize_decibel(wav, target_level, length_masks(lengths))


def sisdr_eval(src, tar, eps=1e-10):
    """evaluation.py:5-10 -- waveform SI-SDR of one utterance, Python float."""
    alpha = (src * tar).sum() / ((tar * tar).sum() + eps)
    proj = alpha * tar
    resid = ((proj - src) ** 2).sum() + eps
    return (10.0 * ((proj * proj).sum() / resid + eps).log10()).item()


# -------------------------------------------------- the un-fused eval step (runner.py:556-602)
def eval_step(preprocessor, head, lengths, wavs, objective="SISDR"):
    """The reference's evaluation step as written: 3-channel STFT, six feature
    tensors, head, `_decode_wav` against the clean level, criterion, per-utterance
    waveform SI-SDR.  ``head`` is a dict(weight=, bias=, activation=, cmvn=, eps=).
    Returns dict(loss, sisdr (B,), wav_predicted (B, Tmax), predicted, offset)."""
    feats_up, feats_down, linear_inp, phase_inp, linear_tar, phase_tar = preprocessor(wavs)
    wav_tar = wavs[:, preprocessor.channel_tar, :]
    predicted, offset = linear_residual_head(feats_down, linear_inp, **head)
    wav_predicted = decode_wav(preprocessor, predicted, phase_inp, lengths, wav_tar)
    hop = preprocessor._win_args["hop_length"]
    masks = length_masks(stft_lengths(lengths, hop))
    if objective == "SISDR":
        loss, _ = sisdr_spectral(predicted, linear_tar, masks)
    elif objective == "L1":
        loss = l1_logspectral((predicted + 1e-10).log(), linear_tar, masks)
    else:
        raise ValueError(objective)
    scores = [sisdr_eval(wav_predicted[b, :n], wav_tar[b, :n]) for b, n in enumerate(lengths.tolist())]
    return {"loss": loss, "sisdr": torch.tensor(scores), "wav_predicted": wav_predicted,
            "predicted": predicted, "offset": offset, "linear_inp": linear_inp,
            "linear_tar": linear_tar, "phase_inp": phase_inp, "feats_down": feats_down}


def recurrent_head(kind, state, features, linears, bidirectional=False, activation="Identity", cmvn=False, eps=1e-6):
    """model.py:37-60 (``LSTM``: log_predicted = act(Linear(lstm(x))), predicted = exp) and model.py:63-91 (``Residual``:
    offset = act(Linear([cmvn](lstm(x)))), predicted = linears * offset) from a reference state dict
    (keys ``lstm.*`` and ``scaling_layer.0.{weight,bias}``)."""
    hidden = state["lstm.weight_hh_l0"].shape[1]
    layers = 1 + max(int(k.split("_l")[1].split("_")[0]) for k in state if k.startswith("lstm.weight_ih_l"))
    lstm = torch.nn.LSTM(input_size=features.shape[-1], hidden_size=hidden, num_layers=layers, batch_first=True,
                         bidirectional=bidirectional)
    lstm.load_state_dict({k[5:]: v for k, v in state.items() if k.startswith("lstm.")})
    h, _ = lstm(features)
    if kind == "Residual" and cmvn:
        h = (h - h.mean(dim=1, keepdim=True)) / (h.std(dim=1, keepdim=True) + eps)
    out = getattr(torch.nn, activation)()(torch.nn.functional.linear(h, state["scaling_layer.0.weight"], state["scaling_layer.0.bias"]))
    if kind == "LSTM":
        return out.exp(), {"log_predicted": out}
    return linears * out, {"offset": out}
