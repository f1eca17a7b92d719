"""CPU: the oracle restatements against (a) outputs of the unmodified reference modules
(tests/golden/signal_path_ref.npz, runner_evaluate_ref.npz -- made by oracle/make_golden.py)
and (b) an independent float64 direct-DFT STFT."""
import os

import numpy as np
import pytest
import torch

from oracle import signal_path as sp
from oracle import stft_f64
from oracle.preprocessor import OnlinePreprocessor, compute_deltas, melscale_fbanks


@pytest.fixture(scope="module")
def ref(golden_dir):
    return np.load(os.path.join(golden_dir, "signal_path_ref.npz"))


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_losses_match_reference(ref):
    masks = T(ref["masks"])
    loss, _ = sp.sisdr_spectral(T(ref["predicted"]), T(ref["linear_tar"]), masks)
    np.testing.assert_allclose(loss.numpy(), ref["SISDR"], rtol=1e-6)
    l1 = sp.l1_logspectral(T(ref["log_predicted"]), T(ref["linear_tar"]), masks)
    np.testing.assert_allclose(l1.numpy(), ref["L1"], rtol=1e-6)
    w = sp.wsd(T(ref["linear_inp"]), T(ref["offset"]), T(ref["linear_tar"]), masks, alpha=0.3, db_interval=50)
    np.testing.assert_allclose(w.numpy(), ref["WSD"], rtol=1e-6)


def test_loss_gradients_match_reference(ref):
    masks = T(ref["masks"])
    p = T(ref["predicted"]).clone().requires_grad_(True)
    sp.sisdr_spectral(p, T(ref["linear_tar"]), masks)[0].backward()
    np.testing.assert_allclose(p.grad.numpy(), ref["SISDR_grad"], rtol=1e-5, atol=1e-8)
    lp = T(ref["log_predicted"]).clone().requires_grad_(True)
    sp.l1_logspectral(lp, T(ref["linear_tar"]), masks).backward()
    np.testing.assert_allclose(lp.grad.numpy(), ref["L1_grad"], rtol=1e-6, atol=1e-9)


def test_length_masks(ref):
    np.testing.assert_array_equal(sp.length_masks(T(ref["stft_len"])).numpy(), ref["masks"])
    assert sp.length_masks(T(ref["stft_len"])).dtype == torch.int64


def test_waveform_level(ref):
    assert sp.sisdr_eval(T(ref["ev_src"]), T(ref["ev_tar"])) == pytest.approx(float(ref["sisdr_eval"]), abs=1e-5)
    assert sp.sisdr_eval(T(ref["ev_src"]), T(ref["ev_src"])) == pytest.approx(float(ref["sisdr_eval_self"]), abs=1e-4)
    m = sp.length_masks(T(ref["nd_len"]))
    a, r = T(ref["nd_audio"]), T(ref["nd_ref"])
    np.testing.assert_allclose(sp.masked_mean(a, m).numpy(), ref["masked_mean"], rtol=1e-6)
    np.testing.assert_allclose(sp.masked_normalize_decibel(a, -25, m).numpy(), ref["nd_scalar"], rtol=1e-6)
    np.testing.assert_allclose(sp.masked_normalize_decibel(a, r, m).numpy(), ref["nd_tensor"], rtol=1e-6)


def test_mixing_and_collate(ref):
    noisy, scaled = sp.add_noise(T(ref["an_speech"]), T(ref["an_noise"]), T(ref["an_snrs"]))
    np.testing.assert_allclose(noisy.numpy(), ref["an_noisy"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(scaled.numpy(), ref["an_scaled"], rtol=1e-6, atol=1e-7)
    noisy2, scaled2 = sp.add_noise(T(ref["an_speech"]), T(ref["an_long_noise"]), T(ref["an_snrs"]))
    np.testing.assert_allclose(noisy2.numpy(), ref["an_noisy2"], rtol=1e-6, atol=1e-7)
    # requested SNR is hit exactly (SURVEY 8c known-answer)
    s = T(ref["an_speech"])
    snr = 10 * torch.log10(s.pow(2).sum() / scaled.pow(2).sum())
    assert snr.item() == pytest.approx(5.0, abs=1e-4)
    np.testing.assert_allclose(sp.normalize_wav_decibel(s[0]).numpy(), ref["norm_db"], rtol=1e-6)
    items = T(ref["co_items"])
    parts = [items[:40], items[40:65], items[65:98]]
    lengths, wavs = sp.collate(parts)
    np.testing.assert_array_equal(lengths.numpy(), ref["co_lengths"])
    np.testing.assert_array_equal(wavs.numpy(), ref["co_wavs"])
    assert wavs.is_contiguous() and wavs.shape == (3, 3, 40)


def test_heads_match_reference(ref):
    pred, off = sp.linear_residual_head(T(ref["lr_feats"]), T(ref["lr_linears"]), T(ref["lr_weight"]), T(ref["lr_bias"]))
    np.testing.assert_allclose(pred.numpy(), ref["lr_predicted"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(off.numpy(), ref["lr_offset"], rtol=1e-5, atol=1e-7)
    lin = sp.linear_head(T(ref["lr_feats"]), T(ref["li_weight"]), T(ref["li_bias"]))
    np.testing.assert_allclose(lin.numpy(), ref["li_predicted"], rtol=1e-5, atol=1e-6)


def test_eval_step_matches_reference_runner(golden_dir):
    """oracle.eval_step == the reference's own Runner.evaluate() (loss and SI-SDR)."""
    g = np.load(os.path.join(golden_dir, "runner_evaluate_ref.npz"))
    pre = OnlinePreprocessor(win_ms=32, hop_ms=16, n_freq=257)
    c = pre.get_feat_config
    pre.feat_list = [c("linear", 0, log=True), c("linear", 0, log=True), c("linear", 0), c("phase", 0),
                     c("linear", 1), c("phase", 1)]
    pre.channel_inp, pre.channel_tar = 0, 1
    head = dict(weight=T(g["weight"]), bias=T(g["bias"]))
    items = [T(g[f"item{i}"]) for i in range(len(g["lengths"]))]
    losses, scores = [], []
    for i in range(0, len(items), 2):
        lengths, wavs = sp.collate(items[i:i + 2])
        with torch.no_grad():
            out = sp.eval_step(pre, head, lengths, wavs)
        losses.append(out["loss"].item())
        scores.append(out["sisdr"].mean().item())
        if i == 0:
            np.testing.assert_allclose(out["wav_predicted"][0].numpy(), g["enhanced0"], rtol=1e-4, atol=1e-6)
    assert np.mean(losses) == pytest.approx(float(g["loss"]), abs=1e-4)
    assert np.mean(scores) == pytest.approx(float(g["scores"][0]), abs=1e-3)


# ----------------------------------------------------------------- STFT contract
@pytest.mark.parametrize("n_freq,win_ms,hop_ms,T_", [(257, 32, 16, 2500), (201, 25, 10, 1777), (513, 64, 16, 3100),
                                                     (257, 32, 16, 600)])
def test_torch_stft_oracle_vs_float64_dft(n_freq, win_ms, hop_ms, T_):
    g = torch.Generator().manual_seed(T_)
    pre = OnlinePreprocessor(win_ms=win_ms, hop_ms=hop_ms, n_freq=n_freq)
    n_fft, hop = pre._win_args["n_fft"], pre._win_args["hop_length"]
    x = torch.randn(2, T_, generator=g) * 0.05
    z32 = torch.view_as_complex(pre._stft(x)).transpose(1, 2).numpy()          # (R, F, K)
    z64 = stft_f64.stft(x.numpy(), n_fft, hop)
    assert z32.shape == z64.shape == (2, T_ // hop + 1, n_freq)
    scale = np.abs(z64).max()
    assert np.abs(z32 - z64).max() / scale < 2e-6
    # inverse: torch.istft oracle vs float64 overlap-add, and the round trip itself
    lin, ph = pre._magphase(torch.view_as_real(torch.from_numpy(z32)))
    y32 = pre.istft(lin, ph).numpy()
    y64 = stft_f64.istft(z64, n_fft, hop)
    assert y32.shape == y64.shape == (2, hop * (T_ // hop))
    assert np.abs(y32 - y64).max() < 2e-6
    assert np.abs(y64 - x.numpy()[:, :y64.shape[1]]).max() < 1e-12


def test_preprocessor_golden_is_stable(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocessor_oracle.npz"))
    for tag, (n_freq, win_ms, hop_ms) in {"n512": (257, 32, 16), "n400": (201, 25, 10), "n1024": (513, 64, 16)}.items():
        pre = OnlinePreprocessor(win_ms=win_ms, hop_ms=hop_ms, n_freq=n_freq)
        c = pre.get_feat_config
        feats = pre(T(g[f"{tag}_wavs"]), [c("linear", 0), c("phase", 0), c("linear", 1, log=True),
                                          c("mel", 0, log=True, delta=2), c("mel", 1, log=True, delta=1, cmvn=True)])
        scale = g[f"{tag}_linear0"].max()
        np.testing.assert_allclose(feats[0].numpy(), g[f"{tag}_linear0"], rtol=1e-4, atol=1e-6 * scale)
        np.testing.assert_allclose(feats[3].numpy(), g[f"{tag}_mel_d2"], rtol=1e-3, atol=1e-3)
        assert feats[3].shape[-1] == 120 and feats[4].shape[-1] == 80
        np.testing.assert_allclose(pre.istft(feats[0], feats[1]).numpy(), g[f"{tag}_istft"], atol=2e-6)


def test_features_against_torchaudio():
    ta = pytest.importorskip("torchaudio")
    x = torch.randn(3, 40, 57)
    np.testing.assert_allclose(compute_deltas(x).numpy(), ta.functional.compute_deltas(x).numpy(), rtol=1e-5, atol=1e-6)
    for n_freq in (201, 257, 513):
        fb = ta.functional.melscale_fbanks(n_freq, 0.0, 8000.0, 40, 16000, norm=None, mel_scale="htk")
        np.testing.assert_allclose(melscale_fbanks(n_freq, 0.0, 8000.0, 40, 16000).numpy(), fb.numpy(), atol=1e-6)


def test_no_wav_call_returns_dummy_features():
    pre = OnlinePreprocessor()
    c = pre.get_feat_config
    out = pre(feat_list=[c("mel", 0, log=True, delta=1, cmvn=True), c("linear", 1)])
    assert out[0].shape[-1] == 80 and out[1].shape[-1] == 201 and out[0].shape[1] == 16000 // 160 + 1


@pytest.mark.parametrize("tag,cls,kw", [("lstm_uni", "LSTM", dict(bidirectional=False)), ("lstm_bi", "LSTM", dict(bidirectional=True)),
                                        ("res_uni", "Residual", dict(bidirectional=False, activation="Sigmoid", cmvn=False)),
                                        ("res_bi_cmvn", "Residual", dict(bidirectional=True, activation="Sigmoid", cmvn=True)),
                                        ("res_relu", "Residual", dict(bidirectional=False, activation="ReLU", cmvn=True))])
def test_recurrent_head_golden_is_reproducible(golden_dir, tag, cls, kw):
    """tests/golden/recurrent_heads_ref.npz (outputs of the reference's model.LSTM / model.Residual, model.py:37-91): the
    restatement in oracle/signal_path.py reproduces outputs and loss from the stored state dict, and the drop-in
    classes accept that state dict by name (strict) -- the contract the GPU parity test builds on."""
    import speech_enhancement_by_s3prl_b200 as se
    gold = np.load(os.path.join(golden_dir, "recurrent_heads_ref.npz"))
    state = {k[len(tag) + 7:]: T(gold[k]) for k in gold.files if k.startswith(f"{tag}_param_")}
    predicted, res = sp.recurrent_head(cls, state, T(gold["feats"]), T(gold["linears"]), **kw)
    np.testing.assert_allclose(predicted.detach().numpy(), gold[f"{tag}_predicted"], rtol=1e-5, atol=1e-6)
    for k, v in res.items():
        np.testing.assert_allclose(v.detach().numpy(), gold[f"{tag}_{k}"], rtol=1e-5, atol=1e-6)
    head = getattr(se, cls)(input_size=9, output_size=9, hidden_size=12, num_layers=2, **kw)
    head.load_state_dict(state)                                    # strict: same names and shapes as the reference


def test_scoring_golden_is_self_consistent(golden_dir):
    """tests/golden/scoring_ref.npz (the reference's own sampler.scoring / matching, LSTM head + L1): the matching formula
    (sampler.py:113-116) restated in numpy reproduces the stored match scores from the stored gradient embeddings, the
    batch-mean embedding is the element-count-weighted mean of the per-utterance ones (objective.py:113-116 averages over
    ALL valid elements), and the parameter layout is the drop-in LSTM's."""
    import speech_enhancement_by_s3prl_b200 as se
    g = np.load(os.path.join(golden_dir, "scoring_ref.npz"))
    per, mean, lengths = g["per_utt"].astype(np.float64), g["mean"].astype(np.float64), g["lengths"]
    q, k = per[:2], per[2:]
    qn = q / (np.sqrt((q ** 2).sum(-1, keepdims=True)) + 1e-12)
    kn = k / (np.sqrt((k ** 2).sum(-1, keepdims=True)) + 1e-12)
    np.testing.assert_allclose(kn @ qn.mean(0), g["match"], rtol=1e-5, atol=1e-6)
    frames = lengths // 160 + 1
    w = frames / frames.sum()
    np.testing.assert_allclose((per * w[:, None]).sum(0), mean[0], rtol=2e-3, atol=1e-6 * np.abs(mean).max())
    head = se.LSTM(input_size=201, output_size=201, hidden_size=24, num_layers=2)
    assert [n for n, _ in head.named_parameters()] == [str(n) for n in g["param_names"]]
    assert per.shape[1] == sum(p.numel() for p in head.parameters())


def test_oracle_training_steps_match_reference_runner_train(golden_dir):
    """tests/golden/runner_train_ref.npz = four optimizer steps of the reference's own Runner.train() (runner.py:431-471:
    preprocessor -> LinearResidual -> SISDR -> backward -> clip_grad_norm_(1.0) -> Adam).  The oracle restatement of that
    step reproduces every loss and the final weights."""
    g = np.load(os.path.join(golden_dir, "runner_train_ref.npz"))
    items = [T(g[f"item{i}"]) for i in range(len(g["lengths"]))]
    lengths, wavs = sp.collate(items)
    pre = OnlinePreprocessor(sample_rate=16000, win_ms=32, hop_ms=16, n_freq=257)
    c = pre.get_feat_config
    w = T(g["w0"]).clone().requires_grad_(True)
    b = T(g["b0"]).clone().requires_grad_(True)
    opt = torch.optim.Adam([w, b], lr=float(g["lr"]), betas=(0.9, 0.999))
    masks = sp.length_masks(sp.stft_lengths(lengths, 256))
    feats, lin_i, lin_t = pre(wavs, [c("linear", 0, log=True), c("linear", 0), c("linear", 1)])
    losses = []
    for _ in range(int(g["steps"])):
        pred, _ = sp.linear_residual_head(feats, lin_i, w, b)
        loss, _ = sp.sisdr_spectral(pred, lin_t, masks)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_([w, b], float(g["grad_clip"]))
        opt.step()
        losses.append(loss.item())
    np.testing.assert_allclose(losses, g["losses"], atol=2e-4)
    np.testing.assert_allclose(w.detach().numpy(), g["w1"], atol=2e-5)
    np.testing.assert_allclose(b.detach().numpy(), g["b1"], atol=2e-5)
