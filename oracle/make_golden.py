"""Generate ``tests/golden/*.npz`` (run in the build container: ``python -m oracle.make_golden``).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

* ``signal_path_ref.npz``  -- outputs of the UNMODIFIED reference modules
  (objective.py, evaluation.py, utils.py, dataset.py, model.py, sampler.py,
  runner.py imported through ``oracle/ref_loader.py``) on small seeded inputs.
  The inputs are stored next to the outputs, so the fixture is self-contained.
* ``runner_evaluate_ref.npz`` -- loss / SI-SDR returned by the reference's own
  ``Runner.evaluate()`` (runner.py:521-622) driven over a 3-batch synthetic
  dataset with the restated preprocessor, the reference ``LinearResidual`` and
  the reference ``SISDR`` criterion.
* ``recurrent_heads_ref.npz`` -- the reference's ``model.LSTM`` and ``model.Residual`` (model.py:37-91) at small sizes:
  state dicts, inputs, outputs and the gradients of a scalar loss w.r.t. EVERY parameter (the LSTM below the
  projection included), so the drop-in heads' projection + multiply / exp -- forward, weight gradient and the input
  gradient that reaches the LSTM -- are pinned by the reference itself.
  (``python -m oracle.make_golden recurrent`` regenerates this file alone.)
* ``runner_train_ref.npz`` -- the reference's own ``Runner.train()`` (runner.py:307-518) driven for four optimizer steps
  (``--optim Adam``, gradient clipping 1.0) over a one-batch synthetic dataset with the restated preprocessor, the reference
  ``LinearResidual`` and ``SISDR``: the loss of every step and the weights before / after.
  (``python -m oracle.make_golden train`` regenerates this file alone.)
* ``scoring_ref.npz`` -- the reference's own ``sampler.scoring()`` (sampler.py:59-110) and ``sampler.matching()``
  (sampler.py:113-116) driven with the restated preprocessor, the reference ``model.LSTM`` and ``objective.L1`` -- the
  combination ``run_active.sh`` names -- on a small ragged batch: per-utterance gradient embeddings over ALL parameters,
  the batch-mean embedding, and the match scores.  (``python -m oracle.make_golden scoring`` regenerates this file alone.)
* ``preprocessor_oracle.npz`` -- outputs of ``oracle/preprocessor.py``
  (torch.stft / torch.istft based) for small inputs; guards against drift of the
  oracle itself across torch versions (this one is NOT a reference output).
"""
import os
import types
from argparse import Namespace

import numpy as np
import torch

from . import ref_loader
from .preprocessor import OnlinePreprocessor

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def speechlike(T, gen):
    """Small non-white test signal: AM harmonics + coloured noise."""
    t = torch.arange(T, dtype=torch.float32) / 16000.0
    f0 = 90.0 + 120.0 * torch.rand(1, generator=gen).item()
    sig = torch.zeros(T)
    for h in range(1, 6):
        sig += torch.sin(2 * np.pi * f0 * h * t + 6.28 * torch.rand(1, generator=gen).item()) / h
    sig *= 0.6 + 0.4 * torch.sin(2 * np.pi * 3.0 * t)
    noise = torch.randn(T, generator=gen)
    sig += 0.1 * torch.cumsum(noise, 0) / (1 + torch.arange(T)) ** 0.5
    return sig


def make_signal_path(ref):
    g = torch.Generator().manual_seed(1337)
    out = {}
    B, Fr, K = 3, 11, 9
    predicted = torch.randn(B, Fr, K, generator=g) * 2.0 + 0.5          # includes negatives (relu path)
    linear_tar = torch.rand(B, Fr, K, generator=g) * 3.0
    linear_inp = linear_tar + torch.randn(B, Fr, K, generator=g)
    log_predicted = torch.randn(B, Fr, K, generator=g)
    offset = torch.sigmoid(torch.randn(B, Fr, K, generator=g))
    stft_len = torch.LongTensor([11, 7, 4])
    masks = ref["sampler"].get_length_masks(stft_len, torch.arange(100))
    out.update(predicted=_np(predicted), linear_tar=_np(linear_tar), linear_inp=_np(linear_inp),
               log_predicted=_np(log_predicted), offset=_np(offset), stft_len=_np(stft_len), masks=_np(masks))
    out["SISDR"] = _np(ref["objective"].SISDR()(predicted, linear_tar, masks)[0])
    out["L1"] = _np(ref["objective"].L1()(log_predicted, linear_tar, masks)[0])
    out["WSD"] = _np(ref["objective"].WSD(alpha=0.3, db_interval=50)(linear_inp, offset, linear_tar, masks)[0])

    # gradients of the two objectives (for the closed-form backward kernels)
    p = predicted.clone().requires_grad_(True)
    ref["objective"].SISDR()(p, linear_tar, masks)[0].backward()
    out["SISDR_grad"] = _np(p.grad)
    lp = log_predicted.clone().requires_grad_(True)
    ref["objective"].L1()(lp, linear_tar, masks)[0].backward()
    out["L1_grad"] = _np(lp.grad)

    src = torch.randn(2000, generator=g) * 0.05
    tar = src * 0.7 + torch.randn(2000, generator=g) * 0.01
    out.update(ev_src=_np(src), ev_tar=_np(tar))
    out["sisdr_eval"] = np.float64(ref["evaluation"].sisdr_eval(src, tar))
    out["sisdr_eval_self"] = np.float64(ref["evaluation"].sisdr_eval(src, src))

    audio = torch.randn(3, 50, generator=g) * 0.1
    refwav = torch.randn(3, 50, generator=g) * 0.02
    wlen = torch.LongTensor([50, 31, 8])
    wmask = ref["sampler"].get_length_masks(wlen, torch.arange(100))
    out.update(nd_audio=_np(audio), nd_ref=_np(refwav), nd_len=_np(wlen))
    out["nd_scalar"] = _np(ref["utils"].masked_normalize_decibel(audio, -25, wmask))
    out["nd_tensor"] = _np(ref["utils"].masked_normalize_decibel(audio, refwav, wmask))
    out["masked_mean"] = _np(ref["utils"].masked_mean(audio, wmask))

    # the reference always mixes one utterance at a time (dataset.py:158, sampler.py:51):
    # its broadcast of snrs (B,) against (B,1) powers only type-checks for B == 1.
    speech = torch.randn(1, 300, generator=g)
    noise = torch.randn(1, 120, generator=g)
    snrs = torch.tensor([5.0])
    noisy, scaled = ref["dataset"].add_noise(speech, noise, snrs)
    out.update(an_speech=_np(speech), an_noise=_np(noise), an_snrs=_np(snrs), an_noisy=_np(noisy), an_scaled=_np(scaled))
    long_noise = torch.randn(1, 500, generator=g)
    noisy2, scaled2 = ref["dataset"].add_noise(speech, long_noise, snrs)
    out.update(an_long_noise=_np(long_noise), an_noisy2=_np(noisy2), an_scaled2=_np(scaled2))
    fake = types.SimpleNamespace(target_level=-25)
    out["norm_db"] = _np(ref["dataset"].OnlineDataset.normalize_wav_decibel(fake, speech[0]))

    items = [torch.randn(n, 3, generator=g) for n in (40, 25, 33)]
    lengths, wavs = ref["dataset"].OnlineDataset.collate_fn(None, items)
    out.update(co_items=np.concatenate([_np(i) for i in items]), co_lengths=_np(lengths), co_wavs=_np(wavs))

    torch.manual_seed(1337)
    head = ref["model"].LinearResidual(input_size=K, output_size=K)
    feats = torch.randn(B, Fr, K, generator=g) * 3 - 4
    pred, res = head(features=feats, linears=linear_inp.abs())
    out.update(lr_weight=_np(head.linear.weight), lr_bias=_np(head.linear.bias), lr_feats=_np(feats),
               lr_linears=_np(linear_inp.abs()), lr_predicted=_np(pred), lr_offset=_np(res["offset"]))
    lin = ref["model"].Linear(K, K, activation="ReLU")
    out.update(li_weight=_np(lin.linear.weight), li_bias=_np(lin.linear.bias),
               li_predicted=_np(lin(features=feats)[0]))
    return out


class _SynthSet(torch.utils.data.Dataset):
    def __init__(self, lengths, seed, ref):
        g = torch.Generator().manual_seed(seed)
        self.items = []
        for n in lengths:
            fake = types.SimpleNamespace(target_level=-25)
            norm = lambda a: ref["dataset"].OnlineDataset.normalize_wav_decibel(fake, a)
            speech = norm(speechlike(n, g))
            noise = norm(torch.randn(n, generator=g))
            noisy, scaled = ref["dataset"].add_noise(speech[None], noise[None], torch.ones(1) * 5.0)
            self.items.append(torch.stack([noisy[0], speech, scaled[0]], dim=-1))
        self._collate = ref["dataset"].OnlineDataset.collate_fn

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]

    def collate_fn(self, samples):
        return self._collate(None, samples)


def make_runner_evaluate(ref):
    """Drive the reference's own Runner.evaluate()."""
    import tempfile
    pre = OnlinePreprocessor(sample_rate=16000, win_ms=32, hop_ms=16, n_freq=257, n_mels=40, n_mfcc=13)
    pre.feat_list = [
        pre.get_feat_config("linear", 0, log=True), pre.get_feat_config("linear", 0, log=True),
        pre.get_feat_config("linear", 0), pre.get_feat_config("phase", 0),
        pre.get_feat_config("linear", 1), pre.get_feat_config("phase", 1)]
    pre.channel_inp, pre.channel_tar = 0, 1
    torch.manual_seed(1337)
    head = ref["model"].LinearResidual(input_size=257, output_size=257)
    args = Namespace(gpu=False, objective="SISDR", dropout=None, dropout2=None, optim="Adam", resume=None,
                     seed=1337, from_waveform=False, from_rawfeature=True, no_metric=False, n_jobs=1,
                     save_best=None, eval_init=False, sync_sampler=False, sampler_device=None,
                     active_sampling=False, pseudo_clean=False, pseudo_noise=False)
    config = {"runner": {"gradient_clipping": 1.0, "eval_metrics": ["sisdr"], "learning_rate": 1e-4,
                         "total_step": 4, "eval_splits": [], "log_step": 100, "media_step": 100,
                         "eval_step": 100, "save_step": 100, "max_keep": 1},
              "objective": {}, "dataloader": {"batch_size": 2, "eval_batch_size": 2}}
    lengths = [4000, 3333, 2900, 4100, 3000, 3601]
    dset = _SynthSet(lengths, seed=7, ref=ref)
    with tempfile.TemporaryDirectory() as tmp:
        runner = ref["runner"].Runner(args, config, pre, torch.nn.Identity(), torch.nn.Identity(), head, tmp)
        runner.set_model()
        args.n_jobs = 0                                   # DataLoader workers (runner.py:211)
        loader = runner.get_dataloader(dset, train=False)
        args.n_jobs = 1                                   # joblib workers (runner.py:599)
        loss, scores, noisy, clean, enhanced = runner.evaluate(loader)
        runner.manager.shutdown()
    out = {"lengths": np.array(lengths), "loss": _np(loss), "scores": _np(scores),
           "weight": _np(head.linear.weight), "bias": _np(head.linear.bias),
           "enhanced0": _np(enhanced[0]), "n_fft": np.int64(512), "hop": np.int64(256)}
    for i, it in enumerate(dset.items):
        out[f"item{i}"] = _np(it)
    return out


def make_recurrent_heads(ref):
    """model.LSTM (log_predicted -> exp) and model.Residual (offset * linears), uni- and bidirectional, with CMVN."""
    g = torch.Generator().manual_seed(4242)
    out = {}
    B, Fr, Din, Dh, K = 3, 17, 9, 12, 9
    feats = torch.randn(B, Fr, Din, generator=g)
    linears = torch.rand(B, Fr, K, generator=g) * 2.0
    out.update(feats=_np(feats), linears=_np(linears))
    cases = {"lstm_uni": ("LSTM", dict(bidirectional=False, activation="Identity")),
             "lstm_bi": ("LSTM", dict(bidirectional=True, activation="Identity")),
             "res_uni": ("Residual", dict(bidirectional=False, activation="Sigmoid", cmvn=False)),
             "res_bi_cmvn": ("Residual", dict(bidirectional=True, activation="Sigmoid", cmvn=True)),
             "res_relu": ("Residual", dict(bidirectional=False, activation="ReLU", cmvn=True))}
    for tag, (cls, kw) in cases.items():
        torch.manual_seed(1337)
        head = getattr(ref["model"], cls)(input_size=Din, output_size=K, hidden_size=Dh, num_layers=2, **kw)
        predicted, res = head(features=feats, linears=linears)
        loss = (predicted * torch.linspace(0.5, 1.5, K)).pow(2).mean() + 0.1 * predicted.mean()
        loss.backward()
        out[f"{tag}_predicted"] = _np(predicted)
        for k, v in res.items():
            out[f"{tag}_{k}"] = _np(v)
        out[f"{tag}_loss"] = _np(loss)
        for name, p in head.named_parameters():
            out[f"{tag}_param_{name}"] = _np(p)
            out[f"{tag}_grad_{name}"] = _np(p.grad)
    return out


def make_runner_train(ref):
    """Drive the reference's own Runner.train() (unmodified) for four steps and record losses and weights."""
    import tempfile
    pre = OnlinePreprocessor(sample_rate=16000, win_ms=32, hop_ms=16, n_freq=257, n_mels=40, n_mfcc=13)
    c = pre.get_feat_config
    pre.feat_list = [c("linear", 0, log=True), c("linear", 0, log=True), c("linear", 0), c("phase", 0), c("linear", 1), c("phase", 1)]
    pre.channel_inp, pre.channel_tar = 0, 1
    torch.manual_seed(1337)
    head = ref["model"].LinearResidual(input_size=257, output_size=257)
    w0, b0 = head.linear.weight.detach().clone(), head.linear.bias.detach().clone()
    args = Namespace(gpu=False, objective="SISDR", dropout=None, dropout2=None, optim="Adam", resume=None,
                     seed=1337, from_waveform=False, from_rawfeature=True, no_metric=False, n_jobs=0,
                     save_best=None, eval_init=False, sync_sampler=False, sampler_device=None,
                     active_sampling=False, pseudo_clean=False, pseudo_noise=False)
    steps = 4
    config = {"runner": {"gradient_clipping": 1.0, "eval_metrics": ["sisdr"], "learning_rate": 1e-3,
                         "total_step": steps, "eval_splits": [], "log_step": 1000, "media_step": 1000,
                         "eval_step": 1000, "save_step": 1000, "max_keep": 1},
              "objective": {}, "dataloader": {"batch_size": 4, "eval_batch_size": 4}}
    lengths = [9000, 8333, 8200, 9100]                       # >= 32 frames each: inside the fused training route's range
    dset = _SynthSet(lengths, seed=11, ref=ref)
    losses = []
    with tempfile.TemporaryDirectory() as tmp:
        runner = ref["runner"].Runner(args, config, pre, torch.nn.Identity(), torch.nn.Identity(), head, tmp)
        runner.set_model()
        runner.criterion.register_forward_hook(lambda mod, inp, out: losses.append(float(out[0].detach())))
        runner.get_dataset = lambda *a, **k: dset            # (instance attribute: the class and its source stay untouched)
        torch.manual_seed(5)                                 # DataLoader shuffle order (one batch per epoch: order only)
        runner.train()
        runner.manager.shutdown()
    out = {"lengths": np.array(lengths), "losses": np.array(losses), "steps": np.int64(steps), "lr": np.float64(1e-3),
           "grad_clip": np.float64(1.0), "w0": _np(w0), "b0": _np(b0), "w1": _np(head.linear.weight), "b1": _np(head.linear.bias)}
    for i, it in enumerate(dset.items):
        out[f"item{i}"] = _np(it)
    return out


def make_scoring(ref):
    """sampler.scoring / matching of the unmodified reference (LSTM head + L1 objective, --from_rawfeature)."""
    g = torch.Generator().manual_seed(777)
    pre = OnlinePreprocessor(sample_rate=16000, win_ms=25, hop_ms=10, n_freq=201, n_mels=40, n_mfcc=13)
    c = pre.get_feat_config
    pre.feat_list = [c("linear", 0, log=True), c("linear", 0, log=True), c("linear", 0), c("phase", 0), c("linear", 1), c("phase", 1)]
    pre.channel_inp, pre.channel_tar = 0, 1
    lengths = torch.LongTensor([3200, 2400, 3200, 1777, 2999])
    T = int(lengths.max())
    wavs = torch.zeros(len(lengths), 3, T)
    for b, n in enumerate(lengths.tolist()):
        clean = speechlike(n, g) * 0.05
        noise = torch.randn(n, generator=g) * 0.02
        wavs[b, 0, :n], wavs[b, 1, :n], wavs[b, 2, :n] = clean + noise, clean, noise
    torch.manual_seed(1337)
    head = ref["model"].LSTM(input_size=201, output_size=201, hidden_size=24, num_layers=2, bidirectional=False)
    crit = ref["objective"].L1()
    args = Namespace(from_waveform=False, from_rawfeature=True, active_layerid=None)
    ascending = torch.arange(ref["sampler"].MAX_POSITIONS_LEN)
    per_utt = ref["sampler"].scoring(args, {}, pre, head, crit, ascending, lengths, wavs)
    mean = ref["sampler"].scoring(args, {}, pre, head, crit, ascending, lengths, wavs, mean=True)
    match = ref["sampler"].matching(per_utt[:2], per_utt[2:])
    out = {"lengths": _np(lengths), "wavs": _np(wavs), "per_utt": _np(per_utt), "mean": _np(mean), "match": _np(match),
           "param_names": np.array([n for n, _ in head.named_parameters()])}
    for name, p in head.named_parameters():
        out[f"param_{name}"] = _np(p)
    return out


def make_preprocessor_oracle():
    g = torch.Generator().manual_seed(2024)
    out = {}
    for tag, (n_freq, win_ms, hop_ms, T) in {"n512": (257, 32, 16, 2500), "n400": (201, 25, 10, 1777),
                                              "n1024": (513, 64, 16, 3100)}.items():
        pre = OnlinePreprocessor(win_ms=win_ms, hop_ms=hop_ms, n_freq=n_freq)
        wavs = torch.stack([speechlike(T, g) * 0.05, torch.randn(T, generator=g) * 0.05], 0)[None]   # (1,2,T)
        feats = pre(wavs, [pre.get_feat_config("linear", 0), pre.get_feat_config("phase", 0),
                           pre.get_feat_config("linear", 1, log=True),
                           pre.get_feat_config("mel", 0, log=True, delta=2),
                           pre.get_feat_config("mel", 1, log=True, delta=1, cmvn=True)])
        wav = pre.istft(feats[0], feats[1])
        out[f"{tag}_wavs"] = _np(wavs)
        for name, f in zip(("linear0", "phase0", "loglinear1", "mel_d2", "mel_d1_cmvn"), feats):
            out[f"{tag}_{name}"] = _np(f)
        out[f"{tag}_istft"] = _np(wav)
    return out


def main():
    import sys
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(1)
    ref = ref_loader.load()
    if "train" in sys.argv[1:]:
        np.savez_compressed(os.path.join(GOLDEN_DIR, "runner_train_ref.npz"), **make_runner_train(ref))
        return
    if "scoring" in sys.argv[1:]:
        np.savez_compressed(os.path.join(GOLDEN_DIR, "scoring_ref.npz"), **make_scoring(ref))
        return
    np.savez_compressed(os.path.join(GOLDEN_DIR, "recurrent_heads_ref.npz"), **make_recurrent_heads(ref))
    if "recurrent" in sys.argv[1:]:
        return
    np.savez_compressed(os.path.join(GOLDEN_DIR, "scoring_ref.npz"), **make_scoring(ref))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "runner_train_ref.npz"), **make_runner_train(ref))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "signal_path_ref.npz"), **make_signal_path(ref))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "runner_evaluate_ref.npz"), **make_runner_evaluate(ref))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "preprocessor_oracle.npz"), **make_preprocessor_oracle())
    for f in sorted(os.listdir(GOLDEN_DIR)):
        print(f, os.path.getsize(os.path.join(GOLDEN_DIR, f)))


if __name__ == "__main__":
    main()
