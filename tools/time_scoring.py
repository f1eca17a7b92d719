"""Active-sampling scoring (sampler.py:59-120): one-pass per-utterance gradient embeddings vs the reference's loop of
backward calls, on BASELINE.json configs[4]'s shapes (12 training + 32 query utterances of up to 10 s, n_fft 400 / hop 160)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_enhancement_by_s3prl_b200 as se
from speech_enhancement_by_s3prl_b200 import sampler_ops, synth
dev = torch.device("cuda", 0)
for nfreq, win, hop_ms in [(201, 25, 10), (257, 32, 16)]:
    pre = se.OnlinePreprocessor(sample_rate=16000, win_ms=win, hop_ms=hop_ms, n_freq=nfreq).to(dev)
    pre.channel_inp, pre.channel_tar = 0, 1
    torch.manual_seed(1337)
    head = se.LinearResidual(input_size=nfreq, output_size=nfreq, precision=1).to(dev)
    crit = se.SISDR()
    lengths, wavs = synth.batch(44, 10.0)
    lengths, wavs = lengths.to(dev), wavs.to(dev)
    c = pre.get_feat_config
    feats, lin_i, lin_t = pre(wavs, [c("linear", 0, log=True), c("linear", 0), c("linear", 1)])
    frames = lengths // pre._win_args["hop_length"] + 1

    def timeit(fn, n):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            out = fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3, out

    ms_fast, g_fast = timeit(lambda: sampler_ops.scoring_batched(head, crit, feats, lin_i, lin_t, frames), 20)
    ms_loop, g_loop = timeit(lambda: sampler_ops.scoring_loop(head, crit, feats, lin_i, lin_t, frames), 3)
    ms_match, s = timeit(lambda: se.matching(g_fast[12:], g_fast[:12]), 50)
    cos = torch.nn.functional.cosine_similarity(g_fast, g_loop, dim=1).min().item()
    print(f"n_freq {nfreq}: 44 x 10 s, {g_fast.shape[1]} parameters: one pass {ms_fast:.2f} ms, per-utterance loop {ms_loop:.1f} ms "
          f"({ms_loop / ms_fast:.0f}x), matching 32 x 12 {ms_match * 1e3:.0f} us, min cosine(fast, loop) {cos:.6f}", flush=True)
