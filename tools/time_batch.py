"""Eval-step time against batch size / utterance length (one-wave effects): python tools/time_batch.py"""
import sys, torch
sys.path.insert(0, '/root/repo')
import speech_enhancement_by_s3prl_b200 as se
from speech_enhancement_by_s3prl_b200 import synth
dev = torch.device('cuda', 0)
pre = se.OnlinePreprocessor(sample_rate=16000, win_ms=32, hop_ms=16, n_freq=257).to(dev)
pre.channel_inp, pre.channel_tar = 0, 1
torch.manual_seed(1337)
head = se.LinearResidual(input_size=257, output_size=257).to(dev)
for B, secs in [(64, 4.0), (128, 4.0), (256, 4.0), (512, 4.0), (64, 16.0), (16, 60.0)]:
    eng = se.EnhancementEngine(pre, head, log_features=True, precision=1)
    lengths, wavs = synth.batch(8, secs)
    lengths, wavs = lengths.repeat(B // 8).to(dev), wavs.repeat(B // 8, 1, 1).to(dev)
    g = eng.capture_bound(lengths, wavs)
    for _ in range(5):
        g["graph"].replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        g["graph"].replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 30
    print(f"B {B} x {secs}s: {ms * 1e3:.1f} us/step  {B * secs / ms * 1e3 / 1e6:.2f} M audio-s/s  sisdr {g['sisdr'].mean().item():.4f}", flush=True)
    del g, eng
