"""One active-sampling scoring pass at the configs[4] shape (12 + 32 utterances of 3-10 s, n_fft 400 / hop 160), eager, so that
`ncu --metrics gpu__time_duration.sum` lists its kernels:  python tools/one_scoring.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, speech_enhancement_by_s3prl_b200 as se
from speech_enhancement_by_s3prl_b200 import sampler_ops, synth
dev = torch.device("cuda", 0)
pre = se.OnlinePreprocessor(sample_rate=16000, win_ms=25, hop_ms=10, n_freq=201).to(dev)
pre.channel_inp, pre.channel_tar = 0, 1
torch.manual_seed(1337)
head = se.LinearResidual(input_size=201, output_size=201, precision=1).to(dev)
crit = se.SISDR()
lengths, wavs = synth.batch(44, 10.0, first_index=900000, min_seconds=3.0)
lengths, wavs = lengths.to(dev), wavs.to(dev)
for _ in range(3):
    grads = sampler_ops.scoring(pre, head, crit, lengths, wavs)
    scores = sampler_ops.matching(grads[12:], grads[:12])
torch.cuda.synchronize()
print("embeddings", tuple(grads.shape), "scores", tuple(scores.shape))
