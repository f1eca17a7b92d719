"""One fused training step at the configs[2] shape (48 utterances of 3-10 s, n_fft 400 / hop 160, LinearResidual(201) + SISDR +
ClipAdam), eager (no graph) so that `ncu --metrics gpu__time_duration.sum` lists its kernels:  python tools/one_train_step.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, speech_enhancement_by_s3prl_b200 as se
from speech_enhancement_by_s3prl_b200 import synth
dev = torch.device("cuda", 0)
pre = se.OnlinePreprocessor(sample_rate=16000, win_ms=25, hop_ms=10, n_freq=201).to(dev)
pre.channel_inp, pre.channel_tar = 0, 1
torch.manual_seed(1337)
head = se.LinearResidual(input_size=201, output_size=201, precision=1).to(dev)
eng = se.EnhancementEngine(pre, head, precision=1)
opt = se.ClipAdam(head.parameters(), lr=1e-4)
crit = se.SISDR()
lengths, wavs = synth.batch(48, 10.0, first_index=700000, min_seconds=3.0)
lengths, wavs = lengths.to(dev), wavs.to(dev)
print("fused route:", eng.fused_training_supported(crit, 48, wavs.shape[2]), flush=True)
for _ in range(3):
    eng.train_step(lengths, wavs, crit, optimizer=opt, grad_clip=1.0)
torch.cuda.synchronize()
