"""Time the training step (runner.py:431-471 on the kernels): preprocessor tensors -> head -> criterion -> backward -> Adam."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_enhancement_by_s3prl_b200 as se
from speech_enhancement_by_s3prl_b200 import synth
dev = torch.device("cuda", 0)
for nfreq, win, hop in [(257, 32, 16), (201, 25, 10)]:
    pre = se.OnlinePreprocessor(sample_rate=16000, win_ms=win, hop_ms=hop, n_freq=nfreq).to(dev)
    pre.channel_inp, pre.channel_tar = 0, 1
    torch.manual_seed(1337)
    for precision in (0, 1):
        torch.manual_seed(1337)
        head = se.LinearResidual(input_size=nfreq, output_size=nfreq, precision=precision).to(dev)
        eng = se.EnhancementEngine(pre, head, log_features=True, precision=precision)
        opt = torch.optim.Adam(head.parameters(), lr=1e-4)
        lengths, wavs = synth.batch(64, 4.0)
        lengths, wavs = lengths.to(dev), wavs.to(dev)
        obj = se.SISDR()
        for _ in range(5):
            loss = eng.train_step(lengths, wavs, obj, optimizer=opt, grad_clip=1.0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            loss = eng.train_step(lengths, wavs, obj, optimizer=opt, grad_clip=1.0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"n_freq {nfreq} hop_ms {hop} precision {precision}: {ms:.3f} ms/step  {256 / ms * 1e3:.0f} audio-s/s  loss {loss.item():.4f}")
        del loss                                   # (keeps the eager autograd graph -- and its default-stream nodes -- alive)
        torch.manual_seed(1337)
        head2 = se.LinearResidual(input_size=nfreq, output_size=nfreq, precision=precision).to(dev)
        eng = se.EnhancementEngine(pre, head2, log_features=True, precision=precision)
        opt2 = torch.optim.Adam(head2.parameters(), lr=1e-4, capturable=True)
        for _ in range(3):
            loss = eng.train_step_graph(lengths, wavs, obj, opt2, grad_clip=1.0)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            loss = eng.train_step_graph(lengths, wavs, obj, opt2, grad_clip=1.0)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"   graph replay: {ms:.3f} ms/step  {256 / ms * 1e3:.0f} audio-s/s  loss {loss.item():.4f}")
        if precision == 1:
            torch.manual_seed(1337)
            head3 = se.LinearResidual(input_size=nfreq, output_size=nfreq, precision=precision).to(dev)
            eng = se.EnhancementEngine(pre, head3, log_features=True, precision=precision)
            opt3 = se.ClipAdam(head3.parameters(), lr=1e-4)
            for _ in range(3):
                loss = eng.train_step_graph(lengths, wavs, obj, opt3, grad_clip=1.0)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                loss = eng.train_step_graph(lengths, wavs, obj, opt3, grad_clip=1.0)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print(f"   graph replay, ClipAdam: {ms:.3f} ms/step  {256 / ms * 1e3:.0f} audio-s/s  loss {loss.item():.4f}")
