#!/usr/bin/env python
"""CTA timeline of one fused evaluation step (debugging aid; uses se_set_trace).

  python tools/trace_step.py [--utt 64] [--seconds 4] [--steps 3]

Prints, per kernel of the last traced step: first/last CTA start, first/last CTA end (us, relative to the
step's first CTA start), mean CTA duration, and the fraction of SM-time the kernel's CTAs cover.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utt", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--eager", action="store_true")
    ap.add_argument("--verbose", "-v", action="store_true")
    args = ap.parse_args()
    import speech_enhancement_by_s3prl_b200 as se
    from speech_enhancement_by_s3prl_b200 import _lib, synth
    dev = torch.device("cuda", 0)
    pre = se.OnlinePreprocessor(sample_rate=16000, win_ms=32, hop_ms=16, n_freq=257).to(dev)
    pre.channel_inp, pre.channel_tar = 0, 1
    torch.manual_seed(1337)
    head = se.LinearResidual(input_size=257, output_size=257).to(dev)
    eng = se.EnhancementEngine(pre, head, log_features=True, precision=1)
    lengths, wavs = synth.batch(args.utt, args.seconds)
    lengths, wavs = lengths.to(dev), wavs.to(dev)
    cap = 1 << 16
    buf = torch.zeros(1 + 4 * cap, dtype=torch.int64, device=dev)
    lib = _lib.load()
    lib.se_set_trace(buf.data_ptr())
    if args.eager:
        run = lambda: eng.eval_step(lengths, wavs)
        run()
    else:
        g = eng.capture_bound(lengths, wavs)
        run = g["graph"].replay
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    buf.zero_()
    torch.cuda.synchronize()
    for _ in range(args.steps):
        run()
    torch.cuda.synchronize()
    lib.se_set_trace(None)
    n = int(buf[0].item())
    rec = buf[1:1 + 4 * n].view(n, 4).cpu().numpy()
    kid = rec[:, 0] >> 32
    t0, t1, sm = rec[:, 2].astype(np.float64), rec[:, 3].astype(np.float64), rec[:, 1]
    names = {1: "K1 stft+stats", 2: "K2 head", 3: "K3 mask_istft", 4: "K3' finalize"}
    # split into steps: records of kernel 1 cluster in time; use the gaps between K1 start times
    order = np.argsort(t0)
    k1 = np.sort(t0[kid == 1])
    per = len(k1) // args.steps
    step_starts = [k1[i * per] for i in range(args.steps)]
    last = step_starts[-1]
    sel = t0 >= last
    base = t0[sel].min()
    print(f"records {n}; last step: {sel.sum()} CTAs; step period {(step_starts[-1] - step_starts[0]) / max(1, args.steps - 1) / 1e3:.1f} us")
    n_sm = len(np.unique(sm))
    for k in (1, 2, 3, 4):
        m = sel & (kid == k)
        if not m.any():
            continue
        a, b = (t0[m] - base) / 1e3, (t1[m] - base) / 1e3
        dur = b - a
        span = b.max() - a.min()
        print(f"{names[k]:16s} CTAs {m.sum():4d}  start {a.min():7.2f}..{a.max():7.2f}  end {b.min():7.2f}..{b.max():7.2f}  "
              f"dur mean {dur.mean():6.2f} min {dur.min():6.2f} max {dur.max():6.2f}  span {span:6.2f}  SMs {len(np.unique(sm[m]))}/{n_sm}")
        if args.verbose:
            blk = (rec[:, 0] & 0xffffffff)[m]
            o = np.argsort(blk)
            pct = np.percentile(b, [5, 25, 50, 75, 95])
            print("    end percentiles 5/25/50/75/95:", " ".join(f"{x:6.2f}" for x in pct))
            print("    end by block (every 16th):", " ".join(f"{x:5.1f}" for x in b[o][::16]))
            print("    start by block (every 16th):", " ".join(f"{x:5.1f}" for x in a[o][::16]))
    marks = {17: "K1: prologue done", 18: "K1: thread 0 frames done", 19: "K1: all frames done", 11: "head: prologue done", 12: "head: stats+bias ready", 13: "head: first stage landed", 14: "head: A normalised",
             15: "head: accumulator ready", 16: "head: tile staged", 41: "fin: dependency released", 42: "fin: gain ready"}
    for k, name in marks.items():
        m = sel & (kid == k)
        if m.any():
            b = (t1[m] - base) / 1e3
            print(f"    {name:28s} min {b.min():7.2f}  median {np.median(b):7.2f}  max {b.max():7.2f}")
    tot = (t1[sel].max() - base) / 1e3
    print(f"step span {tot:.2f} us")


if __name__ == "__main__":
    main()
