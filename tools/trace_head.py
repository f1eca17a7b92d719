#!/usr/bin/env python
"""Per-phase CTA timeline of the fused head alone (debugging aid; uses se_set_trace).

  python tools/trace_head.py [B F D]        default 128 3751 513

Marks recorded by thread 0 of each CTA (head_fused.cu): 11 barriers + tensor memory ready, 12 CMVN scale / shift in
shared memory, 13 first k-block landed, 14 last k-block normalised, 15 accumulators complete, 16 epilogue computed,
finish = stores issued.  Prints the mean time between consecutive marks over all CTAs and the CTAs resident per SM.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_enhancement_by_s3prl_b200 import _lib, ops

B, F, D = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 3751, 513)
dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)
LD = ops.round4(D)
feats = torch.zeros(B, F, LD, device=dev)
feats[..., :D] = torch.randn(B, F, D, device=dev) * 2 - 3
W = ops.round_tf32(ops.pad_weight((torch.randn(D, D, generator=gen) * 0.05).to(dev)))
b = torch.zeros(D, device=dev)
sums = ops.feature_sums(feats, D)
run = lambda: ops.linear_head_tma(feats, D, W, b, "Sigmoid", sums, 1e-6)
for _ in range(3):
    run()
torch.cuda.synchronize()
cap = 1 << 18
buf = torch.zeros(1 + 4 * cap, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.se_set_trace(buf.data_ptr())
run()
torch.cuda.synchronize()
lib.se_set_trace(None)
n = int(buf[0].item())
rec = buf[1:1 + 4 * n].view(n, 4).cpu().numpy()
kid, cta = rec[:, 0] >> 32, rec[:, 0] & 0xffffffff
t0, t1 = rec[:, 2].astype(np.float64), rec[:, 3].astype(np.float64)
base = t0[t0 > 0].min()
n_cta = int(cta.max()) + 1
ids = [11, 12, 13, 14, 15, 16, 2]
tm = {}
for k in ids:
    m = kid == k
    arr = np.full(n_cta, np.nan)
    arr[cta[m]] = t1[m]
    tm[k] = arr
start = np.full(n_cta, np.nan)
m = kid == 2
start[cta[m]] = t0[m]
sm = np.full(n_cta, -1)
sm[cta[m]] = rec[m, 1]
print(f"B {B} F {F} D {D}: {n_cta} CTAs, kernel span {(np.nanmax(tm[2]) - base) / 1e3:.1f} us")
prev, pname = start, "start"
for k in ids:
    d = (tm[k] - prev) / 1e3
    print(f"  {pname:>6s} -> {('finish' if k == 2 else str(k)):>6s}: mean {np.nanmean(d):7.2f} us  p10 {np.nanpercentile(d, 10):7.2f}  p90 {np.nanpercentile(d, 90):7.2f}")
    prev, pname = tm[k], ("finish" if k == 2 else str(k))
dur = (tm[2] - start) / 1e3
print(f"  CTA duration mean {np.nanmean(dur):.2f} us; sum over CTAs / (SMs x span) = "
      f"{np.nansum(dur) / (len(np.unique(sm[sm >= 0])) * (np.nanmax(tm[2]) - base) / 1e3):.2f} CTAs resident per SM")
