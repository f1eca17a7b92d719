"""Time the split-K tcgen05 weight-gradient kernel of the head alone (se_linear_head_bwd_fused: transposing producers +
reduce) over shapes:  python tools/time_head_bwd.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_enhancement_by_s3prl_b200 import ops
dev = torch.device("cuda", 0)


def timeit(fn, n=20):
    g = torch.cuda.CUDAGraph()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(4):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (4 * n) * 1e3


for B, F, D in [(64, 251, 257), (48, 1001, 201), (64, 401, 201), (128, 401, 201), (16, 3751, 257)]:
    gen = torch.Generator().manual_seed(0)
    LD = ops.round4(D)
    feats = torch.zeros(B, F, LD)
    feats[..., :D] = torch.randn(B, F, D, generator=gen) * 2 - 3
    feats = feats.to(dev)
    offset = torch.rand(B, F, LD, device=dev)
    go = torch.randn(B, F, LD, device=dev) * 1e-3
    sums = ops.feature_sums(feats, D)
    t = timeit(lambda: ops.linear_head_bwd_fused(feats, D, sums, 1e-6, offset, go, D, "Sigmoid"))
    gb = B * F * D * 12 / 1e9
    print(f"B {B:4d} F {F:5d} D {D}: {t:8.1f} us   {gb / (t * 1e-6):7.0f} GB/s algorithmic (x, offset, grad read once)", flush=True)

# the same with the SISDR objective's backward folded in (no grad_offset tensor) against the two-kernel form
for B, F, D in [(48, 1001, 201), (64, 251, 257)]:
    gen = torch.Generator().manual_seed(0)
    LD = ops.round4(D)
    feats = torch.zeros(B, F, LD); feats[..., :D] = torch.randn(B, F, D, generator=gen) * 2 - 3
    feats = feats.to(dev)
    offset = torch.rand(B, F, LD, device=dev)
    inp, tar = torch.randn(B, F, LD, device=dev) ** 2, torch.randn(B, F, LD, device=dev) ** 2
    frames = torch.full((B,), F, dtype=torch.int64, device=dev)
    sums = ops.feature_sums(feats, D)
    _, _, g_off, sums3 = ops.sisdr_mask_step(offset, inp, tar, frames, 0, D)
    t2 = timeit(lambda: (ops.sisdr_mask_step(offset, inp, tar, frames, 0, D), ops.linear_head_bwd_fused(feats, D, sums, 1e-6, offset, g_off, D, "Sigmoid")))
    t1 = timeit(lambda: (ops.sisdr_mask_step(offset, inp, tar, frames, 0, D, want_grad=False),
                         ops.linear_head_bwd_sisdr(feats, D, sums, 1e-6, offset, inp, tar, frames, 0, sums3, D, "Sigmoid")))
    print(f"B {B:4d} F {F:5d} D {D}: objective + weight gradient, separate {t2:7.1f} us, folded {t1:7.1f} us", flush=True)
