"""Time the fused TMA / tcgen05 head alone (se_linear_head_fused) over shapes and launch shapes:
python tools/time_head_fused.py"""
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


for B, F, D in [(64, 251, 257), (256, 251, 257), (64, 401, 201), (256, 401, 201), (64, 251, 513), (256, 251, 513), (16, 3751, 513), (128, 3751, 513), (128, 3751, 257)]:
    gen = torch.Generator().manual_seed(0)
    LD = ops.round4(D)
    feats = torch.zeros(B, F, LD)
    feats[..., :D] = torch.randn(B, F, D, generator=gen) * 2 - 3
    feats = feats.to(dev)
    W = ops.round_tf32(ops.pad_weight((torch.randn(D, D, generator=gen) * 0.05).to(dev)))
    b = torch.zeros(D, device=dev)
    sums = ops.feature_sums(feats, D)
    t = timeit(lambda: ops.linear_head_tma(feats, D, W, b, "Sigmoid", sums, 1e-6))
    gb = B * F * D * 8 / 1e9
    print(f"B {B:4d} F {F:5d} D {D}: {t:8.1f} us   {gb / (t * 1e-6):7.0f} GB/s algorithmic   {2 * B * F * D * D / (t * 1e-6) / 1e12:6.1f} TFLOP/s", flush=True)
