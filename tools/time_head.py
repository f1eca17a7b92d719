"""Time the mask-head kernels alone (forward variants, backward SIMT vs tensor core) on the bench shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from speech_enhancement_by_s3prl_b200 import ops
B, F, D = 64, 251, 257
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
feats = (torch.randn(B, F, D, generator=g) * 2 - 3).to(dev)
offset = torch.rand(B, F, D, generator=g).to(dev)
go = torch.randn(B, F, D, generator=g).to(dev)
W = (torch.randn(D, D, generator=g) * 0.05).to(dev)
b = torch.zeros(D, device=dev)
mean, std = ops.cmvn_stats(feats)


def timeit(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for prec in (0, 1):
    t = timeit(lambda: torch.ops.se_b200.linear_head(feats, mean, std, 1e-6, W, b, 2, prec))
    print(f"forward  precision {prec}: {t:8.1f} us")
    t = timeit(lambda: torch.ops.se_b200.linear_head_bwd(feats, mean, std, 1e-6, W, offset, go, 2, prec))
    print(f"backward precision {prec}: {t:8.1f} us")
