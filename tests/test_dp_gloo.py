"""CPU, world_size 2 over gloo: the data-parallel reductions give the single-process numbers."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world_size))
    from speech_enhancement_by_s3prl_b200 import dp
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    g = torch.Generator().manual_seed(0)
    loss = torch.randn(7, generator=g)
    metric = torch.randn(7, generator=g) * 10
    lo, hi = dp.shard_bounds(7, rank, world_size)
    ml, mm, n = dp.global_means(loss[lo:hi], metric[lo:hi])
    # L1: numerator and count must be reduced separately (unequal shard sizes)
    num = torch.tensor([float(loss[lo:hi].abs().sum()), float(hi - lo) * 3], dtype=torch.float64)
    l1 = dp.global_l1(num)
    # gradient averaging
    lin = torch.nn.Linear(3, 2)
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    dp.allreduce_gradients(lin.parameters())
    # running sums of an evaluation pass as se_finalize_metrics_acc keeps them: [sum loss, sum metric, utterances]
    acc = torch.tensor([float(loss[lo:hi].double().sum()), float(metric[lo:hi].double().sum()), float(hi - lo)], dtype=torch.float64)
    al, am, an = dp.means_from_acc(acc)
    assert acc[2].item() == hi - lo                         # the accumulator itself is left untouched
    ret[rank] = (ml.item(), mm.item(), n, l1.item(), lin.weight.grad[0, 0].item(), al.item(), am.item(), an.item())
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_reductions_match_single_process():
    world_size = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world_size, port, ret), nprocs=world_size, join=True)
    g = torch.Generator().manual_seed(0)
    loss = torch.randn(7, generator=g)
    metric = torch.randn(7, generator=g) * 10
    for r in range(world_size):
        ml, mm, n, l1, g00, al, am, an = ret[r]
        assert an == 7 and al == pytest.approx(ml, abs=1e-12) and am == pytest.approx(mm, abs=1e-12)
        assert n == 7
        assert ml == pytest.approx(loss.double().mean().item(), abs=1e-12)
        assert mm == pytest.approx(metric.double().mean().item(), abs=1e-12)
        assert l1 == pytest.approx(loss.abs().sum().item() / 21.0, rel=1e-6)
        assert g00 == pytest.approx(1.5)
