"""Host-side logic of the data-parallel gradient all-reduce, exercised with world_size-2 gloo on CPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_opt(params, tail):
    """The real FusedAdamW on CPU parameters: host-side layout logic only (its arithmetic has no CPU path)."""
    from qavit_b200.optim import FusedAdamW
    opt = FusedAdamW([(f"p{i}", p) for i, p in enumerate(params)], lr=1e-3, tail_elems=tail)
    opt.attach_grads()
    return opt


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qavit_b200.dp import GradAllReducer
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.randn(s)) for s in [(7, 5), (3,), (64, 16), (10,), (33,), (8, 8)]]
    bank = torch.nn.Parameter(torch.full((4,), float(rank + 1)))
    opt = _make_opt(params, 4)
    red = GradAllReducer(opt, n_buckets=3, bank_params=[bank])
    assert opt.grad_prescale == 1.0 / world          # buffers hold SUMS; the mean is folded into the clip pass
    assert 1 <= len(red.buckets) <= 3
    covered = sorted(i for lo, hi, _ in red.buckets for i in range(lo, hi))
    assert covered == list(range(len(params)))
    red.reset()
    x = torch.full((1,), float(rank + 1))
    loss = sum((p * x).sum() for p in params)       # d loss / d p = rank + 1 everywhere
    loss.backward()
    red.finish()
    ok = all(torch.allclose(p.grad * opt.grad_prescale, torch.full_like(p, (1 + world) / 2)) for p in params)
    ok = ok and torch.allclose(bank.data, torch.full((4,), (1 + world) / 2))
    # graph-mode flavour: no hooks, ONE all-reduce of the whole flat buffer with the bank state in its tail
    opt2 = _make_opt(params, 4)
    bank.data.fill_(float(rank + 1))
    red2 = GradAllReducer(opt2, n_buckets=3, bank_params=[bank], overlap=False)
    assert not red2._hooks
    for p in params:
        p.grad.fill_(float(rank + 1))
    red2.reduce_flat()
    ok = ok and all(torch.allclose(p.grad * opt2.grad_prescale, torch.full_like(p, (1 + world) / 2)) for p in params)
    ok = ok and torch.allclose(bank.data, torch.full((4,), (1 + world) / 2))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    world = 2
    port = 29500 + os.getpid() % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
