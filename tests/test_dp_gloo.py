"""CPU, world_size 2, gloo: the flat gradient bucket averages exactly the trainable ZiRa parameters and
leaves replicas identical after an optimiser step (the only collective on the path)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ziragroundingdino_b200 as zb
    from ziragroundingdino_b200.dp import FlatGradBucket
    torch.manual_seed(0)                       # identical replicas
    base = torch.nn.Linear(16, 8)
    branch = zb.RepZeroLinear(16, 8)
    for p in base.parameters():
        p.requires_grad_(False)
    branch.train()
    params = list(branch.parameters())
    bucket = FlatGradBucket(params, world)
    assert bucket.flat.numel() == sum(p.numel() for p in params) == 2 * (16 * 8 + 8) + 1
    opt = torch.optim.SGD(params, lr=0.1)
    torch.manual_seed(100 + rank)              # different data per rank
    x = torch.randn(5, 16)
    y, zl = branch.forward_folded(x, base.weight, base.bias)
    (y.square().mean() + 0.1 * zl).backward()
    local = bucket.flat.clone()
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    bucket.all_reduce()
    want = sum(gathered) / world
    assert torch.allclose(bucket.flat, want, atol=1e-7)
    assert all(p.grad.data_ptr() >= bucket.flat.data_ptr() for p in params)   # grads are views of the bucket
    opt.step()
    after = torch.cat([p.detach().reshape(-1) for p in params])
    both = [torch.zeros_like(after) for _ in range(world)]
    dist.all_gather(both, after)
    assert torch.equal(both[0], both[1])       # replicas stay in lock-step
    # async variant
    opt.zero_grad(set_to_none=False)
    y, zl = branch.forward_folded(x, base.weight, base.bias)
    (y.square().mean() + 0.1 * zl).backward()
    work = bucket.all_reduce(async_op=True)
    bucket.finish(work)
    ret[rank] = 1
    dist.destroy_process_group()


def test_flat_bucket_two_ranks_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: 1, 1: 1}


def test_flat_bucket_single_process():
    import ziragroundingdino_b200 as zb
    from ziragroundingdino_b200.dp import FlatGradBucket
    br = zb.RepZeroLinear(4, 4)
    b = FlatGradBucket(br.parameters(), 1)
    assert b.all_reduce() is None and b.nbytes == (2 * (16 + 4) + 1) * 4
    b.zero()


def test_flat_bucket_detects_detached_grads():
    """optimizer.zero_grad(set_to_none=True) or a re-created Parameter detaches a grad from the bucket; the
    all-reduce must refuse instead of averaging stale zeros."""
    import pytest
    from ziragroundingdino_b200.dp import FlatGradBucket
    lin = torch.nn.Linear(4, 3)
    params = list(lin.parameters())
    bucket = FlatGradBucket(params, 1)
    bucket.all_reduce()                                   # attached: fine
    torch.optim.SGD(params, lr=0.1).zero_grad()           # default set_to_none=True
    with pytest.raises(RuntimeError, match="not a view of the bucket"):
        bucket.all_reduce()
    bucket.attach()
    lin(torch.randn(2, 4)).sum().backward()
    assert bucket.flat.abs().sum() > 0
    bucket.zero_grad()
    assert bucket.flat.abs().sum() == 0
    bucket.all_reduce()
    lin.weight.grad = torch.zeros_like(lin.weight)        # a foreign gradient tensor
    with pytest.raises(RuntimeError):
        bucket.check_attached()
