"""world_size-2 gloo test (CPU) of the data-parallel plumbing in gptst_b200/dp.py: sharded batch, one flat
gradient all-reduce from the end-of-backward callback, None-gradients preserved, 1-vs-2 rank gradient equality."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.Tanh(), torch.nn.Linear(16, 3))
    net.unused = torch.nn.Parameter(torch.ones(5))       # never receives a gradient (like decoder.time_feature1_)
    return net


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from gptst_b200 import dp
    r, _, w = dp.init_from_env("gloo")
    assert (r, w) == (rank, world)
    net = _net()
    if rank == 1:
        with torch.no_grad():
            for p in net.parameters():
                p.add_(1.0)                               # diverge, then re-sync from rank 0
    dp.broadcast_parameters(net)
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
    red = dp.FlatGradAllReduce(net.parameters()).attach()
    for _ in range(2):                                    # two steps: the callback must re-arm itself
        net.zero_grad(set_to_none=True)
        xs, ys = dp.shard_batch(x, rank, world), dp.shard_batch(y, rank, world)
        loss = (net(xs) - ys).abs().mean()
        loss.backward()
    assert red.calls == 2 and net.unused.grad is None
    assert red.last_numel == sum(p.numel() for p in net.parameters()) - 5
    if rank == 0:
        torch.save([p.grad for p in net.parameters() if p.grad is not None], out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_equal_single_process(tmp_path):
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    net = _net()
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
    (net(x) - y).abs().mean().backward()
    want = [p.grad for p in net.parameters() if p.grad is not None]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert torch.allclose(a, b, atol=1e-6), (a - b).abs().max()


def test_single_process_is_a_no_op():
    from gptst_b200 import dp
    net = _net()
    red = dp.FlatGradAllReduce(net.parameters()).attach()
    (net(torch.ones(2, 6))).sum().backward()
    assert red.calls == 0 and red.world == 1
    assert dp.shard_batch(torch.arange(6), 1, 3).tolist() == [1, 4]


def _worker_bucketed(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from gptst_b200 import dp
    dp.init_from_env("gloo")
    net = _net()
    dp.broadcast_parameters(net)
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
    # bucket 0 = last layer (its gradients complete first in backward) + the never-used parameter, bucket 1 = first layer
    red = dp.BucketedGradAllReduce([list(net[2].parameters()) + [net.unused], list(net[0].parameters())]).arm()
    for _ in range(2):
        net.zero_grad(set_to_none=True)
        xs, ys = dp.shard_batch(x, rank, world), dp.shard_batch(y, rank, world)
        (net(xs) - ys).abs().mean().backward()
        local = {id(p): p.grad.clone() for p in net.parameters() if p.grad is not None}
        views = red.reduce()
    assert red.calls == 2 and net.unused.grad is None and id(net.unused) not in views
    assert red.last_numel == sum(p.numel() for p in net.parameters()) - 5
    assert abs(red.scale - 0.5) < 1e-12
    for p in net.parameters():                                   # .grad itself is left untouched (rank-local)
        if p.grad is not None:
            assert torch.equal(p.grad, local[id(p)])
    if rank == 0:
        torch.save([views[id(p)] * red.scale for p in net.parameters() if p.grad is not None], out)
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_reducer_matches_single_process(tmp_path):
    """BucketedGradAllReduce: rank-summed flat views x 1/world == single-process gradients of the global batch."""
    out = str(tmp_path / "gb.pt")
    mp.spawn(_worker_bucketed, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    net = _net()
    x = torch.randn(8, 6, generator=torch.Generator().manual_seed(1))
    y = torch.randn(8, 3, generator=torch.Generator().manual_seed(2))
    (net(x) - y).abs().mean().backward()
    want = [p.grad for p in net.parameters() if p.grad is not None]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.allclose(a, b, atol=1e-6), (a - b).abs().max()
