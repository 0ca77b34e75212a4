"""CPU tests of the multi-GPU host logic with the gloo backend, world_size 2: the flat gradient
bucket all-reduce (data-parallel over independent graph batches) and the graph / row partition helpers.
The kernels themselves never run here (no CPU fallback); the model below is a stand-in with plain Linear layers."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from atmlgraphattentionnetworks_b200.parallel import GradBucket, all_reduce_packed_grads, row_partition, shard_graphs


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                   # replicated model
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ELU(), torch.nn.Linear(5, 3))
    bucket = GradBucket(model.parameters())
    gen = torch.Generator().manual_seed(100)
    x_all, y_all = torch.randn(8, 6, generator=gen), torch.randn(8, 3, generator=gen)
    mine = shard_graphs(8, world, rank)                    # "graphs" = rows here
    for _ in range(2):                                     # two steps: the views must survive zero()/backward()
        bucket.zero()
        loss = torch.nn.functional.mse_loss(model(x_all[mine]), y_all[mine])
        loss.backward()
        bucket.all_reduce_mean()
    out[rank] = bucket.flat.clone()
    dist.destroy_process_group()


def test_grad_bucket_all_reduce_matches_single_process():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ELU(), torch.nn.Linear(5, 3))
    gen = torch.Generator().manual_seed(100)
    x_all, y_all = torch.randn(8, 6, generator=gen), torch.randn(8, 3, generator=gen)
    torch.nn.functional.mse_loss(model(x_all), y_all).backward()          # equal shards => mean of shard means
    want = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert torch.allclose(out[0], out[1])
    assert torch.allclose(out[0], want, rtol=1e-5, atol=1e-7)


class _PackedGrad(torch.autograd.Function):
    """y = x @ w1.T + x @ w2.T with the two weight gradients returned as views of ONE packed buffer (what the GAT layer's
    backward does for its per-head parameters)."""

    @staticmethod
    def forward(ctx, x, w1, w2):
        ctx.save_for_backward(x)
        return x @ w1.t() + x @ w2.t()

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        packed = torch.empty(2, g.shape[1], x.shape[1])
        packed[0] = g.t() @ x
        packed[1] = g.t() @ x
        return None, packed[0], packed[1]


def _packed_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    w1, w2 = torch.randn(3, 6, requires_grad=True), torch.randn(3, 6, requires_grad=True)
    lin = torch.nn.Linear(3, 2)                              # an ordinary parameter next to the packed ones
    gen = torch.Generator().manual_seed(100)
    x_all, y_all = torch.randn(8, 6, generator=gen), torch.randn(8, 2, generator=gen)
    mine = shard_graphs(8, world, rank)
    params = [w1, w2, *lin.parameters()]
    for _ in range(2):
        for p in params:
            p.grad = None
        torch.nn.functional.mse_loss(lin(_PackedGrad.apply(x_all[mine], w1, w2)), y_all[mine]).backward()
        n_buffers = all_reduce_packed_grads(params)
    assert w1.grad.untyped_storage().data_ptr() == w2.grad.untyped_storage().data_ptr()   # adopted as views, no copy
    out[rank] = (n_buffers, torch.cat([p.grad.flatten() for p in params]).clone())
    dist.destroy_process_group()


def test_packed_grad_all_reduce_matches_single_process():
    world = 2
    port = _free_port()
    out = mp.Manager().dict()
    mp.spawn(_packed_worker, args=(world, port, out), nprocs=world, join=True)
    torch.manual_seed(0)
    w1, w2 = torch.randn(3, 6, requires_grad=True), torch.randn(3, 6, requires_grad=True)
    lin = torch.nn.Linear(3, 2)
    gen = torch.Generator().manual_seed(100)
    x_all, y_all = torch.randn(8, 6, generator=gen), torch.randn(8, 2, generator=gen)
    torch.nn.functional.mse_loss(lin(_PackedGrad.apply(x_all, w1, w2)), y_all).backward()
    want = torch.cat([p.grad.flatten() for p in [w1, w2, *lin.parameters()]])
    assert out[0][0] == 3                                    # one packed base + lin.weight + lin.bias
    assert torch.allclose(out[0][1], out[1][1])
    assert torch.allclose(out[0][1], want, rtol=1e-5, atol=1e-7)


def test_partition_helpers():
    assert shard_graphs(10, 4, 1) == [1, 5, 9]
    assert sum(len(shard_graphs(128, 8, r)) for r in range(8)) == 128
    parts = row_partition(2_400_000, 8)
    assert parts[0] == (0, 300_000) and parts[-1][1] == 2_400_000
    parts = row_partition(10, 4)
    assert parts == [(0, 3), (3, 6), (6, 9), (9, 10)]         # ceil blocks: ONE definition with partition.block_size
    assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    import numpy as np
    from atmlgraphattentionnetworks_b200.partition import build_row_partition
    ei = torch.from_numpy(np.random.default_rng(0).integers(0, 10, size=(2, 30)))
    assert [(p.lo, p.hi) for p in (build_row_partition(ei, 10, 4, r) for r in range(4))] == parts
    assert row_partition(3, 4) == [(0, 1), (1, 2), (2, 3), (3, 3)]


def test_grad_bucket_views_single_process():
    torch.manual_seed(1)
    model = torch.nn.Linear(4, 2)
    bucket = GradBucket(model.parameters())
    model(torch.ones(3, 4)).sum().backward()
    assert bucket.flat.abs().sum() > 0
    assert model.weight.grad.data_ptr() == bucket.flat.data_ptr()        # grads are views into the flat buffer
    bucket.all_reduce_mean()                                             # no process group: no-op
    bucket.zero()
    assert float(model.weight.grad.abs().sum()) == 0.0
