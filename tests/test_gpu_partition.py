"""GPU test of the destination-row partitioned layer: the P ranks' stages are run one after the other on ONE GPU with
the collectives emulated by concatenation / summation (no kernels wait on one another), and the assembled forward
output and gradients must match the single-GPU layer at 1e-5.  The real NCCL path is exercised by
tests/test_gpu_partition.py::test_partitioned_layer_nccl when >= 2 GPUs are visible."""
import os

import pytest
import torch

from util import nerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = [  # N, E, F, C, H, concat, world
    (1000, 12000, 64, 128, 4, True, 2),
    (999, 9000, 40, 64, 2, True, 4),
    (700, 8000, 96, 47, 4, False, 3),
    (640, 6000, 33, 5, 3, True, 2),
    (512, 5000, 50, 7, 1, False, 8),
    (3000, 40000, 32, 64, 2, True, 2),      # e // 5 = 8000 in-edges of node 17: a "giant" row (cut into segments)
]


def _single_gpu(layer, x, ei, gout):
    xg = x.clone().requires_grad_(True)
    out = layer(xg, ei)
    out.backward(gout)
    grads = [p.grad.clone() for p in layer.parameters()]
    layer.zero_grad()
    return out.detach(), xg.grad, grads


@pytest.mark.parametrize("dropout", [0.0, 0.6], ids=["nodrop", "philox"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_c%d_h%d_%s_p%d" % (c[0], c[3], c[4], "cat" if c[5] else "mean", c[6]))
def test_partitioned_stages_match_single_gpu(case, dropout):
    import GAT
    from atmlgraphattentionnetworks_b200 import partition as pt
    from atmlgraphattentionnetworks_b200.gat import dropout_mask_tensor
    n, e, f, c, h, concat, world = case
    torch.manual_seed(n + c)
    layer = GAT.GraphAttentionLayer(f, c, num_heads=h, concat=concat, dropout=0.0).to(DEV)
    drop = None
    if dropout:
        # in-kernel dropout in partitioned mode: every rank's stages get the same (p, seed); the single-GPU layer is fed the
        # materialised mask of that seed (global ORIGINAL edge positions key the multipliers on every rank)
        seed = torch.tensor([n * 7919 + c, e], dtype=torch.int64, device=DEV)
        drop = (dropout, seed)
        full_mask = dropout_mask_tensor(dropout, seed, e + n, h)
        layer.mask_hook = lambda shape: full_mask
    with torch.no_grad():
        layer.bias.uniform_(-0.5, 0.5)
    x = torch.randn(n, f, device=DEV)
    ei = torch.randint(0, n, (2, e), device=DEV)
    ei[1, : e // 5] = 17                                    # hub destination (degree > HUB_DEGREE)
    ei[0, e // 5: e // 5 + max(700, e // 8)] = 23           # hub source (CSC side); giant in the last case
    gout = torch.randn(n, h * c if concat else c, device=DEV)
    out_ref, gx_ref, grads_ref = _single_gpu(layer, x, ei, gout)

    geom = (f, c, h, concat)
    heads_mode = (not concat) and h > 1
    with torch.no_grad():
        w, bw, a1, a2, b1, b2 = (t.contiguous() for t in layer._packed())
        bias = layer.bias.detach()
        parts = [pt.build_row_partition(ei, n, world, r) for r in range(world)]
        blk = parts[0].block
        xs = [x[p.lo:p.hi].contiguous() for p in parts]
        proj = [pt.stage_proj(geom, (w, bw, a1, a2, b1, b2), xs[r], blk) for r in range(world)]
        wh_full = torch.cat([p[0] for p in proj])                       # == all_gather_into_tensor
        s_src_full = torch.cat([p[1] for p in proj])
        fwd = [pt.stage_edge_fwd(geom, parts[r], wh_full, s_src_full, proj[r][2], bias, drop) for r in range(world)]
        out = torch.cat([f_[0] for f_ in fwd])
        assert nerr(out.cpu().numpy(), out_ref.cpu().numpy()) <= 1e-5
        prep = [pt.stage_prep(geom, gout[p.lo:p.hi].contiguous(), fwd[r][3] if heads_mode else fwd[r][0], bias, proj[r][2],
                              fwd[r][1], fwd[r][2], blk) for r, p in enumerate(parts)]
        rowrec_full = torch.cat([p[0] for p in prep])
        g_full = torch.cat([p[1] for p in prep])
        csc = [pt.stage_csc(geom, parts[r], proj[r][0][:parts[r].n_own], proj[r][1][:parts[r].n_own], rowrec_full, g_full, drop)
               for r in range(world)]
        g_s_dst = sum(c_[2] for c_ in csc)                              # == reduce_scatter (sum) + own slice below
        fin = [pt.stage_finish(geom, proj[r][0][:parts[r].n_own], a1, a2, csc[r][1],
                               g_s_dst[r * blk: r * blk + parts[r].n_own].contiguous(), csc[r][0]) for r in range(world)]
        pb = [pt.stage_proj_bwd(geom, csc[r][0], xs[r], w, True) for r in range(world)]
        g_x = torch.cat([p[0] for p in pb])
        g_w = sum(p[1] for p in pb)
        g_bw, g_a1, g_a2, g_b1, g_b2 = (sum(f_[k] for f_ in fin) for k in range(5))
        g_bias = sum(p[2] for p in prep)
    assert nerr(g_x.cpu().numpy(), gx_ref.cpu().numpy()) <= 1e-5
    # reference gradients in packed form through autograd of the packing
    pk = layer._packed()
    packed_ref = torch.autograd.grad(pk, list(layer.parameters()), [g_w, g_bw, g_a1, g_a2, g_b1, g_b2], allow_unused=True)
    named = dict(zip([k for k, _ in layer.named_parameters()], zip(packed_ref, grads_ref)))
    scale_a1 = max(float(g.abs().max()) for k, (_, g) in named.items() if "attentions1" in k and k.endswith("weight"))
    for k, (got, want) in named.items():
        if k == "bias":
            assert nerr(g_bias.cpu().numpy(), want.cpu().numpy()) <= 1e-5
            continue
        if "attentions" in k:      # sums of dz cancel inside rows (see test_gpu_parity): judged on the g_a1 scale
            assert float((got - want).abs().max()) <= 2e-5 * max(scale_a1, float(want.abs().max())), k
        else:
            assert nerr(got.cpu().numpy(), want.cpu().numpy()) <= 1e-5, k


def _nccl_worker(rank, world, port, results, second_forward=False, replicated_input=False):
    import torch.distributed as dist
    import GAT
    from atmlgraphattentionnetworks_b200 import partition as pt
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    if second_forward:
        os.environ["B200GAT_PEER_PUSH"] = "1"           # the fused projection + peer-memory all-gather of Wh
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n, e, f, c, h = 2000, 30000, 64, 128, 4
    torch.manual_seed(0)
    layer = GAT.GraphAttentionLayer(f, c, num_heads=h, concat=True, dropout=0.0).to(dev)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(n, f, generator=gen).to(dev)
    ei = torch.randint(0, n, (2, e), generator=gen).to(dev)
    gout = torch.randn(n, h * c, generator=gen).to(dev)
    part = pt.build_row_partition(ei, n, world, rank)
    xo = x[part.lo:part.hi].clone().requires_grad_(True)
    # replicated_input: every rank holds the layer's input for all nodes and projects it itself (no forward exchange)
    out = pt.partitioned_layer_forward(layer, xo, part, x_full=x if replicated_input else None)
    if second_forward:
        # a SECOND forward through the same layer before the first one's backward (an eval forward, a second micro-batch,
        # activation checkpointing): it overwrites the layer's persistent symmetric Wh buffer, which the first forward's
        # backward must not depend on
        with torch.no_grad():
            pt.partitioned_layer_forward(layer, torch.randn_like(xo), part)
        results["peer"] = getattr(layer, "_peer_buf", (None, None))[1] is not None
    out.backward(gout[part.lo:part.hi])
    flat = torch.cat([p.grad.flatten() for p in layer.parameters()])
    dist.all_reduce(flat)
    if rank == 0:
        layer.zero_grad()
        xr = x.clone().requires_grad_(True)
        ref = layer(xr, ei)
        ref.backward(gout)
        flat_ref = torch.cat([p.grad.flatten() for p in layer.parameters()])
        results["out"] = nerr(out.detach().cpu().numpy(), ref[part.lo:part.hi].detach().cpu().numpy())
        results["gx"] = nerr(xo.grad.cpu().numpy(), xr.grad[part.lo:part.hi].cpu().numpy())
        results["gp"] = float((flat - flat_ref).abs().max() / flat_ref.abs().max())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_partitioned_layer_nccl():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, port, results), nprocs=2, join=True)
    assert results["out"] <= 1e-5 and results["gx"] <= 1e-5 and results["gp"] <= 2e-5, dict(results)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_partitioned_layer_peer_push_second_forward_before_backward():
    """forward A, forward B, backward A in peer-push mode gives A's gradients (the saved Wh is a private copy, not a view
    of the symmetric buffer the second forward overwrites)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, port, results, True), nprocs=2, join=True)
    assert results["out"] <= 1e-5 and results["gx"] <= 1e-5 and results["gp"] <= 2e-5, dict(results)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_partitioned_layer_nccl_replicated_input():
    """layer 1 of a static graph: the input features are replicated, every rank projects all rows, nothing is exchanged
    in the forward; results equal the single-GPU layer"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_nccl_worker, args=(2, port, results, False, True), nprocs=2, join=True)
    assert results["out"] <= 1e-5 and results["gx"] <= 1e-5 and results["gp"] <= 2e-5, dict(results)


def _nccl_stack_worker(rank, world, port, results, bf16=False):
    """3-layer stack row-partitioned over 2 GPUs (fused ELU boundaries, kept operand splits, replicated input for layer 1,
    in-kernel dropout with rank 0's seed) against the single-GPU GATStack on rank 0 fed the SAME dropout masks."""
    import torch.distributed as dist
    from atmlgraphattentionnetworks_b200 import partition as pt
    from atmlgraphattentionnetworks_b200.gat import dropout_mask_tensor
    from atmlgraphattentionnetworks_b200.gatnet import GATStack
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n, e, f = 3001, 40000, 64
    spec = [(f, 64, 4, True), (256, 32, 4, True), (128, 7, 2, False)]
    torch.manual_seed(0)
    model = GATStack(spec, dropout=0.0 if bf16 else 0.5).to(dev).train()
    if bf16:      # gathered rows — and the all-gathered copies on the wire — stored as bf16 (tolerance 1e-2, not parity)
        from atmlgraphattentionnetworks_b200.gat import set_gather_dtype
        set_gather_dtype(model, torch.bfloat16)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(n, f, generator=gen).to(dev)
    ei = torch.randint(0, n, (2, e), generator=gen).to(dev)
    gout = torch.randn(n, 7, generator=gen).to(dev)
    part = pt.build_row_partition(ei, n, world, rank)
    pmodel = pt.PartitionedGATStack(model)
    xo = x[part.lo:part.hi].clone().requires_grad_(True)
    out = pmodel(xo, part, x_full=x)
    seeds = [] if bf16 else [c._last_dropout_seed.clone() for c in model.convs]
    out.backward(gout[part.lo:part.hi])
    flat = torch.cat([p.grad.flatten() for p in model.parameters()])
    dist.all_reduce(flat)
    if rank == 0:
        model.zero_grad()
        if bf16:
            set_gather_dtype(model, torch.float32)          # the reference run: fp32 rows on one GPU
        for c, s in zip(model.convs, seeds):
            m = dropout_mask_tensor(0.5, s, e + n, c.num_heads)
            c.mask_hook = (lambda mm: (lambda shape: mm))(m)
        xr = x.clone().requires_grad_(True)
        ref = model(xr, ei)
        ref.backward(gout)
        flat_ref = torch.cat([p.grad.flatten() for p in model.parameters()])
        results["out"] = nerr(out.detach().cpu().numpy(), ref[part.lo:part.hi].detach().cpu().numpy())
        results["gp"] = float((flat - flat_ref).abs().max() / flat_ref.abs().max())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_partitioned_stack_nccl_fused_boundaries_and_dropout():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_nccl_stack_worker, args=(2, port, results), nprocs=2, join=True)
    assert results["out"] <= 1e-5 and results["gp"] <= 2e-5, dict(results)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_partitioned_stack_nccl_bf16_rows_on_the_wire():
    """bf16 gather mode, row-partitioned: the all-gathered Wh / gradient rows are bf16 (half the NVLink bytes); results stay
    within the mode's stated 1e-2 of the fp32 single-GPU stack"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_nccl_stack_worker, args=(2, port, results, True), nprocs=2, join=True)
    assert 1e-6 < results["out"] <= 1e-2 and results["gp"] <= 1e-2, dict(results)
