"""Multi-GPU execution of the path, only where it shards.  One process per GPU, torch.distributed
(NCCL over NVLink on the box, gloo in the CPU tests of the host logic).

  * batches of small graphs (PPI-shaped, CIFAR10-superpixel-shaped; run_gnn_benchmark.py:60-66): graphs are
    independent units -> data parallel, the model is replicated and ONE flat NCCL all-reduce of all parameter
    gradients runs per step (`GradBucket`).
  * one large graph: contiguous destination-row blocks (`row_partition`); the exchange step is an all-gather of the
    projected Wh / s_src rows (and of gout in backward).
  * Cora-sized graphs / heads sweep: replicas only.
"""
import torch
import torch.distributed as dist


def shard_graphs(num_graphs, world_size, rank):
    """Graph ids owned by `rank`: r, r+P, r+2P, ... (equal counts when P divides num_graphs, as the per-graph mean
    loss of run_gnn_benchmark.py:64 needs)."""
    return list(range(rank, num_graphs, world_size))


def row_partition(num_nodes, world_size):
    """Contiguous destination-row blocks of the large-graph scheme: [(begin, end)] per rank.  ONE block definition for the
    whole package: blocks of ceil(N / P) rows (partition.block_size), the last ones shorter or empty — exactly the rows
    partition.build_row_partition gives rank r and the layout of the all-gathered [P * block, D] buffers."""
    from .partition import block_size
    b = block_size(num_nodes, world_size)
    return [(min(r * b, num_nodes), min((r + 1) * b, num_nodes)) for r in range(world_size)]


class GradBucket:
    """All parameter gradients of a replicated model in ONE flat buffer: p.grad are views into it, so the per-step
    exchange is a single all-reduce launch (the CIFAR-shaped net is 38 KB of gradients — latency-bound; the
    PPI-shaped stack 7.4 MB) with no pack / unpack copies."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=dt, device=dev)
        self._slots, off = {}, 0
        for p in self.params:
            self._slots[id(p)] = self.flat[off:off + p.numel()].view_as(p)
            p.grad = self._slots[id(p)]
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None, weight=None):
        """Average (or `weight`-ed sum: pass n_r / N for node-level mean losses) of the gradients over ranks."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        for p in self.params:   # autograd may have replaced a .grad view (e.g. first accumulation): re-pack
            if p.grad is not None and p.grad.data_ptr() != self._slot(p).data_ptr():
                self._slot(p).copy_(p.grad)
                p.grad = self._slot(p)
        if weight is None:
            self.flat.div_(dist.get_world_size(group))
        else:
            self.flat.mul_(float(weight))
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)

    def _slot(self, p):
        return self._slots[id(p)]


def packed_grad_buffers(params):
    """The distinct storages the .grad tensors of `params` live in, each as one flat tensor.  The layer's backward hands
    every per-head gradient back as a VIEW of its packed output buffers (gat.py), so with `zero_grad(set_to_none=True)`
    the gradients of a step are views of a handful of base buffers (7 per GAT layer) plus the plain parameters' own."""
    # group by STORAGE (autograd detaches the views it adopts, so ._base is gone): every storage a .grad lives in is one of
    # the backward's dedicated packed buffers (or a plain parameter's own gradient) and is reduced whole
    bases, seen = [], set()
    for p in params:
        g = p.grad
        if g is None:
            continue
        st = g.untyped_storage()
        if st.data_ptr() in seen:
            continue
        seen.add(st.data_ptr())
        bases.append(torch.empty(0, dtype=g.dtype, device=g.device).set_(st))
    return bases


def scale_packed_grads(params, weight):
    """Multiply every gradient buffer of `params` by `weight` with one multi-tensor kernel (the rank's share of the global
    mean: 1 / P for equal shards, n_r / N for node-level mean losses).  -> the buffers."""
    bases = packed_grad_buffers(params)
    if bases:
        torch._foreach_mul_(bases, float(weight))
    return bases


def all_reduce_packed_grads(params, group=None, weight=None):
    """Data-parallel gradient exchange without a staging copy: the packed gradient buffers of the step
    (packed_grad_buffers) are scaled with one multi-tensor kernel and all-reduced inside one NCCL group (one fused
    launch) — no flat bucket, no accumulate-into-bucket kernels.  Average over ranks by default; `weight` = n_r / N for
    node-level mean losses.  Returns the number of buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    bases = scale_packed_grads(params, (1.0 / dist.get_world_size(group)) if weight is None else float(weight))
    if not bases:
        return 0
    coalesce = None
    if dist.get_backend(group) == "nccl":             # gloo (the CPU tests) has no coalescing: plain loop there
        try:
            from torch.distributed.distributed_c10d import _coalescing_manager as coalesce
        except ImportError:
            coalesce = None
    if coalesce is not None:
        with coalesce(group=group, device=bases[0].device, async_ops=False):
            for b in bases:
                dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
    else:
        for b in bases:
            dist.all_reduce(b, op=dist.ReduceOp.SUM, group=group)
    return len(bases)


class ArenaExchange:
    """Data-parallel gradient exchange of a model whose GAT layers write their parameter gradients into ONE persistent arena
    (gat.assign_grad_arena): per step one in-place scale and ONE all-reduce of the arena — no per-step discovery of gradient
    buffers — plus a coalesced all-reduce of the few parameters outside the GAT layers (GATNet's lin1 / lin2).  Falls back to
    all_reduce_packed_grads for a step in which a layer could not use its arena slice (a parameter already held a gradient)."""

    def __init__(self, module):
        from .gat import GraphAttentionLayer, assign_grad_arena
        self.module = module
        self.layers = [m for m in module.modules() if isinstance(m, GraphAttentionLayer)]
        self.arena = assign_grad_arena(module)
        inside = {id(p) for m in self.layers for p in m.parameters()}
        self.params = [p for p in module.parameters() if p.requires_grad]
        self.others = [p for p in self.params if id(p) not in inside]

    def all_reduce(self, group=None, weight=None):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        if not all(m._grad_store["used"] for m in self.layers):
            all_reduce_packed_grads(self.params, group, weight)
            return
        scale = (1.0 / dist.get_world_size(group)) if weight is None else float(weight)
        self.arena.mul_(scale)
        dist.all_reduce(self.arena, op=dist.ReduceOp.SUM, group=group)
        if self.others:
            all_reduce_packed_grads(self.others, group, weight)
