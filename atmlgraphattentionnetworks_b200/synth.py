"""Synthetic graphs of the five BASELINE.json shapes.  There is no network, so every
benchmark / parity input is generated here, on the CPU, from torch.Generator().manual_seed(seed).

Shapes mirror what the reference's callers feed the layer:
  cora   ~ run_inductive.py:43-61 (Planetoid Cora, NormalizeFeatures)        one graph, 2-layer GATNet
  ppi    ~ BASELINE.json config 2 (24 block-diagonal graphs, 50 feats)       3-layer stack 4/4/6 heads x 256
  cifar  ~ run_gnn_benchmark.py:35-41 (superpixel kNN graphs, batched)       GATNet('GAT','CIFAR10',F)
  large  ~ BASELINE.json config 5 (power-law, 2.4M nodes / 62M edges)        3-layer stack 4 heads x 128
"""
from types import SimpleNamespace

import torch


def _undirected_pairs(n, pairs, gen, lo=0):
    a = torch.randint(0, n, (pairs,), generator=gen, dtype=torch.int64) + lo
    b = torch.randint(0, n, (pairs,), generator=gen, dtype=torch.int64) + lo
    return torch.stack([torch.cat([a, b]), torch.cat([b, a])])


def cora_shaped(seed=0, num_nodes=2708, undirected_pairs=5278, num_features=1433, num_classes=7):
    g = torch.Generator().manual_seed(seed)
    edge_index = _undirected_pairs(num_nodes, undirected_pairs, g)
    x = (torch.rand(num_nodes, num_features, generator=g) < 0.0127).float()
    x = x / x.sum(dim=1, keepdim=True).clamp(min=1.0)          # NormalizeFeatures (run_inductive.py:61)
    y = torch.randint(0, num_classes, (num_nodes,), generator=g)
    return SimpleNamespace(x=x, edge_index=edge_index, y=y, num_graphs=1, name="cora")


def ppi_shaped(seed=0, num_graphs=24, num_nodes=56944, num_edges=818716, num_features=50, num_labels=121,
               keep_graphs=None):
    """Block-diagonal batch; `keep_graphs=k` returns only the first k graphs (bounded CPU-baseline sample)."""
    g = torch.Generator().manual_seed(seed)
    w = torch.rand(num_graphs, generator=g) + 0.3
    sizes = torch.floor(w / w.sum() * num_nodes).long()
    sizes[0] += num_nodes - sizes.sum()
    pairs = torch.floor(sizes.double() / num_nodes * (num_edges // 2)).long()
    pairs[0] += num_edges // 2 - pairs.sum()
    offs = torch.cumsum(sizes, 0) - sizes
    k = num_graphs if keep_graphs is None else int(keep_graphs)
    blocks = [_undirected_pairs(int(sizes[b]), int(pairs[b]), g, lo=int(offs[b])) for b in range(num_graphs)][:k]
    n = int(sizes[:k].sum())
    edge_index = torch.cat(blocks, dim=1)
    gx = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(num_nodes, num_features, generator=gx)[:n].contiguous()
    y = (torch.rand(num_nodes, num_labels, generator=gx) < 0.5).float()[:n].contiguous()
    batch = torch.repeat_interleave(torch.arange(k), sizes[:k])
    return SimpleNamespace(x=x, edge_index=edge_index, y=y, batch=batch, num_graphs=k, name="ppi")


def cifar_shaped(seed=0, num_graphs=128, num_features=5, k=8, num_classes=10):
    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(85, 151, (num_graphs,), generator=g)
    srcs, dsts, lo = [], [], 0
    for b in range(num_graphs):
        nb = int(sizes[b])
        pos = torch.rand(nb, 2, generator=g)
        d = torch.cdist(pos, pos)
        d.fill_diagonal_(float("inf"))
        nbr = d.topk(k, dim=1, largest=False).indices           # [nb,k]: the k nearest sources of each target
        srcs.append(nbr.reshape(-1) + lo)
        dsts.append(torch.arange(nb).repeat_interleave(k) + lo)
        lo += nb
    edge_index = torch.stack([torch.cat(srcs), torch.cat(dsts)])
    n = int(sizes.sum())
    x = torch.rand(n, num_features, generator=g)
    y = torch.randint(0, num_classes, (num_graphs,), generator=g)
    batch = torch.repeat_interleave(torch.arange(num_graphs), sizes)
    return SimpleNamespace(x=x, edge_index=edge_index, y=y, batch=batch, num_graphs=num_graphs, name="cifar")


def select_graphs(data, graph_ids):
    """Re-collate the graphs `graph_ids` of a block-diagonal batch (nodes numbered graph-contiguously, edges never cross
    graphs) into a batch of their own: what a data-parallel rank owns of a global batch (parallel.shard_graphs)."""
    ids = torch.as_tensor(list(graph_ids), dtype=torch.int64)
    g_total = int(data.num_graphs)
    new_id = torch.full((g_total,), -1, dtype=torch.int64)
    new_id[ids] = torch.arange(ids.numel())
    node_keep = new_id[data.batch] >= 0
    order = torch.argsort(new_id[data.batch][node_keep], stable=True)      # nodes grouped by NEW graph id
    old_nodes = torch.nonzero(node_keep).flatten()[order]
    remap = torch.full((data.x.shape[0],), -1, dtype=torch.int64)
    remap[old_nodes] = torch.arange(old_nodes.numel())
    ei = data.edge_index
    edge_keep = node_keep[ei[1]]
    out = SimpleNamespace(x=data.x[old_nodes].contiguous(), edge_index=remap[ei[:, edge_keep]].contiguous(),
                          batch=new_id[data.batch][old_nodes].contiguous(), num_graphs=int(ids.numel()), name=data.name)
    y = data.y
    out.y = (y[ids] if y.shape[0] == g_total and y.shape[0] != data.x.shape[0] else y[old_nodes]).contiguous()
    return out


def powerlaw(seed=0, num_nodes=2_400_000, num_edges=62_000_000, num_features=100, num_classes=47):
    """Both endpoints = perm[floor(N * U^2)] (Zipf-1/2 popularity; in-degree tail exponent ~3)."""
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(num_nodes, generator=g)
    def endpoints():
        u = torch.rand(num_edges, generator=g, dtype=torch.float64)
        return perm[(u * u * num_nodes).long().clamp_(max=num_nodes - 1)]
    edge_index = torch.stack([endpoints(), endpoints()])
    x = torch.randn(num_nodes, num_features, generator=g)
    y = torch.randint(0, num_classes, (num_nodes,), generator=g)
    return SimpleNamespace(x=x, edge_index=edge_index, y=y, num_graphs=1, name="large")


PPI_STACK = [(50, 256, 4, True), (1024, 256, 4, True), (1024, 121, 6, False)]      # (in, out, heads, concat)
LARGE_STACK = [(100, 128, 4, True), (512, 128, 4, True), (512, 47, 4, False)]
