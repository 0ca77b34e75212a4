"""Stand-in for torch_geometric.utils (only what GAT.py:38 and GAT.py:60 call)."""
import torch


def add_self_loops(edge_index, edge_attr=None, fill_value=None, num_nodes=None):
    # PyG semantics: append (n, n) for every node at the END; never dedup / remove existing loops.
    # fill_value is only consulted when edge_attr is given (the reference never passes edge_attr).
    n = int(num_nodes) if num_nodes is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
    loop = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device).unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop], dim=1), edge_attr


def softmax(src, index, ptr=None, num_nodes=None, dim=0):
    # PyG semantics: segment softmax grouped by `index`, max-subtracted, denominator + 1e-16,
    # applied independently to every trailing column.
    n = int(num_nodes) if num_nodes is not None else (int(index.max()) + 1 if index.numel() else 0)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    seg_max = torch.full((n,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
    seg_max = seg_max.scatter_reduce(0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - seg_max.index_select(0, index)).exp()
    seg_sum = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).scatter_add(0, idx, out)
    return out / (seg_sum.index_select(0, index) + 1e-16)
