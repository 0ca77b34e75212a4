"""Stand-in for torch_scatter.scatter_mean (GATNet.py:5,73): per-group mean, empty groups -> 0."""
import torch


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    assert dim == 0 and out is None
    n = int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)
    total = torch.zeros((n,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device).index_add(0, index, src)
    count = torch.zeros(n, dtype=src.dtype, device=src.device).index_add(
        0, index, torch.ones(index.numel(), dtype=src.dtype, device=src.device)).clamp(min=1)
    return total / count.view(-1, *([1] * (src.dim() - 1)))
