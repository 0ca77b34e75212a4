"""Inert placeholder so `from torch_geometric.loader import DataLoader` (GATNet.py:8) succeeds."""


class DataLoader:  # never instantiated on the GAT path
    def __init__(self, *a, **k):
        raise NotImplementedError("loaders are out of scope")
