"""Inert placeholder so `from torch_geometric.datasets import Planetoid` (GATNet.py:6) succeeds."""


class Planetoid:  # never instantiated on the GAT path
    def __init__(self, *a, **k):
        raise NotImplementedError("datasets are out of scope (no network)")
