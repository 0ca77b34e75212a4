"""Inert placeholder so `import torch_geometric.transforms as T` (GATNet.py:4) succeeds."""
