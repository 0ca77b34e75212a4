"""B200-native GAT layer hot path behind the ATMLGraphAttentionNetworks API (GAT.py / GATNet.py)."""
from .gat import GraphAttentionLayer  # noqa: F401
from .gatnet import GATNet, GATStack  # noqa: F401
from .graph import GraphCSR, build_csr  # noqa: F401
