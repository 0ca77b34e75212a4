"""Drop-in replacement for the reference's GATNet.py: `from GATNet import GATNet` (run_inductive.py:9,
run_gnn_benchmark.py:9) resolves to the model built on the B200-native layer."""
from atmlgraphattentionnetworks_b200.gatnet import GATNet  # noqa: F401
from GAT import GraphAttentionLayer  # noqa: F401
