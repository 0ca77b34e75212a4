"""Drop-in replacement for the reference's GAT.py: `from GAT import GraphAttentionLayer` (GATNet.py:9,
run_inductive.py:8, run_heads_experiment.py:7, run_params_experiment.py:7) resolves to the B200-native layer."""
from atmlgraphattentionnetworks_b200.gat import GraphAttentionLayer  # noqa: F401
