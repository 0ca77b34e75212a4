"""Minimal behavioural stand-in for the three torch_geometric symbols the reference's GAT.py uses.

TEST INFRASTRUCTURE ONLY (part of oracle/). torch_geometric / torch_scatter are third-party, un-pinned
dependencies of the reference (no requirements file; environment evidence points at PyG ~2.0.4) and are
not installable in this image.  This package restates their *published* semantics
so that /root/reference/GAT.py and GATNet.py can be imported UNMODIFIED and used as the executable oracle.
Nothing in the product path imports this.
"""
from . import nn, utils, transforms, datasets, loader  # noqa: F401

__version__ = "0.0-standin"
