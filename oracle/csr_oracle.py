"""Integer oracle for the graph structure arrays (test infrastructure).

Canonical edge order follows GAT.py:38 -> [PyG] add_self_loops: full = cat([edge_index, (n,n) for n<N]).
Destination-sorted CSR = STABLE sort of `full` by target, so inside a row the original edge order is kept and
the appended self loop is last.  CSC = STABLE sort of the CSR-ordered edges by source.
"""
import numpy as np


def full_edges(edge_index, num_nodes):
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    loops = np.arange(num_nodes, dtype=np.int64)
    return np.concatenate([ei, np.stack([loops, loops])], axis=1)


def csr_oracle(edge_index, num_nodes):
    """-> dict(rowptr[N+1], col[E'], eid[E'], colptr[N+1], crow[E'], ceid[E'], cpos[E']) as int64."""
    full = full_edges(edge_index, num_nodes)
    src, dst = full[0], full[1]
    perm = np.argsort(dst, kind="stable")
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=num_nodes), out=rowptr[1:])
    col, eid, row = src[perm], perm, dst[perm]
    cperm = np.argsort(col, kind="stable")
    colptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(col, minlength=num_nodes), out=colptr[1:])
    return dict(rowptr=rowptr, col=col, eid=eid, colptr=colptr, crow=row[cperm], ceid=eid[cperm], cpos=cperm)
