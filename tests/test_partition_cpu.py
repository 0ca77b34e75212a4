"""CPU test of the row-partition graph builder (torch ops) against the integer CSR/CSC oracle: every rank's local
CSR / CSC must be exactly the matching slice of the single-GPU canonical arrays (bit-exact index work)."""
import numpy as np
import pytest
import torch

from atmlgraphattentionnetworks_b200.partition import block_size, build_row_partition, gather_layout
from oracle.csr_oracle import csr_oracle


@pytest.mark.parametrize("n,e,world", [(50, 400, 2), (101, 900, 4), (64, 0, 8), (37, 300, 3), (5, 40, 8)])
def test_row_partition_is_a_slice_of_the_canonical_csr(n, e, world):
    rng = np.random.default_rng(n + e + world)
    ei = rng.integers(0, n, size=(2, e))
    if e:
        ei[1, : e // 3] = 3                                    # a hub destination
        ei[:, 5:15] = ei[:, 20:30]                             # duplicates
    want = csr_oracle(ei, n)
    b = block_size(n, world)
    seen_rows = 0
    for r in range(world):
        p = build_row_partition(torch.from_numpy(ei), n, world, r)
        lo, hi = min(r * b, n), min((r + 1) * b, n)
        assert (p.lo, p.hi, p.block) == (lo, hi, b)
        seen_rows += p.n_own
        a0, a1 = want["rowptr"][lo], want["rowptr"][hi]
        assert np.array_equal(p.rowptr.numpy().astype(np.int64), want["rowptr"][lo:hi + 1] - a0)
        assert np.array_equal(p.col.numpy().astype(np.int64), want["col"][a0:a1])
        assert np.array_equal(p.eid.numpy().astype(np.int64), want["eid"][a0:a1])
        c0, c1 = want["colptr"][lo], want["colptr"][hi]
        assert np.array_equal(p.colptr.numpy().astype(np.int64), want["colptr"][lo:hi + 1] - c0)
        assert np.array_equal(p.crow.numpy().astype(np.int64), want["crow"][c0:c1])
        assert np.array_equal(p.ceid.numpy().astype(np.int64), want["ceid"][c0:c1])
    assert seen_rows == n


def test_row_partition_degree_classes():
    """hub rows / columns (degree > HUB_DEGREE) of every rank's block, their rowend / colend view and the largest degrees
    (include/b200gat.h: b200gat_graph.hub_rows ...), against the canonical arrays."""
    from atmlgraphattentionnetworks_b200._abi import HUB_DEGREE
    n, e, world = 90, 5000, 3
    rng = np.random.default_rng(5)
    ei = rng.integers(0, n, size=(2, e))
    ei[1, :1500] = 7                                           # hub destination (rank 0)
    ei[0, 1500:2400] = 64                                      # hub source (rank 2)
    want = csr_oracle(ei, n)
    b = block_size(n, world)
    found = [0, 0]
    for r in range(world):
        p = build_row_partition(torch.from_numpy(ei), n, world, r)
        lo, hi = min(r * b, n), min((r + 1) * b, n)
        for which, (ptr_k, hubs, ends, mx) in enumerate((("rowptr", p.hub_rows, p.rowend, p.max_in_degree),
                                                         ("colptr", p.hub_cols, p.colend, p.max_out_degree))):
            ptr = want[ptr_k][lo:hi + 1] - want[ptr_k][lo]
            deg = np.diff(ptr)
            assert np.array_equal(np.sort(hubs.numpy()), np.nonzero(deg > HUB_DEGREE)[0])
            assert np.array_equal(ends.numpy().astype(np.int64), np.where(deg > HUB_DEGREE, ptr[:-1], ptr[1:]))
            assert mx == deg.max()
            found[which] += int(hubs.numel())
    assert found == [1, 1]


def test_gather_layout():
    assert gather_layout((100, 128, 4, True)) == (True, 512, 512, 128)       # gout gathered directly
    assert gather_layout((512, 47, 4, False)) == (False, 48, 48, 0)          # mean mode: one padded [C] row per node
    assert gather_layout((9, 5, 3, True)) == (False, 24, 24, 8)              # unaligned heads: padded copy
    assert gather_layout((64, 7, 1, False)) == (False, 8, 8, 8)              # H == 1 is concat-like
