"""CPU tests of host-side logic added in round 2 (no GPU, no CUDA library calls): batch padding for captured steps, graph
selection for data-parallel strong splits, the bench's workload / config plumbing and the reference staging recipe."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_pad_batch_layout_and_limits():
    from atmlgraphattentionnetworks_b200 import synth
    from atmlgraphattentionnetworks_b200.capture import pad_batch
    d = synth.cifar_shaped(num_graphs=5, seed=1)
    n, e = d.x.shape[0], d.edge_index.shape[1]
    p = pad_batch(d, n + 10, e + 25, 8)
    assert p.x.shape == (n + 10, 5) and p.edge_index.shape == (2, e + 25) and p.num_graphs == 8
    assert torch.equal(p.x[:n], d.x) and float(p.x[n:].abs().max()) == 0.0           # dummy nodes carry zero features
    pad = p.edge_index[:, e:]
    assert torch.equal(pad[0], pad[1]) and int(pad.min()) >= n and int(pad.max()) < n + 10   # self loops on dummy nodes only
    assert int(torch.bincount(pad[0] - n).max()) <= 3                                  # spread round-robin (no hub row)
    assert torch.equal(p.batch[:n], d.batch) and set(p.batch[n:].tolist()) == {5}      # dummy nodes form one extra graph
    assert p.y.tolist() == d.y.tolist() + [-100] * 3                                   # F.nll_loss skips the dummy graphs
    assert d.x.shape[0] == n and d.edge_index.shape[1] == e                            # the input batch is not modified
    with pytest.raises(ValueError):
        pad_batch(d, n - 1, e, 5)
    with pytest.raises(ValueError):
        pad_batch(d, n, e + 1, 5)                                                      # edges to pad but no dummy node
    c = synth.cora_shaped(num_nodes=50, undirected_pairs=60, num_features=7)
    q = pad_batch(c, 64, 150)
    assert q.y[50:].tolist() == [-100] * 14 and q.edge_index.shape[1] == 150           # node-level labels are padded too


def test_select_graphs_recollates_a_block_diagonal_batch():
    from atmlgraphattentionnetworks_b200 import synth
    from atmlgraphattentionnetworks_b200.parallel import shard_graphs
    d = synth.cifar_shaped(num_graphs=12, seed=2)
    parts = [synth.select_graphs(d, shard_graphs(12, 3, r)) for r in range(3)]
    assert sum(p.x.shape[0] for p in parts) == d.x.shape[0] and sum(p.edge_index.shape[1] for p in parts) == d.edge_index.shape[1]
    for r, p in enumerate(parts):
        assert p.num_graphs == 4 and torch.equal(p.y, d.y[shard_graphs(12, 3, r)])
        assert int(p.edge_index.min()) >= 0 and int(p.edge_index.max()) < p.x.shape[0]
        assert torch.equal(p.batch[p.edge_index[0]], p.batch[p.edge_index[1]])         # edges stay inside their graph
        assert torch.equal(p.batch, torch.sort(p.batch).values)                        # nodes stay graph-contiguous
        # the first selected graph is the original graph r: same features in the same order
        g0 = d.x[d.batch == r]
        assert torch.equal(p.x[: g0.shape[0]], g0)
    node = synth.select_graphs(synth.ppi_shaped(keep_graphs=4), [1, 3])                 # node-level labels follow the nodes
    assert node.y.shape[0] == node.x.shape[0]


def test_bench_config_is_the_same_object_for_both_arms():
    """`config` describes the WORKLOAD only, so the reference arm prints the b200 arm's dict (the driver's same_config)"""
    import bench
    data, spec, _, desc = bench.make_workload("ppi", 0, sample=2)
    a, flush_a = bench.describe_config("ppi", desc, data, spec, 1, "dp1")
    b, _ = bench.describe_config("ppi", desc, data, spec, 1, "anything")
    assert a == b and not flush_a and "inputs larger than L2" in a["l2"]
    small, _, _, sdesc = bench.make_workload("cora", 0)
    c, flush_c = bench.describe_config("cora", sdesc, small, [(1433, 8, 8, True), (64, 7, 1, False)], 1, "single")
    assert flush_c and "flushed" in c["l2"]
    # bf16 rows halve the gathered bytes and add the bf16 copies (separately toleranced mode)
    f32 = bench.algorithmic_bytes("b200gat_edge_fwd", 1000, 20000, 64, 128, 4, True, True, cached=False)[0]
    b16 = bench.algorithmic_bytes("b200gat_edge_fwd", 1000, 20000, 64, 128, 4, True, True, cached=False, row_b=2)[0]
    assert f32 - b16 == 20000 * 2 * 512


def test_reference_staging_recipe(tmp_path):
    from oracle import stage_reference
    src = tmp_path / "ref"
    src.mkdir()
    for name in stage_reference.FILES:
        (src / name).write_text(f"# {name}\n")
    dest = tmp_path / "_ref"
    assert stage_reference.stage(str(src), str(dest)) == str(dest)
    man = json.load(open(dest / "MANIFEST.json"))
    assert sorted(man["sha256"]) == sorted(stage_reference.FILES)
    assert (dest / "GAT.py").read_text() == "# GAT.py\n"                                # byte for byte
    assert stage_reference.stage(str(tmp_path / "missing"), str(dest)) is None           # GPU box: nothing to stage from


def test_gather_dtype_switch_and_arena_layout():
    import GAT
    from atmlgraphattentionnetworks_b200.gat import grad_arena_numel, set_gather_dtype
    layer = GAT.GraphAttentionLayer(50, 121, num_heads=6, concat=False)
    assert layer.gather_dtype == torch.float32
    set_gather_dtype(layer, torch.bfloat16)
    assert layer.gather_dtype == torch.bfloat16
    with pytest.raises(ValueError):
        set_gather_dtype(layer, torch.float16)
    dp = 6 * 124
    assert grad_arena_numel(layer) == dp * 50 + 3 * dp + 2 * 6 + 121
