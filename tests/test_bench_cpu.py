"""CPU tests of the measurement code: the algorithmic-byte model bench.py reports rooflines against must reproduce the
table of SURVEY.md §8d, and the reference arm (`bench.py --impl reference`, the CPU oracle port on a bounded sample) must
print the contract's JSON line without a GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GB = 1e9
# (N, E', F, C, H, concat, need_gx, cached) -> (K1, K2, K3, K4) in GB, SURVEY.md §8d "Model values"
SURVEY_ROWS = {
    "C2-L1 50->4x256 cat": ((56944, 875660, 50, 256, 4, True, False, True), (0.247, 0.473, 1.177, 0.245)),
    "C2-L2 1024->4x256 cat": ((56944, 875660, 1024, 256, 4, True, True, True), (0.473, 0.473, 1.177, 0.708)),
    "C2-L3 1024->6x121 mean": ((56944, 875660, 1024, 121, 6, False, True, True), (0.404, 0.366, 0.566, 0.638)),
    "C4 H=8 50->8x64 cat": ((56944, 875660, 50, 64, 8, True, False, True), (0.132, 0.242, 0.601, 0.128)),
    "C5-L1 100->4x128 cat": ((2_400_000, 64_400_000, 100, 128, 4, True, False, False), (5.95, 138.2, 156.1, 5.88)),
    "C5-L2 512->4x128 cat": ((2_400_000, 64_400_000, 512, 128, 4, True, True, False), (9.91, 138.2, 156.1, 14.75)),
    "C5-L3 512->4x47 mean": ((2_400_000, 64_400_000, 512, 47, 4, False, True, False), (6.80, 52.1, 22.5, 11.64)),
}


@pytest.mark.parametrize("name", list(SURVEY_ROWS), ids=lambda s: s.split()[0])
def test_algorithmic_bytes_reproduce_the_survey_table(name):
    import bench
    (n, ep, f, c, h, concat, need_gx, cached), want = SURVEY_ROWS[name]
    ops = ("b200gat_proj_fwd", "b200gat_edge_fwd", "b200gat_edge_bwd", "b200gat_proj_bwd")
    for op, w in zip(ops, want):
        got, bound = bench.algorithmic_bytes(op, n, ep, f, c, h, concat, need_gx, cached)
        assert bound == ("tensor" if "proj" in op else "hbm")
        assert abs(got / GB - w) <= 0.012 * w + 0.0006, (name, op, got / GB, w)


def test_gemm_flops_model():
    import bench
    # C2-L2: 358.5 GF forward + backward, 2 N F D (2 + gx) + 4 N D
    n, f, c, h = 56944, 1024, 256, 4
    total = bench.gemm_flops("b200gat_proj_fwd", n, f, c, h, True) + bench.gemm_flops("b200gat_proj_bwd", n, f, c, h, True)
    assert abs(total / 1e9 - 358.5) < 1.0


def test_reference_arm_prints_the_contract_line_without_a_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cifar",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "edges/s" and line["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
                "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    from oracle import ref_loader
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["nodes"] > 0 and "workload" in line["config"] and "l2" in line["config"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0


def test_graph_cache_keys_on_tensor_identity_and_version(monkeypatch):
    """Host logic of graph.GraphCache (no GPU: the CSR build is stubbed): a hit needs the same tensor object at the same
    in-place version; a refilled buffer (run_gnn_benchmark.py:60-63 style loaders) replaces its stale entry instead of
    growing the cache; a new tensor object is a new graph."""
    import torch
    from atmlgraphattentionnetworks_b200 import graph
    built = []
    monkeypatch.setattr(graph, "build_csr", lambda ei, n, validate=True: built.append((id(ei), ei._version, n)) or object())
    cache = graph.GraphCache(capacity=3)
    ei = torch.zeros((2, 5), dtype=torch.int64)
    a = cache.get(ei, 4)
    assert cache.get(ei, 4) is a and (cache.hits, cache.misses) == (1, 1)
    ei.add_(1)                                             # in-place refill bumps the version: rebuild, stale entry dropped
    b = cache.get(ei, 4)
    assert b is not a and cache.misses == 2 and len(cache._entries) == 1
    assert cache.get(ei, 5) is not b and len(cache._entries) == 1      # other node count: also a different graph
    others = [torch.zeros((2, 3), dtype=torch.int64) for _ in range(4)]
    for t in others:
        cache.get(t, 4)
    assert len(cache._entries) == 3                        # capacity
    del others, t
    cache.get(ei, 5)
    assert all(ref() is not None for ref, *_ in cache._entries)        # dead tensors are purged on access
    assert len(built) == cache.misses


def test_graph_cache_accepts_inference_mode_tensors(monkeypatch):
    """A tensor created under torch.inference_mode() (an eval loop doing batch.to(device) inside it) has no version
    counter: the cache keys it on identity alone instead of raising."""
    import torch
    from atmlgraphattentionnetworks_b200 import graph
    monkeypatch.setattr(graph, "build_csr", lambda ei, n, validate=True: object())
    cache = graph.GraphCache()
    with torch.inference_mode():
        ei = torch.zeros((2, 5), dtype=torch.int64)
        a = cache.get(ei, 4)
        assert cache.get(ei, 4) is a
    assert cache.get(ei, 4) is a and (cache.hits, cache.misses) == (2, 1)
