"""CPU tests: the oracle (port + closed form + CSR oracle) against the fixtures produced by the unmodified
reference (tests/golden/make_golden.py), and against the reference itself when /root/reference is present."""
import numpy as np
import pytest
import torch

from oracle import closed_form, csr_oracle, ref_loader
from oracle.gat_port import PortGATNet
from util import (ACT_FILES, FP32_TOL, GRAD_KEYS, HUBREF_FILES, LAYER_FILES, NET_FILES, case_id, load, nerr, packed_grads,
                  port_layer_from_golden)


def test_fixture_inventory():
    assert len(LAYER_FILES) >= 18 and len(NET_FILES) >= 3 and len(ACT_FILES) >= 5 and len(HUBREF_FILES) >= 4


def test_hub_fixtures_hold_the_degree_classes():
    """the hubref fixtures really contain rows / columns above B200GAT_HUB_DEGREE and B200GAT_GIANT_DEGREE"""
    from atmlgraphattentionnetworks_b200._abi import HUB_DEGREE
    for path in HUBREF_FILES:
        g = load(path)
        n = g["x"].shape[0]
        indeg = np.bincount(g["edge_index"][1], minlength=n) + 1
        outdeg = np.bincount(g["edge_index"][0], minlength=n) + 1
        lim = 4096 if "giant" in case_id(path) else HUB_DEGREE
        assert indeg.max() > lim and outdeg.max() > lim, case_id(path)


@pytest.mark.parametrize("path", LAYER_FILES + ACT_FILES + HUBREF_FILES, ids=case_id)
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_port_matches_reference_fixture(path, prec):
    g = load(path)
    dt = torch.float32 if prec == "f32" else torch.float64
    layer = port_layer_from_golden(g, dt)
    layer.train()
    if "mask" in g:
        mask = torch.from_numpy(g["mask"])
        layer.mask_hook = lambda shape: mask
    else:
        layer.dropout_val = 0.0
    x = torch.from_numpy(g["x"]).to(dt).requires_grad_(True)
    out = layer(x, torch.from_numpy(g["edge_index"]))
    out.backward(torch.from_numpy(g["gout"]).to(dt))
    tol = FP32_TOL if prec == "f32" else 1e-12
    assert nerr(out.detach().numpy(), g["out_" + prec]) <= tol
    got = packed_grads(layer, x.grad)
    for k in GRAD_KEYS:
        assert nerr(got[k], g[k + "_" + prec]) <= tol, k


@pytest.mark.parametrize("path", LAYER_FILES + HUBREF_FILES, ids=case_id)
def test_closed_form_matches_reference_fixture(path):
    g = load(path)
    out, cache = closed_form.forward(g["x"], g["edge_index"], g["W"], g["bw"], g["a1"], g["b1"], g["a2"], g["b2"],
                                     g["bias"], bool(g["concat"]), mask=g.get("mask"))
    assert nerr(out, g["out_f64"]) <= 1e-12
    gr = closed_form.backward(g["gout"], cache)
    for k in GRAD_KEYS:
        assert nerr(gr[k[2:]], g[k + "_f64"]) <= 1e-9, k   # (cancellation in near-zero sums on big_logits)
    # fp32 reference sits within the stated tolerance of f64 truth (noise floor)
    assert nerr(g["out_f32"], g["out_f64"]) <= FP32_TOL


@pytest.mark.parametrize("path", NET_FILES, ids=case_id)
def test_port_net_matches_reference_fixture(path):
    from types import SimpleNamespace
    g = load(path)
    ds = {"net_cora": "Cora", "net_cifar_f3": "CIFAR10", "net_pubmed": "Pubmed"}[case_id(path)]
    net = PortGATNet("GAT", ds, g["x"].shape[1])
    sd = {k[len("param:"):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param:")}
    assert list(net.state_dict().keys()) == list(sd.keys())          # state_dict key set AND order are contract
    net.load_state_dict(sd)
    net.eval()
    data = SimpleNamespace(x=torch.from_numpy(g["x"]), edge_index=torch.from_numpy(g["edge_index"]),
                           batch=torch.from_numpy(g["batch"]) if "batch" in g else None)
    out = net(data)
    loss = torch.nn.functional.nll_loss(out, torch.from_numpy(g["y"]))
    loss.backward()
    assert nerr(out.detach().numpy(), g["out_f32"]) <= FP32_TOL
    for k, p in net.named_parameters():
        assert nerr(p.grad.numpy(), g["grad:" + k + "_f32"]) <= 2e-5, k


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_port_matches_live_reference_and_init_rng():
    """Same seed => identical parameters (RNG consumption order GAT.py:19-25) and identical fwd/bwd."""
    ref_gat, ref_net = ref_loader.load()
    from oracle.gat_port import PortGraphAttentionLayer
    torch.manual_seed(5)
    a = ref_gat.GraphAttentionLayer(13, 6, num_heads=4, concat=True, dropout=0.0)
    torch.manual_seed(5)
    b = PortGraphAttentionLayer(13, 6, num_heads=4, concat=True, dropout=0.0)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(90, 13, generator=gen)
    ei = torch.randint(0, 90, (2, 500), generator=gen)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    oa, ob = a(xa, ei), b(xb, ei)
    oa.square().sum().backward()
    ob.square().sum().backward()
    assert nerr(ob.detach().numpy(), oa.detach().numpy()) <= 1e-6
    assert nerr(xb.grad.numpy(), xa.grad.numpy()) <= 1e-5
    torch.manual_seed(9)
    na = ref_net.GATNet("GAT", "Cora", 20)
    torch.manual_seed(9)
    nb = PortGATNet("GAT", "Cora", 20)
    for (ka, va), (kb, vb) in zip(na.state_dict().items(), nb.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


def test_csr_oracle_properties():
    rng = np.random.default_rng(0)
    n, e = 50, 400
    ei = rng.integers(0, n, size=(2, e))
    g = csr_oracle.csr_oracle(ei, n)
    full = csr_oracle.full_edges(ei, n)
    assert g["rowptr"][0] == 0 and g["rowptr"][-1] == e + n
    for i in range(n):
        seg = slice(g["rowptr"][i], g["rowptr"][i + 1])
        eids = g["eid"][seg]
        assert np.all(full[1][eids] == i) and np.all(np.diff(eids) > 0)      # stable: original order kept
        assert eids[-1] == e + i and g["col"][seg][-1] == i                 # appended self loop is last
    assert np.array_equal(np.sort(g["eid"]), np.arange(e + n))
    for j in range(n):
        seg = slice(g["colptr"][j], g["colptr"][j + 1])
        assert np.all(full[0][g["ceid"][seg]] == j) and np.all(np.diff(g["crow"][seg]) >= 0)
    empty = csr_oracle.csr_oracle(np.zeros((2, 0), dtype=np.int64), 5)
    assert np.array_equal(empty["rowptr"], np.arange(6)) and np.array_equal(empty["col"], np.arange(5))
