"""GPU tests of whole-step CUDA-graph capture (atmlgraphattentionnetworks_b200/capture.py): the
captured train step / eval forward must reproduce the eager loop of the reference's trainers (run_gnn_benchmark.py:58-66: a
NEW batch and edge_index every step; run_inductive.py:74-95: train step + eval forward per epoch), including the CSR / CSC
build that runs inside the graph, and the capacity padding must be exact."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from util import nerr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _to_dev(d):
    return SimpleNamespace(x=d.x.to(DEV), edge_index=d.edge_index.to(DEV), y=d.y.to(DEV),
                           batch=d.batch.to(DEV) if getattr(d, "batch", None) is not None else None, num_graphs=d.num_graphs)


def test_sync_free_csr_build_matches_the_synchronous_build_and_checks_lazily():
    from atmlgraphattentionnetworks_b200.graph import build_csr
    gen = torch.Generator().manual_seed(0)
    ei = torch.randint(0, 5000, (2, 60000), generator=gen).to(DEV)
    a, b = build_csr(ei, 5000), build_csr(ei, 5000, sync=False)
    for k, t in a.arrays().items():
        assert torch.equal(t, b.arrays()[k]), k
    b.check()
    bad = ei.clone()
    bad[0, 7] = 5000
    g = build_csr(bad, 5000, sync=False)          # no raise, no sync ...
    with pytest.raises(IndexError):
        g.check()                                 # ... the index check is deferred


def test_pad_batch_is_exact():
    """dummy nodes / self loops / graphs change neither the real outputs, nor the loss, nor any gradient"""
    import GATNet
    from atmlgraphattentionnetworks_b200 import synth
    from atmlgraphattentionnetworks_b200.capture import pad_batch
    d = _to_dev(synth.cifar_shaped(num_graphs=12, seed=3))
    n, e = d.x.shape[0], d.edge_index.shape[1]
    p = pad_batch(d, n + 37, e + 501, d.num_graphs + 3)
    torch.manual_seed(0)
    net = GATNet.GATNet("GAT", "CIFAR10", 5).to(DEV).train()

    def run(data):
        net.zero_grad()
        out = net(data)
        loss = F.nll_loss(out, data.y)
        loss.backward()
        return out.detach(), float(loss), torch.cat([q.grad.flatten() for q in net.parameters()]).clone()
    o1, l1, g1 = run(d)
    o2, l2, g2 = run(p)
    assert p.x.shape[0] == n + 37 and p.edge_index.shape[1] == e + 501 and o2.shape[0] == d.num_graphs + 3
    assert nerr(o2[:d.num_graphs].cpu().numpy(), o1.cpu().numpy()) <= 1e-6 and abs(l1 - l2) <= 1e-6 * abs(l1)
    assert nerr(g2.cpu().numpy(), g1.cpu().numpy()) <= 1e-5


def test_captured_train_step_with_a_new_batch_every_step_matches_the_eager_loop():
    import GATNet
    from atmlgraphattentionnetworks_b200 import synth
    from atmlgraphattentionnetworks_b200.capture import capture_train_step, pad_batch
    batches = [_to_dev(synth.cifar_shaped(num_graphs=16, seed=s)) for s in range(4)]
    n_cap = max(b.x.shape[0] for b in batches) + 8
    e_cap = max(b.edge_index.shape[1] for b in batches) + 64
    padded = [pad_batch(b, n_cap, e_cap, 17) for b in batches]

    def make():
        torch.manual_seed(1)
        net = GATNet.GATNet("GAT", "CIFAR10", 5).to(DEV).train()
        opt = torch.optim.Adam(net.parameters(), lr=5e-3, weight_decay=5e-4, fused=True, capturable=True)
        return net, opt
    net_e, opt_e = make()
    eager = []
    for b in batches:                                       # the reference's loop, run_gnn_benchmark.py:60-66
        opt_e.zero_grad(set_to_none=True)
        loss = F.nll_loss(net_e(b), b.y)
        loss.backward()
        opt_e.step()
        eager.append(float(loss.detach()))
    net_c, opt_c = make()
    state0 = {k: v.clone() for k, v in net_c.state_dict().items()}
    step = capture_train_step(net_c, opt_c, lambda out, y: F.nll_loss(out, y), padded[0])
    # the warm-up / capture runs already stepped the optimizer: rewind parameters and Adam state in place
    net_c.load_state_dict(state0)
    for st in opt_c.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()
    got = []
    for b in padded:
        loss = step(x=b.x, edge_index=b.edge_index, y=b.y, batch=b.batch)
        got.append(float(loss))                             # reading the loss is the caller's choice, not the step's
    step.check()
    assert step.replays == len(padded)
    assert np.allclose(got, eager, rtol=2e-5, atol=0), (got, eager)
    for (k, a), (_, b) in zip(net_c.state_dict().items(), net_e.state_dict().items()):
        assert nerr(a.cpu().numpy(), b.cpu().numpy()) <= 1e-4, k


def test_captured_eval_forward_and_new_graph_per_replay():
    import GATNet
    from atmlgraphattentionnetworks_b200 import synth
    from atmlgraphattentionnetworks_b200.capture import capture_eval_forward
    d = _to_dev(synth.cora_shaped(num_nodes=900, undirected_pairs=2000, num_features=100))
    torch.manual_seed(2)
    net = GATNet.GATNet("GAT", "Cora", 100).to(DEV).eval()
    with torch.no_grad():
        want = net(d).clone()
    fwd = capture_eval_forward(net, d)
    assert nerr(fwd().cpu().numpy(), want.cpu().numpy()) <= 1e-6
    perm = torch.randperm(900, device=DEV)
    ei2 = perm[d.edge_index]                                # another graph with the same edge count
    with torch.no_grad():
        want2 = net(SimpleNamespace(x=d.x, edge_index=ei2)).clone()
    got2 = fwd(edge_index=ei2)
    assert nerr(got2.cpu().numpy(), want2.cpu().numpy()) <= 1e-6 and nerr(got2.cpu().numpy(), want.cpu().numpy()) > 1e-3
    with pytest.raises(ValueError):
        fwd(edge_index=ei2[:, :-1])                         # shapes are static: pad_batch is the way


def test_captured_training_draws_a_fresh_dropout_mask_every_replay():
    """Cora-style training (feature dropout 0.6 + in-kernel attention dropout 0.6): the seed words are produced by an RNG
    kernel INSIDE the graph, so two replays on the same inputs give different losses; eval replays are deterministic."""
    import GATNet
    from atmlgraphattentionnetworks_b200 import synth
    from atmlgraphattentionnetworks_b200.capture import capture_train_step
    d = _to_dev(synth.cora_shaped(num_nodes=800, undirected_pairs=1800, num_features=64))
    torch.manual_seed(3)
    net = GATNet.GATNet("GAT", "Cora", 64).to(DEV).train()
    opt = torch.optim.Adam(net.parameters(), lr=0.0, fused=True, capturable=True)       # lr 0: only the masks change
    step = capture_train_step(net, opt, lambda out, y: F.nll_loss(out, y), d)
    losses = [float(step()) for _ in range(4)]
    assert len(set(losses)) == 4, losses


# ------------------------------------------------------------------ fused readout head (GATNet.py:72-75) vs torch ops
@pytest.mark.parametrize("shape", [(15052, 128, 64, 64, 10), (1000, 7, 33, 20, 3), (50, 60, 64, 64, 10)],
                         ids=["cifar128", "odd", "empty_graphs"])
@pytest.mark.parametrize("act_in", [False, True], ids=["plain", "elu_on_load"])
def test_readout_head_matches_torch_ops(shape, act_in):
    """scatter_mean -> lin1 -> relu -> lin2 -> log_softmax as one fused op, forward and all five gradients, against the
    same composition in float64 torch ops (unsorted batch vector, empty graphs, ELU applied on load)."""
    from atmlgraphattentionnetworks_b200.gatnet import readout_head
    n, g, f, hd, k = shape
    gen = torch.Generator().manual_seed(n + g)
    x = torch.randn(n, f, generator=gen)
    batch = torch.randint(0, g if g < n else n // 2, (n,), generator=gen)          # unsorted; some graphs stay empty
    lin1, lin2 = torch.nn.Linear(f, hd), torch.nn.Linear(hd, k)
    gl = torch.randn(g, k, generator=gen)

    def ref(dt):
        xr = x.to(dt).requires_grad_(True)
        l1, l2 = torch.nn.Linear(f, hd).to(dt), torch.nn.Linear(hd, k).to(dt)
        l1.load_state_dict({a: b.to(dt) for a, b in lin1.state_dict().items()})
        l2.load_state_dict({a: b.to(dt) for a, b in lin2.state_dict().items()})
        xa = F.elu(xr) if act_in else xr
        tot = torch.zeros(g, f, dtype=dt).index_add(0, batch, xa)
        cnt = torch.zeros(g, dtype=dt).index_add(0, batch, torch.ones(n, dtype=dt)).clamp(min=1)
        xa.retain_grad()
        out = F.log_softmax(l2(F.relu(l1(tot / cnt.unsqueeze(1)))), dim=1)
        out.backward(gl.to(dt))
        # the fused op returns the gradient w.r.t. act(x) when act_in (the producing layer applies act')
        gx = xa.grad if act_in else xr.grad
        return [out.detach(), gx, l1.weight.grad, l1.bias.grad, l2.weight.grad, l2.bias.grad]
    want = ref(torch.float64)
    xg = x.to(DEV).requires_grad_(True)
    l1, l2 = lin1.to(DEV), lin2.to(DEV)
    out = readout_head(xg, batch.to(DEV), l1, l2, num_graphs=g, act_in=act_in)
    out.backward(gl.to(DEV))
    got = [out.detach(), xg.grad, l1.weight.grad, l1.bias.grad, l2.weight.grad, l2.bias.grad]
    for name, a, b in zip(("logp", "g_x", "g_w1", "g_b1", "g_w2", "g_b2"), got, want):
        assert nerr(a.cpu().numpy(), b.numpy()) <= 1e-5, name


def test_captured_step_with_a_static_graph_builds_the_csr_once():
    """full-graph training (run_inductive.py:74-95): static_graph=True keeps the ingestion out of the replayed graph, gives
    the same losses as the dynamic capture on the same inputs (lr = 0, eval-free model without dropout) and refuses a new
    edge_index"""
    from atmlgraphattentionnetworks_b200 import synth
    from atmlgraphattentionnetworks_b200.capture import CapturedStep
    from atmlgraphattentionnetworks_b200.gatnet import GATStack
    d = _to_dev(synth.cora_shaped(num_nodes=700, undirected_pairs=1500, num_features=48))
    torch.manual_seed(4)
    model = GATStack([(48, 8, 4, True), (32, 7, 1, False)], dropout=0.0).to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=0.0, fused=True, capturable=True)

    def fn(x, edge_index, y):
        opt.zero_grad(set_to_none=True)
        loss = F.nll_loss(F.log_softmax(model(x, edge_index), dim=1), y)
        loss.backward()
        opt.step()
        return loss.detach()
    inputs = dict(x=d.x, edge_index=d.edge_index, y=d.y)
    dyn, sta = CapturedStep(fn, inputs), CapturedStep(fn, inputs, static_graph=True)
    a, b = float(dyn()), float(sta())
    assert abs(a - b) <= 1e-6 * abs(a)
    x2 = torch.randn_like(d.x)
    assert abs(float(dyn(x=x2)) - float(sta(x=x2))) <= 1e-6 * abs(a) and abs(float(sta()) - b) > 1e-4
    with pytest.raises(ValueError):
        sta(edge_index=d.edge_index)
