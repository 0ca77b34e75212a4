"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference/GAT.py, GATNet.py) on
oracle/pyg_standin.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference ships no tests / golden vectors of its own, so these fixtures — outputs and
autograd gradients of the reference's own Python in fp32 (the 1e-5 target) and fp64 (truth) on seeded inputs —
are what pins the oracle and the CUDA path.  Dropout masks are injected by patching
torch.nn.functional.dropout while the reference's message() runs (GAT.py:61), in ORIGINAL [edges;loops] order.
"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402

# name, N, E, F_in, C, H, concat, p, special
LAYER_CASES = [
    ("h3c5_cat", 60, 300, 11, 5, 3, True, 0.0, ""),
    ("h3c5_mean", 60, 300, 11, 5, 3, False, 0.0, ""),
    ("h1c7_mean", 80, 400, 64, 7, 1, False, 0.0, ""),
    ("h8c8_cat", 120, 700, 33, 8, 8, True, 0.0, ""),
    ("h8c8_cat_drop", 120, 700, 33, 8, 8, True, 0.6, ""),
    ("h8c3_mean_drop", 90, 500, 64, 3, 8, False, 0.6, ""),
    ("h4c64_cat", 150, 1200, 50, 64, 4, True, 0.0, ""),
    ("h2c121_mean", 70, 420, 40, 121, 2, False, 0.0, ""),
    ("h4c47_mean", 64, 380, 24, 47, 4, False, 0.0, ""),
    ("h16c2_cat", 50, 260, 9, 2, 16, True, 0.0, ""),
    ("h32c8_cat", 40, 200, 12, 8, 32, True, 0.0, ""),
    ("h2c256_cat", 48, 400, 20, 256, 2, True, 0.0, ""),
    ("f3_cifar", 100, 800, 3, 8, 8, True, 0.0, ""),
    ("f5_cifar", 100, 800, 5, 8, 8, True, 0.0, ""),
    ("no_edges", 17, 0, 6, 4, 2, True, 0.0, ""),
    ("dups_loops_isolated", 40, 160, 10, 6, 2, True, 0.0, "dups"),
    ("big_logits", 50, 300, 8, 8, 4, True, 0.0, "big"),
    ("hub", 300, 900, 16, 16, 2, True, 0.0, "hub"),
]


# Degree classes (include/b200gat.h: B200GAT_HUB_DEGREE = 512, B200GAT_GIANT_DEGREE = 4096): graphs with a destination and a
# source of 600-700 edges ("hub") and of ~5000 edges ("giant").  Written as hubref_*.npz: they pin the ORACLE on these
# shapes (tests/test_oracle_cpu.py); the CUDA hub / giant schedules are then checked against the oracle on seeded graphs
# of the same kind (tests/test_gpu_parity.py: the *_hub cases).
HUBREF_CASES = [
    ("h4c16_cat", 400, 6000, 12, 16, 4, True, 0.0, "hubs"),
    ("h4c12_mean_drop", 400, 6000, 12, 12, 4, False, 0.6, "hubs"),
    ("giant_h2c16_cat", 500, 12000, 8, 16, 2, True, 0.0, "giant"),
    ("giant_h3c8_mean", 500, 12000, 8, 8, 3, False, 0.0, "giant"),
]

# run_act_func_experiment.py:13-74,111: the same layer with another logit activation.  name, N, E, F, C, H, concat, p, act
ACT_CASES = [
    ("logsigmoid_h8c8_cat_drop", 120, 700, 33, 8, 8, True, 0.6, "log_sigmoid"),
    ("tanh_h8c8_cat_drop", 120, 700, 33, 8, 8, True, 0.6, "tanh"),
    ("logsigmoid_h1c7_mean", 80, 400, 64, 7, 1, False, 0.0, "log_sigmoid"),
    ("tanh_h4c16_mean", 90, 600, 20, 16, 4, False, 0.0, "tanh"),
    ("tanh_h2c64_cat", 100, 900, 40, 64, 2, True, 0.0, "tanh"),
    # the experiment's third variant, nn.Softmax() on the [E', H] logits: implicit dim = 1, i.e. ACROSS THE HEADS of an edge
    ("softmax_h8c8_cat_drop", 120, 700, 33, 8, 8, True, 0.6, "softmax"),
    ("softmax_h4c16_mean", 90, 600, 20, 16, 4, False, 0.0, "softmax"),
    ("softmax_h2c64_cat", 100, 900, 40, 64, 2, True, 0.0, "softmax"),
    ("softmax_h1c7_mean", 80, 400, 64, 7, 1, False, 0.0, "softmax"),      # one head: every logit becomes 1 (uniform attention)
    ("softmax_hub_h3c8_cat", 400, 6000, 12, 8, 3, True, 0.0, "softmax"),  # hub destination / source (> 512 edges)
]
ACTIVATIONS = {"log_sigmoid": torch.nn.LogSigmoid, "tanh": torch.nn.Tanh, "softmax": torch.nn.Softmax}


def make_graph(name, n, e, special, gen):
    if e == 0:
        return torch.zeros(2, 0, dtype=torch.int64)
    ei = torch.randint(0, n, (2, e), generator=gen)
    if special == "dups":
        ei[:, 10:20] = ei[:, 0:10]                       # duplicate edges are kept (counted twice)
        ei[1, 20:30] = ei[0, 20:30]                      # pre-existing self loops are kept too
        keep = (ei != n - 1).all(dim=0)                  # node n-1 isolated: its row is the appended self loop only
        ei = ei[:, keep]
    if special == "hub":
        ei[1, : e // 2] = 7                              # one destination with ~450 in-edges
    if special == "hubs":
        ei[1, :700] = 7                                  # destination with 700 in-edges, source with 600 out-edges
        ei[0, 700:1300] = 11
    if special == "giant":
        ei[1, :5000] = 3                                 # > 4096 in-edges / out-edges: cut into segments by the kernels
        ei[0, 5000:9500] = 5
    return ei


def pack(layer):
    H = layer.num_heads
    st = lambda ms, attr: torch.stack([getattr(m, attr).detach() for m in ms])
    return dict(W=st(layer.ws, "weight"), bw=st(layer.ws, "bias"),
                a1=st(layer.attentions1, "weight").reshape(H, -1), b1=st(layer.attentions1, "bias").reshape(H),
                a2=st(layer.attentions2, "weight").reshape(H, -1), b2=st(layer.attentions2, "bias").reshape(H),
                bias=layer.bias.detach())


def grads(layer, xg):
    H = layer.num_heads
    st = lambda ms, attr: torch.stack([getattr(m, attr).grad for m in ms])
    return dict(g_x=xg, g_W=st(layer.ws, "weight"), g_bw=st(layer.ws, "bias"),
                g_a1=st(layer.attentions1, "weight").reshape(H, -1), g_b1=st(layer.attentions1, "bias").reshape(H),
                g_a2=st(layer.attentions2, "weight").reshape(H, -1), g_b2=st(layer.attentions2, "bias").reshape(H),
                g_bias=layer.bias.grad)


class patched_dropout:
    """Replace F.dropout by multiplication with a supplied keep-multiplier mask (identity if mask is None)."""

    def __init__(self, mask):
        self.mask = mask

    def __enter__(self):
        self.orig = torch.nn.functional.dropout
        mask = self.mask
        def fake(inp, p=0.5, training=True, inplace=False):
            return inp if mask is None else inp * mask.to(inp.dtype)
        torch.nn.functional.dropout = fake

    def __exit__(self, *a):
        torch.nn.functional.dropout = self.orig


def run_layer(ref_gat, case, act=None):
    name, n, e, f, c, h, concat, p, special = case
    gen = torch.Generator().manual_seed(sum(map(ord, name)))
    torch.manual_seed(sum(map(ord, name)) + 1)
    if act is None:
        layer = ref_gat.GraphAttentionLayer(f, c, num_heads=h, concat=concat, dropout=p)
    else:      # ref_gat is the reference's run_act_func_experiment module here
        layer = ref_gat.GraphAttentionLayerActivationTest(f, c, num_heads=h, concat=concat, dropout=p,
                                                         activation_function=ACTIVATIONS[act]())
    with torch.no_grad():
        layer.bias.uniform_(-0.5, 0.5)                   # zero at init (GAT.py:32-35); made non-trivial here
        if special == "big":
            for m in list(layer.attentions1) + list(layer.attentions2):
                m.weight.mul_(60.0)                      # logits of magnitude ~1e2: softmax stability
    ei = make_graph(name, n, e, special, gen)
    x = torch.randn(n, f, generator=gen)
    ep = ei.shape[1] + n
    mask = None
    if p > 0:
        mask = (torch.rand(ep, h, generator=gen) >= p).float() / (1.0 - p)
    d_out = c * h if concat else c
    gout = torch.randn(n, d_out, generator=gen)
    rec = dict(x=x, edge_index=ei, gout=gout, concat=np.array(concat), p=np.array(p), **pack(layer))
    if mask is not None:
        rec["mask"] = mask
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        lay = layer.to(dt)
        lay.train()
        lay.zero_grad()
        xx = x.detach().clone().to(dt).requires_grad_(True)
        with patched_dropout(mask):
            out = lay(xx, ei)
        out.backward(gout.to(dt))
        rec["out_" + tag] = out.detach()
        for k, v in grads(lay, xx.grad).items():
            rec[k + "_" + tag] = v
    return {k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in rec.items()}


def run_net(ref_net, dataset, f, n, e, graphs, seed, classes=7):
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    net = ref_net.GATNet("GAT", dataset, f)
    x = torch.rand(n, f, generator=gen)
    if graphs:
        per = n // graphs
        ei = torch.cat([torch.randint(0, per, (2, e // graphs), generator=gen) + b * per for b in range(graphs)], 1)
        batch = torch.arange(n) // per
        y = torch.randint(0, 10, (graphs,), generator=gen)
    else:
        ei, batch = torch.randint(0, n, (2, e), generator=gen), None
        y = torch.randint(0, classes, (n,), generator=gen)
    rec = dict(x=x, edge_index=ei, y=y)
    if batch is not None:
        rec["batch"] = batch
    for k, v in net.state_dict().items():
        rec["param:" + k] = v.clone()
    net.eval()                                           # dropout off (feature dropout GATNet.py:78 and GAT.py:61)
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = net.to(dt)
        m.zero_grad()
        out = m(SimpleNamespace(x=x.to(dt), edge_index=ei, batch=batch))
        loss = torch.nn.functional.nll_loss(out, y)
        loss.backward()
        rec["out_" + tag] = out.detach()
        rec["loss_" + tag] = loss.detach()
        for k, p_ in m.named_parameters():
            rec["grad:" + k + "_" + tag] = p_.grad.clone()
    return {k: (v.detach().numpy() if torch.is_tensor(v) else v) for k, v in rec.items()}


def main():
    ref_gat, ref_net = ref_loader.load()
    for case in ([] if "--only-act-softmax" in sys.argv else HUBREF_CASES):
        rec = run_layer(ref_gat, case)
        np.savez_compressed(os.path.join(HERE, f"hubref_{case[0]}.npz"), **rec)
        print("hubref", case[0], {k: v.shape for k, v in rec.items() if k in ("x", "edge_index", "out_f32")})
    if "--only-hubref" in sys.argv:
        return
    for case in ([] if "--only-act-softmax" in sys.argv else LAYER_CASES):
        rec = run_layer(ref_gat, case)
        np.savez_compressed(os.path.join(HERE, f"layer_{case[0]}.npz"), **rec)
        print("layer", case[0], {k: v.shape for k, v in rec.items() if k in ("x", "edge_index", "out_f32")})
    ref_act = ref_loader.load_act_experiment()
    for (name, n, e, f, c, h, concat, p, act) in ACT_CASES:
        if "--only-act-softmax" in sys.argv and act != "softmax":
            continue
        rec = run_layer(ref_act, (name, n, e, f, c, h, concat, p, "hubs" if "_hub_" in name else ""), act=act)
        rec["activation"] = np.array(act)
        np.savez_compressed(os.path.join(HERE, f"actlayer_{name}.npz"), **rec)
        print("actlayer", name, act)
    if "--only-act-softmax" in sys.argv:
        return
    np.savez_compressed(os.path.join(HERE, "net_cora.npz"), **run_net(ref_net, "Cora", 37, 150, 700, 0, 11))
    np.savez_compressed(os.path.join(HERE, "net_cifar_f3.npz"), **run_net(ref_net, "CIFAR10", 3, 160, 1280, 8, 12))
    np.savez_compressed(os.path.join(HERE, "net_pubmed.npz"), **run_net(ref_net, "Pubmed", 21, 130, 600, 0, 13, classes=3))
    print("done")


if __name__ == "__main__":
    main()
