"""Shared helpers for the parity tests."""
import glob
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LAYER_FILES = sorted(glob.glob(os.path.join(GOLDEN, "layer_*.npz")))
NET_FILES = sorted(glob.glob(os.path.join(GOLDEN, "net_*.npz")))
ACT_FILES = sorted(glob.glob(os.path.join(GOLDEN, "actlayer_*.npz")))   # run_act_func_experiment.py layer, other logit activations
# graphs with hub (> 512) and giant (> 4096) degrees, from the unmodified reference: they pin the ORACLE on those shapes
HUBREF_FILES = sorted(glob.glob(os.path.join(GOLDEN, "hubref_*.npz")))
ACTIVATIONS = {"log_sigmoid": torch.nn.LogSigmoid, "tanh": torch.nn.Tanh, "softmax": torch.nn.Softmax}
GRAD_KEYS = ("g_x", "g_W", "g_bw", "g_a1", "g_b1", "g_a2", "g_b2", "g_bias")
FP32_TOL = 1e-5   # north-star: fp32 outputs within 1e-5 relative (max|a-b| <= tol * max|b|)


def case_id(path):
    return os.path.basename(path)[:-4]


def load(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def nerr(a, b):
    """normalised max error max|a-b| / max|b| (the pass criterion of SURVEY.md §8c)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    bmax = float(np.abs(b).max())
    if bmax == 0.0:
        # exactly-zero reference (e.g. attention-parameter gradients of a graph with self loops only, where the
        # softmax is degenerate): the error is judged absolutely, on the scale of the O(1) inputs
        return float(np.abs(a).max())
    return float(np.abs(a - b).max()) / bmax


def port_layer_from_golden(g, dtype=torch.float32):
    """Build oracle.gat_port.PortGraphAttentionLayer holding the fixture's parameters."""
    from oracle.gat_port import PortGraphAttentionLayer
    H, C, F = g["W"].shape
    act = ACTIVATIONS[str(g["activation"])]() if "activation" in g else None
    layer = PortGraphAttentionLayer(F, C, num_heads=H, concat=bool(g["concat"]), dropout=float(g["p"]),
                                    activation_function=act)
    load_packed(layer, g)
    return layer.to(dtype)


def load_packed(layer, g):
    H = g["W"].shape[0]
    with torch.no_grad():
        for h in range(H):
            layer.ws[h].weight.copy_(torch.from_numpy(g["W"][h]))
            layer.ws[h].bias.copy_(torch.from_numpy(g["bw"][h]))
            layer.attentions1[h].weight.copy_(torch.from_numpy(g["a1"][h:h + 1]))
            layer.attentions1[h].bias.copy_(torch.from_numpy(g["b1"][h:h + 1]))
            layer.attentions2[h].weight.copy_(torch.from_numpy(g["a2"][h:h + 1]))
            layer.attentions2[h].bias.copy_(torch.from_numpy(g["b2"][h:h + 1]))
        layer.bias.copy_(torch.from_numpy(g["bias"]))


def packed_grads(layer, xg):
    H = layer.num_heads
    st = lambda ms, attr: torch.stack([getattr(m, attr).grad for m in ms]).detach().cpu().numpy()
    return dict(g_x=xg.detach().cpu().numpy(), g_W=st(layer.ws, "weight"), g_bw=st(layer.ws, "bias"),
                g_a1=st(layer.attentions1, "weight").reshape(H, -1), g_b1=st(layer.attentions1, "bias").reshape(H),
                g_a2=st(layer.attentions2, "weight").reshape(H, -1), g_b2=st(layer.attentions2, "bias").reshape(H),
                g_bias=layer.bias.grad.detach().cpu().numpy())


def attention_term_sums(module):
    """Hooks for an oracle module (call BEFORE its forward): for every attentions1/2[h] Linear under `module` record, after
    the backward, the sum of ABSOLUTE terms of the reductions its gradients are —
        weight.grad[c] = sum_n g_s[n] Wh[n,c]  ->  max_c sum_n |g_s[n]| |Wh[n,c]|        bias.grad = sum_n g_s[n]  ->  sum_n |g_s[n]|
    Deep layers' Wh rows share a large common component (ELU outputs are not centred) while sum_n g_s[n] ~ 0 (softmax shift
    invariance), so these sums cancel by factors of 100-1000: fp32 can only resolve them to a few units of roundoff OF THE
    SUMMED MAGNITUDES, which is what the parity tests allow on top of the 1e-5 bar (the fp32 reference itself sits there).
    -> dict filled in place: {"<prefix>attentions1.<h>.weight": terms, "...bias": terms}"""
    sums, cap = {}, {}
    for name, lin in module.named_modules():
        if ".attentions" not in "." + name or not isinstance(lin, torch.nn.Linear):
            continue

        def fwd(mod, inp, out, name=name):
            cap[name] = inp[0].detach()

        def bwd(mod, gin, gout, name=name):
            g = gout[0].detach().abs()
            sums[name + ".weight"] = float((g * cap.pop(name).abs()).sum(0).max())
            sums[name + ".bias"] = float(g.sum())
        lin.register_forward_hook(fwd)
        lin.register_full_backward_hook(bwd)
    return sums


FP32_SUM_ULPS = 4 * 2.0 ** -24     # 4 units of fp32 roundoff of the summed magnitudes
