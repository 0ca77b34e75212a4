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
ACTIVATIONS = {"log_sigmoid": torch.nn.LogSigmoid, "tanh": torch.nn.Tanh}
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
