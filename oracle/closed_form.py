"""Independent closed-form restatement of one GAT layer, forward AND backward, in numpy float64 over
destination-sorted CSR (TEST INFRASTRUCTURE).

Forward follows SURVEY.md §3(C) (= GAT.py:37-67 + the [PyG] semantics of add_self_loops / propagate /
utils.softmax); backward follows the closed form of SURVEY.md §3(D) — the reference has no backward code, it
differentiates the forward with autograd (run_inductive.py:84).  This file shares no code with gat_port.py
(which goes through autograd), so agreement between the two, and with the fixtures generated from the
unmodified reference, pins the arithmetic the CUDA kernels implement (they use exactly this formulation:
recompute alpha from s_src, s_dst and the row max / sum; Drow = <G, O>).

Parameters are passed packed: W [H,C,F], bw [H,C], a1 [H,C], b1 [H], a2 [H,C], b2 [H], bias [H*C] or [C].
`mask` (optional) is the dropout keep-multiplier [E',H] in ORIGINAL [edges ; loops] order (GAT.py:61).
"""
import numpy as np

from .csr_oracle import csr_oracle

NEG_SLOPE = 0.2  # GAT.py:30


def _seg_reduce(ufunc, vals, rowptr):
    # every row is non-empty (each node has its appended self loop), so reduceat is safe
    return ufunc.reduceat(vals, rowptr[:-1], axis=0)


def forward(x, edge_index, W, bw, a1, b1, a2, b2, bias, concat, mask=None):
    x = np.asarray(x, dtype=np.float64)
    W, bw, a1, b1, a2, b2, bias = (np.asarray(t, dtype=np.float64) for t in (W, bw, a1, b1, a2, b2, bias))
    n = x.shape[0]
    H, C, _ = W.shape
    g = csr_oracle(edge_index, n)
    rowptr, col, eid = g["rowptr"], g["col"], g["eid"]
    deg = np.diff(rowptr)
    row = np.repeat(np.arange(n), deg)
    Wh = np.einsum("nf,hcf->nhc", x, W) + bw[None]                    # GAT.py:43
    s_src = np.einsum("nhc,hc->nh", Wh, a1) + b1[None]                # GAT.py:44 (gathered at the source j)
    s_dst = np.einsum("nhc,hc->nh", Wh, a2) + b2[None]                # GAT.py:45 (gathered at the target i)
    z = s_dst[row] + s_src[col]                                       # GAT.py:57
    e = np.where(z > 0, z, NEG_SLOPE * z)                             # GAT.py:58
    rmax = _seg_reduce(np.maximum, e, rowptr) if n else np.zeros((0, H))
    p = np.exp(e - rmax[row])
    rsum = _seg_reduce(np.add, p, rowptr) if n else np.zeros((0, H))
    alpha = p / (rsum[row] + 1e-16)                                   # GAT.py:60 [PyG]
    m = np.ones_like(alpha) if mask is None else np.asarray(mask, dtype=np.float64)[eid]
    alpha_t = alpha * m                                               # GAT.py:61
    O = np.zeros((n, H, C))
    np.add.at(O, row, alpha_t[:, :, None] * Wh[col])                  # GAT.py:62 + aggr='add'
    out = (O.reshape(n, H * C) if concat else O.mean(axis=1)) + bias  # GAT.py:63-66,54
    cache = dict(g=g, row=row, Wh=Wh, s_src=s_src, s_dst=s_dst, z=z, alpha=alpha, m=m, O=O, x=x, W=W, a1=a1, a2=a2,
                 rmax=rmax, rsum=rsum, concat=concat)
    return out, cache


def backward(gout, cache):
    """gout [N, D_out] -> dict of gradients (x, W, bw, a1, b1, a2, b2, bias)."""
    c = cache
    g, row, Wh, alpha, m, O, x, W, a1, a2, z = (c[k] for k in ("g", "row", "Wh", "alpha", "m", "O", "x", "W", "a1",
                                                                "a2", "z"))
    col = g["col"]
    n, H, C = Wh.shape
    gout = np.asarray(gout, dtype=np.float64)
    G = gout.reshape(n, H, C) if c["concat"] else np.repeat(gout[:, None, :] / H, H, axis=1)
    dalpha = m * np.einsum("ehc,ehc->eh", G[row], Wh[col])
    Drow = np.einsum("nhc,nhc->nh", G, O)
    dz = alpha * (dalpha - Drow[row]) * np.where(z > 0, 1.0, NEG_SLOPE)
    g_s_dst = np.zeros((n, H)); np.add.at(g_s_dst, row, dz)
    g_s_src = np.zeros((n, H)); np.add.at(g_s_src, col, dz)
    gWh = np.zeros((n, H, C)); np.add.at(gWh, col, (alpha * m)[:, :, None] * G[row])
    gT = gWh + g_s_src[:, :, None] * a1[None] + g_s_dst[:, :, None] * a2[None]
    return dict(
        x=np.einsum("nhc,hcf->nf", gT, W), W=np.einsum("nhc,nf->hcf", gT, x), bw=gT.sum(axis=0),
        a1=np.einsum("nh,nhc->hc", g_s_src, Wh), a2=np.einsum("nh,nhc->hc", g_s_dst, Wh),
        b1=g_s_src.sum(axis=0), b2=g_s_dst.sum(axis=0), bias=gout.sum(axis=0),
        # intermediates, exposed so kernel-level tests can check each stage
        Drow=Drow, g_s_src=g_s_src, g_s_dst=g_s_dst, gT=gT)
