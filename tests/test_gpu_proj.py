"""GPU parity of the projection ABI (b200gat_proj_fwd / b200gat_proj_bwd, GAT.py:42-52 and its autograd) called
directly through the C ABI, against float64 matmuls of the same inputs.

The tensor-core path computes an fp32-accurate product from two fp16 planes per operand ("3xFP16 split",
csrc/proj_tc.cu); these cases walk every kernel variant (K-major / MN-major operands, 128- and 256-wide tiles, fused
logit epilogue, split-K atomic epilogue), ragged M / N / K tails, the x_split hand-over from forward to backward and
operands whose magnitudes are far from 1 (gradients ~1e-7, features ~1e4).  Bar: max|a-b| <= 1e-5 max|b| (FP32_TOL).
"""
import ctypes

import pytest
import torch

from util import FP32_TOL

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _nerr(got, want):
    return float((got.double() - want).abs().max() / want.abs().max().clamp_min(1e-300))


def _run(n, f, c, h, x_scale=1.0, g_scale=1.0, keep_split=True, need_gx=True, seed=0):
    from atmlgraphattentionnetworks_b200 import _abi
    from atmlgraphattentionnetworks_b200.gat import _layer_struct, _workspace
    lib = _abi.lib()
    layer = _layer_struct(f, c, h, True)
    cp = layer.c_pad
    dp = h * cp
    gen = torch.Generator(device="cpu").manual_seed(seed)
    x = (torch.randn(n, f, generator=gen) * x_scale).to(DEV)
    w = torch.randn(h, cp, f, generator=gen) / f ** 0.5
    w[:, c:, :] = 0                                   # pad rows of the packed weight are zero (include/b200gat.h)
    w = w.reshape(dp, f).to(DEV)
    vecs = torch.randn(3, h, cp, generator=gen)
    vecs[:, :, c:] = 0
    bw, a1, a2 = (v.reshape(dp).contiguous().to(DEV) for v in vecs)
    b1, b2 = torch.randn(h, generator=gen).to(DEV), torch.randn(h, generator=gen).to(DEV)
    # rows of very different magnitude, as real gradients have
    gt = (torch.randn(n, dp, generator=gen) * g_scale * torch.rand(n, 1, generator=gen).pow(2)).to(DEV)
    stream = torch.cuda.current_stream().cuda_stream
    wh = torch.empty(n, dp, device=DEV)
    s_src, s_dst = torch.empty(n, h, device=DEV), torch.empty(n, h, device=DEV)
    wsb = int(lib.b200gat_proj_fwd_workspace_bytes(ctypes.byref(layer), n))
    ws = _workspace(wsb, DEV)
    sb = int(lib.b200gat_proj_split_bytes(ctypes.byref(layer), n)) if keep_split else 0
    split = _workspace(sb, DEV) if sb else None
    pa = _abi.ProjFwdArgs(layer, n, x.data_ptr(), f, w.data_ptr(), bw.data_ptr(), a1.data_ptr(), a2.data_ptr(),
                          b1.data_ptr(), b2.data_ptr(), wh.data_ptr(), s_src.data_ptr(), s_dst.data_ptr(),
                          ws.data_ptr(), wsb, split.data_ptr() if sb else None, sb)
    _abi.check(lib.b200gat_proj_fwd(ctypes.byref(pa), stream), "proj_fwd")
    gx = torch.empty(n, f, device=DEV) if need_gx else None
    gw = torch.full((dp, f), 7.0, device=DEV)
    wsb2 = int(lib.b200gat_proj_bwd_workspace_bytes(ctypes.byref(layer), n))
    ws2 = _workspace(wsb2, DEV)
    pb = _abi.ProjBwdArgs(layer, n, gt.data_ptr(), x.data_ptr(), f, w.data_ptr(), gx.data_ptr() if need_gx else None, f,
                          gw.data_ptr(), ws2.data_ptr(), wsb2, split.data_ptr() if sb else None, sb)
    _abi.check(lib.b200gat_proj_bwd(ctypes.byref(pb), stream), "proj_bwd")
    torch.cuda.synchronize()
    xd, wd, gd = x.double(), w.double(), gt.double()
    wh_ref = xd @ wd.t() + bw.double()
    whh = wh_ref.view(n, h, cp)
    errs = {
        "wh": _nerr(wh, wh_ref),
        "s_src": _nerr(s_src, (whh * a1.double().view(h, cp)).sum(-1) + b1.double()),
        "s_dst": _nerr(s_dst, (whh * a2.double().view(h, cp)).sum(-1) + b2.double()),
        "g_w": _nerr(gw, gd.t() @ xd),
    }
    if need_gx:
        errs["g_x"] = _nerr(gx, gd @ wd)
    return errs


CASES = [
    # n, f, c, h, kwargs                                      what it reaches
    (512, 64, 32, 4, {}),                                   # BN=128, fused logits (Cp | 128), one k-block
    (1024, 64, 8, 8, {}),                                   # the smoke() geometry: Dp = 64 (half an n-tile), 8 heads per thread
    (3000, 50, 256, 4, {}),                                 # PPI layer 1: K tail 50, Cp = 256 logits across two warps
    (2048, 256, 64, 4, {}),                                 # BN=256, several heads per epilogue thread
    (1000, 1024, 256, 4, {}),                               # PPI layer 2: K = 1024, gX through MN-major W planes
    (1531, 1024, 121, 6, {}),                               # PPI layer 3: Cp = 124 (unfused logits), Dp = 744 (N tail)
    (2708, 1433, 8, 8, {}),                                 # Cora layer 1: K tail 1433, Dp = 64
    (5000, 100, 128, 4, {"need_gx": False}),                # large-graph layer 1: first layer, no gX
    (4097, 512, 47, 4, {}),                                 # large-graph layer 3: Cp = 48, Dp = 192, M tail of 1 row
    (20000, 96, 16, 8, {}),                                 # deep split-K (many nodes, small tile count)
    (3000, 72, 40, 3, {"keep_split": False}),               # split recomputed in backward (x_split = NULL)
    (3000, 200, 64, 2, {"x_scale": 3e4, "g_scale": 1e-7}),  # operand magnitudes far from 1: the power-of-two scales
    (3000, 200, 64, 2, {"x_scale": 1e-12, "g_scale": 1e9}),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"n{c[0]}_f{c[1]}_c{c[2]}_h{c[3]}" + "".join(f"_{k}" for k in c[4]))
def test_projection_matches_float64(case):
    n, f, c, h, kw = case
    errs = _run(n, f, c, h, **kw)
    for k, e in errs.items():
        assert e <= FP32_TOL, (k, e, errs)


def test_projection_zero_and_constant_operands():
    """all-zero gradient (scale falls back to 1) and a constant tensor (hi exact, lo zero)."""
    errs = _run(1024, 64, 32, 4, g_scale=0.0)
    assert errs["g_w"] == 0.0 and errs["g_x"] == 0.0, errs
    assert errs["wh"] <= FP32_TOL
