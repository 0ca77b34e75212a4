"""Numerical error of the tensor-core projection GEMMs alone (forward X·W^T, backward gX = gT·W, gW = gT^T·X) against f64,
through the product layer on a graph of self loops only (alpha = 1: out = Wh + bias, gWh = gout), for a chosen TMEM chunk
length:   B200GAT_TC_KC=<k-blocks per accumulation chunk> python tools/gemm_error.py
Prints the normalised max error  max|a - ref| / max|ref|  of out, gX and gW (the parity bar is 1e-5)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from atmlgraphattentionnetworks_b200.gat import GraphAttentionLayer   # noqa: E402


def run(n, f, c, h, scale=1.0, seed=0):
    torch.manual_seed(seed)
    dev = torch.device("cuda:0")
    layer = GraphAttentionLayer(f, c, h, concat=True, dropout=0.0).to(dev)
    x = (torch.randn(n, f, device=dev) * scale).requires_grad_(True)
    ei = torch.zeros(2, 0, dtype=torch.long, device=dev)
    out = layer(x, ei)
    g = torch.randn_like(out)
    out.backward(g)
    w = torch.cat([layer.ws[k].weight for k in range(h)], 0).double()
    b = torch.cat([layer.ws[k].bias for k in range(h)], 0).double() + layer.bias.double()
    ref = x.detach().double() @ w.t() + b
    gx = g.double() @ w
    gw = g.double().t() @ x.detach().double()
    gw_got = torch.cat([layer.ws[k].weight.grad for k in range(h)], 0).double()

    def err(a, r):
        return float((a.double() - r).abs().max() / r.abs().max())
    return err(out.detach(), ref), err(x.grad, gx), err(gw_got, gw)


if __name__ == "__main__":
    print("B200GAT_TC_KC =", os.environ.get("B200GAT_TC_KC", "(default)"))
    for (n, f, c, h, s) in ((56944, 1024, 256, 4, 1.0), (56944, 1024, 256, 4, 0.05), (20000, 4096, 256, 4, 1.0), (200000, 512, 128, 4, 1.0),
                            (56944, 1024, 124, 6, 1.0)):
        e = [max(v) for v in zip(*(run(n, f, c, h, s, seed) for seed in range(2)))]
        print(f"  N {n:7d} F {f:5d} {h} x {c:3d} scale {s:4.2f}:  out {e[0]:.2e}   gX {e[1]:.2e}   gW {e[2]:.2e}")
