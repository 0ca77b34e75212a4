"""Drop-in `GATNet` (reference: GATNet.py:13-87), GAT branch on the B200-native layer.

Constructor table and forward glue follow GATNet.py:17-37 / :60-87 exactly (same attribute names => same
state_dict keys).  The ELU between two GAT layers is fused into the layer boundary (the producer keeps its
pre-activation output, the consumer applies ELU while loading its operand, the producer's backward applies ELU' — see
include/b200gat.h B200GAT_ACT_*) whenever no feature dropout sits between them; the other glue ops (feature dropout,
per-graph mean readout, lin1/lin2, log_softmax) are PyTorch CUDA ops.
The GCN branch (GATNet.py:38-58) is a third-party comparison baseline (torch_geometric.nn.GCNConv) and is only
available when torch_geometric is installed.
"""
import ctypes
import os

import torch
import torch.nn.functional as F

from . import _abi
from .gat import GraphAttentionLayer, _call, _workspace

# dataset -> (conv2 out_channels, conv2 heads, conv2 concat, dropout)           GATNet.py:17-37
_GAT_TABLE = {
    "CIFAR10": (8, 8, True, 0.0),
    "Cora": (7, 1, False, 0.6),
    "Citeseer": (6, 1, False, 0.6),
    "Pubmed": (3, 8, False, 0.6),
    "AmazonComp": (10, 8, False, 0.6),
    "AmazonPhotos": (8, 8, False, 0.6),
}
_GCN_OUT = {"CIFAR10": 64, "Cora": 7, "Citeseer": 6, "Pubmed": 3, "AmazonComp": 10, "AmazonPhotos": 8}


def segment_mean(x, batch, num_graphs=None):
    """torch_scatter.scatter_mean(x, batch, dim=0) (GATNet.py:73): per-graph mean, empty groups -> 0.  `num_graphs` (PyG
    batches carry it as data.num_graphs) avoids the host synchronisation of batch.max().item()."""
    if num_graphs is not None:
        groups = int(num_graphs)
    else:
        groups = int(batch.max().item()) + 1 if batch.numel() else 0
    total = torch.zeros((groups, x.shape[1]), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    count = torch.zeros(groups, dtype=x.dtype, device=x.device).index_add_(
        0, batch, torch.ones(batch.numel(), dtype=x.dtype, device=x.device)).clamp_(min=1)
    return total / count.unsqueeze(1)


class ReadoutHeadFunction(torch.autograd.Function):
    """GATNet.py:72-75 fused (b200gat_readout_fwd / _bwd): log_softmax(lin2(relu(lin1(scatter_mean(act(x), batch))))).
    x [N, F] CUDA float32; act_in: x is the PRE-activation output of the last GAT layer (produced with act_out) and ELU is
    applied while it is loaded — the gradient returned for x is then the one w.r.t. ELU(x), as that layer expects."""

    @staticmethod
    def forward(ctx, x, batch, w1, b1, w2, b2, num_graphs, act_in):
        lib = _abi.lib()
        dev = x.device
        x, batch = x.contiguous(), batch.contiguous()
        w1, b1, w2, b2 = (t.contiguous() for t in (w1, b1, w2, b2))
        n, f = x.shape
        g, hd, k = int(num_graphs), w1.shape[0], w2.shape[0]
        f32 = dict(dtype=torch.float32, device=dev)
        pooled, counts = torch.empty((g, f), **f32), torch.empty(g, **f32)
        hid, logp = torch.empty((g, hd), **f32), torch.empty((g, k), **f32)
        status = torch.empty(1, dtype=torch.int32, device=dev)
        geom = _abi.ReadoutGeom(n, g, f, hd, k)
        with torch.cuda.device(dev):
            a = _abi.ReadoutFwdArgs(geom, x.data_ptr(), x.stride(0) if n else f, batch.data_ptr(), w1.data_ptr(), b1.data_ptr(),
                                    w2.data_ptr(), b2.data_ptr(), pooled.data_ptr(), counts.data_ptr(), hid.data_ptr(),
                                    logp.data_ptr(), status.data_ptr(), _abi.ACT_ELU if act_in else _abi.ACT_NONE)
            _call("b200gat_readout_fwd", lib.b200gat_readout_fwd, a, torch.cuda.current_stream(dev).cuda_stream, (f, hd, k))
        ctx.save_for_backward(batch, w1, w2, pooled, counts, hid, logp)
        ctx.dims = (n, g, f, hd, k)
        ctx.status = status            # nodes whose graph id is outside [0, num_graphs): readable lazily (status.item())
        return logp

    @staticmethod
    def backward(ctx, g_logp):
        batch, w1, w2, pooled, counts, hid, logp = ctx.saved_tensors
        n, g, f, hd, k = ctx.dims
        lib = _abi.lib()
        dev = g_logp.device
        g_logp = g_logp.contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        g_x = torch.empty((n, f), **f32) if ctx.needs_input_grad[0] else None
        g_w1, g_b1, g_w2, g_b2 = torch.empty((hd, f), **f32), torch.empty(hd, **f32), torch.empty((k, hd), **f32), torch.empty(k, **f32)
        ws_bytes = g * (k + hd + f) * 4
        ws = _workspace(ws_bytes, dev)
        with torch.cuda.device(dev):
            a = _abi.ReadoutBwdArgs(_abi.ReadoutGeom(n, g, f, hd, k), batch.data_ptr(), w1.data_ptr(), w2.data_ptr(),
                                    pooled.data_ptr(), counts.data_ptr(), hid.data_ptr(), logp.data_ptr(), g_logp.data_ptr(),
                                    g_x.data_ptr() if g_x is not None else None, f, g_w1.data_ptr(), g_b1.data_ptr(),
                                    g_w2.data_ptr(), g_b2.data_ptr(), ws.data_ptr(), ws_bytes)
            _call("b200gat_readout_bwd", lib.b200gat_readout_bwd, a, torch.cuda.current_stream(dev).cuda_stream, (f, hd, k))
        return g_x, None, g_w1, g_b1, g_w2, g_b2, None, None


def readout_head(x, batch, lin1, lin2, num_graphs=None, act_in=False):
    """GATNet.py:72-75 on the fused kernels.  num_graphs (PyG batches carry it as data.num_graphs) avoids the host
    synchronisation of batch.max().item()."""
    groups = int(num_graphs) if num_graphs is not None else (int(batch.max().item()) + 1 if batch.numel() else 0)
    return ReadoutHeadFunction.apply(x, batch, lin1.weight, lin1.bias, lin2.weight, lin2.bias, groups, bool(act_in))


class GATNet(torch.nn.Module):
    """GATNet.py:13 — GATNet(model_name, dataset_name, num_features); forward(data) with data.x, data.edge_index
    (and data.batch for CIFAR10)."""

    def __init__(self, model_name, dataset_name, num_features):
        super().__init__()
        self.dataset_name = dataset_name
        self.model_name = model_name
        if model_name == "GAT":
            if dataset_name in _GAT_TABLE:
                out2, heads2, concat2, p = _GAT_TABLE[dataset_name]
                self.conv1 = GraphAttentionLayer(num_features, 8, num_heads=8, concat=True, dropout=p)
                self.conv2 = GraphAttentionLayer(64, out2, num_heads=heads2, concat=concat2, dropout=p)
                if dataset_name == "CIFAR10":
                    self.lin1 = torch.nn.Linear(64, 64)
                    self.lin2 = torch.nn.Linear(64, 10)
        elif model_name == "GCN":
            try:
                from torch_geometric.nn import GCNConv
            except ImportError as e:   # pragma: no cover - PyG is not installable in this image
                raise NotImplementedError(
                    "the GCN comparison branch (GATNet.py:38-58) needs torch_geometric.nn.GCNConv, which is a "
                    "third-party baseline outside the GAT hot path") from e
            if dataset_name in _GCN_OUT:
                self.conv1 = GCNConv(num_features, 64)
                self.conv2 = GCNConv(64, _GCN_OUT[dataset_name])
                if dataset_name == "CIFAR10":
                    self.lin1 = torch.nn.Linear(64, 64)
                    self.lin2 = torch.nn.Linear(64, 10)

    def forward(self, data):
        x, edge_index = data.x, data.edge_index
        act = F.relu if self.model_name == "GCN" else F.elu
        if self.dataset_name == "CIFAR10":                     # GATNet.py:62-76
            if self.model_name == "GAT":
                # elu(conv1) is fused into the conv1 -> conv2 boundary, elu(conv2) into the readout's load of x, and
                # scatter_mean -> lin1 -> relu -> lin2 -> log_softmax (GATNet.py:73-75) run as ONE fused op
                x, amax = self.conv1.forward_fused(x, edge_index, act_out=True)
                x = self.conv2.forward_fused(x, edge_index, act_in=True, act_out=True, x_amax=amax,
                                             producer_link=self.conv1.last_link)[0]
                return readout_head(x, data.batch.long(), self.lin1, self.lin2, getattr(data, "num_graphs", None), act_in=True)
            else:
                x = act(self.conv1(x, edge_index))
                x = act(self.conv2(x, edge_index))
            x = segment_mean(x, data.batch, getattr(data, "num_graphs", None))
            x = F.relu(self.lin1(x))
            return F.log_softmax(self.lin2(x), dim=1)
        x = F.dropout(x, p=0.6, training=self.training)        # GATNet.py:78
        if self.model_name == "GAT" and not self.training:     # no feature dropout in between: fuse elu(conv1)
            x, amax = self.conv1.forward_fused(x, edge_index, act_out=True)
            x = self.conv2.forward_fused(x, edge_index, act_in=True, x_amax=amax, producer_link=self.conv1.last_link)[0]
            return F.log_softmax(x, dim=1)
        x = act(self.conv1(x, edge_index))
        x = F.dropout(x, p=0.6, training=self.training)
        x = self.conv2(x, edge_index)
        return F.log_softmax(x, dim=1)


class GATStack(torch.nn.Module):
    """Plain stack of GraphAttentionLayer + ELU between layers: the composition used for the PPI-shaped and
    large-graph BASELINE configs (the reference has no 3-layer model).  spec = [(in, out, heads, concat)]."""

    def __init__(self, spec, dropout=0.0):
        super().__init__()
        self.convs = torch.nn.ModuleList(
            [GraphAttentionLayer(i, o, num_heads=h, concat=c, dropout=dropout) for (i, o, h, c) in spec])

    def forward(self, x, edge_index):
        pending, amax, link = False, None, None      # pending: x is a pre-activation tensor whose ELU the next layer applies
        last = len(self.convs) - 1
        for k, conv in enumerate(self.convs):
            fuse_out = k < last and conv.can_fuse_activation_out() and not os.environ.get("B200GAT_NO_FUSE_ACT")
            x, amax = conv.forward_fused(x, edge_index, act_in=pending, act_out=fuse_out, x_amax=amax if pending else None,
                                         producer_link=link if pending else None)
            pending, link = fuse_out, conv.last_link
            if k < last and not fuse_out:
                x = F.elu(x)
        return x
