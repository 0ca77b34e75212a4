"""torch-CPU port of the reference GAT layer and GATNet glue (TEST INFRASTRUCTURE / CPU baseline only).

This restates, op for op, what /root/reference/GAT.py issues through torch + PyG on a CPU (materialised
per-edge tensors, autograd for the backward) WITHOUT importing torch_geometric, so that it can travel to
the GPU box (where /root/reference does not exist) and serve as
  * the checker for the CUDA path in tests/ and __graft_entry__.smoke(), and
  * the "port" CPU baseline in bench.py (`cpu_baseline`, `--impl reference`).
It is pinned against the unmodified reference by tests/test_oracle_cpu.py (direct comparison when
/root/reference is present) and by tests/golden/*.npz (fixtures produced by the reference itself).

Reference lines followed:
  parameters / init order ........ GAT.py:8-35
  self loops ..................... GAT.py:38  -> [PyG] add_self_loops (append, no dedup)
  per-head Linear x3 + stacking .. GAT.py:42-52
  gather / message / aggregate ... GAT.py:53-67 -> [PyG] propagate: _j lifted by edge_index[0], _i by edge_index[1],
                                   utils.softmax (max-subtracted, +1e-16), scatter-sum over targets
  GATNet glue .................... GATNet.py:13-37 (constructor table), :60-87 (forward)
"""
import torch
import torch.nn.functional as F


def segment_softmax(val, index, num_nodes):
    """[PyG] utils.softmax(src, index): grouped by target, per trailing column (GAT.py:60)."""
    idx = index.view(-1, 1).expand_as(val)
    seg_max = torch.full((num_nodes, val.shape[1]), float("-inf"), dtype=val.dtype)
    seg_max = seg_max.scatter_reduce(0, idx, val.detach(), reduce="amax", include_self=True)
    ex = (val - seg_max.index_select(0, index)).exp()
    seg_sum = torch.zeros((num_nodes, val.shape[1]), dtype=val.dtype).scatter_add(0, idx, ex)
    return ex / (seg_sum.index_select(0, index) + 1e-16)


class PortGraphAttentionLayer(torch.nn.Module):
    """Same constructor, parameter registration order and state_dict keys as GAT.py:8-35."""

    def __init__(self, input_channels, output_channels, num_heads=1, concat=False, dropout=0.6, activation_function=None):
        super().__init__()
        # run_act_func_experiment.py:15,37: the same layer with another logit activation (default GAT.py:30)
        self.attention_relu = activation_function if activation_function is not None else torch.nn.LeakyReLU(negative_slope=0.2)
        self.input_channels, self.output_channels = input_channels, output_channels
        self.num_heads, self.dropout_val, self.concat = num_heads, dropout, concat
        self.ws = torch.nn.ModuleList()
        self.attentions1 = torch.nn.ModuleList()
        self.attentions2 = torch.nn.ModuleList()
        for _ in range(num_heads):  # RNG consumption order: Linear x3 default init, then xavier x3 (GAT.py:19-25)
            w = torch.nn.Linear(input_channels, output_channels)
            a1 = torch.nn.Linear(output_channels, 1)
            a2 = torch.nn.Linear(output_channels, 1)
            for lin in (w, a1, a2):
                torch.nn.init.xavier_uniform_(lin.weight)
            self.ws.append(w)
            self.attentions1.append(a1)
            self.attentions2.append(a2)
        self.bias = torch.nn.Parameter(torch.zeros(output_channels * num_heads if concat else output_channels))
        self.mask_hook = None  # optional callable (E', H) -> keep-multiplier tensor, replaces F.dropout (parity tests)

    def forward(self, x, edge_index):
        n = x.size(0)
        loops = torch.arange(n, dtype=edge_index.dtype).unsqueeze(0).repeat(2, 1)
        full = torch.cat([edge_index, loops], dim=1)                       # GAT.py:38
        src, dst = full[0], full[1]
        wh, s_src, s_dst = [], [], []
        for h in range(self.num_heads):                                     # GAT.py:42-48
            t = self.ws[h](x)
            wh.append(t)
            s_src.append(self.attentions1[h](t))
            s_dst.append(self.attentions2[h](t))
        wh = torch.stack(wh).transpose(0, 1)                                # [N,H,C]   GAT.py:49-50
        s_src = torch.stack(s_src).squeeze(-1).T                            # [N,H]     GAT.py:51
        s_dst = torch.stack(s_dst).squeeze(-1).T                            # [N,H]     GAT.py:52
        x_j = wh.index_select(0, src)                                       # [E',H,C]  propagate/_collect
        z = s_dst.index_select(0, dst) + s_src.index_select(0, src)         # GAT.py:57
        e = self.attention_relu(z)                                          # GAT.py:58,30 / run_act_func_experiment.py:62
        alpha = segment_softmax(e, dst, n)                                  # GAT.py:60
        if self.mask_hook is not None:
            alpha = alpha * self.mask_hook(alpha.shape).to(alpha.dtype)
        else:
            alpha = F.dropout(alpha, p=self.dropout_val, training=self.training)   # GAT.py:61
        msg = x_j * alpha.unsqueeze(-1)                                     # GAT.py:62
        msg = msg.reshape(msg.shape[0], -1) if self.concat else msg.mean(dim=1)    # GAT.py:63-66
        out = torch.zeros((n, msg.shape[1]), dtype=msg.dtype).index_add(0, dst, msg)  # aggr='add' GAT.py:9
        return out + self.bias                                              # GAT.py:54


_GAT_TABLE = {  # dataset -> ((heads1, concat1, p1), (out2, heads2, concat2, p2))     GATNet.py:17-37
    "CIFAR10": ((8, True, 0.0), (8, 8, True, 0.0)),
    "Cora": ((8, True, 0.6), (7, 1, False, 0.6)),
    "Citeseer": ((8, True, 0.6), (6, 1, False, 0.6)),
    "Pubmed": ((8, True, 0.6), (3, 8, False, 0.6)),
    "AmazonComp": ((8, True, 0.6), (10, 8, False, 0.6)),
    "AmazonPhotos": ((8, True, 0.6), (8, 8, False, 0.6)),
}


def segment_mean(x, batch):
    """torch_scatter.scatter_mean(x, batch, dim=0) (GATNet.py:73): empty groups -> 0."""
    g = int(batch.max()) + 1 if batch.numel() else 0
    total = torch.zeros((g, x.shape[1]), dtype=x.dtype).index_add(0, batch, x)
    cnt = torch.zeros(g, dtype=x.dtype).index_add(0, batch, torch.ones(batch.numel(), dtype=x.dtype)).clamp(min=1)
    return total / cnt.unsqueeze(1)


class PortGATNet(torch.nn.Module):
    """GAT branch of GATNet.py (same attribute names => same state_dict keys)."""

    def __init__(self, model_name, dataset_name, num_features):
        super().__init__()
        if model_name != "GAT":
            raise NotImplementedError("the GCN comparison branch (GATNet.py:38-58) is out of scope")
        self.dataset_name, self.model_name = dataset_name, model_name
        (h1, c1, p1), (o2, h2, c2, p2) = _GAT_TABLE[dataset_name]
        self.conv1 = PortGraphAttentionLayer(num_features, 8, num_heads=h1, concat=c1, dropout=p1)
        self.conv2 = PortGraphAttentionLayer(64, o2, num_heads=h2, concat=c2, dropout=p2)
        if dataset_name == "CIFAR10":
            self.lin1 = torch.nn.Linear(64, 64)
            self.lin2 = torch.nn.Linear(64, 10)

    def forward(self, data):
        x, edge_index = data.x, data.edge_index
        if self.dataset_name == "CIFAR10":                                  # GATNet.py:62-76
            x = F.elu(self.conv1(x, edge_index))
            x = F.elu(self.conv2(x, edge_index))
            x = segment_mean(x, data.batch)
            x = F.relu(self.lin1(x))
            return F.log_softmax(self.lin2(x), dim=1)
        x = F.dropout(x, p=0.6, training=self.training)                     # GATNet.py:78
        x = F.elu(self.conv1(x, edge_index))
        x = F.dropout(x, p=0.6, training=self.training)
        x = self.conv2(x, edge_index)
        return F.log_softmax(x, dim=1)


class PortStack(torch.nn.Module):
    """Bench-only composition used for the PPI-shaped / large-graph configs: a plain stack of
    GraphAttentionLayer + ELU between layers.  `spec` = [(in, out, heads, concat), ...]."""

    def __init__(self, spec, dropout=0.0):
        super().__init__()
        self.convs = torch.nn.ModuleList(
            [PortGraphAttentionLayer(i, o, num_heads=h, concat=c, dropout=dropout) for (i, o, h, c) in spec])

    def forward(self, x, edge_index):
        for k, conv in enumerate(self.convs):
            x = conv(x, edge_index)
            if k + 1 < len(self.convs):
                x = F.elu(x)
        return x
