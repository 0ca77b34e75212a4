"""Stand-in for torch_geometric.nn: MessagePassing (aggr='add' only) + inert GCNConv/GATConv names."""
import inspect
import torch


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2):
        super().__init__()
        if aggr != "add" or flow != "source_to_target":
            raise NotImplementedError("stand-in implements aggr='add', flow='source_to_target' only")
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        # PyG __collect__: `foo_j` lifts kwargs['foo'] (tuple element 0) by edge_index[0] (source),
        # `foo_i` lifts (tuple element 1) by edge_index[1] (target); `index` is edge_index[1].
        src, dst = edge_index[0], edge_index[1]
        dim = self.node_dim
        call, dim_size = {}, None
        for name in inspect.signature(self.message).parameters:
            if name == "index":
                call[name] = dst
            elif name.endswith("_j") or name.endswith("_i"):
                val = kwargs[name[:-2]]
                which = 0 if name.endswith("_j") else 1
                if isinstance(val, (tuple, list)):
                    val = val[which]
                call[name] = val.index_select(dim, src if which == 0 else dst)
                if which == 1 or dim_size is None:
                    dim_size = val.size(dim)
            else:
                call[name] = kwargs[name]
        msg = self.message(**call)
        out = torch.zeros((dim_size,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        return out.index_add(0, dst, msg)  # scatter-sum over targets, dim_size = N

    def message(self, x_j):  # pragma: no cover - overridden by the reference layer
        return x_j


class _Absent(torch.nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("torch_geometric is not installed; this name is an inert placeholder")


class GCNConv(_Absent):
    pass


class GATConv(_Absent):
    pass
