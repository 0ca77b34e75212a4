"""Whole-step CUDA-graph capture as a product feature (reference loops: run_inductive.py:74-95 — one
train step + one eval forward per epoch on a static graph — and run_gnn_benchmark.py:58-66 — a NEW batch, hence a new
edge_index, every step).

On Cora- / CIFAR-sized inputs a GAT step is ~40 kernels of a few microseconds each: launched one by one from Python the
step costs 1.7-2.5 ms, replayed as one CUDA graph 0.3-0.4 ms.  `CapturedStep` captures an arbitrary step function over
STATIC input buffers:

    step = CapturedStep(fn, dict(x=x, edge_index=ei, y=y))       # warm-up on a side stream, then capture
    loss = step(x=new_x, edge_index=new_ei, y=new_y)             # copies into the static buffers + ONE graph launch

What makes a per-batch edge_index capturable: the CSR / CSC build of the static `edge_index` buffer runs INSIDE the graph
(graph.build_csr(sync=False): the same kernels, no host read-back, every buffer from the graph's private pool), so each
replay re-sorts whatever edge_index was copied in.  Shapes are static: batches of varying size are padded to a fixed
capacity with `pad_batch` (isolated dummy nodes, dummy self loops on them, labels = ignore_index) — exact, the padding
contributes nothing to the loss or to any gradient.  No `.item()` / host sync happens inside a step; losses are read back by
the caller when it wants them (`CapturedStep.check()` also validates the last edge_index lazily).

`capture_train_step` / `capture_eval_forward` wrap the two loops of the reference's trainers for GATNet-like modules.
"""
from types import SimpleNamespace

import torch

from .graph import GLOBAL_CACHE, build_csr


class CapturedStep:
    """fn(**static_inputs) -> tensor or tuple of tensors, captured once, replayed per call.

    inputs: dict name -> CUDA tensor (example batch; defines the static shapes).  `edge_index` (name configurable) is
    ingested inside the graph; `num_nodes` defaults to inputs['x'].shape[0].  Anything fn allocates lives in the graph's
    private memory pool, so the TMA descriptors and pointers baked into the captured launches stay valid."""

    def __init__(self, fn, inputs, *, edge_index_key="edge_index", num_nodes=None, warmup=3, cache=None, static_graph=False):
        """static_graph: the graph never changes (full-graph training, run_inductive.py:74-95 passes the same
        data.edge_index every epoch): its CSR / CSC is built ONCE, outside the captured graph (with the degree classes and
        the index check of the synchronous build), and replays skip the ingestion; passing a new edge_index then raises."""
        dev = next(iter(inputs.values())).device
        if dev.type != "cuda":
            raise ValueError("CapturedStep needs CUDA tensors")
        self.fn, self.device = fn, dev
        self.static = {k: v.detach().clone() for k, v in inputs.items()}
        self.edge_index_key = edge_index_key if edge_index_key in inputs else None
        self.num_nodes = int(num_nodes if num_nodes is not None else inputs["x"].shape[0]) if self.edge_index_key else None
        self.cache = cache if cache is not None else GLOBAL_CACHE
        self.csr = None
        self.static_graph = bool(static_graph) and self.edge_index_key is not None
        if self.static_graph:
            ei = self.static[self.edge_index_key]
            self.csr = self.cache.put(ei, self.num_nodes, build_csr(ei, self.num_nodes))
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(max(int(warmup), 1)):
                    self._run()
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.outputs = self._run()
        self.replays = 0

    def _run(self):
        if self.edge_index_key is not None and not self.static_graph:
            ei = self.static[self.edge_index_key]
            self.csr = self.cache.put(ei, self.num_nodes, build_csr(ei, self.num_nodes, sync=False))
        return self.fn(**self.static)

    def __call__(self, **inputs):
        """copy the given inputs into the static buffers (same shapes; omitted ones keep their content) and replay"""
        for k, v in inputs.items():
            if self.static_graph and k == self.edge_index_key:
                raise ValueError("this step was captured with static_graph=True: its CSR is built once; capture with "
                                 "static_graph=False to feed a new edge_index per step")
            dst = self.static[k]
            if v.shape != dst.shape or v.dtype != dst.dtype:
                raise ValueError(f"{k}: expected {tuple(dst.shape)} {dst.dtype}, got {tuple(v.shape)} {v.dtype} "
                                 "(pad variable-size batches to the captured capacity with capture.pad_batch)")
            dst.copy_(v, non_blocking=True)
        self.graph.replay()
        self.replays += 1
        return self.outputs

    def check(self):
        """one device->host read: raise IndexError if the LAST edge_index held out-of-range indices"""
        if self.csr is not None and not self.static_graph:
            self.csr.check()
        return self


def pad_batch(data, num_nodes, num_edges, num_graphs=None, ignore_index=-100):
    """Pad a (collated) batch to a fixed capacity so that batches of different size share one captured graph.
    Adds isolated dummy nodes (zero features) up to `num_nodes`, dummy edges up to `num_edges` as self loops spread over
    the dummy nodes, and — for graph-level tasks — puts the dummy nodes into dummy graphs up to `num_graphs` whose labels
    are `ignore_index` (F.nll_loss skips them); node-level labels of dummy nodes are `ignore_index` too.  The real
    nodes' outputs, the loss and every gradient are unchanged.  Needs at least one dummy node when edges must be padded."""
    x, ei = data.x, data.edge_index
    n, e = x.shape[0], ei.shape[1]
    if n > num_nodes or e > num_edges:
        raise ValueError(f"batch ({n} nodes, {e} edges) exceeds the capacity ({num_nodes}, {num_edges})")
    nd, ed = num_nodes - n, num_edges - e
    if ed and not nd:
        raise ValueError("padding edges needs at least one dummy node: raise the node capacity")
    dev = x.device
    out = SimpleNamespace(**vars(data))
    out.x = torch.cat([x, x.new_zeros((nd, x.shape[1]))]) if nd else x
    if ed:
        loops = n + torch.arange(ed, device=dev, dtype=ei.dtype) % nd
        out.edge_index = torch.cat([ei, torch.stack([loops, loops])], dim=1)
    graph_level = getattr(data, "batch", None) is not None and data.y.shape[0] != n
    if getattr(data, "batch", None) is not None:
        g = int(data.num_graphs)
        g_cap = int(num_graphs) if num_graphs is not None else g + (1 if nd else 0)
        if g_cap < g + (1 if nd else 0):
            raise ValueError("num_graphs capacity too small for the dummy graph")
        out.batch = torch.cat([data.batch, torch.full((nd,), g, dtype=data.batch.dtype, device=dev)]) if nd else data.batch
        out.num_graphs = g_cap
        if graph_level:
            out.y = torch.cat([data.y, torch.full((g_cap - g,) + tuple(data.y.shape[1:]), ignore_index, dtype=data.y.dtype, device=dev)])
    if not graph_level and nd:
        out.y = torch.cat([data.y, torch.full((nd,) + tuple(data.y.shape[1:]), ignore_index, dtype=data.y.dtype, device=dev)])
    return out


def _data_inputs(data):
    keys = [k for k in ("x", "edge_index", "y", "batch") if getattr(data, k, None) is not None]
    return {k: getattr(data, k) for k in keys}


def capture_train_step(model, optimizer, loss_fn, data, warmup=3, static_graph=False):
    """One train step of run_inductive.py:75-85 / run_gnn_benchmark.py:60-66 — zero_grad, forward, loss, backward,
    optimizer step — as ONE graph launch.  model(data_like) takes a namespace with x / edge_index (/ batch / num_graphs);
    `optimizer` must be capturable (torch.optim.Adam(..., capturable=True)).  -> CapturedStep; call it with the fields of
    each new batch (x=..., edge_index=..., y=..., batch=...); returns the static loss tensor."""
    for group in optimizer.param_groups:
        if not group.get("capturable", False):
            raise ValueError("capture_train_step needs a capturable optimizer: torch.optim.Adam(..., capturable=True)")
    num_graphs = getattr(data, "num_graphs", None)

    def fn(**t):
        optimizer.zero_grad(set_to_none=True)
        out = model(SimpleNamespace(x=t["x"], edge_index=t["edge_index"], batch=t.get("batch"), num_graphs=num_graphs))
        loss = loss_fn(out, t["y"])
        loss.backward()
        optimizer.step()
        return loss.detach()
    return CapturedStep(fn, _data_inputs(data), warmup=warmup, static_graph=static_graph)


def capture_eval_forward(model, data, warmup=2, static_graph=False):
    """The per-epoch evaluation forward of run_inductive.py:87-95 as one graph launch.  The module must be in eval() mode
    when this is called; -> CapturedStep returning the static output tensor."""
    if model.training:
        raise ValueError("call model.eval() before capturing the evaluation forward")
    num_graphs = getattr(data, "num_graphs", None)

    def fn(**t):
        with torch.no_grad():
            return model(SimpleNamespace(x=t["x"], edge_index=t["edge_index"], batch=t.get("batch"), num_graphs=num_graphs))
    inputs = {k: v for k, v in _data_inputs(data).items() if k != "y"}
    return CapturedStep(fn, inputs, warmup=warmup, static_graph=static_graph)
