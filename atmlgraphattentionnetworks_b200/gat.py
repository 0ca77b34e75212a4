"""Drop-in `GraphAttentionLayer` (reference: GAT.py:8-67) on the B200-native kernels.

Constructor, forward signature, parameter registration order (hence seed-identical init and state_dict keys:
`bias`, `ws.{h}.weight/bias`, `attentions1.{h}.weight/bias`, `attentions2.{h}.weight/bias`) follow GAT.py:8-35.
forward(x, edge_index) (GAT.py:37) runs
    csr_build (cached)  ->  proj_fwd  ->  edge_fwd           and in backward       edge_bwd  ->  proj_bwd
through the C ABI in include/b200gat.h.  CUDA tensors only — there is no CPU fallback.
"""
import ctypes

import torch

from . import _abi
from .graph import GLOBAL_CACHE, GraphCSR

NEGATIVE_SLOPE = 0.2   # GAT.py:30


def _ptr(t):
    return None if t is None else t.data_ptr()


def _layer_struct(f_in, c, h, concat, logit_activation=0, negative_slope=NEGATIVE_SLOPE):
    return _abi.Layer(int(f_in), int(c), int(h), (int(c) + 3) // 4 * 4, 1 if concat else 0, float(negative_slope),
                      int(logit_activation), 0)


def _call(name, fn, args, stream, tag):
    """One ABI call; when _abi.timing is a list, bracket it with CUDA events on the launching stream."""
    if _abi.timing is None:
        _abi.check(fn(ctypes.byref(args), stream), name)
        return
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    _abi.check(fn(ctypes.byref(args), stream), name)
    end.record()
    _abi.timing.append((name, tag, start, end))


class _timed:
    """`with _timed("nccl_all_gather_wh", tag):` — when _abi.timing is a list, bracket a region of the current stream (a
    collective, a barrier) with CUDA events exactly like an ABI call, so bench.py can name each exchange's exposed time."""

    def __init__(self, name, tag):
        self.name, self.tag = name, tag

    def __enter__(self):
        if _abi.timing is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if _abi.timing is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            _abi.timing.append((self.name, self.tag, self.start, end))
        return False


# OPT-IN (B200GAT_OVERLAP_GW=1).  Measured on one B200: PPI-shaped step 5.30 -> 5.20 ms (the gW GEMM of layer k + 1 overlaps
# layer k's edge backward), but the 2.4 M-node graph's step went 177 -> 234 ms — the GB-sized buffers that now live across two
# streams defeat the caching allocator's reuse.  Off by default.
_OVERLAP_GW = __import__("os").environ.get("B200GAT_OVERLAP_GW", "0") == "1"
# The producing layer's prep pass fused into this layer's gX GEMM epilogue (b200gat_proj_bwd_args.fuse_prep, BoundaryLink) is
# OPT-IN (B200GAT_FUSE_PREP=1).  Measured on one B200 (PPI-shaped layers): edge_bwd 0.66 -> 0.50 ms per layer as intended, but
# the gX GEMM 0.64 -> 1.24 ms — the GEMM's epilogue warps sit on the tensor pipe's critical path (the MMA warp stalls as soon as
# both TMEM buffers are full), and the extra tile of `out` they must pull in (even L2-prefetched and double-buffered in
# registers) drops the tensor pipe from 62 % to 16 % active (profiles/r2p_*).  Parity-green (the stack tests run it), kept for a
# design with dedicated loader warps.
_NO_FUSE_PREP = __import__("os").environ.get("B200GAT_FUSE_PREP", "0") != "1"
_side_streams = {}
_join_pending = set()


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def _queue_join(dev, main, side):
    """main waits for side once, when the running backward pass ends (autograd engine final callback — what DDP uses to
    finalise its buckets); outside a backward pass (a direct .backward of the Function in tests) join immediately."""
    key = (dev.type, dev.index)
    if key in _join_pending:
        return

    def join():
        _join_pending.discard(key)
        main.wait_stream(side)
    try:
        _join_pending.add(key)
        torch.autograd.Variable._execution_engine.queue_callback(join)
    except RuntimeError:
        join()


_ws_cache = {}


def _ws_bytes(kind, fn, layer, n):
    """workspace / split sizes depend on (geometry, N) only: one ctypes query per distinct pair instead of one per call"""
    key = (kind, layer.in_channels, layer.out_channels, layer.heads, layer.concat, n)
    v = _ws_cache.get(key)
    if v is None:
        v = _ws_cache[key] = int(fn(ctypes.byref(layer), n))
        if len(_ws_cache) > 4096:
            _ws_cache.clear()
    return v


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def dropout_mask_tensor(p, seed, num_edges, heads):
    """The keep-multipliers the kernels generate for (p, seed) as an [E', H] tensor in ORIGINAL edge order
    (b200gat_dropout_mask) — for tests and debugging; the training path never materialises it."""
    out = torch.empty((int(num_edges), int(heads)), dtype=torch.float32, device=seed.device)
    with torch.cuda.device(seed.device):
        d = _abi.dropout_struct((p, seed))
        _abi.check(_abi.lib().b200gat_dropout_mask(ctypes.byref(d), int(num_edges), int(heads), out.data_ptr(),
                                                   torch.cuda.current_stream(seed.device).cuda_stream), "b200gat_dropout_mask")
    return out


def _split_mask(mask):
    """mask argument of the stage functions -> (tensor or None, (p, seed) or None)"""
    return (None, mask) if isinstance(mask, tuple) else (mask, None)


class BoundaryLink:
    """Hand-over between two consecutive layers whose ELU boundary is fused (producer called with act_out, consumer with
    act_in).  The producer's forward fills in what its prep pass needs; in the backward the CONSUMER's gX GEMM runs that pass
    in its epilogue (b200gat_proj_bwd_args.fuse_prep) and leaves the row records / g_bias here for the producer's
    b200gat_edge_bwd (rowrec_in) — one streaming pass over [N, D] less per layer boundary."""
    __slots__ = ("layer", "n", "d_out", "out", "bias", "s_dst", "rowmax", "rowsum", "rows16", "rowrec", "g_bias", "done")

    def __init__(self):
        self.done = False
        self.layer = None


class GATLayerFunction(torch.autograd.Function):
    """x [N,F], bias [D_out], the layer's persistent packed parameter storage `packed` = (W [Dp,F], bw/a1/a2 [Dp],
    b1/b2 [H]; the per-head parameters are views of it) and the 6H per-head parameters themselves (*params, in the
    order of GraphAttentionLayer._head_parameters: they only route the gradients) -> (out [N, D_out], out_amax).

    fuse = (act_in, act_out, x_amax) describes the layer boundary fusions (include/b200gat.h, B200GAT_ACT_*):
      act_in   the layer consumes ELU(x): x is the PRE-activation output of the previous layer
      act_out  the returned `out` is pre-activation and its consumer applies ELU while loading it; the gradient that
               comes back is therefore d/d ELU(out) and the backward multiplies it by ELU'(out).  Only the internal
               compositions (GATStack, GATNet) set this, and they hand the tensor to nothing but an act_in layer.
      x_amax   int32[1] device word: bound of max|x| from the producing layer (skips one pass over x)
    out_amax (int32[1], bit pattern of max|out|) is returned for the next layer's x_amax.
    """

    @staticmethod
    def forward(ctx, x, bias, graph, geom, mask, fuse, packed, *params):
        # mask: None | [E', H] keep-multiplier tensor (parity tests) | (p, seed int64[2] device tensor): in-kernel Philox
        drop = mask if isinstance(mask, tuple) else None
        mask = None if drop is not None else mask
        w, bw, a1, a2, b1, b2 = packed
        f_in, c, h, concat = geom
        act_in, act_out, x_amax, logit = fuse[:4]
        # bf16 storage of the gathered rows (GraphAttentionLayer.gather_dtype): only the plain configuration has kernels for
        # it (LeakyReLU logits, no dropout / mask); anything else keeps the fp32 rows — never less precise than asked
        rows16 = bool(len(fuse) > 4 and fuse[4]) and drop is None and mask is None and logit[0] == _abi.LOGIT_LEAKY_RELU
        link_in, link_out = (fuse[5], fuse[6]) if len(fuse) > 6 else (None, None)
        lib = _abi.lib()
        dev = x.device
        n = x.shape[0]
        layer = _layer_struct(f_in, c, h, concat, *logit)
        cp = layer.c_pad
        dp = h * cp
        d_out = h * c if concat else c
        heads_mode = (not concat) and h > 1
        x = x.contiguous()
        bias = bias.contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        wh = torch.empty((n, dp), **f32)
        s_src = torch.empty((n, h), **f32)
        s_dst = torch.empty((n, h), **f32)
        out = torch.empty((n, d_out), **f32)
        rowmax = torch.empty((n, h), **f32)
        rowsum = torch.empty((n, h), **f32)
        o_heads = torch.empty((n, dp), **f32) if heads_mode else None
        out_amax = torch.empty(1, dtype=torch.int32, device=dev)      # zeroed by b200gat_edge_fwd itself
        wh16 = torch.empty((n, dp), dtype=torch.bfloat16, device=dev) if rows16 else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            ws_bytes = _ws_bytes("pf", lib.b200gat_proj_fwd_workspace_bytes, layer, n)
            ws = _workspace(ws_bytes, dev)
            # the tensor-core operand split of x is kept for the backward's gW GEMM when a backward will run
            split_bytes = _ws_bytes("ps", lib.b200gat_proj_split_bytes, layer, n) if any(ctx.needs_input_grad) else 0
            x_split = _workspace(split_bytes, dev) if split_bytes else None
            pa = _abi.ProjFwdArgs(layer, n, x.data_ptr(), x.stride(0) if n else f_in, w.data_ptr(), bw.data_ptr(),
                                  a1.data_ptr(), a2.data_ptr(), b1.data_ptr(), b2.data_ptr(), wh.data_ptr(),
                                  s_src.data_ptr(), s_dst.data_ptr(), ws.data_ptr(), ws_bytes,
                                  _ptr(x_split), split_bytes, _abi.ACT_ELU if act_in else _abi.ACT_NONE, _ptr(x_amax))
            pa.wh_bf16 = _ptr(wh16)
            _call("b200gat_proj_fwd", lib.b200gat_proj_fwd, pa, stream, geom)
            ea = _abi.EdgeFwdArgs(layer, graph.c_struct(), wh.data_ptr(), s_src.data_ptr(), s_dst.data_ptr(),
                                  bias.data_ptr(), _ptr(mask), out.data_ptr(), d_out, rowmax.data_ptr(),
                                  rowsum.data_ptr(), _ptr(o_heads), out_amax.data_ptr(), _abi.dropout_struct(drop), _ptr(wh16))
            _call("b200gat_edge_fwd", lib.b200gat_edge_fwd, ea, stream, geom)
            _abi.launches += 2
        ctx.graph, ctx.geom, ctx.mask, ctx.act, ctx.logit = graph, geom, mask, (bool(act_in), bool(act_out)), logit
        ctx.drop = drop
        ctx.rows16 = rows16
        ctx.link_in = link_in if (link_in is not None and link_in.layer is not None and link_in.n == n) else None
        ctx.link_out = None
        if link_out is not None and act_out and not heads_mode and any(ctx.needs_input_grad):
            link_out.layer, link_out.n, link_out.d_out, link_out.rows16 = layer, n, d_out, rows16
            # (detached aliases only: `out` itself will own this node — holding it here would tie ctx -> link -> out ->
            #  grad_fn -> ctx into a reference cycle and keep every saved tensor of the step alive until the GC runs)
            link_out.out, link_out.bias, link_out.s_dst, link_out.rowmax, link_out.rowsum = (
                out.detach(), bias.detach(), s_dst, rowmax, rowsum)
            ctx.link_out = link_out
        # the kernels read the packed storage through raw pointers and the per-head Parameters alias it through `.data =`,
        # which does not share version counters: remember the Parameters' versions so that an in-place update between this
        # forward and its backward (optimizer.step, load_state_dict, ...) is an error, as it is in the reference
        ctx.param_versions = tuple(p._version for p in params)
        ctx.params = params
        ctx.grad_store = fuse[7] if len(fuse) > 7 else None
        ctx.save_for_backward(x, w, a1, a2, bias, wh, s_src, s_dst, rowmax, rowsum, out if not heads_mode else o_heads,
                              x_split)
        ctx.mark_non_differentiable(out_amax)
        return out, out_amax

    @staticmethod
    def backward(ctx, gout, _g_amax):
        x, w, a1, a2, bias, wh, s_src, s_dst, rowmax, rowsum, fwd_out, x_split = ctx.saved_tensors
        graph, (f_in, c, h, concat), mask = ctx.graph, ctx.geom, ctx.mask
        act_in, act_out = ctx.act
        if tuple(p._version for p in ctx.params) != ctx.param_versions:
            raise RuntimeError("one of the variables needed for gradient computation has been modified by an inplace "
                               "operation: a GraphAttentionLayer parameter changed between forward and backward")
        lib = _abi.lib()
        dev = x.device
        n = x.shape[0]
        layer = _layer_struct(f_in, c, h, concat, *ctx.logit)
        cp = layer.c_pad
        dp = h * cp
        d_out = h * c if concat else c
        heads_mode = (not concat) and h > 1
        gout = gout.contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        g_t = torch.empty((n, dp), **f32)
        # parameter-gradient outputs: slices of the layer's PERSISTENT gradient arena when it has one (assign_grad_arena: one
        # flat buffer per model => the data-parallel exchange is ONE all-reduce with no per-step bookkeeping) and no parameter
        # already holds a gradient (the arena is overwritten, not accumulated: with an existing .grad — gradient
        # accumulation, a layer used twice — fresh buffers are used and autograd adds them)
        store = ctx.grad_store
        if store is not None and store["flat"].device == dev and all(p.grad is None for p in ctx.params):
            g_w, g_bw, g_a1, g_a2, g_b1, g_b2, g_bias = (store[k] for k in ("g_w", "g_bw", "g_a1", "g_a2", "g_b1", "g_b2", "g_bias"))
            store["used"] = True
        else:
            if store is not None:
                store["used"] = False
            g_bw = torch.empty(dp, **f32)
            g_a1 = torch.empty(dp, **f32)
            g_a2 = torch.empty(dp, **f32)
            g_b1 = torch.empty(h, **f32)
            g_b2 = torch.empty(h, **f32)
            g_bias = torch.empty(d_out, **f32)
            g_w = torch.empty((dp, f_in), **f32)
        need_gx = ctx.needs_input_grad[0]
        g_x = torch.empty((n, f_in), **f32) if need_gx else None
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            ws_bytes = _ws_bytes("eb", lib.b200gat_edge_bwd_workspace_bytes, layer, n)
            ws = _workspace(ws_bytes, dev)
            # gT goes straight into the tensor-core operand format when the projection backward runs there
            gs_bytes = _ws_bytes("es", lib.b200gat_edge_bwd_split_bytes, layer, n)
            g_split = _workspace(gs_bytes, dev) if gs_bytes else None
            # the across-heads softmax logits couple the heads of an edge in the backward: the one variant with a per-edge
            # scratch buffer ([E', H] floats, include/b200gat.h)
            scratch = (_workspace(graph.num_edges * h * 4, dev) if ctx.logit[0] == _abi.LOGIT_HEAD_SOFTMAX else None)
            # the prep pass of THIS layer already ran in the consuming layer's gX GEMM (BoundaryLink): gout is G itself
            lo = ctx.link_out if (ctx.link_out is not None and ctx.link_out.done) else None
            if lo is not None:
                if store is not None and store.get("used"):
                    g_bias.copy_(lo.g_bias)             # keep the arena slice as THE bias gradient
                else:
                    g_bias = lo.g_bias
            ea = _abi.EdgeBwdArgs(layer, graph.c_struct(), gout.data_ptr(), d_out,
                                  None if heads_mode else fwd_out.data_ptr(), d_out,
                                  fwd_out.data_ptr() if heads_mode else None, bias.data_ptr(),
                                  wh.data_ptr(), s_src.data_ptr(), s_dst.data_ptr(), rowmax.data_ptr(),
                                  rowsum.data_ptr(), _ptr(mask), a1.data_ptr(), a2.data_ptr(),
                                  g_t.data_ptr(), g_bw.data_ptr(), g_a1.data_ptr(), g_a2.data_ptr(),
                                  g_b1.data_ptr(), g_b2.data_ptr(), g_bias.data_ptr(), ws.data_ptr(), ws_bytes,
                                  _abi.ACT_ELU if act_out else _abi.ACT_NONE, _ptr(g_split), gs_bytes,
                                  _abi.dropout_struct(ctx.drop), _ptr(scratch), scratch.numel() if scratch is not None else 0,
                                  1 if ctx.rows16 else 0, lo.rowrec.data_ptr() if lo is not None else None)
            _call("b200gat_edge_bwd", lib.b200gat_edge_bwd, ea, stream, ctx.geom)
            ws2_bytes = _ws_bytes("pb", lib.b200gat_proj_bwd_workspace_bytes, layer, n)

            # the PRODUCING layer's prep pass rides in this layer's gX GEMM when the pair qualifies (BoundaryLink)
            li = ctx.link_in
            fuse_prep = None
            if (li is not None and need_gx and gs_bytes and act_in and not li.rows16 and not _NO_FUSE_PREP and
                    lib.b200gat_proj_bwd_can_fuse_prep(ctypes.byref(layer), n, ctypes.byref(li.layer))):
                li.rowrec = torch.empty((n, int(li.layer.heads), 4), **f32)
                li.g_bias = torch.empty(li.d_out, **f32)
                fuse_prep = _abi.EdgeBwdPrepArgs(li.layer, n, None, 0, li.out.data_ptr(), li.d_out, None, li.bias.data_ptr(),
                                                 li.s_dst.data_ptr(), li.rowmax.data_ptr(), li.rowsum.data_ptr(),
                                                 li.rowrec.data_ptr(), None, li.g_bias.data_ptr(), _abi.ACT_ELU, None)

            def proj_bwd(parts, strm):
                ws2 = _workspace(ws2_bytes, dev)
                pb = _abi.ProjBwdArgs(layer, n, g_t.data_ptr(), x.data_ptr(), x.stride(0) if n else f_in, w.data_ptr(),
                                      _ptr(g_x), f_in, g_w.data_ptr(), ws2.data_ptr(), ws2_bytes,
                                      _ptr(x_split), x_split.numel() if x_split is not None else 0,
                                      _abi.ACT_ELU if act_in else _abi.ACT_NONE, _ptr(g_split), gs_bytes, parts,
                                      ctypes.addressof(fuse_prep) if (fuse_prep is not None and parts != _abi.PROJ_BWD_GW) else None)
                _call("b200gat_proj_bwd", lib.b200gat_proj_bwd, pb, strm, ctx.geom)
                return ws2
            # gX feeds the previous layer's backward; gW feeds nobody until the optimizer.  When this layer has a
            # predecessor waiting (need_gx) and the tensor-core path runs (gs_bytes), gW is issued on a side stream so that
            # it overlaps the predecessor's edge backward (memory-bound, while the GEMM is tensor-bound); the streams are
            # joined by an autograd-engine callback at the end of the backward pass.  Not when a parameter already holds a
            # gradient: AccumulateGrad would then add into it on the main stream before the side stream has finished.
            overlap = (need_gx and gs_bytes and _abi.timing is None and _OVERLAP_GW
                       and all(p.grad is None for p in ctx.params))
            if overlap:
                main = torch.cuda.current_stream(dev)
                side = _side_stream(dev)
                ready = torch.cuda.Event()
                ready.record(main)
                proj_bwd(_abi.PROJ_BWD_GX, stream)
                side.wait_event(ready)
                with torch.cuda.stream(side):
                    ws_side = proj_bwd(_abi.PROJ_BWD_GW, side.cuda_stream)
                for t in (g_w, g_t, g_split, x_split, x, ws_side):
                    if t is not None:
                        t.record_stream(side)
                _queue_join(dev, main, side)
            else:
                proj_bwd(0, stream)
            if fuse_prep is not None:
                li.done = True
            _abi.launches += 2
        # per-head gradients are views of the packed buffers, in the order of _head_parameters (autograd takes them as
        # the parameters' .grad without a copy when no gradient is accumulated yet)
        gw3, gbw2, ga12, ga22 = g_w.view(h, cp, f_in), g_bw.view(h, cp), g_a1.view(h, cp), g_a2.view(h, cp)
        per_head = []
        for k in range(h):
            per_head += [gw3[k, :c], gbw2[k, :c], ga12[k:k + 1, :c], g_b1[k:k + 1], ga22[k:k + 1, :c], g_b2[k:k + 1]]
        # (a FRESH alias of g_bias: autograd adopts an incoming gradient as .grad without a copy only if nobody else holds
        #  that tensor object — the arena's dict does hold the slice itself)
        return (g_x, g_bias.view(-1), None, None, None, None, None, *per_head)


_DEFAULT_GATHER_DTYPE = torch.bfloat16 if __import__("os").environ.get("B200GAT_GATHER_DTYPE", "") in ("bf16", "bfloat16") else torch.float32


def grad_arena_numel(layer):
    c, h, f = layer.output_channels, layer.num_heads, layer.input_channels
    dp = h * ((c + 3) // 4 * 4)
    d_out = h * c if layer.concat else c
    return dp * f + 3 * dp + 2 * h + d_out


def assign_grad_arena(module):
    """Give every GraphAttentionLayer under `module` a slice of ONE persistent flat fp32 buffer for its packed parameter
    gradients (g_w [Dp, F], g_bw / g_a1 / g_a2 [Dp], g_b1 / g_b2 [H], g_bias [D_out]).  The backward then writes into the same
    memory every step, the per-head .grad tensors are views of it, and a data-parallel step exchanges the whole model's GAT
    gradients with one all-reduce of the returned tensor (parallel.ArenaExchange).  -> the flat tensor."""
    layers = [m for m in module.modules() if isinstance(m, GraphAttentionLayer)]
    if not layers:
        raise ValueError("no GraphAttentionLayer under this module")
    dev = layers[0].bias.device
    flat = torch.zeros(sum(grad_arena_numel(m) for m in layers), dtype=torch.float32, device=dev)
    off = 0
    for m in layers:
        c, h, f = m.output_channels, m.num_heads, m.input_channels
        dp = h * ((c + 3) // 4 * 4)
        d_out = h * c if m.concat else c
        store = {"flat": flat, "used": False}
        for key, shape in (("g_w", (dp, f)), ("g_bw", (dp,)), ("g_a1", (dp,)), ("g_a2", (dp,)), ("g_b1", (h,)), ("g_b2", (h,)),
                           ("g_bias", (d_out,))):
            numel = 1
            for v in shape:
                numel *= v
            store[key] = flat[off:off + numel].view(shape)
            off += numel
        m._grad_store = store
    return flat


def set_gather_dtype(module, dtype):
    """Switch every GraphAttentionLayer under `module` to fp32 (default) or bf16 storage of the gathered rows."""
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("gather dtype must be torch.float32 or torch.bfloat16")
    for m in module.modules():
        if isinstance(m, GraphAttentionLayer):
            m.gather_dtype = dtype
    return module


class GraphAttentionLayer(torch.nn.Module):
    """GAT.py:8 — GraphAttentionLayer(input_channels, output_channels, num_heads=1, concat=False, dropout=0.6)."""

    def __init__(self, input_channels, output_channels, num_heads=1, concat=False, dropout=0.6):
        super().__init__()
        self.input_channels = input_channels
        self.output_channels = output_channels
        self.num_heads = num_heads
        self.dropout_val = dropout
        self.ws = torch.nn.ModuleList()
        self.attentions1 = torch.nn.ModuleList()
        self.attentions2 = torch.nn.ModuleList()
        for _ in range(num_heads):   # RNG order of GAT.py:19-25: three default Linear inits, then three xavier draws
            head_transform = torch.nn.Linear(input_channels, output_channels)
            attention1 = torch.nn.Linear(output_channels, 1)
            attention2 = torch.nn.Linear(output_channels, 1)
            torch.nn.init.xavier_uniform_(head_transform.weight)
            torch.nn.init.xavier_uniform_(attention1.weight)
            torch.nn.init.xavier_uniform_(attention2.weight)
            self.ws.append(head_transform)
            self.attentions1.append(attention1)
            self.attentions2.append(attention2)
        self.concat = concat
        if not concat:
            self.bias = torch.nn.Parameter(torch.zeros(output_channels))
        else:
            self.bias = torch.nn.Parameter(torch.zeros(output_channels * num_heads))
        self.graph_cache = GLOBAL_CACHE
        self.mask_hook = None   # parity tests: callable (E', H) -> keep-multiplier [E', H] in ORIGINAL edge order
        self.last_link = None   # BoundaryLink of the last forward_fused(act_out=True) call, for the consuming layer
        self._grad_store = None  # persistent gradient arena slices (assign_grad_arena)
        self._store = None      # persistent packed parameter storage (see _packed_storage)
        # the function applied to the edge logits before the softmax: (B200GAT_LOGIT_* code, negative slope); GAT.py:30
        self.logit_activation = (_abi.LOGIT_LEAKY_RELU, NEGATIVE_SLOPE)
        # storage of the rows the edge kernels GATHER (Wh forward, the gradient rows backward): torch.float32 (default:
        # the reference's arithmetic, 1e-5 parity) or torch.bfloat16 (half the gather / NVLink bytes, tolerance 1e-2 —
        # DESIGN.md §4.11; set_gather_dtype(model, torch.bfloat16) switches every layer of a model)
        self.gather_dtype = _DEFAULT_GATHER_DTYPE

    # ---- persistent packed storage: the kernels read ONE [Dp, F] / [Dp] / [H] set of arrays per layer (head-major, rows
    # padded to c_pad with zeros).  The per-head Linear parameters of GAT.py:19-25 (and the state_dict keys that come with
    # them) are VIEWS of that storage, so optimizers and load_state_dict update it in place and no per-step packing or
    # gradient un-packing kernels run.  Anything that re-creates the parameter tensors (.to(), .cuda(), assignment) is
    # detected by the pointer check and the storage is rebuilt from the current values.
    def _head_parameters(self):
        out = []
        for k in range(self.num_heads):
            out += [self.ws[k].weight, self.ws[k].bias, self.attentions1[k].weight, self.attentions1[k].bias,
                    self.attentions2[k].weight, self.attentions2[k].bias]
        return out

    def _packed_storage(self):
        c, h, f = self.output_channels, self.num_heads, self.input_channels
        cp = (c + 3) // 4 * 4
        st = self._store
        ok = st is not None and st[0].device == self.bias.device
        if ok:
            w3, bw2, a12, a22, b1, b2 = st[1]
            for k in range(h):
                if (self.ws[k].weight.data_ptr() != w3[k].data_ptr() or self.ws[k].bias.data_ptr() != bw2[k].data_ptr() or
                        self.attentions1[k].weight.data_ptr() != a12[k].data_ptr() or
                        self.attentions2[k].weight.data_ptr() != a22[k].data_ptr() or
                        self.attentions1[k].bias.data_ptr() != b1[k:].data_ptr() or
                        self.attentions2[k].bias.data_ptr() != b2[k:].data_ptr()):
                    ok = False
                    break
        if not ok:
            dev = self.bias.device
            f32 = dict(dtype=torch.float32, device=dev)
            w3, bw2, a12, a22 = (torch.zeros((h, cp, f), **f32), torch.zeros((h, cp), **f32), torch.zeros((h, cp), **f32),
                                 torch.zeros((h, cp), **f32))
            b1, b2 = torch.zeros(h, **f32), torch.zeros(h, **f32)
            with torch.no_grad():
                for k in range(h):
                    w3[k, :c].copy_(self.ws[k].weight)
                    bw2[k, :c].copy_(self.ws[k].bias)
                    a12[k, :c].copy_(self.attentions1[k].weight[0])
                    a22[k, :c].copy_(self.attentions2[k].weight[0])
                    b1[k:k + 1].copy_(self.attentions1[k].bias)
                    b2[k:k + 1].copy_(self.attentions2[k].bias)
                    self.ws[k].weight.data = w3[k, :c]
                    self.ws[k].bias.data = bw2[k, :c]
                    self.attentions1[k].weight.data = a12[k:k + 1, :c]
                    self.attentions2[k].weight.data = a22[k:k + 1, :c]
                    self.attentions1[k].bias.data = b1[k:k + 1]
                    self.attentions2[k].bias.data = b2[k:k + 1]
            st = (w3, (w3, bw2, a12, a22, b1, b2),
                  (w3.view(h * cp, f), bw2.view(-1), a12.view(-1), a22.view(-1), b1, b2))
            self._store = st
        return st[2]

    # ---- functional packing through autograd (torch.stack / pad): used by the row-partitioned multi-GPU path and tests
    def _packed(self):
        c, h = self.output_channels, self.num_heads
        pad = (c + 3) // 4 * 4 - c
        w = torch.stack([m.weight for m in self.ws])                              # [H, C, F]
        bw = torch.stack([m.bias for m in self.ws])                               # [H, C]
        a1 = torch.cat([m.weight for m in self.attentions1])                      # [H, C]
        a2 = torch.cat([m.weight for m in self.attentions2])
        b1 = torch.cat([m.bias for m in self.attentions1])                        # [H]
        b2 = torch.cat([m.bias for m in self.attentions2])
        if pad:
            w = torch.nn.functional.pad(w, (0, 0, 0, pad))
            bw, a1, a2 = (torch.nn.functional.pad(t, (0, pad)) for t in (bw, a1, a2))
        return w.reshape(h * (c + pad), -1), bw.reshape(-1), a1.reshape(-1), a2.reshape(-1), b1, b2

    def _dropout_mask(self, num_edges, device):
        """GAT.py:61: F.dropout on the [E', H] coefficients (original edge order), scaled 1/(1-p), not renormalised.
        -> None (eval / p == 0), the hook's [E', H] tensor (parity tests: the reference's own masks), or (p, seed): the
        kernels generate the mask themselves (Philox keyed on (seed, edge, head), include/b200gat.h b200gat_dropout) —
        forward and backward regenerate the same bits and no [E', H] tensor is ever allocated.  The two seed words are drawn
        on the device from torch's CUDA generator (torch.manual_seed applies; capturable in a CUDA graph)."""
        if self.mask_hook is not None:
            return self.mask_hook((num_edges, self.num_heads)).to(device=device, dtype=torch.float32).contiguous()
        p = float(self.dropout_val)
        if not self.training or p <= 0.0:
            return None
        seed = torch.randint(-2 ** 63, 2 ** 63 - 1, (2,), dtype=torch.int64, device=device)
        self._last_dropout_seed = seed
        return (min(p, 1.0), seed)

    def forward(self, x, edge_index, graph=None):
        """GAT.py:37 — forward(x, edge_index) -> [N, H*C] (concat) or [N, C]."""
        return self.forward_fused(x, edge_index, graph=graph)[0]

    def can_fuse_activation_out(self):
        """the ELU that follows this layer can be deferred to its consumer (concat-like layers only, edge_bwd contract)"""
        return bool(self.concat) or self.num_heads == 1

    def forward_fused(self, x, edge_index, graph=None, act_in=False, act_out=False, x_amax=None, producer_link=None):
        """Internal composition entry (GATStack / GATNet): -> (out, out_amax).  act_in: x is a pre-activation tensor
        produced by a layer called with act_out, and ELU(x) is what is projected; act_out: return the pre-activation
        output (see GATLayerFunction).  Never hand an act_out tensor to anything but an act_in layer."""
        if act_out and not self.can_fuse_activation_out():
            raise ValueError("act_out needs a concat-like layer")
        if not x.is_cuda:
            raise _abi.B200GatError("GraphAttentionLayer runs on CUDA tensors only (B200-native path, no CPU fallback)")
        if x.dtype != torch.float32:
            raise TypeError(f"x must be float32, got {x.dtype}")
        if x.dim() != 2 or x.shape[1] != self.input_channels:
            raise ValueError(f"x must be [N, {self.input_channels}], got {tuple(x.shape)}")
        n = x.shape[0]
        if graph is None:
            graph = self.graph_cache.get(edge_index, n)
        elif not isinstance(graph, GraphCSR) or graph.num_nodes != n:
            raise ValueError("graph does not match x")
        mask = self._dropout_mask(graph.num_edges, x.device)
        if x.dtype != torch.float32 or self.bias.dtype != torch.float32 or not self.bias.is_cuda:
            raise TypeError("GraphAttentionLayer parameters must be float32 CUDA tensors (move the module with .to(device))")
        packed = self._packed_storage()
        geom = (self.input_channels, self.output_channels, self.num_heads, bool(self.concat))
        # producer_link: the BoundaryLink of the layer that produced x (act_in); self.last_link: this layer's, for its consumer
        self.last_link = BoundaryLink() if act_out else None
        return GATLayerFunction.apply(x, self.bias, graph, geom, mask,
                                      (bool(act_in), bool(act_out), x_amax, tuple(self.logit_activation),
                                       self.gather_dtype == torch.bfloat16, producer_link if act_in else None, self.last_link,
                                       self._grad_store),
                                      packed, *self._head_parameters())


def _logit_code(fn):
    """torch activation module -> (B200GAT_LOGIT_* code, negative slope) — run_act_func_experiment.py:111."""
    if isinstance(fn, torch.nn.LeakyReLU):
        return _abi.LOGIT_LEAKY_RELU, float(fn.negative_slope)
    if isinstance(fn, torch.nn.LogSigmoid):
        return _abi.LOGIT_LOGSIGMOID, NEGATIVE_SLOPE
    if isinstance(fn, torch.nn.Tanh):
        return _abi.LOGIT_TANH, NEGATIVE_SLOPE
    if isinstance(fn, torch.nn.Softmax):
        # run_act_func_experiment.py:111 passes nn.Softmax() (dim=None): on the 2-D [E', H] logit tensor torch's implicit
        # choice is dim=1, i.e. a softmax ACROSS THE HEADS of one edge (uniform attention with one head)
        if fn.dim not in (None, 1, -1):
            raise NotImplementedError("nn.Softmax(dim=0) on the [E', H] logits would normalise over ALL edges of the graph; "
                                      "only the across-heads form the reference's experiment runs (dim None / 1 / -1) is offered")
        return _abi.LOGIT_HEAD_SOFTMAX, NEGATIVE_SLOPE
    raise NotImplementedError(f"logit activation {type(fn).__name__} is not offered by the fused kernels "
                              "(LeakyReLU, LogSigmoid, Tanh, Softmax are)")


class GraphAttentionLayerActivationTest(GraphAttentionLayer):
    """run_act_func_experiment.py:13-74 — the same layer with a configurable logit activation
    (`activation_function`, default LeakyReLU(0.2)): same parameters, init order and state_dict keys."""

    def __init__(self, input_channels, output_channels, num_heads=1, concat=False, dropout=0.6,
                 activation_function=None):
        super().__init__(input_channels, output_channels, num_heads=num_heads, concat=concat, dropout=dropout)
        self.attention_relu = activation_function if activation_function is not None else torch.nn.LeakyReLU(negative_slope=0.2)
        self.logit_activation = _logit_code(self.attention_relu)
