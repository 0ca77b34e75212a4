"""Destination-row partitioned execution of ONE large graph over P GPUs (BASELINE config 5).

Rank r owns the contiguous node block [r*B, min((r+1)*B, N)), B = ceil(N/P):
  forward  per layer : project own rows -> ALL-GATHER Wh (+ s_src) -> fused edge forward over the local CSR
                       (destinations = own block, sources = global ids into the gathered Wh)
  backward per layer : row records / Drow of own rows -> ALL-GATHER the gatherable gradient rows (+ the 16-byte row
                       records) -> CSC pass over own SOURCE rows (complete gWh / g_s_src for owned rows, no
                       reduce-scatter of the big [N, D] tensor) -> REDUCE-SCATTER g_s_dst [N, H] -> finish + projection
                       backward on own rows; parameter gradients are partial sums (all-reduced by the caller).
With random power-law edges the halo of a block is ~all nodes, so a full all-gather (not a sparse halo exchange) is the
right primitive.  The per-rank stages are the same C-ABI kernels as the single-GPU layer (staged entry points
b200gat_edge_bwd_prep / _csc / _finish); the collectives are torch.distributed (NCCL over NVLink / NVSwitch).

Graph partitioning is a one-off setup step of a static graph and is written with torch ops (usable on CPU tensors, which
is how tests/test_partition_cpu.py checks it against the CSR oracle).
"""
import ctypes

import torch
import torch.distributed as dist

from . import _abi
from .gat import _call, _layer_struct, _ptr, _split_mask, _timed, _workspace


# ------------------------------------------------------------------------------------------------ graph partition
class RowPartition:
    __slots__ = ("num_nodes", "num_input_edges", "world", "rank", "block", "lo", "hi", "rowptr", "col", "eid", "colptr", "crow", "ceid",
                 "hub_rows", "hub_cols", "rowend", "colend", "max_in_degree", "max_out_degree", "_struct")

    @property
    def n_own(self):
        return self.hi - self.lo

    @property
    def padded_rows(self):
        return self.block * self.world

    def c_struct(self):
        return self._struct


def block_size(num_nodes, world):
    return (num_nodes + world - 1) // world


def build_row_partition(edge_index, num_nodes, world, rank):
    """Local CSR (edges whose DESTINATION is in the block; rows re-based to the block, sources global) and local CSC
    (edges whose SOURCE is in the block; columns re-based, destinations global) of [edge_index ; self loops].
    Order inside a row / column = the single-GPU canonical order (stable sort of the original edge list)."""
    dev = edge_index.device
    n, e = int(num_nodes), int(edge_index.shape[1])
    b = block_size(n, world)
    lo, hi = min(rank * b, n), min((rank + 1) * b, n)
    loops = torch.arange(n, dtype=torch.int64, device=dev)
    src = torch.cat([edge_index[0], loops])
    dst = torch.cat([edge_index[1], loops])
    pos = torch.arange(e + n, dtype=torch.int64, device=dev)

    def build(keys_own, other, ids):
        order = torch.sort(keys_own, stable=True).indices
        counts = torch.bincount(keys_own, minlength=hi - lo)
        ptr = torch.zeros(hi - lo + 1, dtype=torch.int64, device=dev)
        torch.cumsum(counts, 0, out=ptr[1:])
        return ptr.to(torch.int32), other[order].to(torch.int32).contiguous(), ids[order].to(torch.int32).contiguous()

    in_dst = (dst >= lo) & (dst < hi)
    rowptr, col, eid = build(dst[in_dst] - lo, src[in_dst], pos[in_dst])
    # CSC must be stable w.r.t. the CSR order (destination-major), as in the single-GPU build
    in_src = (src >= lo) & (src < hi)
    s_src, s_dst, s_pos = src[in_src], dst[in_src], pos[in_src]
    by_dst = torch.sort(s_dst, stable=True).indices
    s_src, s_dst, s_pos = s_src[by_dst], s_dst[by_dst], s_pos[by_dst]
    colptr, crow, ceid = build(s_src - lo, s_dst, s_pos)

    p = RowPartition()
    p.num_nodes, p.num_input_edges, p.world, p.rank, p.block, p.lo, p.hi = n, e, world, rank, b, lo, hi
    p.rowptr, p.col, p.eid, p.colptr, p.crow, p.ceid = rowptr, col, eid, colptr, crow, ceid
    p._struct = None
    # scheduling by degree (include/b200gat.h: b200gat_graph.hub_rows): own rows / columns longer than HUB_DEGREE
    def hubs(ptr):
        deg = ptr[1:] - ptr[:-1]
        is_hub = deg > _abi.HUB_DEGREE
        return (torch.nonzero(is_hub).flatten().to(torch.int32), torch.where(is_hub, ptr[:-1], ptr[1:]).contiguous(),
                int(deg.max()) if deg.numel() else 0)
    (p.hub_rows, p.rowend, p.max_in_degree), (p.hub_cols, p.colend, p.max_out_degree) = hubs(rowptr), hubs(colptr)
    if rowptr.is_cuda:
        nr, nc = int(p.hub_rows.numel()), int(p.hub_cols.numel())
        p._struct = _abi.Graph(hi - lo, int(col.numel()), rowptr.data_ptr(), col.data_ptr(), eid.data_ptr(),
                               colptr.data_ptr(), crow.data_ptr(), ceid.data_ptr(), n,   # span: one big graph
                               p.hub_rows.data_ptr() if nr else None, nr, p.rowend.data_ptr() if nr else None,
                               p.hub_cols.data_ptr() if nc else None, nc, p.colend.data_ptr() if nc else None,
                               p.max_in_degree, p.max_out_degree)
    return p


# ------------------------------------------------------------------------------------------------ per-rank stages
def _geom(geom):
    f_in, c, h, concat = geom
    layer = _layer_struct(f_in, c, h, concat)
    cp = layer.c_pad
    return layer, f_in, c, h, concat, cp, h * cp, (h * c if concat else c), ((not concat) and h > 1)


class PeerBuffer:
    """The full [P * block, Dp] Wh buffer of one layer in SYMMETRIC memory (torch.distributed._symmetric_memory): every
    rank holds one and can address every other rank's copy over NVLink.  The projection kernel stores each tile it
    produces into all of them (b200gat_proj_fwd_args.wh_peers), so the all-gather of Wh happens inside the GEMM's store
    epilogue, tile by tile, instead of as an NCCL collective after it; `barrier()` (signal pads, on the current stream)
    is the only exchange left before the edge kernel reads the buffer."""

    def __init__(self, block, dp, device, group):
        import torch.distributed._symmetric_memory as symm
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world, self.block, self.dp = dist.get_rank(group), dist.get_world_size(group), block, dp
        self.tensor = symm.empty((self.world * block, dp), dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.tensor, group)
        self.ptrs = [int(p) for p in self.handle.buffer_ptrs]
        assert self.ptrs[self.rank] == self.tensor.data_ptr()

    def own_rows(self):
        return self.tensor[self.rank * self.block:(self.rank + 1) * self.block]

    def peer_ptrs(self):
        off = self.rank * self.block * self.dp * 4
        return [p + off for k, p in enumerate(self.ptrs) if k != self.rank]

    def barrier(self):
        self.handle.barrier()


def peer_push_enabled(world):
    """B200GAT_PEER_PUSH=1 / 0 forces the fused projection + peer-memory all-gather on / off; default: on for 2 GPUs
    (measured: 122.7 -> 118.4 ms per step of the 2.4 M-node graph), NCCL all-gather for more."""
    import os
    env = os.environ.get("B200GAT_PEER_PUSH")
    return env == "1" if env is not None else world == 2


def stage_proj(geom, params, x_own, block, peer=None, act_in=False, x_amax=None, keep_split=False, rows16=False):
    """-> wh [block, Dp] (rows >= n_own zero), s_src [block, H], s_dst [n_own, H]   (GAT.py:42-52 on own rows)
    [, x_split when keep_split: the tensor-core operand split of x kept for stage_proj_bwd, or None on the CUDA-core path].
    peer: a PeerBuffer — wh is then the own row block INSIDE the gathered buffer and the kernel also stores it into every
    other rank's copy (rows >= n_own of the block are left untouched: no node id points at them).
    act_in / x_amax: the layer-boundary fusion of b200gat_proj_fwd (x is a pre-activation tensor, ELU applied on load)."""
    lib = _abi.lib()
    layer, f_in, c, h, concat, cp, dp, d_out, heads_mode = _geom(geom)
    w, bw, a1, a2, b1, b2 = params
    dev, n = x_own.device, x_own.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    wh = peer.own_rows() if peer is not None else torch.zeros((block, dp), **f32)
    s_src = torch.zeros((block, h), **f32)
    s_dst = torch.empty((n, h), **f32)
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws_bytes = int(lib.b200gat_proj_fwd_workspace_bytes(ctypes.byref(layer), n))
    ws = _workspace(ws_bytes, dev)
    split_bytes = int(lib.b200gat_proj_split_bytes(ctypes.byref(layer), n)) if keep_split else 0
    x_split = _workspace(split_bytes, dev) if split_bytes else None
    pa = _abi.ProjFwdArgs(layer, n, x_own.data_ptr(), x_own.stride(0) if n else f_in, w.data_ptr(), bw.data_ptr(),
                          a1.data_ptr(), a2.data_ptr(), b1.data_ptr(), b2.data_ptr(), wh.data_ptr(), s_src.data_ptr(),
                          s_dst.data_ptr(), ws.data_ptr(), ws_bytes, _ptr(x_split), split_bytes,
                          _abi.ACT_ELU if act_in else _abi.ACT_NONE, _ptr(x_amax))
    if peer is not None:
        ptrs = peer.peer_ptrs()
        for k, ptr in enumerate(ptrs):
            pa.wh_peers[k] = ptr
        pa.num_peers = len(ptrs)
    # rows16: ALSO write a bf16 copy of the own rows — the buffer that is all-gathered and gathered from in the bf16 mode
    wh16 = torch.zeros((block, dp), dtype=torch.bfloat16, device=dev) if rows16 else None
    pa.wh_bf16 = _ptr(wh16)
    _call("b200gat_proj_fwd", lib.b200gat_proj_fwd, pa, stream, geom)
    if rows16:
        return wh, s_src, s_dst, x_split, wh16
    if keep_split:
        return wh, s_src, s_dst, x_split
    return wh, s_src, s_dst


def stage_edge_fwd(geom, part, wh_full, s_src_full, s_dst_own, bias, mask, out_amax=None, wh16_full=None):
    """-> out [n_own, D_out], rowmax, rowsum [n_own, H], o_heads or None   (GAT.py:53-67 for the own destination rows)
    out_amax: optional int32[1] device word <- bit pattern of max|out| over the OWN rows (the next layer's x_amax)"""
    lib = _abi.lib()
    layer, f_in, c, h, concat, cp, dp, d_out, heads_mode = _geom(geom)
    dev, n = wh_full.device, part.n_own
    f32 = dict(dtype=torch.float32, device=dev)
    out = torch.empty((n, d_out), **f32)
    rowmax = torch.empty((n, h), **f32)
    rowsum = torch.empty((n, h), **f32)
    o_heads = torch.empty((n, dp), **f32) if heads_mode else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    mask, drop = _split_mask(mask)      # [E', H] tensor in ORIGINAL (global) edge order, or (p, seed): in-kernel Philox
    # wh16_full: the all-gathered bf16 rows (the fp32 pointer is then only a placeholder of the own rows: never gathered)
    ea = _abi.EdgeFwdArgs(layer, part.c_struct(), wh_full.data_ptr(), s_src_full.data_ptr(), s_dst_own.data_ptr(),
                          bias.data_ptr(), _ptr(mask), out.data_ptr(), d_out, rowmax.data_ptr(), rowsum.data_ptr(),
                          _ptr(o_heads), _ptr(out_amax), _abi.dropout_struct(drop), _ptr(wh16_full))
    _call("b200gat_edge_fwd", lib.b200gat_edge_fwd, ea, stream, geom)
    return out, rowmax, rowsum, o_heads


def gather_layout(geom, act_out=False):
    """Which rows the backward all-gathers: (direct, width, ldg, head_stride).  direct: gout itself is gatherable (never
    with a deferred output activation: the gathered rows are then gout * ELU'(out), written by the prep stage)."""
    layer, f_in, c, h, concat, cp, dp, d_out, heads_mode = _geom(geom)
    concat_like = concat or h == 1
    if concat_like and c % 4 == 0 and not act_out:
        return True, d_out, d_out, c
    if concat_like:
        return False, dp, dp, cp
    return False, cp, cp, 0


def stage_prep(geom, gout_own, fwd_out_own, bias, s_dst_own, rowmax, rowsum, block, act_out=False, rows16=False):
    """-> rowrec [block, H, 4], g_rows [block, width] (zero padded), g_bias partial [D_out]
    act_out: gout is d/d ELU(out) of a layer whose activation was deferred to its consumer (concat-like layers)"""
    lib = _abi.lib()
    layer, f_in, c, h, concat, cp, dp, d_out, heads_mode = _geom(geom)
    dev, n = gout_own.device, gout_own.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    direct, width, _, _ = gather_layout(geom, act_out or rows16)      # bf16 rows: always the (bf16) copy
    rowrec = torch.zeros((block, h, 4), **f32)
    g_bias = torch.empty(d_out, **f32)
    g16 = None
    if rows16:
        g_rows = g16 = torch.zeros((block, width), dtype=torch.bfloat16, device=dev)
        g_pad = None
    elif direct:
        if n == block:
            g_rows = gout_own
        else:
            g_rows = torch.zeros((block, width), **f32)
            g_rows[:n].copy_(gout_own)
        g_pad = None
    else:
        g_rows = torch.zeros((block, width), **f32)
        g_pad = g_rows
    stream = torch.cuda.current_stream(dev).cuda_stream
    pa = _abi.EdgeBwdPrepArgs(layer, n, gout_own.data_ptr(), d_out,
                              None if heads_mode else fwd_out_own.data_ptr(), d_out,
                              fwd_out_own.data_ptr() if heads_mode else None, bias.data_ptr(),
                              s_dst_own.data_ptr(), rowmax.data_ptr(), rowsum.data_ptr(), rowrec.data_ptr(),
                              _ptr(g_pad), g_bias.data_ptr(), _abi.ACT_ELU if act_out else _abi.ACT_NONE, _ptr(g16))
    _call("b200gat_edge_bwd_prep", lib.b200gat_edge_bwd_prep, pa, stream, geom)
    return rowrec, g_rows, g_bias


def stage_csc(geom, part, wh_own, s_src_own, rowrec_full, g_full, mask, act_out=False, rows16=False):
    """-> g_wh [n_own, Dp], g_s_src [n_own, H], g_s_dst_full [P*block, H] (this rank's partial sums for ALL nodes)"""
    lib = _abi.lib()
    layer, f_in, c, h, concat, cp, dp, d_out, heads_mode = _geom(geom)
    dev, n = wh_own.device, part.n_own
    f32 = dict(dtype=torch.float32, device=dev)
    _, _, ldg, hs = gather_layout(geom, act_out or rows16)
    g_wh = torch.empty((n, dp), **f32)
    g_s_src = torch.empty((n, h), **f32)
    g_s_dst_full = torch.zeros((part.padded_rows, h), **f32)
    stream = torch.cuda.current_stream(dev).cuda_stream
    mask, drop = _split_mask(mask)
    ca = _abi.EdgeBwdCscArgs(layer, n, part.colptr.data_ptr(), part.crow.data_ptr(), part.ceid.data_ptr(),
                             wh_own.data_ptr(), s_src_own.data_ptr(), rowrec_full.data_ptr(), _ptr(mask),
                             None if rows16 else g_full.data_ptr(), ldg, hs, g_wh.data_ptr(), g_s_src.data_ptr(), g_s_dst_full.data_ptr(),
                             part.num_nodes, part.hub_cols.data_ptr() if part.hub_cols.numel() else None,
                             int(part.hub_cols.numel()), part.colend.data_ptr() if part.hub_cols.numel() else None,
                             part.max_out_degree, _abi.dropout_struct(drop), g_full.data_ptr() if rows16 else None)
    _call("b200gat_edge_bwd_csc", lib.b200gat_edge_bwd_csc, ca, stream, geom)
    return g_wh, g_s_src, g_s_dst_full


def stage_finish(geom, wh_own, a1, a2, g_s_src, g_s_dst_own, g_wh, want_split=False):
    """g_wh -> gT in place; -> (g_bw, g_a1, g_a2 [Dp], g_b1, g_b2 [H]) partial sums of the own block [, g_split].
    want_split: gT is written directly as the tensor-core operand split stage_proj_bwd consumes (no fp32 gT, no split
    pass); g_split is None when the geometry's projection backward runs on the CUDA cores (g_wh then holds fp32 gT)."""
    lib = _abi.lib()
    layer, f_in, c, h, concat, cp, dp, d_out, heads_mode = _geom(geom)
    dev, n = wh_own.device, g_wh.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    g_bw, g_a1, g_a2 = (torch.empty(dp, **f32) for _ in range(3))
    g_b1, g_b2 = (torch.empty(h, **f32) for _ in range(2))
    stream = torch.cuda.current_stream(dev).cuda_stream
    gs_bytes = int(lib.b200gat_edge_bwd_split_bytes(ctypes.byref(layer), n)) if want_split else 0
    g_split = _workspace(gs_bytes, dev) if gs_bytes else None
    fa = _abi.EdgeBwdFinishArgs(layer, n, wh_own.data_ptr(), a1.data_ptr(), a2.data_ptr(), g_s_src.data_ptr(),
                                g_s_dst_own.data_ptr(), g_wh.data_ptr(), g_bw.data_ptr(), g_a1.data_ptr(),
                                g_a2.data_ptr(), g_b1.data_ptr(), g_b2.data_ptr(), _ptr(g_split), gs_bytes)
    _call("b200gat_edge_bwd_finish", lib.b200gat_edge_bwd_finish, fa, stream, geom)
    if want_split:
        return g_bw, g_a1, g_a2, g_b1, g_b2, g_split
    return g_bw, g_a1, g_a2, g_b1, g_b2


def stage_proj_bwd(geom, g_t, x_own, w, need_gx, x_split=None, g_split=None, act_in=False):
    """x_split / g_split: the operand splits kept by stage_proj / written by stage_finish (skip the split passes)"""
    lib = _abi.lib()
    layer, f_in, c, h, concat, cp, dp, d_out, heads_mode = _geom(geom)
    dev, n = x_own.device, x_own.shape[0]
    f32 = dict(dtype=torch.float32, device=dev)
    g_w = torch.empty((dp, f_in), **f32)
    g_x = torch.empty((n, f_in), **f32) if need_gx else None
    stream = torch.cuda.current_stream(dev).cuda_stream
    ws_bytes = int(lib.b200gat_proj_bwd_workspace_bytes(ctypes.byref(layer), n))
    ws = _workspace(ws_bytes, dev)
    pb = _abi.ProjBwdArgs(layer, n, g_t.data_ptr(), x_own.data_ptr(), x_own.stride(0) if n else f_in, w.data_ptr(),
                          _ptr(g_x), f_in, g_w.data_ptr(), ws.data_ptr(), ws_bytes,
                          _ptr(x_split), x_split.numel() if x_split is not None else 0,
                          _abi.ACT_ELU if act_in else _abi.ACT_NONE,
                          _ptr(g_split), g_split.numel() if g_split is not None else 0)
    _call("b200gat_proj_bwd", lib.b200gat_proj_bwd, pb, stream, geom)
    return g_x, g_w


# ------------------------------------------------------------------------------------------------ collectives
def all_gather_rows(own_padded, group=None):
    """[block, W] per rank -> [P*block, W]; row r*block + k of the result is global node r*block + k."""
    world = dist.get_world_size(group)
    full = torch.empty((world * own_padded.shape[0],) + tuple(own_padded.shape[1:]), dtype=own_padded.dtype,
                       device=own_padded.device)
    dist.all_gather_into_tensor(full, own_padded.contiguous(), group=group)
    return full


def reduce_scatter_rows(full, block, group=None):
    """[P*block, W] partial sums per rank -> own [block, W] total."""
    own = torch.empty((block,) + tuple(full.shape[1:]), dtype=full.dtype, device=full.device)
    dist.reduce_scatter_tensor(own, full.contiguous(), op=dist.ReduceOp.SUM, group=group)
    return own


class PartitionedGATFunction(torch.autograd.Function):
    """One GAT layer on the own row block of a partitioned graph.  Returned parameter gradients are the own block's
    partial sums: all-reduce (SUM) them over ranks (parallel.GradBucket.all_reduce_mean(weight=1.0))."""

    @staticmethod
    def forward(ctx, x_own, w, bw, a1, a2, b1, b2, bias, part, geom, mask, group, peer=None, x_full=None,
                fuse=(False, False, None)):
        # fuse = (act_in, act_out, x_amax): the layer-boundary fusions of gat.GATLayerFunction — the ELU between two layers
        # is applied by the CONSUMER while it loads its operand, the producer's backward multiplies by ELU'(out).
        # -> (out, out_amax): out_amax (int32[1]) bounds max|out| of the OWN rows for the next layer's operand scale
        act_in, act_out, x_amax = fuse[:3]
        # bf16 storage of the gathered rows — which in this mode are also what travels over NVLink (half the wire bytes);
        # only the plain configuration has kernels for it (no dropout / mask)
        rows16 = bool(len(fuse) > 3 and fuse[3]) and mask is None
        # x_full: the layer's input for ALL nodes, replicated on every rank (the static input features of layer 1).  The
        # rank then projects all N rows itself and NOTHING is exchanged in this layer's forward: for the 2.4 M-node graph
        # a 100 -> 512 projection of every node costs 2.5 ms, its all-gather (4.3 GB received per rank) 6-7 ms.
        if x_own.shape[0] != part.n_own:
            raise ValueError(f"x_own has {x_own.shape[0]} rows, the partition's own block [{part.lo}, {part.hi}) has "
                             f"{part.n_own} (blocks are ceil(N / P) rows: partition.block_size)")
        x_own = x_own.contiguous()
        w, bw, a1, a2, b1, b2, bias = (t.contiguous() for t in (w, bw, a1, a2, b1, b2, bias))
        with torch.cuda.device(x_own.device):
            if peer is not None and x_full is None:   # nobody may still be reading this buffer (an earlier forward's edge
                with _timed("peer_barrier", geom):   # kernel) when the first remote tile lands: one more signal-pad barrier
                    peer.barrier()
            if x_full is not None:
                n_all = x_full.shape[0]
                res = stage_proj(geom, (w, bw, a1, a2, b1, b2), x_full.contiguous(), n_all, None, rows16=rows16)
                wh_full, s_src_full, s_dst_full = res[:3]
                wh16_full = res[4] if rows16 else None
                wh_pad, s_src_pad = wh_full[part.lo:], s_src_full[part.lo:]        # own rows first (only [:n_own] is used)
                s_dst = s_dst_full[part.lo:part.hi]
                peer, x_split = None, None
            else:
                res = stage_proj(geom, (w, bw, a1, a2, b1, b2), x_own, part.block, peer, act_in=act_in, x_amax=x_amax,
                                 keep_split=True, rows16=rows16)
                wh_pad, s_src_pad, s_dst, x_split = res[:4]
                wh16_full = None
            if x_full is not None:
                pass
            elif peer is not None:        # Wh went to every GPU from inside the projection kernel: wait for everybody's tiles
                with _timed("peer_barrier", geom):
                    peer.barrier()
                wh_full = peer.tensor
            elif rows16:
                with _timed("all_gather_wh", geom):
                    wh16_full = all_gather_rows(res[4], group)
                wh_full = wh_pad                # placeholder: the kernels gather from wh16_full
            else:
                with _timed("all_gather_wh", geom):
                    wh_full = all_gather_rows(wh_pad, group)
            if x_full is None:
                with _timed("all_gather_s_src", geom):
                    s_src_full = all_gather_rows(s_src_pad, group)
            out_amax = torch.zeros(1, dtype=torch.int32, device=x_own.device)
            out, rowmax, rowsum, o_heads = stage_edge_fwd(geom, part, wh_full, s_src_full, s_dst, bias, mask, out_amax,
                                                          wh16_full)
        n = part.n_own
        ctx.part, ctx.geom, ctx.mask, ctx.group = part, geom, mask, group
        ctx.act = (bool(act_in), bool(act_out))
        ctx.rows16 = rows16
        # peer mode: wh_pad is a view of the layer's persistent symmetric buffer, which the NEXT forward through this layer
        # overwrites through raw pointers (no autograd version bump) — an eval forward, a second micro-batch or activation
        # checkpointing between this forward and its backward would silently corrupt the saved Wh.  Keep a private copy.
        wh_saved = wh_pad[:n].clone() if peer is not None else wh_pad[:n]
        ctx.save_for_backward(x_own, w, a1, a2, bias, wh_saved, s_src_pad[:n], s_dst, rowmax, rowsum,
                              out if o_heads is None else o_heads, x_split)
        ctx.mark_non_differentiable(out_amax)
        return out, out_amax

    @staticmethod
    def backward(ctx, gout, _g_amax):
        x_own, w, a1, a2, bias, wh_own, s_src_own, s_dst, rowmax, rowsum, fwd_out, x_split = ctx.saved_tensors
        part, geom, mask, group = ctx.part, ctx.geom, ctx.mask, ctx.group
        act_in, act_out = ctx.act
        gout = gout.contiguous()
        with torch.cuda.device(gout.device):
            rows16 = ctx.rows16
            rowrec, g_rows, g_bias = stage_prep(geom, gout, fwd_out, bias, s_dst, rowmax, rowsum, part.block, act_out, rows16)
            with _timed("all_gather_g", geom):
                g_full = all_gather_rows(g_rows, group)
            with _timed("all_gather_rowrec", geom):
                rowrec_full = all_gather_rows(rowrec, group)
            g_wh, g_s_src, g_s_dst_full = stage_csc(geom, part, wh_own, s_src_own, rowrec_full, g_full, mask, act_out, rows16)
            with _timed("reduce_scatter_g_s_dst", geom):
                g_s_dst_own = reduce_scatter_rows(g_s_dst_full, part.block, group)
            g_bw, g_a1, g_a2, g_b1, g_b2, g_split = stage_finish(geom, wh_own, a1, a2, g_s_src, g_s_dst_own, g_wh,
                                                                 want_split=True)
            g_x, g_w = stage_proj_bwd(geom, g_wh, x_own, w, ctx.needs_input_grad[0], x_split, g_split, act_in)
        return g_x, g_w, g_bw, g_a1, g_a2, g_b1, g_b2, g_bias, None, None, None, None, None, None, None


def partitioned_layer_forward(layer, x_own, part, group=None, x_full=None, act_in=False, act_out=False, x_amax=None,
                              return_amax=False):
    """Run a GraphAttentionLayer module on the own block of a row-partitioned graph.  Attention dropout (GAT.py:61) is
    generated inside the kernels from (seed, ORIGINAL edge position, head): every rank uses rank 0's two seed words (one
    16-byte broadcast), so an edge gets the same multiplier in the forward of the rank owning its destination and in the
    backward of the rank owning its source — as on one GPU with the same seed."""
    mask = None
    if layer.mask_hook is not None:
        mask = layer.mask_hook((part.num_nodes + part.num_input_edges, layer.num_heads)).to(device=x_own.device,
                                                                                          dtype=torch.float32).contiguous()
    elif layer.training and float(layer.dropout_val) > 0.0:
        seed = torch.randint(-2 ** 63, 2 ** 63 - 1, (2,), dtype=torch.int64, device=x_own.device)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast(seed, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        layer._last_dropout_seed = seed
        mask = (min(float(layer.dropout_val), 1.0), seed)
    w, bw, a1, a2, b1, b2 = layer._packed()
    geom = (layer.input_channels, layer.output_channels, layer.num_heads, bool(layer.concat))
    if act_out and not layer.can_fuse_activation_out():
        raise ValueError("act_out needs a concat-like layer")
    # (the bf16 rows of the bf16 gather mode travel by NCCL: the peer push ships fp32 tiles)
    rows16 = layer.gather_dtype == torch.bfloat16 and mask is None
    peer = None if (x_full is not None or rows16) else _peer_buffer(layer, geom, part, x_own, group)
    out, amax = PartitionedGATFunction.apply(x_own, w, bw, a1, a2, b1, b2, layer.bias, part, geom, mask, group, peer, x_full,
                                             (bool(act_in), bool(act_out), x_amax, layer.gather_dtype == torch.bfloat16))
    return (out, amax) if return_amax else out


def _peer_buffer(layer, geom, part, x_own, group):
    """The layer's symmetric Wh buffer (created once per layer and partition shape), or None: single rank, projection on
    the CUDA-core path, B200GAT_PEER_PUSH=0, or symmetric memory unavailable (then the NCCL all-gather runs instead)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world <= 1 or world - 1 > 7 or not peer_push_enabled(world):
        return None
    lstruct, _, _, _, _, _, dp, _, _ = _geom(geom)
    key = (part.block, dp, world)
    cached = getattr(layer, "_peer_buf", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    buf = None
    try:
        # the same decision on every rank (the rendezvous is collective): the smallest block must also be on the
        # tensor-core path
        n_last = part.num_nodes - (world - 1) * part.block
        lib = _abi.lib()
        if n_last > 0 and all(int(lib.b200gat_proj_split_bytes(ctypes.byref(lstruct), m)) != 0 for m in (part.block, n_last)):
            buf = PeerBuffer(part.block, dp, x_own.device, group)
    except Exception as exc:                      # symmetric memory not available on this system: NCCL all-gather instead
        import sys
        print(f"[b200gat] peer-memory all-gather disabled ({type(exc).__name__}: {exc}); using NCCL all_gather",
              file=sys.stderr)
        buf = None
    layer._peer_buf = (key, buf)
    return buf


class PartitionedGATStack(torch.nn.Module):
    """GATStack (plain layers + ELU) executed row-partitioned: forward(x_own, part) -> out_own."""

    def __init__(self, stack):
        super().__init__()
        self.stack = stack

    def forward(self, x_own, part, group=None, x_full=None):
        """x_full: the input features of ALL nodes when every rank holds them (a static graph's features are loaded once):
        layer 1 then exchanges nothing in its forward (PartitionedGATFunction)."""
        convs = self.stack.convs
        pending, amax = False, None          # pending: x_own is a pre-activation tensor whose ELU the next layer applies
        last = len(convs) - 1
        for k, conv in enumerate(convs):
            fuse_out = k < last and conv.can_fuse_activation_out()
            # (x_amax bounds the OWN rows only — each rank projects its own rows, so that is all the operand scale needs)
            x_own, amax = partitioned_layer_forward(conv, x_own, part, group, x_full if k == 0 else None, act_in=pending,
                                                    act_out=fuse_out, x_amax=amax if pending else None, return_amax=True)
            pending = fuse_out
            if k < last and not fuse_out:
                x_own = torch.nn.functional.elu(x_own)
        return x_own
