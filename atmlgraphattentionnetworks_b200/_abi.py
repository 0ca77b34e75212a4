"""ctypes binding of include/b200gat.h (the C-ABI boundary of the hot path).

The library is the ONLY implementation of the path: there is no CPU or PyTorch fallback.  If libb200gat.so is missing
or does not load, every op raises — loudly — instead of degrading.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200gat.so")
ABI_VERSION = 15

_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class Graph(C.Structure):
    _fields_ = [("num_nodes", C.c_int64), ("num_edges", C.c_int64),
                ("rowptr", C.c_void_p), ("col", C.c_void_p), ("eid", C.c_void_p),
                ("colptr", C.c_void_p), ("crow", C.c_void_p), ("ceid", C.c_void_p), ("span", C.c_int64),
                ("hub_rows", C.c_void_p), ("num_hub_rows", C.c_int64), ("rowend", C.c_void_p),
                ("hub_cols", C.c_void_p), ("num_hub_cols", C.c_int64), ("colend", C.c_void_p),
                ("max_in_degree", C.c_int64), ("max_out_degree", C.c_int64)]


HUB_DEGREE = 512   # B200GAT_HUB_DEGREE


class Layer(C.Structure):
    _fields_ = [("in_channels", C.c_int64), ("out_channels", C.c_int64), ("heads", C.c_int64), ("c_pad", C.c_int64),
                ("concat", C.c_int32), ("negative_slope", C.c_float), ("logit_activation", C.c_int32),
                ("reserved", C.c_int32)]


class Dropout(C.Structure):
    """b200gat_dropout: in-kernel attention dropout (p, device pointer to two uint64 seed words)"""
    _fields_ = [("p", C.c_float), ("reserved", C.c_float), ("seed", C.c_void_p)]


def dropout_struct(drop):
    """drop: None or (p, seed tensor int64[2] on the device) -> Dropout"""
    if drop is None:
        return Dropout(0.0, 0.0, None)
    p, seed = drop
    return Dropout(float(p), 0.0, seed.data_ptr())


class ProjFwdArgs(C.Structure):
    _fields_ = [("layer", Layer), ("num_nodes", C.c_int64),
                ("x", C.c_void_p), ("ldx", C.c_int64),
                ("w", C.c_void_p), ("bw", C.c_void_p), ("a1", C.c_void_p), ("a2", C.c_void_p),
                ("b1", C.c_void_p), ("b2", C.c_void_p),
                ("wh", C.c_void_p), ("s_src", C.c_void_p), ("s_dst", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("x_split", C.c_void_p), ("x_split_bytes", C.c_size_t),
                ("x_activation", C.c_int32), ("x_amax", C.c_void_p),
                ("wh_peers", C.c_void_p * 7), ("num_peers", C.c_int32), ("wh_bf16", C.c_void_p)]


class EdgeFwdArgs(C.Structure):
    _fields_ = [("layer", Layer), ("graph", Graph),
                ("wh", C.c_void_p), ("s_src", C.c_void_p), ("s_dst", C.c_void_p), ("bias", C.c_void_p),
                ("mask", C.c_void_p),
                ("out", C.c_void_p), ("ldo", C.c_int64),
                ("rowmax", C.c_void_p), ("rowsum", C.c_void_p), ("o_heads", C.c_void_p), ("out_amax", C.c_void_p),
                ("dropout", Dropout), ("wh_bf16", C.c_void_p)]


class EdgeBwdArgs(C.Structure):
    _fields_ = [("layer", Layer), ("graph", Graph),
                ("gout", C.c_void_p), ("ldgo", C.c_int64),
                ("out", C.c_void_p), ("ldo", C.c_int64),
                ("o_heads", C.c_void_p), ("bias", C.c_void_p),
                ("wh", C.c_void_p), ("s_src", C.c_void_p), ("s_dst", C.c_void_p),
                ("rowmax", C.c_void_p), ("rowsum", C.c_void_p), ("mask", C.c_void_p),
                ("a1", C.c_void_p), ("a2", C.c_void_p),
                ("g_t", C.c_void_p), ("g_bw", C.c_void_p), ("g_a1", C.c_void_p), ("g_a2", C.c_void_p),
                ("g_b1", C.c_void_p), ("g_b2", C.c_void_p), ("g_bias", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("out_activation", C.c_int32), ("g_t_split", C.c_void_p), ("g_t_split_bytes", C.c_size_t),
                ("dropout", Dropout), ("edge_scratch", C.c_void_p), ("edge_scratch_bytes", C.c_size_t),
                ("gather_bf16", C.c_int32), ("rowrec_in", C.c_void_p)]


class EdgeBwdPrepArgs(C.Structure):
    _fields_ = [("layer", Layer), ("num_rows", C.c_int64),
                ("gout", C.c_void_p), ("ldgo", C.c_int64), ("out", C.c_void_p), ("ldo", C.c_int64),
                ("o_heads", C.c_void_p), ("bias", C.c_void_p),
                ("s_dst", C.c_void_p), ("rowmax", C.c_void_p), ("rowsum", C.c_void_p),
                ("rowrec", C.c_void_p), ("g_pad", C.c_void_p), ("g_bias", C.c_void_p), ("out_activation", C.c_int32),
                ("g_pad_bf16", C.c_void_p)]


class EdgeBwdCscArgs(C.Structure):
    _fields_ = [("layer", Layer), ("num_rows", C.c_int64),
                ("colptr", C.c_void_p), ("crow", C.c_void_p), ("ceid", C.c_void_p),
                ("wh", C.c_void_p), ("s_src", C.c_void_p), ("rowrec", C.c_void_p), ("mask", C.c_void_p),
                ("g", C.c_void_p), ("ldg", C.c_int64), ("g_head_stride", C.c_int64),
                ("g_wh", C.c_void_p), ("g_s_src", C.c_void_p), ("g_s_dst", C.c_void_p), ("span", C.c_int64),
                ("hub_cols", C.c_void_p), ("num_hub_cols", C.c_int64), ("colend", C.c_void_p),
                ("max_out_degree", C.c_int64), ("dropout", Dropout), ("g_bf16", C.c_void_p)]


class EdgeBwdFinishArgs(C.Structure):
    _fields_ = [("layer", Layer), ("num_rows", C.c_int64),
                ("wh", C.c_void_p), ("a1", C.c_void_p), ("a2", C.c_void_p),
                ("g_s_src", C.c_void_p), ("g_s_dst", C.c_void_p), ("g_t", C.c_void_p),
                ("g_bw", C.c_void_p), ("g_a1", C.c_void_p), ("g_a2", C.c_void_p),
                ("g_b1", C.c_void_p), ("g_b2", C.c_void_p), ("g_t_split", C.c_void_p), ("g_t_split_bytes", C.c_size_t)]


class ProjBwdArgs(C.Structure):
    _fields_ = [("layer", Layer), ("num_nodes", C.c_int64),
                ("g_t", C.c_void_p), ("x", C.c_void_p), ("ldx", C.c_int64), ("w", C.c_void_p),
                ("g_x", C.c_void_p), ("ldgx", C.c_int64), ("g_w", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("x_split", C.c_void_p), ("x_split_bytes", C.c_size_t),
                ("x_activation", C.c_int32), ("g_t_split", C.c_void_p), ("g_t_split_bytes", C.c_size_t),
                ("parts", C.c_int32), ("fuse_prep", C.c_void_p)]


class ReadoutGeom(C.Structure):
    _fields_ = [("num_nodes", C.c_int64), ("num_graphs", C.c_int64), ("in_channels", C.c_int64), ("hidden", C.c_int64),
                ("classes", C.c_int64)]


class ReadoutFwdArgs(C.Structure):
    _fields_ = [("geom", ReadoutGeom), ("x", C.c_void_p), ("ldx", C.c_int64), ("batch", C.c_void_p),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("pooled", C.c_void_p), ("counts", C.c_void_p), ("hidden_out", C.c_void_p), ("logp", C.c_void_p),
                ("status", C.c_void_p), ("x_activation", C.c_int32)]


class ReadoutBwdArgs(C.Structure):
    _fields_ = [("geom", ReadoutGeom), ("batch", C.c_void_p), ("w1", C.c_void_p), ("w2", C.c_void_p),
                ("pooled", C.c_void_p), ("counts", C.c_void_p), ("hidden_out", C.c_void_p), ("logp", C.c_void_p),
                ("g_logp", C.c_void_p), ("g_x", C.c_void_p), ("ldgx", C.c_int64),
                ("g_w1", C.c_void_p), ("g_b1", C.c_void_p), ("g_w2", C.c_void_p), ("g_b2", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


ACT_NONE, ACT_ELU = 0, 1
PROJ_BWD_GX, PROJ_BWD_GW = 1, 2
LOGIT_LEAKY_RELU, LOGIT_LOGSIGMOID, LOGIT_TANH, LOGIT_HEAD_SOFTMAX = 0, 1, 2, 3

_SIGNATURES = {
    "b200gat_abi_version": (C.c_int, []),
    "b200gat_launch_count": (C.c_uint64, []),
    "b200gat_last_error": (C.c_int, [C.c_char_p, C.c_size_t]),
    "b200gat_dropout_mask": (C.c_int, [C.POINTER(Dropout), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "b200gat_csr_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "b200gat_csr_build": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 7 +
                          [C.c_void_p, C.c_size_t, C.c_void_p]),
    "b200gat_hub_rows": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200gat_proj_fwd_workspace_bytes": (C.c_size_t, [C.POINTER(Layer), C.c_int64]),
    "b200gat_proj_split_bytes": (C.c_size_t, [C.POINTER(Layer), C.c_int64]),
    "b200gat_proj_fwd": (C.c_int, [C.POINTER(ProjFwdArgs), C.c_void_p]),
    "b200gat_edge_fwd": (C.c_int, [C.POINTER(EdgeFwdArgs), C.c_void_p]),
    "b200gat_edge_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(Layer), C.c_int64]),
    "b200gat_edge_bwd_split_bytes": (C.c_size_t, [C.POINTER(Layer), C.c_int64]),
    "b200gat_edge_bwd": (C.c_int, [C.POINTER(EdgeBwdArgs), C.c_void_p]),
    "b200gat_edge_bwd_prep": (C.c_int, [C.POINTER(EdgeBwdPrepArgs), C.c_void_p]),
    "b200gat_edge_bwd_csc": (C.c_int, [C.POINTER(EdgeBwdCscArgs), C.c_void_p]),
    "b200gat_edge_bwd_finish": (C.c_int, [C.POINTER(EdgeBwdFinishArgs), C.c_void_p]),
    "b200gat_proj_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(Layer), C.c_int64]),
    "b200gat_proj_bwd": (C.c_int, [C.POINTER(ProjBwdArgs), C.c_void_p]),
    "b200gat_proj_bwd_can_fuse_prep": (C.c_int, [C.POINTER(Layer), C.c_int64, C.POINTER(Layer)]),
    "b200gat_readout_fwd": (C.c_int, [C.POINTER(ReadoutFwdArgs), C.c_void_p]),
    "b200gat_readout_bwd": (C.c_int, [C.POINTER(ReadoutBwdArgs), C.c_void_p]),
}

_lib = None
launches = 0   # number of kernel-launching ABI calls made by this process
timing = None  # when set to a list, gat.py appends (op_name, layer_tag, start_event, end_event) per ABI call


def launch_count():
    """kernels launched by libb200gat.so in this process (bench.py reports the delta over the timed region)."""
    return int(lib().b200gat_launch_count())


class B200GatError(RuntimeError):
    pass


def lib():
    """Load libb200gat.so (once).  Raises if it is absent: the CUDA extension IS the product path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise B200GatError(
            f"{LIB_PATH} not found: build it with `python -m atmlgraphattentionnetworks_b200.build` "
            "(there is no CPU / PyTorch fallback for the GAT hot path)")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)   # AttributeError if the header and the library disagree
        fn.restype, fn.argtypes = res, args
    got = handle.b200gat_abi_version()
    if got != ABI_VERSION:
        raise B200GatError(f"libb200gat.so ABI version {got} != binding version {ABI_VERSION}: rebuild the library")
    _lib = handle
    return _lib


def last_error():
    buf = C.create_string_buffer(512)
    lib().b200gat_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc, what):
    if rc != 0:
        kind = "argument error" if rc < 0 else "CUDA error"
        raise B200GatError(f"{what} failed ({kind} {rc}): {last_error()}")
