"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/b200gat.h declares (no
compute calls), the ctypes structs mirror the header, the drop-in modules keep the reference's names / signatures /
state_dict keys / seeded init, and the product path refuses CPU tensors (no fallback)."""
import ctypes
import inspect
import os
import re

import pytest
import torch

from atmlgraphattentionnetworks_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200gat.h")


@pytest.fixture(scope="module")
def built_lib():
    from atmlgraphattentionnetworks_b200.build import build
    build()
    return _abi.lib()


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200gat_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 11
    raw = ctypes.CDLL(_abi.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in include/b200gat.h but not exported"
    assert set(syms) == set(_abi._SIGNATURES), "ctypes binding and header disagree"
    assert built_lib.b200gat_abi_version() == _abi.ABI_VERSION
    m = re.search(r"#define\s+B200GAT_ABI_VERSION\s+(\d+)", open(HEADER).read())
    assert int(m.group(1)) == _abi.ABI_VERSION


def test_argument_errors_are_reported_without_a_gpu(built_lib):
    # NULL args / bad geometry are rejected on the host before any launch (negative code + message)
    assert built_lib.b200gat_edge_fwd(None, None) == -1
    assert "NULL" in _abi.last_error()
    bad = _abi.ProjFwdArgs()
    bad.layer = _abi.Layer(4, 6, 2, 6, 1, 0.2)     # c_pad must be round_up(6, 4) = 8
    assert built_lib.b200gat_proj_fwd(ctypes.byref(bad), None) == -2
    assert "c_pad" in _abi.last_error()
    with pytest.raises(_abi.B200GatError):
        _abi.check(-2, "demo")


def test_struct_layouts_match_header(tmp_path):
    """Compile include/b200gat.h with gcc and compare sizeof / offsetof of every struct field with the ctypes mirror."""
    import subprocess
    structs = {"b200gat_graph": _abi.Graph, "b200gat_layer": _abi.Layer, "b200gat_proj_fwd_args": _abi.ProjFwdArgs,
               "b200gat_edge_fwd_args": _abi.EdgeFwdArgs, "b200gat_edge_bwd_args": _abi.EdgeBwdArgs,
               "b200gat_proj_bwd_args": _abi.ProjBwdArgs, "b200gat_edge_bwd_prep_args": _abi.EdgeBwdPrepArgs,
               "b200gat_edge_bwd_csc_args": _abi.EdgeBwdCscArgs, "b200gat_edge_bwd_finish_args": _abi.EdgeBwdFinishArgs,
               "b200gat_dropout": _abi.Dropout, "b200gat_readout_geom": _abi.ReadoutGeom,
               "b200gat_readout_fwd_args": _abi.ReadoutFwdArgs, "b200gat_readout_bwd_args": _abi.ReadoutBwdArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200gat.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_dropin_signatures_and_state_dict_match_reference_port():
    import GAT
    import GATNet
    from oracle.gat_port import PortGATNet, PortGraphAttentionLayer
    sig = inspect.signature(GAT.GraphAttentionLayer.__init__)
    assert list(sig.parameters) == ["self", "input_channels", "output_channels", "num_heads", "concat", "dropout"]
    assert [sig.parameters[k].default for k in ("num_heads", "concat", "dropout")] == [1, False, 0.6]   # GAT.py:8
    assert list(inspect.signature(GAT.GraphAttentionLayer.forward).parameters)[:3] == ["self", "x", "edge_index"]
    assert list(inspect.signature(GATNet.GATNet.__init__).parameters) == ["self", "model_name", "dataset_name",
                                                                          "num_features"]
    torch.manual_seed(3)
    a = GAT.GraphAttentionLayer(13, 6, num_heads=4, concat=True, dropout=0.0)
    torch.manual_seed(3)
    b = PortGraphAttentionLayer(13, 6, num_heads=4, concat=True, dropout=0.0)
    assert list(a.state_dict()) == list(b.state_dict())
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k]), k                      # same RNG consumption order (GAT.py:19-25)
    for ds, f in (("Cora", 20), ("CIFAR10", 3), ("Pubmed", 9), ("Citeseer", 5), ("AmazonComp", 7), ("AmazonPhotos", 7)):
        torch.manual_seed(4)
        na = GATNet.GATNet("GAT", ds, f)
        torch.manual_seed(4)
        nb = PortGATNet("GAT", ds, f)
        assert list(na.state_dict()) == list(nb.state_dict())
        for k, v in na.state_dict().items():
            assert torch.equal(v, nb.state_dict()[k]), (ds, k)
    assert len(GATNet.GATNet("GAT", "Cora", 1433).state_dict()) == 56
    assert sum(p.numel() for p in GATNet.GATNet("GAT", "Cora", 1433).parameters()) == 92462


def test_packed_parameters_round_trip():
    import GAT
    torch.manual_seed(0)
    layer = GAT.GraphAttentionLayer(5, 3, num_heads=2, concat=False, dropout=0.0)
    w, bw, a1, a2, b1, b2 = layer._packed()
    assert w.shape == (8, 5) and bw.shape == (8,) and a1.shape == (8,) and b1.shape == (2,)
    assert torch.equal(w[0:3], layer.ws[0].weight) and torch.equal(w[4:7], layer.ws[1].weight)
    assert torch.all(w[3] == 0) and torch.all(w[7] == 0) and bw[3] == 0 and a2[7] == 0
    assert torch.equal(a1[4:7], layer.attentions1[1].weight[0]) and b2[1] == layer.attentions2[1].bias[0]
    w.sum().backward()
    assert torch.all(layer.ws[1].weight.grad == 1)


def test_persistent_packed_storage_tracks_the_per_head_parameters():
    """The kernels read one packed [Dp, F] / [Dp] / [H] set of arrays; the per-head Linear parameters are views of it."""
    import GAT
    torch.manual_seed(0)
    layer = GAT.GraphAttentionLayer(5, 3, num_heads=2, concat=True, dropout=0.0)
    before = {k: v.clone() for k, v in layer.state_dict().items()}
    w, bw, a1, a2, b1, b2 = layer._packed_storage()
    assert w.shape == (8, 5) and bw.shape == (8,) and b1.shape == (2,)
    for k, v in layer.state_dict().items():
        assert torch.equal(v, before[k]), k                               # tying does not change any value
    assert torch.equal(w[4:7], layer.ws[1].weight) and torch.all(w[3] == 0) and torch.all(w[7] == 0)
    with torch.no_grad():                                                  # an optimizer's in-place update is seen
        layer.ws[1].weight.add_(1.0)
        layer.attentions2[0].bias.fill_(7.0)
    w2, *_, b2b = layer._packed_storage()
    assert w2.data_ptr() == w.data_ptr()                                   # no rebuild
    assert torch.equal(w2[4:7], before["ws.1.weight"] + 1.0) and float(b2b[0]) == 7.0
    sd = {k: torch.full_like(v, 0.25) for k, v in layer.state_dict().items()}
    layer.load_state_dict(sd)                                              # load_state_dict copies in place
    assert torch.all(layer._packed_storage()[0][0:3] == 0.25) and torch.all(layer._packed_storage()[0][3] == 0)
    layer.ws[0].weight.data = torch.ones(3, 5)                             # what .to() / .cuda() do: new tensors
    w3 = layer._packed_storage()[0]
    assert w3.data_ptr() != w.data_ptr() and torch.all(w3[0:3] == 1.0) and torch.all(w3[4:7] == 0.25)
    assert layer.ws[0].weight.data_ptr() == w3.data_ptr()


def test_no_cpu_fallback():
    import GAT
    layer = GAT.GraphAttentionLayer(4, 4, num_heads=2, concat=True, dropout=0.0)
    with pytest.raises(_abi.B200GatError):
        layer(torch.randn(5, 4), torch.zeros(2, 0, dtype=torch.int64))
    from atmlgraphattentionnetworks_b200.graph import build_csr
    with pytest.raises(_abi.B200GatError):
        build_csr(torch.zeros(2, 3, dtype=torch.int64), 4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "atmlgraphattentionnetworks_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    for f in ("GAT.py", "GATNet.py"):
        assert "oracle" not in open(os.path.join(ROOT, f)).read()


def test_shipped_library_is_sm_100a_tcgen05_code(built_lib):
    """The hot GEMMs are hand-written Blackwell code: the shipped .so holds sm_100a SASS with tcgen05.mma (UTCHMMA, incl. the
    cta_group::2 form), TMA tensor loads (UTMALDG), tcgen05.ld (LDTM), tcgen05.commit (UTCBAR) and red.global.add (REDG) —
    the mnemonics B200_PROFILING.md names (the per-kernel table: tools/sass_summary.py -> profiles/r3_sass_summary.md)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", _abi.LIB_PATH], capture_output=True, text=True).stdout
    # every cubin that holds a kernel is sm_100a (nvcc's link step adds one EMPTY default-arch cubin: no functions in it)
    arch, kernels_per_arch = None, {}
    for line in sass.splitlines():
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch = m.group(1)
        elif "Function :" in line:
            kernels_per_arch[arch] = kernels_per_arch.get(arch, 0) + 1
    assert set(kernels_per_arch) == {"sm_100a"}, kernels_per_arch
    assert kernels_per_arch["sm_100a"] > 100
    for mnemonic, least in (("UTCHMMA", 100), ("UTCHMMA.2CTA", 10), ("UTMALDG.2D", 100), ("LDTM", 50), ("UTCBAR", 10),
                            ("REDG.E.ADD.F32", 90)):
        assert sass.count(mnemonic) >= least, f"{mnemonic}: {sass.count(mnemonic)} < {least}"
