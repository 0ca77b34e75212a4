"""Import the reference's GAT.py / GATNet.py UNMODIFIED on top of oracle/pyg_standin.

Test infrastructure (see oracle/__init__.py).  The files come from /root/reference where it exists (the build
container) and otherwise from oracle/_ref/, the git-ignored byte-for-byte staging copy that oracle/stage_reference.py
makes so that the reference itself — not only the port — can run on the GPU box (bench.py --impl reference,
cpu_baseline kind "reference").  Used by tests/golden/make_golden.py, bench.py's CPU legs and the CPU tests.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _reference_dir():
    env = os.environ.get("B200GAT_REFERENCE_DIR")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_HERE, "_ref")):
        if os.path.isfile(os.path.join(cand, "GAT.py")):
            return cand
    return "/root/reference"


REFERENCE_DIR = _reference_dir()
_STANDIN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyg_standin")
_cache = {}


def available():
    return os.path.isfile(os.path.join(REFERENCE_DIR, "GAT.py"))


def _load(name, alias):
    spec = importlib.util.spec_from_file_location(alias, os.path.join(REFERENCE_DIR, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Returns (ref_GAT_module, ref_GATNet_module).  The repo's own root-level GAT/GATNet modules are untouched:
    the reference modules are registered under private aliases and `GAT` is swapped in only while the
    reference's GATNet.py executes its `from GAT import GraphAttentionLayer` (GATNet.py:9)."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_DIR}")
    saved = {k: sys.modules.get(k) for k in ("GAT", "torch_geometric", "torch_scatter")}
    sys.path.insert(0, _STANDIN)
    try:
        for k in [m for m in sys.modules if m == "torch_geometric" or m.startswith("torch_geometric.")
                  or m == "torch_scatter"]:
            del sys.modules[k]
        ref_gat = _load("GAT", "_reference_GAT")
        sys.modules["GAT"] = ref_gat
        ref_net = _load("GATNet", "_reference_GATNet")
    finally:
        sys.path.remove(_STANDIN)
        if saved["GAT"] is not None:
            sys.modules["GAT"] = saved["GAT"]
        else:
            sys.modules.pop("GAT", None)
    _cache["mods"] = (ref_gat, ref_net)
    return _cache["mods"]


def load_act_experiment():
    """The reference's run_act_func_experiment.py module, UNMODIFIED (its experiment lives in main(), which is not
    run): gives GraphAttentionLayerActivationTest (run_act_func_experiment.py:13-74)."""
    if "act" in _cache:
        return _cache["act"]
    if not os.path.isfile(os.path.join(REFERENCE_DIR, "run_act_func_experiment.py")):
        raise FileNotFoundError(f"reference not found under {REFERENCE_DIR}")
    sys.path.insert(0, _STANDIN)
    try:
        for k in [m for m in sys.modules if m == "torch_geometric" or m.startswith("torch_geometric.")
                  or m == "torch_scatter"]:
            del sys.modules[k]
        _cache["act"] = _load("run_act_func_experiment", "_reference_run_act_func_experiment")
    finally:
        sys.path.remove(_STANDIN)
    return _cache["act"]


class RefStack:
    """Bench-only composition of the REFERENCE's own GraphAttentionLayer (GAT.py:8) + F.elu between layers, for the
    PPI-shaped / large-graph configs (the reference has no 3-layer model).  spec = [(in, out, heads, concat)]."""

    def __new__(cls, spec, dropout=0.0):
        import torch
        ref_gat, _ = load()

        class _Stack(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.convs = torch.nn.ModuleList(
                    [ref_gat.GraphAttentionLayer(i, o, num_heads=h, concat=c, dropout=dropout) for (i, o, h, c) in spec])

            def forward(self, x, edge_index):
                for k, conv in enumerate(self.convs):
                    x = conv(x, edge_index)
                    if k + 1 < len(self.convs):
                        x = torch.nn.functional.elu(x)
                return x
        return _Stack()
