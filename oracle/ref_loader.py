"""Import the reference's GAT.py / GATNet.py UNMODIFIED from /root/reference on top of oracle/pyg_standin.

Test infrastructure (see oracle/__init__.py).  /root/reference only exists in the build container, so this
module is used by tests/golden/make_golden.py and by CPU tests that skip when the reference is absent.
"""
import importlib.util
import os
import sys

REFERENCE_DIR = os.environ.get("B200GAT_REFERENCE_DIR", "/root/reference")
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
