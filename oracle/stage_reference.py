"""Recipe: stage the reference's own Python — UNMODIFIED — under oracle/_ref/ so that it can travel to the GPU box.

    python -m oracle.stage_reference            (also run by __graft_entry__.build() when /root/reference is present)

/root/reference exists only in the build container.  The reference is interpreted Python, so the thing a compiled
reference would leave in oracle/_ref/ (a .so) is here the three modules the hot path lives in, copied byte for byte:
GAT.py, GATNet.py and run_act_func_experiment.py (its private copy of the layer).  oracle/_ref/ is git-ignored (the
sources never enter this repository's history) but NOT gpurun-ignored, so `bench.py --impl reference` and the CPU
baseline on the GPU box run the reference's own files on top of oracle/pyg_standin (kind: "reference") instead of the
port.  MANIFEST.json records the sha256 of every staged file.  TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = "/root/reference"
FILES = ("GAT.py", "GATNet.py", "run_act_func_experiment.py")


def stage(source=SOURCE, dest=DEST):
    """-> dest, or None when the reference is not present (GPU box: whatever was staged earlier is used as is)."""
    if not os.path.isfile(os.path.join(source, "GAT.py")):
        return None
    os.makedirs(dest, exist_ok=True)
    manifest = {}
    for name in FILES:
        src = os.path.join(source, name)
        shutil.copyfile(src, os.path.join(dest, name))
        manifest[name] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump({"source": source, "sha256": manifest}, f, indent=1)
    return dest


if __name__ == "__main__":
    print(stage())
