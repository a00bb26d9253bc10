"""Helper (not a test): run one of the reference's UNMODIFIED scripts from oracle/_ref as `__main__`.

    python tests/_run_ref_script.py {b200|reference} <script.py> [script args...]

`b200` puts the namespace-package overlay in front of the reference copy, so the script's
`models.gat_layer` resolves to the B200 layer while every other module is the reference's own; `reference` runs the same
script on the reference layer.  Offline stand-ins for pytorch_lightning / torch_geometric / matplotlib / igraph come last.
After the script returns, one line `REF_SCRIPT_INFO {...}` reports which GATLayer class the script's model used and how
many of our kernels were launched."""
import json
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
OVERLAY = os.path.join(ROOT, "gat-pytorch_b200", "overlay")
SHIMS = os.path.join(OVERLAY, "shims")


def setup_path(layer):
    paths = ([OVERLAY, ROOT] if layer == "b200" else []) + [REF, SHIMS]
    sys.path[:] = paths + [p for p in sys.path if p not in paths and os.path.abspath(p or ".") != os.path.dirname(os.path.abspath(__file__))]


def main():
    layer, script, args = sys.argv[1], sys.argv[2], sys.argv[3:]
    setup_path(layer)
    sys.argv = [os.path.join(REF, script)] + args
    runpy.run_path(os.path.join(REF, script), run_name="__main__")
    import models.gat_layer as layer_mod
    info = {"layer_module": layer_mod.GATLayer.__module__, "layer_file": os.path.relpath(layer_mod.__file__, ROOT), "launches": 0}
    if layer == "b200":
        from gat_pytorch_b200 import _lib
        info["launches"] = int(_lib.load().gat_launch_count())
    print("REF_SCRIPT_INFO " + json.dumps(info), flush=True)


if __name__ == "__main__":
    main()
