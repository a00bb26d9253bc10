#!/usr/bin/env python
"""Recipe for `oracle/_ref/`: a byte-for-byte copy of the reference's own Python files -- TEST INFRASTRUCTURE / CPU BASELINE.

    python oracle/make_ref.py [--reference /root/reference]

The reference (loodvn/gat-pytorch) is pure Python with no build step, so "compiling the reference from where it lies"
is a file copy.  The copies are written ONLY to `oracle/_ref/`, which is git-ignored (no reference source ever enters
this repository's history) but not gpurun-ignored, so the unmodified files travel to the GPU box next to the built
`.so` -- exactly like a compiled `oracle/_ref/*.so` would.  `__graft_entry__.build()` runs this recipe whenever the
reference checkout is present.

What uses `oracle/_ref/` (and nothing else may):
  * `bench.py --impl reference` and the `cpu_baseline` leg: import `oracle/_ref/models/gat_layer.py` -- the reference's own
    `GATLayer`, unmodified -- and time it on the host cores (`cpu_baseline.kind = "reference"`);
  * `tests/test_gpu_reference_callers.py`: run the reference's unmodified `train.py` / `vis.py` and task models on the
    B200 layer (namespace-package overlay + offline stand-ins) and compare training curves with the reference layer.
The product path (`gat-pytorch_b200/`) never imports it.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")

# hot path + its callers (8-b "who calls it", 8-f2) + the one small checkpoint vis.py needs
FILES = [
    "models/gat_layer.py", "models/utils.py",
    "models/GATModel.py", "models/planetoid_gat.py", "models/ppi_gat.py", "models/pattern_gat.py",
    "train.py", "vis.py", "run_config.py", "data_utils.py",
    "visualisation/entropy_histograms.py", "visualisation/neighbourhood_attention_weights.py",
    "visualisation/weight_histograms.py",
    "checkpoints/Cora-100epochs.ckpt",
]


def make(reference: str = "/root/reference", dest: str = DEST, quiet: bool = False) -> dict:
    if not os.path.isfile(os.path.join(reference, "models", "gat_layer.py")):
        raise FileNotFoundError(f"no reference checkout at {reference}")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(reference, rel), os.path.join(dest, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(dst):
            os.chmod(dst, 0o644)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        with open(dst, "rb") as fh:
            manifest[rel] = hashlib.sha256(fh.read()).hexdigest()
    with open(os.path.join(dest, "MANIFEST.json"), "w") as fh:
        json.dump({"source": reference, "sha256": manifest}, fh, indent=1, sort_keys=True)
    if not quiet:
        print(f"oracle/_ref: {len(manifest)} files copied unmodified from {reference}")
    return manifest


def available(dest: str = DEST) -> bool:
    return os.path.isfile(os.path.join(dest, "models", "gat_layer.py"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    try:
        make(args.reference)
    except FileNotFoundError as exc:
        print(exc, file=sys.stderr)
        sys.exit(1)
