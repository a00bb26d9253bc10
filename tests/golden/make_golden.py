"""Generate the golden fixtures by running the REFERENCE itself (build container only).

    python tests/golden/make_golden.py

1. Extracts the layer weights of the four checkpoints the reference ships
   (`/root/reference/checkpoints/*.ckpt`, Lightning 1.2.x pickles; two stub classes are enough
   to unpickle them -- SURVEY.md section 5.4) into `ckpt_weights.npz`.
2. For every case of `oracle/cases.py::small_cases()` imports
   `/root/reference/models/gat_layer.py::GATLayer`, loads the case's weights, and runs
   forward + autograd backward (upstream dL/dout AND dL/dalpha) twice: as shipped (fp32) and
   with `.double()` (fp64).  Sampled rows of every result plus full-tensor sums go to
   `golden.npz`.

`/root/reference` does not exist on the GPU box; nothing at test time reads it -- only this
script does.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def extract_checkpoints():
    for mod, cls in [("pytorch_lightning.callbacks.early_stopping", "EarlyStopping"),
                     ("pytorch_lightning.callbacks.model_checkpoint", "ModelCheckpoint")]:
        parts = mod.split(".")
        for i in range(1, len(parts) + 1):
            sys.modules.setdefault(".".join(parts[:i]), types.ModuleType(".".join(parts[:i])))
        setattr(sys.modules[mod], cls, type(cls, (), {}))
    out = {}
    for tag in ["Cora", "Citeseer", "Pubmed", "PATTERN"]:
        ck = torch.load(f"{REF}/checkpoints/{tag}-100epochs.ckpt", map_location="cpu", weights_only=False)
        for k, v in ck["state_dict"].items():
            out[f"{tag}.{k}"] = v.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "ckpt_weights.npz"), **out)
    print("ckpt_weights.npz:", len(out), "tensors")


def run_reference(case, double):
    from models.gat_layer import GATLayer  # the reference's own file
    nh, f = case["nh"], case["f"]
    layer = GATLayer(case["x"].shape[1], f, nh, case["concat"], dropout=0,
                     add_self_loops=case["add_self_loops"], bias=case["bias"] is not None,
                     const_attention=case["const_attention"])
    layer.device = "cpu"
    with torch.no_grad():
        layer.W.weight.copy_(torch.from_numpy(case["W"]))
        if not case["const_attention"]:
            layer.a.weight.copy_(torch.from_numpy(case["a"]))
        if case["bias"] is not None:
            layer.bias_param.copy_(torch.from_numpy(case["bias"]))
    x = torch.from_numpy(case["x"])
    if double:
        layer, x = layer.double(), x.double()
    x.requires_grad_(True)
    out, (ei2, alpha) = layer(x, torch.from_numpy(case["edge_index"]), return_attention_weights=True)
    from cases import upstream_grads
    go, ga = upstream_grads(case, out.shape[0], out.shape[1], alpha.shape[0])
    loss = (out * torch.from_numpy(go).to(out.dtype)).sum()
    if alpha.requires_grad:
        loss = loss + (alpha * torch.from_numpy(ga).to(alpha.dtype)).sum()
    loss.backward()
    res = dict(out=out, alpha=alpha, gx=x.grad, gW=layer.W.weight.grad)
    if not case["const_attention"]:
        res["ga"] = layer.a.weight.grad
    if case["bias"] is not None:
        res["gb"] = layer.bias_param.grad
    return {k: v.detach().numpy() for k, v in res.items()}, ei2.numpy()


def main():
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    extract_checkpoints()
    from cases import small_cases, sample_idx
    store = {}
    for case in small_cases():
        name = case["name"]
        for double in (False, True):
            res, ei2 = run_reference(case, double)
            tag = "f64" if double else "f32"
            for k, v in res.items():
                v2 = v.reshape(v.shape[0], -1)
                rows, cols = sample_idx(v2.shape[0], 64), sample_idx(v2.shape[1], 160)
                store[f"{name}/{tag}/{k}"] = v2[np.ix_(rows, cols)].astype(np.float64 if double else np.float32)
                store[f"{name}/{tag}/{k}_sum"] = np.float64(v2.astype(np.float64).sum())
                store[f"{name}/{tag}/{k}_abs"] = np.float64(np.abs(v2.astype(np.float64)).sum())
                store[f"{name}/{tag}/{k}_max"] = np.float64(np.abs(v2.astype(np.float64)).max()) if v2.size else np.float64(0)
        # integer path: rewritten edge list (sampled columns + checksums) and degree counts
        cols = sample_idx(ei2.shape[1], 512)
        store[f"{name}/ei_cols"] = ei2[:, cols].astype(np.int64)
        store[f"{name}/ei_shape"] = np.array(ei2.shape, dtype=np.int64)
        w = (np.arange(ei2.shape[1], dtype=np.int64) % 1000003) + 1
        store[f"{name}/ei_checksum"] = np.array([(ei2[0].astype(np.int64) * w).sum(), (ei2[1].astype(np.int64) * w).sum()], dtype=np.int64)
        deg = np.bincount(ei2[1], minlength=case["x"].shape[0])
        store[f"{name}/deg_checksum"] = np.array([(deg * (np.arange(deg.size) % 1000003 + 1)).sum(), deg.max()], dtype=np.int64)
        print(name, "E'=", ei2.shape[1], flush=True)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **store)
    print("golden.npz:", len(store), "arrays,", os.path.getsize(os.path.join(HERE, "golden.npz")) / 1e6, "MB")


if __name__ == "__main__":
    main()
