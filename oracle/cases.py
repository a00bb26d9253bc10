"""Parity cases shared by `tests/golden/make_golden.py` and the parity tests -- TEST
INFRASTRUCTURE (see the header of `gat_oracle.py`).

A case is one `GATLayer.forward` call: inputs, weights and flags.  Inputs are regenerated
from seeds (`gat-pytorch_b200/synth.py`); weights come from the committed checkpoint extract
`tests/golden/ckpt_weights.npz` (Cora, Citeseer, Pubmed, PATTERN -- the four checkpoints the
reference ships) or from a numpy-seeded Xavier init (PPI, products,
whose checkpoints are absent / do not exist).  Inputs to layers beyond the first are produced
by chaining the oracle through the reference's inter-layer glue (GATModel.py:120-151: layer ->
skip -> ELU), then truncated to 8 mantissa bits so the regenerated fp32 inputs are
bit-identical on every machine.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
GOLDEN_DIR = os.path.join(_ROOT, "tests", "golden")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


synth = _load("gat_b200_synth", os.path.join(_ROOT, "gat-pytorch_b200", "synth.py"))
oracle = _load("gat_oracle", os.path.join(_HERE, "gat_oracle.py"))

_CKPT_NAME = {"cora": "Cora", "citeseer": "Citeseer", "pubmed": "Pubmed", "pattern": "PATTERN"}


def ckpt_weights():
    return np.load(os.path.join(GOLDEN_DIR, "ckpt_weights.npz"))


def model_weights(name):
    """[(W, a, skip_or_None)] per layer."""
    shapes = synth.LAYER_SHAPES[name]
    if name in _CKPT_NAME:
        z, tag, out, j = ckpt_weights(), _CKPT_NAME[name], [], 0
        for i in range(len(shapes)):
            skip = None
            if synth.SKIP[name][i]:
                skip = z[f"{tag}.skip_layer_list.{j}.weight"]
                j += 1
            out.append((z[f"{tag}.gat_layer_list.{i}.W.weight"], z[f"{tag}.gat_layer_list.{i}.a.weight"], skip))
        return out
    out = []
    for i, (W, a) in enumerate(synth.seeded_weights(name)):
        skip = "identity" if synth.SKIP[name][i] else None   # PPI L1: Identity skip (GATModel.py:107-108)
        out.append((W, a, skip))
    return out


def _truncate(x):
    """Keep sign, exponent and the top 8 mantissa bits (bf16-representable fp32)."""
    bits = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFF0000)
    return bits.view(np.float32)


def _elu(x):
    return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))


def model_cases(name, **gen_kwargs):
    """One case per layer of the named config, inputs chained through the oracle."""
    x, ei = synth.GENERATORS[name](**gen_kwargs)
    weights = model_weights(name)
    shapes = synth.LAYER_SHAPES[name]
    cases = []
    for i, ((f_in, nh, f, concat), (W, a, skip)) in enumerate(zip(shapes, weights)):
        x = _truncate(x) if i else np.ascontiguousarray(x, dtype=np.float32)
        cases.append(dict(name=f"{name}_L{i}", x=x, edge_index=ei, W=W, a=a, nh=nh, f=f, concat=concat,
                          add_self_loops=True, const_attention=False, bias=None))
        if i + 1 == len(shapes):
            break
        fw = oracle.forward(x, ei, W, a, nh, f, concat, True)
        h = fw["out"]
        if skip is not None:                                      # GATModel.py:135-145
            sk = x.astype(np.float64) if isinstance(skip, str) else x.astype(np.float64) @ skip.astype(np.float64).T
            h = h + (sk if concat else sk.reshape(-1, nh, f).mean(axis=1))
        x = _elu(h).astype(np.float32)                            # GATModel.py:148-149
    return cases


def adversarial_cases():
    """Edge cases the reference exhibits (SURVEY.md section 8-a 'behavioural edge cases', 8-c)."""
    rng = np.random.default_rng(7)
    x, ei = synth.adversarial()
    f_in = x.shape[1]

    def wa(nh, f, scale=1.0):
        return (synth.xavier_uniform(rng, nh * f, f_in) * scale, synth.xavier_uniform(rng, nh, 2 * nh * f) * scale)

    base = dict(x=x, edge_index=ei, add_self_loops=True, const_attention=False, bias=None)
    cases = []
    W, a = wa(4, 8)
    cases.append(dict(base, name="adv_concat", W=W, a=a, nh=4, f=8, concat=True))
    W, a = wa(3, 5)
    cases.append(dict(base, name="adv_mean_oddF", W=W, a=a, nh=3, f=5, concat=False))
    W, a = wa(3, 7)
    cases.append(dict(base, name="adv_concat_oddF", W=W, a=a, nh=3, f=7, concat=True))
    W, a = wa(1, 1)
    cases.append(dict(base, name="adv_1x1", W=W, a=a, nh=1, f=1, concat=False))
    W, a = wa(2, 4)
    cases.append(dict(base, name="adv_noloops", W=W, a=a, nh=2, f=4, concat=True, add_self_loops=False))
    cases.append(dict(base, name="adv_const", W=W, a=None, nh=2, f=4, concat=True, const_attention=True))
    cases.append(dict(base, name="adv_bias", W=W, a=a, nh=2, f=4, concat=True,
                      bias=rng.standard_normal(8).astype(np.float32)))
    W, a = wa(6, 40)
    cases.append(dict(base, name="adv_wide", W=W, a=a, nh=6, f=40, concat=True))
    # epsilon-dominated softmax: huge logits so exp underflows towards the +1e-8 (section 0-4)
    W, a = wa(4, 8, scale=60.0)
    cases.append(dict(base, name="adv_eps_dominated", W=W, a=a, nh=4, f=8, concat=True))
    # tied global maxima: 3-valued one-hot features -> few distinct logits (section 9.2)
    xt = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=x.shape[0])]
    Wt = synth.xavier_uniform(rng, 8, 3)
    at = synth.xavier_uniform(rng, 2, 16)
    cases.append(dict(base, name="adv_ties", x=xt, W=Wt, a=at, nh=2, f=4, concat=True))
    # int32 edge_index is accepted by the reference (section 8-a)
    W, a = wa(2, 4)
    cases.append(dict(base, name="adv_int32", edge_index=ei.astype(np.int32), W=W, a=a, nh=2, f=4, concat=True))
    return cases


def small_cases():
    """Cases cheap enough for the CPU suite and the golden fixtures."""
    cases = adversarial_cases()
    cases += model_cases("cora")
    cases += model_cases("pubmed")
    cases += model_cases("ppi")
    cases += model_cases("pattern", graphs=16)
    cases += model_cases("products", scale=1.0 / 256)
    return cases


def upstream_grads(case, n_out_rows, out_cols, n_edges):
    """Seeded upstream gradients dL/dout and dL/dalpha (training feeds both: SURVEY 0-7)."""
    rng = np.random.default_rng(abs(hash_name(case["name"])) % (2 ** 32))
    go = rng.standard_normal((n_out_rows, out_cols)).astype(np.float32)
    ga = rng.standard_normal((n_edges, case["nh"])).astype(np.float32)
    return go, ga


def hash_name(s):
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def sample_idx(n, k=96):
    return np.unique(np.linspace(0, max(n - 1, 0), num=min(n, k)).astype(np.int64)) if n else np.zeros(0, np.int64)
