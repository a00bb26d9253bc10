"""Layers beyond the edge kernels' limits (num_heads > 8, more than 1024 floats per padded row): the reference constructor
accepts any shape (gat_layer.py:13), so the drop-in processes such layers in head groups (gat_layer.py::_GATWideFunction).
Checked against the fp64 oracle like every other case: rewritten edge list bit-exact, out / alpha / every gradient within
1e-5 tensor-relative, with and without an upstream dL/dalpha, concat and head mean, const_attention, tied maxima."""
import numpy as np
import pytest
import torch

import cases
import gat_oracle as O
from test_gpu_parity import run_cuda, run_oracle, make_layer

pytestmark = pytest.mark.gpu


def _wide_cases():
    from gat_pytorch_b200 import synth
    rng = np.random.default_rng(23)
    x, ei = synth.adversarial()
    f_in = x.shape[1]

    def wa(nh, f):
        return synth.xavier_uniform(rng, nh * f, f_in), synth.xavier_uniform(rng, nh, 2 * nh * f)
    base = dict(x=x, edge_index=ei, add_self_loops=True, const_attention=False, bias=None)
    out = []
    for name, nh, f, concat in (("wide_12heads", 12, 8, True), ("wide_row1200", 4, 300, True), ("wide_16x70_mean", 16, 70, False),
                                ("wide_9heads_oddF_mean", 9, 5, False), ("wide_10x128", 10, 128, True)):
        W, a = wa(nh, f)
        out.append(dict(base, name=name, W=W, a=a, nh=nh, f=f, concat=concat))
    W, _ = wa(12, 8)
    out.append(dict(base, name="wide_const", W=W, a=None, nh=12, f=8, concat=True, const_attention=True))
    xt = np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=x.shape[0])]
    out.append(dict(base, name="wide_ties", x=xt, W=synth.xavier_uniform(rng, 40, 3), a=synth.xavier_uniform(rng, 10, 80), nh=10, f=4,
                    concat=True))
    return {c["name"]: c for c in out}


WIDE = None


def _case(name):
    global WIDE
    if WIDE is None:
        WIDE = _wide_cases()
    return WIDE[name]


@pytest.mark.parametrize("name", ["wide_12heads", "wide_row1200", "wide_16x70_mean", "wide_9heads_oddF_mean", "wide_10x128", "wide_const",
                                  "wide_ties"])
def test_wide_layer_matches_oracle(name):
    case = _case(name)
    got, ei2 = run_cuda(case)
    fw, want = run_oracle(case)
    assert np.array_equal(ei2, fw["edge_index"])
    errs = {k: O.rel_err(got[k], want[k]) for k in want}
    assert all(e <= 1e-5 for e in errs.values()), errs


@pytest.mark.parametrize("name", ["wide_12heads", "wide_16x70_mean", "wide_ties"])
def test_wide_layer_backward_without_attention_gradient(name):
    """The common training path (nothing consumes the returned attention): rowdot + one source-major pass per head group."""
    case = _case(name)
    layer = make_layer(case)
    x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    out = layer(x, torch.from_numpy(case["edge_index"]).cuda())
    fw = O.forward(case["x"], case["edge_index"].astype(np.int64), case["W"], case["a"], case["nh"], case["f"], case["concat"], True)
    go, _ = cases.upstream_grads(case, out.shape[0], out.shape[1], 1)
    (out * torch.from_numpy(go).cuda()).sum().backward()
    gr = O.backward(fw, go, None)
    errs = {"out": O.rel_err(out.detach().cpu().numpy(), fw["out"]), "gx": O.rel_err(x.grad.cpu().numpy(), gr["x"]),
            "gW": O.rel_err(layer.W.weight.grad.cpu().numpy(), gr["W"]), "ga": O.rel_err(layer.a.weight.grad.cpu().numpy(), gr["a"])}
    assert all(e <= 1e-5 for e in errs.values()), errs


def test_wide_layer_dropout_keeps_expectation():
    """Attention dropout in a grouped layer: masks differ between head groups (Philox offset = group) and the mean over draws
    approaches the eval output."""
    case = _case("wide_12heads")
    layer = make_layer(case, dropout=0.5)
    x = torch.from_numpy(case["x"]).cuda()
    ei = torch.from_numpy(case["edge_index"]).cuda()
    layer.eval()
    with torch.no_grad():
        ref = layer(x, ei)
        layer.train()
        torch.manual_seed(0)
        one = layer(x, ei)
        mean = sum(layer(x, ei) for _ in range(300)) / 300
    assert not torch.allclose(one, ref)
    d = one.view(-1, 12, 8) / ref.view(-1, 12, 8).abs().clamp(min=1e-6)
    assert not torch.allclose(d[:, :8], d[:, 4:12])       # the two groups did not draw the same masks
    assert ((mean - ref).abs().mean() / ref.abs().mean()).item() < 0.08


def test_wide_layer_refuses_what_it_does_not_implement():
    case = _case("wide_12heads")
    layer = make_layer(case)
    layer.output_activation = "elu"
    with pytest.raises(NotImplementedError):
        layer(torch.from_numpy(case["x"]).cuda(), torch.from_numpy(case["edge_index"]).cuda())


# ---- the four-head head-mean backward kernel (csrc/edge_bwd_hm.cuh) over every chunk count it is instantiated for ----
@pytest.mark.parametrize("f", [18, 20, 29, 32, 36, 45, 47, 52, 61, 64])
@pytest.mark.parametrize("graph", ["adversarial", "products"])
def test_head_mean_four_heads_every_width(f, graph):
    """NH = 4, concat = False, F = 18 .. 64 (5 .. 16 float4 chunks per head: chunks-per-lane 2, 3, 4, exact and ragged), on the
    adversarial graph (duplicates, existing loops, isolated nodes) and on a products-shaped graph with hub rows (> 256 edges: the
    cooperative long-row launch runs next to the new kernel).  Fused backward (nothing consumes the attention) against the oracle."""
    from gat_pytorch_b200 import synth
    rng = np.random.default_rng(100 + f)
    if graph == "adversarial":
        x, ei = synth.adversarial()
    else:
        x, ei = synth.products(scale=1.0 / 512)
        x = x[:, :24].copy()
    f_in = x.shape[1]
    case = dict(name=f"hm4_{graph}_{f}", x=x, edge_index=ei, add_self_loops=True, const_attention=False, bias=None,
                W=synth.xavier_uniform(rng, 4 * f, f_in), a=synth.xavier_uniform(rng, 4, 8 * f), nh=4, f=f, concat=False)
    layer = make_layer(case)
    xt = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    out = layer(xt, torch.from_numpy(case["edge_index"]).cuda())
    fw = O.forward(case["x"], case["edge_index"].astype(np.int64), case["W"], case["a"], 4, f, False, True)
    go, _ = cases.upstream_grads(case, out.shape[0], out.shape[1], 1)
    (out * torch.from_numpy(go).cuda()).sum().backward()
    gr = O.backward(fw, go, None)
    errs = {"out": O.rel_err(out.detach().cpu().numpy(), fw["out"]), "gx": O.rel_err(xt.grad.cpu().numpy(), gr["x"]),
            "gW": O.rel_err(layer.W.weight.grad.cpu().numpy(), gr["W"]), "ga": O.rel_err(layer.a.weight.grad.cpu().numpy(), gr["a"])}
    assert all(e <= 1e-5 for e in errs.values()), errs


@pytest.mark.parametrize("name,nh,f,concat", [("narrow_1x1", 1, 1, False), ("narrow_1x7", 1, 7, False), ("narrow_2x3_mean", 2, 3, False),
                                              ("narrow_8x3_mean", 8, 3, False), ("narrow_4x4", 4, 4, True), ("narrow_3x5", 3, 5, True)])
@pytest.mark.parametrize("graph", ["adversarial", "pattern"])
def test_narrow_rows_with_widened_groups(name, nh, f, concat, graph):
    """Rows of 1 .. 16 chunks on SMALL graphs run with lane groups wider than the row needs (pick_group(chunks, n_rows)): every
    group width from the minimal one up to 32, on the 97-node adversarial graph (G = 32 everywhere) and a PATTERN batch; both
    backward families against the oracle."""
    from gat_pytorch_b200 import synth
    rng = np.random.default_rng(hash((name, graph)) % 2 ** 32)
    x, ei = synth.adversarial() if graph == "adversarial" else synth.pattern(graphs=24)
    f_in = x.shape[1]
    case = dict(name=f"{name}_{graph}", x=x, edge_index=ei, add_self_loops=True, const_attention=False, bias=None,
                W=synth.xavier_uniform(rng, nh * f, f_in), a=synth.xavier_uniform(rng, nh, 2 * nh * f), nh=nh, f=f, concat=concat)
    got, ei2 = run_cuda(case)
    fw, want = run_oracle(case)
    assert np.array_equal(ei2, fw["edge_index"])
    errs = {k: O.rel_err(got[k], want[k]) for k in want}
    tol = 5e-5 if (nh, f) == (1, 1) else 1e-5          # 1x1 rows: the reference's own fp32 run is 2.6e-5 from its fp64 run (pattern_L3)
    assert all(e <= tol for e in errs.values()), errs
    layer = make_layer(case)
    xt = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    out = layer(xt, torch.from_numpy(case["edge_index"]).cuda())
    go, _ = cases.upstream_grads(case, out.shape[0], out.shape[1], 1)
    (out * torch.from_numpy(go).cuda()).sum().backward()
    gr = O.backward(fw, go, None)
    errs = {"gx": O.rel_err(xt.grad.cpu().numpy(), gr["x"]), "gW": O.rel_err(layer.W.weight.grad.cpu().numpy(), gr["W"]),
            "ga": O.rel_err(layer.a.weight.grad.cpu().numpy(), gr["a"])}
    assert all(e <= tol for e in errs.values()), errs
