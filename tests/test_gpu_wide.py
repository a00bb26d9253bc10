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
