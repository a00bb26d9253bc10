"""Property tests (hypothesis) of the oracle's integer path and forward on random small graphs -- including, where the
reference checkout is present, a direct comparison with the REFERENCE's own functions (`models/utils.py`,
`models/gat_layer.py`) on every drawn input: the oracle is pinned on the reference beyond the committed golden cases."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

import gat_oracle as O

REF = "/root/reference"
HAVE_REF = os.path.isfile(os.path.join(REF, "models", "gat_layer.py"))


def _reference_modules():
    """The reference's `models` package next to whatever `models` is already importable (the overlay is not on sys.path in
    the test process)."""
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return importlib.import_module("models.utils"), importlib.import_module("models.gat_layer")


@st.composite
def edge_lists(draw):
    n = draw(st.integers(1, 24))
    e = draw(st.integers(1, 90))
    src = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    dst = draw(st.lists(st.integers(0, n - 1), min_size=e, max_size=e))
    return n, np.array([src, dst], dtype=np.int64)


@settings(max_examples=150, deadline=None)
@given(edge_lists())
def test_rewrite_and_csr_invariants(g):
    n, ei = g
    ei2 = O.add_remaining_self_loops(ei)
    n_idx = int(ei.max()) + 1
    keep = ei[0] != ei[1]
    assert np.array_equal(ei2[:, :keep.sum()], ei[:, keep])                       # kept edges, input order (utils.py:61,65)
    assert np.array_equal(ei2[0, keep.sum():], np.arange(n_idx)) and np.array_equal(ei2[1, keep.sum():], np.arange(n_idx))
    assert np.array_equal(O.add_remaining_self_loops(ei2), ei2)                   # idempotent (GATModel.py:166)
    rowptr, col, eid = O.csr_by_target(ei2, n)
    assert np.array_equal(np.diff(rowptr), O.in_degrees(ei2, n)) and rowptr[-1] == ei2.shape[1]
    assert np.array_equal(np.sort(eid), np.arange(ei2.shape[1]))
    d = ei2[1][eid]
    assert np.all((d[1:] > d[:-1]) | ((d[1:] == d[:-1]) & (eid[1:] > eid[:-1])))  # grouped by target, stable inside a row
    assert np.array_equal(col, ei2[0][eid])
    rowptr_t, col_t, pos_t = O.csr_by_source(ei2, n, eid)
    e_t = eid[pos_t]
    s = ei2[0][e_t]
    assert np.all((s[1:] > s[:-1]) | ((s[1:] == s[:-1]) & (e_t[1:] > e_t[:-1])))
    assert np.array_equal(col_t, ei2[1][e_t]) and np.array_equal(np.diff(rowptr_t), np.bincount(ei2[0], minlength=n))


@pytest.mark.skipif(not HAVE_REF, reason="reference checkout not present")
@settings(max_examples=60, deadline=None)
@given(edge_lists(), st.integers(1, 3), st.integers(1, 5), st.booleans(), st.integers(0, 2 ** 31 - 1))
def test_oracle_equals_the_reference_on_random_graphs(g, nh, f, concat, seed):
    """utils.add_remaining_self_loops and GATLayer.forward of the reference (run in fp64) against the oracle, input by input."""
    ref_utils, ref_layer = _reference_modules()
    n, ei = g
    ei_t = torch.from_numpy(ei)
    ref_ei2 = ref_utils.add_remaining_self_loops(ei_t)
    assert np.array_equal(ref_ei2.numpy(), O.add_remaining_self_loops(ei))
    n_rows = max(n, int(ei.max()) + 1)
    rng = np.random.default_rng(seed)
    f_in = 3
    x = rng.standard_normal((n_rows, f_in))
    torch.manual_seed(seed % 1000)
    layer = ref_layer.GATLayer(f_in, f, nh, concat, dropout=0, add_self_loops=True).double()
    layer.device = "cpu"
    xt = torch.from_numpy(x).requires_grad_(True)
    out, (ei_ret, alpha) = layer(xt, ei_t, return_attention_weights=True)
    fw = O.forward(x, ei, layer.W.weight.detach().numpy(), layer.a.weight.detach().numpy(), nh, f, concat, True)
    assert np.array_equal(ei_ret.numpy(), fw["edge_index"])
    assert O.rel_err(out.detach().numpy(), fw["out"]) < 1e-9
    assert O.rel_err(alpha.detach().numpy(), fw["alpha"]) < 1e-9
    # the backward the reference leaves to autograd -- duplicate edges in these draws tie the global max(), whose gradient
    # torch splits evenly over the arg-max set
    go, ga = rng.standard_normal(fw["out"].shape), rng.standard_normal(fw["alpha"].shape)
    ((out * torch.from_numpy(go)).sum() + (alpha * torch.from_numpy(ga)).sum()).backward()
    gr = O.backward(fw, go, ga)
    # (the residual ~2e-8 is nn.LeakyReLU holding its slope 0.01 as an fp32 constant)
    assert O.rel_err(xt.grad.numpy(), gr["x"]) < 1e-7
    assert O.rel_err(layer.W.weight.grad.numpy(), gr["W"]) < 1e-7
    assert O.rel_err(layer.a.weight.grad.numpy(), gr["a"]) < 1e-7
