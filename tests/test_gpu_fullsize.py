"""BASELINE.json's full-size configuration (ogbn-products-shaped: 2 449 029 nodes, 61 859 140 edges) through
size-independent properties, plus a mid-size parity run against the torch port of the reference formulation
executed in fp64 on the same GPU (the reference's own formulation cannot allocate the full graph).

Integer work (Kernel 1) is checked bit-exactly with independent torch ops; floating point within the 1e-5
tensor-relative bar of BASELINE.json's north_star.  Everything goes through the drop-in GATLayer / the C ABI.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _rel(got, want):
    return ((got.double() - want.double()).abs().max() / want.double().abs().max().clamp(min=1e-30)).item()


@pytest.fixture(scope="module")
def products_full():
    import gat_pytorch_b200 as g
    x, ei = g.synth.products(scale=1.0)
    return torch.from_numpy(x).cuda(), torch.from_numpy(ei).cuda()


def test_full_size_structure_is_the_reference_rewrite_and_a_stable_sort(products_full):
    """utils.py:47-72 + SURVEY.md 9.3 at full size: rewritten list, degree counts, stable CSR and CSR^T, inverse maps,
    idempotence -- each checked with independent torch integer ops, bit-exact."""
    import gat_pytorch_b200 as g
    x, ei = products_full
    n = x.size(0)
    st = g.build_structure(ei, n, True)
    ei2 = st.edge_index
    n_idx = int(ei.max().item()) + 1
    keep = ei[0] != ei[1]
    n_keep = int(keep.sum().item())
    assert st.n_edges == n_keep + n_idx == ei2.size(1)
    assert torch.equal(ei2[:, :n_keep], ei[:, keep])                                   # kept edges, input order
    loops = torch.arange(n_idx, device="cuda", dtype=ei.dtype)
    assert torch.equal(ei2[0, n_keep:], loops) and torch.equal(ei2[1, n_keep:], loops)  # then the loops 0..N_idx-1
    # degree counts == GATModel.py:196-201
    deg = torch.bincount(ei2[1], minlength=n)
    assert torch.equal((st.rowptr[1:] - st.rowptr[:-1]).long(), deg)
    assert int(st.rowptr[0]) == 0 and int(st.rowptr[-1]) == st.n_edges
    # CSR by target: a permutation, grouped by target, stable inside a row, col = src
    eid = st.eid.long()
    assert torch.equal(torch.sort(eid).values, torch.arange(st.n_edges, device="cuda"))
    dst_sorted = ei2[1][eid]
    assert bool(((dst_sorted[1:] > dst_sorted[:-1]) | ((dst_sorted[1:] == dst_sorted[:-1]) & (eid[1:] > eid[:-1]))).all())
    assert torch.equal(st.col.long(), ei2[0][eid])
    # CSR by source: slot j holds the edge eid[pos_t[j]]
    pos_t = st.pos_t.long()
    e_t = eid[pos_t]
    src_sorted = ei2[0][e_t]
    assert bool(((src_sorted[1:] > src_sorted[:-1]) | ((src_sorted[1:] == src_sorted[:-1]) & (e_t[1:] > e_t[:-1]))).all())
    assert torch.equal(st.col_t.long(), ei2[1][e_t])
    assert torch.equal((st.rowptr_t[1:] - st.rowptr_t[:-1]).long(), torch.bincount(ei2[0], minlength=n))
    assert torch.equal(pos_t[st.tpos.long()], torch.arange(st.n_edges, device="cuda"))
    # idempotence (GATModel.py:166 feeds the rewritten list to the next layer)
    st2 = g.build_structure(ei2, n, True)
    assert torch.equal(st2.edge_index, ei2)
    assert torch.equal(st2.rowptr, st.rowptr) and torch.equal(st2.col, st.col)


def test_full_size_layer_properties(products_full):
    """Hidden-layer shape (256 -> 4 x 64, concat) on the full graph: attention rows sum to Z/(Z+1e-8); sampled output rows
    equal sum alpha*Wh[src] recomputed in fp64; the backward is linear in the upstream gradient and bitwise deterministic."""
    import gat_pytorch_b200 as g
    _, ei = products_full
    n = int(ei.max().item()) + 1
    torch.manual_seed(0)
    x = torch.randn(n, 256, device="cuda")
    layer = g.GATLayer(256, 64, 4, True, dropout=0.0, add_self_loops=True).cuda()
    xg = x.clone().requires_grad_(True)
    out, (ei2, alpha) = layer(xg, ei, return_attention_weights=True)
    assert torch.isfinite(out).all() and torch.isfinite(alpha).all()
    # every node has a self-loop, Z >= exp(0.01*(l - M)) of it: the row sums are 1 - 1e-8/(Z + 1e-8)
    row_sum = torch.zeros(n, 4, device="cuda", dtype=torch.float64).index_add_(0, ei2[1], alpha.double())
    assert (row_sum - 1.0).abs().max().item() < 1e-5
    assert alpha.min().item() >= 0.0
    # sampled rows (the 8 largest in-degrees and 2000 random ones) against fp64 sum alpha*Wh[src]
    deg = torch.bincount(ei2[1], minlength=n)
    rows = torch.cat([torch.topk(deg, 8).indices, torch.randint(0, n, (2000,), device="cuda")]).unique()
    mask = torch.zeros(n, dtype=torch.bool, device="cuda")
    mask[rows] = True
    sel = mask[ei2[1]]
    src, dst, al = ei2[0][sel], ei2[1][sel], alpha[sel].double()
    wh = (x.double() @ layer.W.weight.detach().double().t()).view(n, 4, 64)
    want = torch.zeros(n, 4, 64, device="cuda", dtype=torch.float64).index_add_(0, dst, al[:, :, None] * wh[src])
    assert _rel(out.detach().view(n, 4, 64)[rows], want[rows]) < TOL
    del wh, want
    # backward: linear in the upstream gradient, deterministic
    go = torch.randn_like(out)

    def grads(scale):
        layer.W.weight.grad = layer.a.weight.grad = None
        xi = x.clone().requires_grad_(True)
        o = layer(xi, ei)
        (o * (go * scale)).sum().backward()
        return xi.grad, layer.W.weight.grad.clone(), layer.a.weight.grad.clone()

    g1, g1b, g2 = grads(1.0), grads(1.0), grads(2.0)
    for a, b in zip(g1, g1b):
        assert torch.equal(a, b)                       # bitwise reproducible
    for a, b in zip(g1, g2):
        assert _rel(b, 2.0 * a) < 1e-6
    assert all(torch.isfinite(t).all() for t in g1)


@pytest.mark.parametrize("layer_idx", [0, 1, 2])
def test_mid_size_products_matches_the_reference_formulation_in_fp64(layer_idx):
    """products-shaped graph at 1/32 scale (76 532 nodes, ~2.0 M edges): every layer shape of the config against the torch
    port of gat_layer.py:53-140 run in fp64 on the GPU with autograd (out, alpha, dx, dW, da), 1e-5 tensor-relative."""
    import gat_pytorch_b200 as g
    import torch_port
    x_np, ei_np = g.synth.products(scale=1.0 / 32)
    f_in, nh, f, concat = g.synth.LAYER_SHAPES["products"][layer_idx]
    w_np, a_np = g.synth.seeded_weights("products")[layer_idx]
    ei = torch.from_numpy(ei_np).cuda()
    n = x_np.shape[0]
    torch.manual_seed(layer_idx)
    x = torch.from_numpy(x_np).cuda() if layer_idx == 0 else torch.randn(n, f_in, device="cuda")
    layer = g.GATLayer(f_in, f, nh, concat, dropout=0.0, add_self_loops=True).cuda()
    with torch.no_grad():
        layer.W.weight.copy_(torch.from_numpy(w_np))
        layer.a.weight.copy_(torch.from_numpy(a_np))
    xg = x.clone().requires_grad_(True)
    out, (ei2, alpha) = layer(xg, ei, return_attention_weights=True)
    go = torch.randn_like(out)
    (out * go).sum().backward()
    # the reference formulation, fp64
    xr = x.double().requires_grad_(True)
    wr = layer.W.weight.detach().double().requires_grad_(True)
    ar = layer.a.weight.detach().double().requires_grad_(True)
    out_r, ei_r, alpha_r = torch_port.layer_forward(xr, ei, wr, ar, nh, f, concat, True)
    (out_r * go.double()).sum().backward()
    assert torch.equal(ei2, ei_r)
    assert _rel(out.detach(), out_r.detach()) < TOL
    assert _rel(alpha.detach(), alpha_r.detach()) < TOL
    assert _rel(xg.grad, xr.grad) < TOL
    assert _rel(layer.W.weight.grad, wr.grad) < TOL
    assert _rel(layer.a.weight.grad, ar.grad) < TOL
