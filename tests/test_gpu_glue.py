"""Output glue: skip add, ELU and the next layer's input dropout folded into the kernel that writes a layer's
output must equal the reference's op-by-op sequence (GATModel.py:126-149) -- value and every gradient, including the one
flowing into the skip branch.  The reference side is the SAME CUDA layer with the glue switched off and torch ops around it, so
what is tested is exactly the fusion; the layer itself is pinned against the oracle in test_gpu_parity.py."""
import types

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases
from test_gpu_parity import make_layer

pytestmark = pytest.mark.gpu

TOL = 5e-6


def _rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


@pytest.mark.parametrize("name", ["products_L1", "ppi_L1", "pattern_L1", "cora_L0", "adv_concat_oddF", "products_L2", "pubmed_L1",
                                  "adv_mean_oddF"])
@pytest.mark.parametrize("with_alpha", [False, True])
def test_skip_and_elu_in_the_output_kernel(name, with_alpha, small_cases):
    """y = ELU(layer(x) + skip): concat layers with F % 4 == 0 take the edge-kernel epilogue + rowdot adjoint, head-mean /
    padded layers the merge kernel + the element-wise adjoint; with_alpha adds an upstream dL/dalpha (three-pass backward)."""
    case = small_cases[name]
    ei = torch.from_numpy(case["edge_index"]).cuda()
    n = case["x"].shape[0]
    d_out = case["nh"] * case["f"] if case["concat"] else case["f"]
    rng = np.random.default_rng(7)
    skip0 = torch.from_numpy(rng.standard_normal((n, d_out)).astype(np.float32) * 0.5).cuda()
    res = []
    for fused in (False, True):
        layer = make_layer(case)
        x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
        skip = skip0.clone().requires_grad_(True)
        if fused:
            layer.output_activation = "elu"
            r = layer(x, ei, return_attention_weights=with_alpha, skip=skip)
        else:
            r = layer(x, ei, return_attention_weights=with_alpha)
        out, alpha = (r[0], r[1][1]) if with_alpha else (r, None)
        if not fused:
            out = F.elu(out + skip)
        go, ga = cases.upstream_grads(case, out.shape[0], out.shape[1], alpha.shape[0] if with_alpha else 1)
        loss = (out * torch.from_numpy(go).cuda()).sum()
        if with_alpha:
            loss = loss + (alpha * torch.from_numpy(ga).cuda()).sum()
        loss.backward()
        res.append({"out": out.detach(), "gx": x.grad, "gW": layer.W.weight.grad, "ga": layer.a.weight.grad, "gskip": skip.grad})
    for k in res[0]:
        assert _rel(res[1][k], res[0][k]) <= TOL, (name, with_alpha, k, _rel(res[1][k], res[0][k]))


@pytest.mark.parametrize("name", ["products_L1", "cora_L0", "products_L2", "adv_concat_oddF"])
def test_output_dropout_mask_is_bernoulli_and_shared_by_the_backward(name, small_cases):
    """y = dropout_p(ELU(layer(x) + skip)) in training mode: every stored element is 0 or E/(1-p), the keep rate is 1-p, two
    forwards draw different masks, eval mode applies none, and the backward uses the mask of ITS forward (gradients equal to
    autograd through the same composition with that mask as a constant)."""
    case = small_cases[name]
    p = 0.4
    ei = torch.from_numpy(case["edge_index"]).cuda()
    n = case["x"].shape[0]
    d_out = case["nh"] * case["f"] if case["concat"] else case["f"]
    rng = np.random.default_rng(11)
    skip0 = torch.from_numpy(rng.standard_normal((n, d_out)).astype(np.float32) * 0.5).cuda()
    layer = make_layer(case)
    layer.output_activation, layer.output_dropout = "elu", p
    layer.eval()
    with torch.no_grad():
        y0 = layer(torch.from_numpy(case["x"]).cuda(), ei, skip=skip0)           # eval: no dropout
    layer.train()
    x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    skip = skip0.clone().requires_grad_(True)
    y = layer(x, ei, skip=skip)
    big = y0.abs() > 1e-6
    ratio = (y.detach() / y0)[big]
    kept = ratio.abs() > 0.5
    assert torch.allclose(ratio[kept], torch.full_like(ratio[kept], 1.0 / (1.0 - p)), rtol=2e-5, atol=0)
    assert torch.all(ratio[~kept] == 0)
    rate = kept.double().mean().item()
    assert abs(rate - (1.0 - p)) < 4.0 * np.sqrt(p * (1 - p) / kept.numel()) + 1e-3, rate
    with torch.no_grad():
        y_again = layer(torch.from_numpy(case["x"]).cuda(), ei, skip=skip0)
    assert not torch.equal(y_again != 0, y.detach() != 0), "two training forwards drew the same mask"
    go, _ = cases.upstream_grads(case, n, d_out, 1)
    # elements with |E| <= 1e-6 cannot be told kept from dropped from the outside: they get no upstream gradient on either side
    go = torch.where(big, torch.from_numpy(go).cuda(), torch.zeros_like(y0))
    (y * go).sum().backward()
    got = {"gx": x.grad, "gW": layer.W.weight.grad, "ga": layer.a.weight.grad, "gskip": skip.grad}
    # the same composition in torch with the drawn mask as a constant
    mask = torch.where(big, (y.detach() != 0).float(), torch.ones_like(y0)) / (1.0 - p)
    go_eff = go
    ref = make_layer(case)
    ref.train()
    x2 = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    skip2 = skip0.clone().requires_grad_(True)
    (F.elu(ref(x2, ei) + skip2) * mask * go_eff).sum().backward()
    # (a fresh fused pass would draw a fresh mask, hence the torch composition with the drawn mask as the reference)
    want = {"gx": x2.grad, "gW": ref.W.weight.grad, "ga": ref.a.weight.grad, "gskip": skip2.grad}
    for k in want:
        assert _rel(got[k], want[k]) <= 2e-5, (name, k, _rel(got[k], want[k]))


def _toy_model(shapes, add_skip, dropout, seed=3):
    """A stand-in for the attributes of `GATModel` that `model_forward` reads (GATModel.py:42-116): B200 layers, the
    reference's Identity / Linear skip layers."""
    from gat_pytorch_b200 import GATLayer
    torch.manual_seed(seed)
    layers, skips = [], []
    nh_per, f_per, concat_per = [1], [shapes[0][0]], []
    for i, (f_in, nh, f, concat) in enumerate(shapes):
        layers.append(GATLayer(f_in, f, nh, concat, dropout=dropout, add_self_loops=True, bias=False))
        nh_per.append(nh)
        f_per.append(f)
        concat_per.append(concat)
        if add_skip[i]:
            skips.append(torch.nn.Identity() if f_in == nh * f else torch.nn.Linear(f_in, nh * f, bias=False))
    m = torch.nn.Module()
    m.gat_layer_list = torch.nn.ModuleList(layers)
    m.skip_layer_list = torch.nn.ModuleList(skips)
    m.add_skip_connection, m.heads_concat_per_layer = list(add_skip), concat_per
    m.num_heads_per_layer, m.head_output_features_per_layer, m.dropout = nh_per, f_per, dropout
    return m.cuda()


def _reference_forward(model, x, edge_index, want_attention):
    """GATModel.forward / forward_and_return_attention (GATModel.py:118-187), op by op, on the same modules."""
    attention, skip_count = [], 0
    n_layers = len(model.gat_layer_list)
    for i in range(n_layers):
        layer_input = x
        x = F.dropout(x, p=model.dropout, training=model.training)
        if want_attention:
            x, (edge_index, alpha) = model.gat_layer_list[i](x, edge_index, return_attention_weights=True)
            attention.append(alpha)
        else:
            x = model.gat_layer_list[i](x, edge_index)
        if model.add_skip_connection[i]:
            skip_output = model.skip_layer_list[skip_count](layer_input)
            skip_count += 1
            if model.heads_concat_per_layer[i]:
                x = x + skip_output
            else:
                skip_output = skip_output.view(-1, model.num_heads_per_layer[i + 1], model.head_output_features_per_layer[i + 1])
                x = x + skip_output.mean(dim=1)
        if i != n_layers - 1:
            x = F.elu(x)
    return x, edge_index, attention


@pytest.mark.parametrize("config", ["ppi", "pattern", "cora"])
@pytest.mark.parametrize("want_attention", [False, True])
def test_model_forward_equals_the_reference_sequence(config, want_attention):
    """glue.model_forward == GATModel.forward / forward_and_return_attention on the PPI (identity skip on layer 1, head-mean
    output), PATTERN (Linear skips everywhere, head-mean 1x1 output) and Cora stacks, eval mode and training mode with p = 0:
    output, attention, and the gradients of every parameter (skip projections included) and of the input."""
    from gat_pytorch_b200 import model_forward, synth
    shapes = {"ppi": [(50, 4, 64, True), (256, 4, 64, True), (256, 6, 121, False)],
              "pattern": [(3, 4, 12, True), (48, 4, 24, True), (96, 4, 12, True), (48, 1, 1, False)],
              "cora": [(1433, 8, 8, True), (64, 1, 7, False)]}[config]
    add_skip = {"ppi": [False, True, False], "pattern": [True, True, True, True], "cora": [False, False]}[config]
    x0, ei = {"ppi": lambda: synth.ppi(), "pattern": lambda: synth.pattern(graphs=8), "cora": lambda: synth.cora()}[config]()
    x0 = torch.from_numpy(np.ascontiguousarray(x0[:, :shapes[0][0]], dtype=np.float32)).cuda()
    ei = torch.from_numpy(ei).cuda()
    model = _toy_model(shapes, add_skip, dropout=0.0)
    data = types.SimpleNamespace(x=None, edge_index=ei)
    res = []
    for fused in (False, True):
        model.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        data.x = x
        if fused:
            r = model_forward(model, data, True if want_attention else None)
            out, att = (r[0], r[2]) if want_attention else (r, [])
        else:
            out, _, att = _reference_forward(model, x, ei, want_attention)
        g = torch.Generator(device="cuda").manual_seed(5)
        loss = (out * torch.randn(out.shape, device="cuda", generator=g)).sum()
        for a in att:
            loss = loss + (a * torch.randn(a.shape, device="cuda", generator=g)).sum() * 0.1
        loss.backward()
        grads = {f"g:{k}": v.grad.clone() for k, v in model.named_parameters()}
        res.append(dict(out=out.detach(), gx=x.grad.clone(), **grads, **{f"alpha{i}": a.detach() for i, a in enumerate(att)}))
    assert set(res[0]) == set(res[1])
    for k in res[0]:
        assert _rel(res[1][k], res[0][k]) <= 2e-5, (config, want_attention, k, _rel(res[1][k], res[0][k]))
    # the layers come back with their own settings (the unchanged callers use the same modules)
    assert all(l.output_activation is None and l.output_dropout == 0.0 for l in model.gat_layer_list)


def test_model_forward_training_dropout_statistics():
    """Cora stack in training mode (p = 0.6 on the inputs AND on the attention, run_config.py): the folded dropout keeps the
    expectation -- the mean output over many draws approaches the reference sequence's mean over many draws."""
    from gat_pytorch_b200 import model_forward, synth
    shapes, add_skip = [(1433, 8, 8, True), (64, 1, 7, False)], [False, False]
    x0, ei = synth.cora()
    x0, ei = torch.from_numpy(x0).cuda(), torch.from_numpy(ei).cuda()
    model = _toy_model(shapes, add_skip, dropout=0.6)
    model.train()
    data = types.SimpleNamespace(x=x0, edge_index=ei)
    torch.manual_seed(0)
    with torch.no_grad():
        reps = 200
        a = sum(model_forward(model, data) for _ in range(reps)) / reps
        b = sum(_reference_forward(model, x0, ei, False)[0] for _ in range(reps)) / reps
        spread = torch.stack([_reference_forward(model, x0, ei, False)[0] for _ in range(20)]).std(dim=0).mean().item()
    # two independent means of `reps` draws differ by ~ spread * sqrt(2 / reps) per element
    assert (a - b).abs().mean().item() < 3.0 * spread * np.sqrt(2.0 / reps), ((a - b).abs().mean().item(), spread)


def test_glue_refuses_what_it_does_not_implement(small_cases):
    case = small_cases["adv_bias"]
    layer = make_layer(case)
    x = torch.from_numpy(case["x"]).cuda()
    ei = torch.from_numpy(case["edge_index"]).cuda()
    d_out = case["nh"] * case["f"] if case["concat"] else case["f"]
    with pytest.raises(ValueError):
        layer(x, ei, skip=torch.zeros(x.shape[0], d_out, device="cuda"))        # bias sits between the layer and the glue
    case = small_cases["adv_concat"]
    layer = make_layer(case)
    with pytest.raises(RuntimeError):
        layer(torch.from_numpy(case["x"]).cuda(), torch.from_numpy(case["edge_index"]).cuda(),
              skip=torch.zeros(3, 3, device="cuda"))                              # wrong shape
    layer.output_dropout = 1.5
    layer.train()
    with pytest.raises(ValueError):
        layer(torch.from_numpy(case["x"]).cuda(), torch.from_numpy(case["edge_index"]).cuda())


@pytest.mark.parametrize("name", ["adv_concat", "adv_mean_oddF", "adv_ties", "cora_L0", "cora_L1", "pubmed_L1", "ppi_L1", "ppi_L2",
                                  "pattern_L1", "products_L1", "products_L2"])
def test_fused_attention_norm_matches_the_materialised_one(name, small_cases):
    """SURVEY.md 8-f3, fused form: layer.attention_norm computes sum |alpha*deg - 1| / E' from the score terms and its gradient
    rides in the one-pass backward.  Must equal attention_norm(edge_index', [alpha]) on the returned attention (itself pinned
    against GATModel.calc_attention_norm in test_gpu_parity.py) in value and in every gradient of task loss + lambda * norm."""
    from gat_pytorch_b200 import attention_norm
    case = small_cases[name]
    ei = torch.from_numpy(case["edge_index"]).cuda()
    lam = 3.0
    res = []
    for fused in (False, True):
        layer = make_layer(case)
        x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
        if fused:
            layer.attention_norm = True
            out = layer(x, ei)
            norm = layer.attention_norm_value
        else:
            out, (ei2, alpha) = layer(x, ei, return_attention_weights=True)
            norm = attention_norm(ei2, [alpha])
        go, _ = cases.upstream_grads(case, out.shape[0], out.shape[1], 1)
        ((out * torch.from_numpy(go).cuda()).sum() + lam * norm).backward()
        res.append({"norm": norm.detach().reshape(1), "out": out.detach(), "gx": x.grad, "gW": layer.W.weight.grad, "ga": layer.a.weight.grad})
    for k in res[0]:
        assert _rel(res[1][k], res[0][k]) <= 1e-5, (name, k, _rel(res[1][k], res[0][k]))
    # the regulariser really reached the gradients (the comparison above is not between two copies of the task gradient)
    layer = make_layer(case)
    x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    out = layer(x, ei)
    (out * torch.from_numpy(go).cuda()).sum().backward()
    assert _rel(layer.a.weight.grad, res[1]["ga"]) > 1e-3


def test_model_forward_with_fused_norm_equals_calc_attention_norm():
    """The PPI training step's ingredients (ppi_gat.py:22-33): forward_and_return_attention + calc_attention_norm, against
    model_forward(..., attention_norm=True) which needs no attention tensor."""
    from gat_pytorch_b200 import attention_norm, model_forward, synth
    shapes, add_skip = [(50, 4, 64, True), (256, 4, 64, True), (256, 6, 121, False)], [False, True, False]
    x0, ei = synth.ppi()
    x0, ei = torch.from_numpy(x0).cuda(), torch.from_numpy(ei).cuda()
    model = _toy_model(shapes, add_skip, dropout=0.0)
    res = []
    for fused in (False, True):
        model.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        data = types.SimpleNamespace(x=x, edge_index=ei)
        if fused:
            out, norm = model_forward(model, data, attention_norm=True)
        else:
            out, ei2, att = _reference_forward(model, x, ei, True)
            norm = attention_norm(ei2, att)
        g = torch.Generator(device="cuda").manual_seed(5)
        ((out * torch.randn(out.shape, device="cuda", generator=g)).sum() + 2.0 * norm).backward()
        res.append(dict(out=out.detach(), norm=norm.detach().reshape(1), gx=x.grad.clone(),
                        **{f"g:{k}": v.grad.clone() for k, v in model.named_parameters()}))
    for k in res[0]:
        assert _rel(res[1][k], res[0][k]) <= 2e-5, (k, _rel(res[1][k], res[0][k]))


def test_micro_f1_equals_sklearn():
    """gat_pytorch_b200.micro_f1 == the call PPI_GAT makes every step (ppi_gat.py:38), on PPI-shaped logits / labels and the edge
    cases (nothing predicted, nothing true)."""
    from sklearn.metrics import f1_score
    from gat_pytorch_b200 import micro_f1
    g = torch.Generator().manual_seed(3)
    for n, c, bias in ((4800, 121, 0.0), (4800, 121, -1.5), (97, 7, 0.3), (1, 1, 0.0)):
        out = torch.randn((n, c), generator=g) + bias
        y = (torch.rand((n, c), generator=g) < 0.3).float()
        want = f1_score(y_pred=out.numpy() > 0, y_true=y.numpy(), average="micro", zero_division=0)
        assert abs(micro_f1(out.cuda(), y.cuda()) - want) < 1e-12, (n, c, bias)
    z = torch.zeros((10, 5))
    assert micro_f1((z - 1).cuda(), z.cuda()) == 0.0
