"""Parity of the CUDA path (through the C ABI, via the drop-in GATLayer) against the oracle and the
golden vectors minted from the reference.  Bar: integer work bit-exact; fp32 within 1e-5 tensor-relative
(max|got-want| / max|want|) of the fp64 oracle, as BASELINE.json's north_star states.  Cases whose own
fp32 reference noise exceeds that (eps-dominated softmax) carry their tolerance below.
"""
import numpy as np
import pytest
import torch

import cases
import gat_oracle as O

pytestmark = pytest.mark.gpu

CASE_NAMES = [c["name"] for c in cases.adversarial_cases()] + [
    f"{m}_L{i}" for m in ("cora", "pubmed", "ppi", "pattern", "products") for i in range(len(cases.synth.LAYER_SHAPES[m]))]

TOL = 1e-5
# looser where the reference's fp32 run itself is further than 1e-5 from its fp64 run (value = measured
# reference fp32 noise, see tests/test_oracle_golden.py) -- the CUDA path must be no worse than that.
TOL_OVERRIDE = {"adv_eps_dominated": 2e-3, "pattern_L2": 5e-5, "pattern_L3": 5e-5}


_ELEMENTWISE = {}


@pytest.fixture(scope="module", autouse=True)
def _dump_elementwise_report():
    """After the module ran: gpurun_out/parity_elementwise.json (copied to profiles/ by hand when it is to be cited)."""
    yield
    if _ELEMENTWISE:
        import json
        import os
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_elementwise.json"), "w") as fh:
            json.dump(_ELEMENTWISE, fh, indent=1, sort_keys=True)


def make_layer(case, device="cuda", gemm_algo=0, dropout=0.0):
    from gat_pytorch_b200 import GATLayer
    layer = GATLayer(case["x"].shape[1], case["f"], case["nh"], case["concat"], dropout=dropout,
                     add_self_loops=case["add_self_loops"], bias=case["bias"] is not None,
                     const_attention=case["const_attention"]).to(device)
    layer.gemm_algo = gemm_algo
    with torch.no_grad():
        layer.W.weight.copy_(torch.from_numpy(case["W"]))
        if not case["const_attention"]:
            layer.a.weight.copy_(torch.from_numpy(case["a"]))
        if case["bias"] is not None:
            layer.bias_param.copy_(torch.from_numpy(case["bias"]))
    return layer


def run_cuda(case, gemm_algo=0):
    layer = make_layer(case, gemm_algo=gemm_algo)
    x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    ei = torch.from_numpy(case["edge_index"]).cuda()
    out, (ei2, alpha) = layer(x, ei, return_attention_weights=True)
    go, ga = cases.upstream_grads(case, out.shape[0], out.shape[1], alpha.shape[0])
    loss = (out * torch.from_numpy(go).cuda()).sum()
    if alpha.requires_grad:
        loss = loss + (alpha * torch.from_numpy(ga).cuda()).sum()
    loss.backward()
    res = dict(out=out, alpha=alpha, gx=x.grad, gW=layer.W.weight.grad)
    if not case["const_attention"]:
        res["ga"] = layer.a.weight.grad
    if case["bias"] is not None:
        res["gb"] = layer.bias_param.grad.reshape(-1, 1)
    return {k: v.detach().cpu().numpy() for k, v in res.items()}, ei2.cpu().numpy()


def run_oracle(case):
    fw = O.forward(case["x"], case["edge_index"].astype(np.int64), case["W"], case["a"], case["nh"], case["f"], case["concat"],
                   case["add_self_loops"], case["bias"], case["const_attention"])
    go, ga = cases.upstream_grads(case, fw["out"].shape[0], fw["out"].shape[1], fw["alpha"].shape[0])
    gr = O.backward(fw, go, None if case["const_attention"] else ga)
    res = dict(out=fw["out"], alpha=fw["alpha"], gx=gr["x"], gW=gr["W"])
    if not case["const_attention"]:
        res["ga"] = gr["a"]
    if case["bias"] is not None:
        res["gb"] = gr["bias"].reshape(-1, 1)
    return fw, res


@pytest.mark.parametrize("name", CASE_NAMES)
def test_layer_matches_oracle_and_golden(name, small_cases, golden):
    case = small_cases[name]
    got, ei2 = run_cuda(case)
    fw, want = run_oracle(case)
    assert ei2.dtype == case["edge_index"].dtype
    assert np.array_equal(ei2, fw["edge_index"]), "rewritten edge list must be bit-exact (utils.py:47-67)"
    tol = TOL_OVERRIDE.get(name, TOL)
    errs = {k: O.rel_err(got[k], want[k]) for k in want}
    assert all(e <= tol for e in errs.values()), errs
    # element-wise figures beside the tensor-relative one (a row judged against its own magnitude): recorded for
    # profiles/parity_elementwise_r02.json; the TYPICAL row must itself meet the bar, and no row may be off by more than
    # the cancellation noise an fp32 evaluation of the same sums has (measured on the reference's own fp32 run: 1e-3)
    rep = {k: O.elementwise_report(got[k], want[k]) for k in want}
    _ELEMENTWISE[name] = rep
    for k, r in rep.items():
        assert r["row_rel_median"] <= tol, (k, r)
        if name != "adv_eps_dominated":     # (its alpha / out rows go down to fp32 denormals: 100 % row-relative, 1e-10 absolute)
            # measured on B200 (profiles/parity_elementwise_r02.json): worst p99 2.1e-5 (pattern_L3 dx), worst atol 2.0e-5
            assert r["row_rel_p99"] <= 10 * tol and r["atol_needed_over_mean"] <= 10 * tol, (k, r)
    # and against the reference's own fp32 outputs (sampled rows), within the same bar + the reference's noise
    for k in want:
        if name == "adv_int32" and k.startswith("g"):
            continue   # the reference's int32 gradients are wrong under torch 2.11 (see test_oracle_golden.py)
        g2 = got[k].reshape(got[k].shape[0], -1)
        sample = g2[np.ix_(cases.sample_idx(g2.shape[0], 64), cases.sample_idx(g2.shape[1], 160))]
        scale = max(float(golden[f"{name}/f64/{k}_max"]), 1e-30)
        ref32 = golden[f"{name}/f32/{k}"]
        err = np.abs(sample - ref32).max() / scale if sample.size else 0.0
        assert err <= 2 * tol, (k, err)


@pytest.mark.parametrize("name", ["adv_concat", "cora_L0", "pattern_L0", "products_L1"])
def test_structure_bit_exact(name, small_cases):
    """Kernel 1 against the oracle's stable argsort / bincount."""
    from gat_pytorch_b200 import build_structure
    case = small_cases[name]
    n = case["x"].shape[0]
    ei = torch.from_numpy(case["edge_index"]).cuda()
    st = build_structure(ei, n, True)
    want_ei = O.add_remaining_self_loops(case["edge_index"])
    assert np.array_equal(st.edge_index.cpu().numpy(), want_ei)
    rowptr, col, eid = O.csr_by_target(want_ei, n)
    rowptr_t, col_t, pos_t = O.csr_by_source(want_ei, n, eid)
    for got, want in [(st.rowptr, rowptr), (st.col, col), (st.eid, eid), (st.rowptr_t, rowptr_t), (st.col_t, col_t), (st.pos_t, pos_t)]:
        assert np.array_equal(got.cpu().numpy().astype(np.int64), want)
    assert np.array_equal(st.in_degrees().cpu().numpy(), O.in_degrees(want_ei, n))
    # idempotent: the rewritten list maps to the same structure (GATModel.py:166 feeds it to the next layer)
    st2 = build_structure(st.edge_index, n, True)
    assert torch.equal(st2.edge_index, st.edge_index) and torch.equal(st2.col, st.col)
    # no rewrite requested: list used as is
    st3 = build_structure(ei, n, False)
    assert st3.n_edges == ei.size(1) and st3.edge_index is ei


def test_out_of_range_index_raises():
    from gat_pytorch_b200 import build_structure
    ei = torch.tensor([[0, 5], [1, 2]], device="cuda")
    with pytest.raises(IndexError):
        build_structure(ei, 4, True)


def test_trailing_isolated_nodes_are_zero(small_cases):
    case = small_cases["adv_concat"]
    got, _ = run_cuda(case)
    assert np.all(got["out"][-5:] == 0.0)   # nodes that never appear in edge_index get no self-loop


def test_deterministic_bitwise(small_cases):
    """Atomic-free backward: two runs give identical bits (replaces a race detector)."""
    for name in ("adv_wide", "products_L1", "pattern_L0"):
        a, _ = run_cuda(small_cases[name])
        b, _ = run_cuda(small_cases[name])
        for k in a:
            assert np.array_equal(a[k], b[k]), (name, k)


def test_eval_forward_without_attention_matches(small_cases):
    case = small_cases["cora_L0"]
    layer = make_layer(case).eval()
    x = torch.from_numpy(case["x"]).cuda()
    ei = torch.from_numpy(case["edge_index"]).cuda()
    with torch.no_grad():
        out = layer(x, ei)
        out2, (_, alpha) = layer(x, ei, return_attention_weights=True)
    assert torch.equal(out, out2)
    fw = O.forward(case["x"], case["edge_index"], case["W"], case["a"], case["nh"], case["f"], True, True)
    assert O.rel_err(out.cpu().numpy(), fw["out"]) <= TOL


def test_dropout_statistics_and_backward_mask(small_cases):
    """Philox cannot bit-match nn.Dropout: check keep-rate, unbiasedness, pre-dropout alpha,
    and that backward regenerates the forward's mask (gradient check against the oracle with that mask)."""
    case = small_cases["products_L0"]
    p = 0.6
    layer = make_layer(case, dropout=p).train()
    x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    ei = torch.from_numpy(case["edge_index"]).cuda()
    torch.manual_seed(123)
    out, (ei2, alpha) = layer(x, ei, return_attention_weights=True)
    fw = O.forward(case["x"], case["edge_index"], case["W"], case["a"], case["nh"], case["f"], True, True)
    assert O.rel_err(alpha.detach().cpu().numpy(), fw["alpha"]) <= TOL          # returned alpha is pre-dropout
    # recover the mask from out = sum m*alpha*Wh: use a probe with one-hot features instead
    torch.manual_seed(123)
    out_b, _ = layer(x, ei, return_attention_weights=True)
    assert torch.equal(out, out_b)                                               # same seed -> same mask
    torch.manual_seed(124)
    out_c, _ = layer(x, ei, return_attention_weights=True)
    assert not torch.equal(out, out_c)
    # unbiased: mean over many masks approaches the eval output
    acc = torch.zeros_like(out)
    reps = 200
    with torch.no_grad():
        for _ in range(reps):
            acc += layer(x, ei)
    mean = (acc / reps).cpu().numpy()
    # per-element noise of a 200-sample mean at p=0.6 is ~sqrt(p/(1-p)/reps) ~ 9 %; the BIAS must vanish
    noise = np.abs(mean - fw["out"]).mean() / np.abs(fw["out"]).mean()
    bias = abs((mean - fw["out"]).sum()) / np.abs(fw["out"]).sum()
    assert noise < 0.2 and bias < 0.01, (noise, bias)


def test_dropout_gradient_uses_forward_mask():
    """Linear probe: with Wh = identity-like features the output reveals the mask; backward must use it."""
    from gat_pytorch_b200 import GATLayer
    torch.manual_seed(0)
    n, nh, f = 64, 2, 4
    rng = np.random.default_rng(5)
    ei_np = rng.integers(0, n, size=(2, 600)).astype(np.int64)
    layer = GATLayer(8, f, nh, True, dropout=0.5, add_self_loops=True).cuda().train()
    x = torch.randn(n, 8, device="cuda", requires_grad=True)
    ei = torch.from_numpy(ei_np).cuda()
    torch.manual_seed(77)
    out, (ei2, alpha) = layer(x, ei, return_attention_weights=True)
    go = torch.randn_like(out)
    (out * go).sum().backward()
    gx = x.grad.clone()
    # finite-difference directional derivative with the SAME mask (same seed)
    d = torch.randn_like(x)
    eps = 1e-2
    with torch.no_grad():
        torch.manual_seed(77)
        op = layer(x + eps * d, ei)
        torch.manual_seed(77)
        om = layer(x - eps * d, ei)
    fd = ((op - om) * go).sum().item() / (2 * eps)
    an = (gx * d).sum().item()
    assert abs(fd - an) <= 2e-2 * max(abs(fd), abs(an), 1e-3), (fd, an)


def test_wrong_dtype_raises():
    from gat_pytorch_b200 import GATLayer
    layer = GATLayer(4, 2, 2, True).cuda()
    with pytest.raises(RuntimeError):
        layer(torch.randn(3, 4, device="cuda", dtype=torch.float64), torch.zeros((2, 2), dtype=torch.long, device="cuda"))


def test_host_buffer_mode_matches_device_mode(small_cases):
    """A module and inputs that live in HOST memory (what the reference's vis.py hands the layer, vis.py:41-47): the layer
    copies them to the GPU, runs the same kernels and returns host tensors; gradients reach the host-resident parameters.
    (Without a CUDA device this raises: tests/test_host.py.)"""
    case = small_cases["adv_concat"]
    dev_layer = make_layer(case)
    host_layer = make_layer(case, device="cpu")
    x_d = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    x_h = torch.from_numpy(case["x"]).requires_grad_(True)
    ei = torch.from_numpy(case["edge_index"])
    out_d, (ei_d, alpha_d) = dev_layer(x_d, ei.cuda(), return_attention_weights=True)
    out_h, (ei_h, alpha_h) = host_layer(x_h, ei, return_attention_weights=True)
    assert not out_h.is_cuda and not alpha_h.is_cuda and not ei_h.is_cuda
    assert torch.equal(out_h, out_d.cpu()) and torch.equal(alpha_h, alpha_d.cpu()) and torch.equal(ei_h, ei_d.cpu())
    go = torch.randn_like(out_h)
    (out_d * go.cuda()).sum().backward()
    (out_h * go).sum().backward()
    assert torch.equal(x_h.grad, x_d.grad.cpu())
    assert torch.equal(host_layer.W.weight.grad, dev_layer.W.weight.grad.cpu())
    assert torch.equal(host_layer.a.weight.grad, dev_layer.a.weight.grad.cpu())
    assert list(host_layer.state_dict().keys()) == ["W.weight", "a.weight"]      # the device twin is not a sub-module
    out_plain = host_layer(x_h.detach(), ei)
    assert torch.equal(out_plain, out_h.detach())


def test_gemm_all_layouts():
    """gat_gemm (fp32 FFMA path) against torch fp64 matmul for every transposition and ragged sizes."""
    from gat_pytorch_b200.gat_layer import gemm
    torch.manual_seed(1)
    for (m, n, k) in [(1, 1, 1), (70, 33, 129), (257, 64, 1433), (8, 256, 20000), (300, 100, 5000)]:
        for ta in (False, True):
            for tb in (False, True):
                a = torch.randn((k, m) if ta else (m, k), device="cuda")
                b = torch.randn((n, k) if tb else (k, n), device="cuda")
                c = torch.empty((m, n), device="cuda")
                gemm(ta, tb, m, n, k, a, a.stride(0), b, b.stride(0), c, n, algo=1)
                want = (a.double().T if ta else a.double()) @ (b.double().T if tb else b.double())
                err = (c.double() - want).abs().max().item() / want.abs().max().item()
                assert err < 2e-6, (m, n, k, ta, tb, err)


def test_gemm_tcgen05_3xtf32():
    """Kernel 2 (tcgen05 + TMA, hi/lo split in shared memory) against torch fp64: fp32-grade accuracy on
    ragged M/N/K tails, every N tile width."""
    from gat_pytorch_b200 import _lib
    from gat_pytorch_b200.gat_layer import gemm
    lib = _lib.load()
    torch.manual_seed(2)
    for (m, n, k) in [(128, 64, 32), (300, 192, 100), (1000, 72, 520), (4097, 256, 1024), (257, 128, 36), (5000, 1024, 48),
                      (20000, 256, 256)]:
        assert lib.gat_gemm_tc_supported(0, 1, m, n, k, k, k, n)
        a = torch.randn((m, k), device="cuda")
        b = torch.randn((n, k), device="cuda")
        c = torch.full((m, n), float("nan"), device="cuda")
        gemm(False, True, m, n, k, a, k, b, k, c, n, algo=2)
        want = a.double() @ b.double().T
        err = ((c.double() - want).abs().max() / want.abs().max()).item()
        assert err < 4e-6, (m, n, k, err)      # K = 1024: 3.1e-6, dominated by the tensor core's round-toward-zero accumulator
    # TN: both operands MN-major, split-K with a fixed-order reduction (dW = dWh^T x)
    for (m, n, k) in [(128, 64, 32), (256, 256, 4096), (192, 256, 10000), (256, 100, 5000), (64, 72, 3333), (260, 136, 70000)]:
        assert lib.gat_gemm_tc_supported(1, 0, m, n, k, m, n, n)
        a = torch.randn((k, m), device="cuda")
        b = torch.randn((k, n), device="cuda")
        c = torch.full((m, n), float("nan"), device="cuda")
        gemm(True, False, m, n, k, a, m, b, n, c, n, algo=2)
        want = a.double().T @ b.double()
        err = ((c.double() - want).abs().max() / want.abs().max()).item()
        # long contractions: the bar is the tensor core's round-toward-zero fp32 accumulator over the 2048 rows of one split
        # (3.3e-6 .. 4.2e-6 at K = 70 000 depending on the data), not the operand split
        assert err < 6e-6, ("TN", m, n, k, err)
    assert not lib.gat_gemm_tc_supported(0, 1, 100, 64, 1433, 1433, 1433, 64)     # Cora: K*4 bytes is not a 16-byte multiple


def test_project_allgather_multi_destination():
    """Fused projection -> all-gather kernel, exercised on ONE GPU: the destinations are two local buffers standing in
    for the ranks' gathered buffers (a peer pointer is an ordinary global address to the kernel).  Every destination
    must receive exactly gat_project_fwd's slab at row_offset, and no row outside the slab may be touched."""
    import ctypes
    from gat_pytorch_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(3)
    for (rows, f_in, nh, fp, lo) in [(1000, 100, 4, 64, 0), (777, 256, 4, 48, 1300), (130, 64, 8, 8, 5)]:
        dp = nh * fp
        x = torch.randn(rows, f_in, device="cuda")
        w = torch.randn(dp, f_in, device="cuda") * 0.1
        a_src, a_tgt = torch.randn(nh, dp, device="cuda"), torch.randn(nh, dp, device="cuda")
        want_wh = torch.empty(rows, dp, device="cuda")
        want_s, want_t = torch.empty(rows, nh, device="cuda"), torch.empty(rows, nh, device="cuda")
        _lib.call("gat_project_fwd", x.data_ptr(), rows, f_in, f_in, 0, w.data_ptr(), f_in, dp, a_src.data_ptr(), a_tgt.data_ptr(), nh,
                  want_wh.data_ptr(), want_s.data_ptr(), want_t.data_ptr(), 2, None, 0, torch.cuda.current_stream().cuda_stream)
        total = lo + rows + 37
        dests = [torch.full((total, dp), float("nan"), device="cuda") for _ in range(2)]
        s_src, s_tgt = torch.empty(rows, nh, device="cuda"), torch.empty(rows, nh, device="cuda")
        arr = (ctypes.c_void_p * 2)(*[d.data_ptr() for d in dests])
        _lib.call("gat_project_fwd_allgather", x.data_ptr(), rows, f_in, f_in, 0, w.data_ptr(), f_in, dp, a_src.data_ptr(), a_tgt.data_ptr(),
                  nh, arr, 2, lo, s_src.data_ptr(), s_tgt.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ref = (x.double() @ w.double().T)
        assert ((want_wh.double() - ref).abs().max() / ref.abs().max()).item() < 3e-6
        for d in dests:
            assert torch.equal(d[lo:lo + rows], want_wh)
            assert torch.isnan(d[:lo]).all() and torch.isnan(d[lo + rows:]).all()
        assert torch.equal(s_src, want_s) and torch.equal(s_tgt, want_t)


def test_gemm_pair_cta_group2():
    """Large-M / short-K NT products run on the persistent CTA-pair kernel (csrc/gemm_pair.cu: cta_group::2, B resident, A through
    tensor memory, score terms as 16 extra MMA columns).  Against torch fp64 on ragged M / N / K, with the ELU operand and
    the ELU' output multiplier, and the multi-destination (fused all-gather) store with a row offset."""
    import ctypes
    from gat_pytorch_b200 import _lib
    from gat_pytorch_b200.gat_layer import gemm
    torch.manual_seed(7)
    st = torch.cuda.current_stream().cuda_stream
    for (m, dp, k, nh, act) in [(16384, 256, 256, 4, False), (20001, 256, 100, 4, False), (19133, 192, 256, 4, True),
                                (33000, 64, 64, 8, False), (16500, 128, 48, 1, False), (70001, 100, 256, 2, True)]:
        x = torch.randn(m, k, device="cuda")
        w = torch.randn(dp, k, device="cuda") / k ** 0.5
        a_src, a_tgt = torch.randn(nh, dp, device="cuda") / dp ** 0.5, torch.randn(nh, dp, device="cuda") / dp ** 0.5
        wh = torch.full((m, dp), float("nan"), device="cuda")
        s_src, s_tgt = torch.full((m, nh), float("nan"), device="cuda"), torch.full((m, nh), float("nan"), device="cuda")
        _lib.call("gat_project_fwd", x.data_ptr(), m, k, k, int(act), w.data_ptr(), k, dp, a_src.data_ptr(), a_tgt.data_ptr(), nh,
                  wh.data_ptr(), s_src.data_ptr(), s_tgt.data_ptr(), 2, None, 0, st)
        xd = torch.nn.functional.elu(x.double()) if act else x.double()
        ref = xd @ w.double().T
        rel = lambda got, want: ((got.double() - want).abs().max() / want.abs().max()).item()
        assert rel(wh, ref) < 2e-6, (m, dp, k, rel(wh, ref))
        # the reference forms the logits' node terms from Wh (gat_layer.py:76-82); here they are x (A W)^T, 16 more MMA columns
        assert rel(s_src, ref @ a_src.double().T) < 2e-6 and rel(s_tgt, ref @ a_tgt.double().T) < 2e-6, (m, dp, k)
    for (m, n, k, act, mul) in [(16384, 256, 256, False, False), (20001, 100, 256, False, True), (50000, 256, 192, True, True),
                                (16385, 8, 16, False, False), (25000, 72, 252, False, False), (400000, 256, 256, False, True)]:
        a = torch.randn(m, k, device="cuda")
        b = torch.randn(n, k, device="cuda")
        c = torch.full((m, n), float("nan"), device="cuda")
        msrc = torch.randn(m, n, device="cuda") if mul else None
        gemm(False, True, m, n, k, a, k, b, k, c, n, algo=2, act_a=act, mul_elu_grad=msrc)
        want = (torch.nn.functional.elu(a.double()) if act else a.double()) @ b.double().T
        if mul:
            want = want * torch.where(msrc > 0, torch.ones_like(msrc), msrc.exp()).double()
        err = ((c.double() - want).abs().max() / want.abs().max()).item()
        assert err < 2e-6, (m, n, k, act, mul, err)
    # TN (dW = dWh^T x over >= 65536 node rows): A through tensor memory (transposed on the way in), B MN-major in shared memory,
    # fp64 slots every 2048 rows.  The bar is the accumulator's round-toward-zero over 2048 rows (same as the one-CTA kernel).
    for (m, n, k, act) in [(256, 256, 70000, False), (192, 256, 100003, False), (256, 100, 66000, True), (64, 72, 131072, False),
                           (8, 8, 65536, False)]:
        a = torch.randn(k, m, device="cuda")
        b = torch.randn(k, n, device="cuda")
        c = torch.full((m, n), float("nan"), device="cuda")
        gemm(True, False, m, n, k, a, m, b, n, c, n, algo=2, act_b=act)
        want = a.double().T @ (torch.nn.functional.elu(b.double()) if act else b.double())
        err = ((c.double() - want).abs().max() / want.abs().max()).item()
        assert err < 8e-6, ("TN", m, n, k, act, err)
    # fused projection -> all-gather through the pair kernel: two destinations, slab at a row offset
    rows, f_in, nh, fp, lo = 20000, 256, 4, 64, 1300
    dp = nh * fp
    x = torch.randn(rows, f_in, device="cuda")
    w = torch.randn(dp, f_in, device="cuda") * 0.1
    a_src, a_tgt = torch.randn(nh, dp, device="cuda"), torch.randn(nh, dp, device="cuda")
    want_wh = torch.empty(rows, dp, device="cuda")
    want_s, want_t = torch.empty(rows, nh, device="cuda"), torch.empty(rows, nh, device="cuda")
    _lib.call("gat_project_fwd", x.data_ptr(), rows, f_in, f_in, 0, w.data_ptr(), f_in, dp, a_src.data_ptr(), a_tgt.data_ptr(), nh,
              want_wh.data_ptr(), want_s.data_ptr(), want_t.data_ptr(), 2, None, 0, st)
    dests = [torch.full((lo + rows + 37, dp), float("nan"), device="cuda") for _ in range(2)]
    s_src, s_tgt = torch.empty(rows, nh, device="cuda"), torch.empty(rows, nh, device="cuda")
    arr = (ctypes.c_void_p * 2)(*[d.data_ptr() for d in dests])
    _lib.call("gat_project_fwd_allgather", x.data_ptr(), rows, f_in, f_in, 0, w.data_ptr(), f_in, dp, a_src.data_ptr(), a_tgt.data_ptr(),
              nh, arr, 2, lo, s_src.data_ptr(), s_tgt.data_ptr(), st)
    torch.cuda.synchronize()
    for d in dests:
        assert torch.equal(d[lo:lo + rows], want_wh)
        assert torch.isnan(d[:lo]).all() and torch.isnan(d[lo + rows:]).all()
    assert torch.equal(s_src, want_s) and torch.equal(s_tgt, want_t)


@pytest.mark.parametrize("name", ["adv_concat", "adv_mean_oddF", "adv_1x1", "adv_wide", "adv_ties", "adv_eps_dominated", "adv_bias",
                                  "cora_L0", "cora_L1", "pubmed_L1", "ppi_L2", "pattern_L0", "pattern_L2",
                                  "products_L0", "products_L1", "products_L2"])
def test_backward_without_attention_gradient(name, small_cases):
    """Nothing consumes the returned attention (GATModel.forward, PATTERN training, inference): the backward then runs
    as rowdot + ONE fused source-major pass (gat_edge_bwd_fused) instead of main + rowsum + finish.  Same bar."""
    case = small_cases[name]
    layer = make_layer(case)
    x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    ei = torch.from_numpy(case["edge_index"]).cuda()
    out = layer(x, ei)
    fw = O.forward(case["x"], case["edge_index"].astype(np.int64), case["W"], case["a"], case["nh"], case["f"], case["concat"],
                   case["add_self_loops"], case["bias"], case["const_attention"])
    go, _ = cases.upstream_grads(case, fw["out"].shape[0], fw["out"].shape[1], fw["alpha"].shape[0])
    (out * torch.from_numpy(go).cuda()).sum().backward()
    gr = O.backward(fw, go, None)
    tol = TOL_OVERRIDE.get(name, TOL)
    errs = {"out": O.rel_err(out.detach().cpu().numpy(), fw["out"]), "gx": O.rel_err(x.grad.cpu().numpy(), gr["x"]),
            "gW": O.rel_err(layer.W.weight.grad.cpu().numpy(), gr["W"]), "ga": O.rel_err(layer.a.weight.grad.cpu().numpy(), gr["a"])}
    assert all(e <= tol for e in errs.values()), errs


@pytest.mark.parametrize("name", ["products_L1", "cora_L1", "pattern_L1", "ppi_L1"])
def test_fused_input_elu_matches_explicit_elu(name, small_cases):
    """Opt-in glue fusion: layer.input_activation = "elu" must equal layer(F.elu(x)) -- the F.elu
    GATModel.forward applies between layers (GATModel.py:148-149) -- in the output and in every gradient, on both GEMM
    paths (tcgen05 for the aligned shapes, FFMA otherwise)."""
    import torch.nn.functional as F
    case = small_cases[name]
    rng = np.random.default_rng(11)
    x_np = (rng.standard_normal(case["x"].shape) * 1.5).astype(np.float32)     # both signs, so ELU matters
    ei = torch.from_numpy(case["edge_index"]).cuda()
    res = []
    for fused in (False, True):
        layer = make_layer(case)
        layer.input_activation = "elu" if fused else None
        x = torch.from_numpy(x_np).cuda().requires_grad_(True)
        out = layer(x if fused else F.elu(x), ei)
        go, _ = cases.upstream_grads(case, out.shape[0], out.shape[1], 1)
        (out * torch.from_numpy(go).cuda()).sum().backward()
        res.append({"out": out.detach(), "gx": x.grad, "gW": layer.W.weight.grad, "ga": layer.a.weight.grad})
    for k in res[0]:
        a, b = res[0][k].double(), res[1][k].double()
        err = ((a - b).abs().max() / a.abs().max().clamp(min=1e-30)).item()
        assert err <= 2e-6, (k, err)


@pytest.mark.parametrize("name", ["products_L1", "cora_L0", "pattern_L1", "ppi_L1", "adv_concat_oddF", "adv_wide"])
def test_fused_output_elu_matches_explicit_elu(name, small_cases):
    """Opt-in glue fusion on the output side (the north star's "...then ELU"): layer.output_activation =
    "elu" must equal F.elu(layer(x)) in the output and every gradient -- ELU in the edge kernel's epilogue, its adjoint in
    the backward's per-node pass (out recovered as log1p(h))."""
    import torch.nn.functional as F
    case = small_cases[name]
    ei = torch.from_numpy(case["edge_index"]).cuda()
    for with_alpha in (False, True):
        res = []
        for fused in (False, True):
            layer = make_layer(case)
            layer.output_activation = "elu" if fused else None
            x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
            if with_alpha:
                out, (_, alpha) = layer(x, ei, return_attention_weights=True)
            else:
                out, alpha = layer(x, ei), None
            if not fused:
                out = F.elu(out)
            go, ga = cases.upstream_grads(case, out.shape[0], out.shape[1], alpha.shape[0] if with_alpha else 1)
            loss = (out * torch.from_numpy(go).cuda()).sum()
            if with_alpha:
                loss = loss + (alpha * torch.from_numpy(ga).cuda()).sum()
            loss.backward()
            res.append({"out": out.detach(), "gx": x.grad, "gW": layer.W.weight.grad, "ga": layer.a.weight.grad})
        for k in res[0]:
            a, b = res[0][k].double(), res[1][k].double()
            err = ((a - b).abs().max() / a.abs().max().clamp(min=1e-30)).item()
            assert err <= 5e-6, (name, with_alpha, k, err)


def test_attention_norm_matches_reference_formula(small_cases):
    """gat_pytorch_b200.attention_norm == GATModel.calc_attention_norm (GATModel.py:189-234), restated
    here with the reference's own torch ops (scatter_add degrees, broadcast back, |alpha*deg - 1|_1 / E', mean over
    layers), value and gradient w.r.t. every attention tensor."""
    from gat_pytorch_b200 import attention_norm
    for name in ("adv_concat", "cora_L0", "products_L1"):
        case = small_cases[name]
        layer = make_layer(case)
        x = torch.from_numpy(case["x"]).cuda()
        ei = torch.from_numpy(case["edge_index"]).cuda()
        _, (ei2, alpha) = layer(x, ei, return_attention_weights=True)
        att = [alpha.detach().clone().requires_grad_(True), (alpha.detach() * 0.5 + 0.01).requires_grad_(True)]
        got = attention_norm(ei2, att)
        got.backward()
        # the reference formulation
        ref_att = [a.detach().clone().double().requires_grad_(True) for a in att]
        dst = ei2[1]
        ones = torch.ones(dst.numel(), dtype=torch.float64, device="cuda")
        deg = torch.zeros(dst.numel(), dtype=torch.float64, device="cuda").scatter_add_(0, dst, ones).index_select(0, dst)
        want = sum(torch.norm(a * deg[:, None] - 1.0, p=1) / dst.numel() for a in ref_att) / len(ref_att)
        want.backward()
        assert abs(got.item() - want.item()) <= 1e-6 * max(abs(want.item()), 1e-30), (name, got.item(), want.item())
        for a, r in zip(att, ref_att):
            err = ((a.grad.double() - r.grad).abs().max() / r.grad.abs().max()).item()
            assert err <= 1e-6, (name, err)


@pytest.mark.gpu
def test_visualisation_feed_matches_reference_loops(small_cases):
    """the per-node attention entropy / degree-scaled weights the vis scripts build with one edge-list mask
    per node (entropy_histograms.py:103-115, weight_histograms.py:74-87; restated with scipy.stats.entropy in
    oracle/gat_oracle.py) against one CSR pass on the GPU.  Bar: 1e-5 tensor-relative."""
    from gat_pytorch_b200 import degree_scaled_attention, neighbourhood_entropy
    for name in ("adv_concat", "adv_mean_oddF", "adv_ties", "cora_L0", "ppi_L2"):
        if name not in small_cases:
            continue
        case = small_cases[name]
        layer = make_layer(case)
        x = torch.from_numpy(case["x"]).cuda()
        ei = torch.from_numpy(case["edge_index"]).cuda()
        _, (ei2, alpha) = layer(x, ei, return_attention_weights=True)
        n = x.size(0)
        ent, uni = neighbourhood_entropy(ei2, alpha)
        scaled = degree_scaled_attention(ei2, alpha)
        want_ent, want_uni = O.neighbourhood_entropy(ei2.cpu().numpy(), alpha.detach().cpu().numpy(), n)
        want_scaled = O.degree_scaled_attention(ei2.cpu().numpy(), alpha.detach().cpu().numpy(), n)
        assert O.rel_err(ent.cpu().numpy(), want_ent) < 1e-5, name
        assert O.rel_err(uni.cpu().numpy(), want_uni) < 1e-5, name
        assert scaled.shape == want_scaled.shape
        assert O.rel_err(scaled.cpu().numpy(), want_scaled) < 1e-5, name
        # star-plot feed (neighbourhood_attention_weights.py:45-60): neighbours' ids bit-exact, edge widths to fp32 rounding
        from gat_pytorch_b200 import neighbourhood_attention
        ei_np, al_np = ei2.cpu().numpy(), alpha.detach().cpu().numpy()
        picks = [int(v) for v in np.unique(ei_np[1])[::max(1, n // 12)][:12]]
        for head in (0, alpha.size(1) - 1):
            got = neighbourhood_attention(ei2, alpha, picks, head)
            want = O.neighbourhood_attention(ei_np, al_np, picks, head)
            assert len(got) == len(want)
            for (gs, gw), (ws_, ww) in zip(got, want):
                assert np.array_equal(gs.cpu().numpy(), ws_), name
                assert np.allclose(gw.cpu().numpy(), ww, rtol=2e-6, atol=0), name


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["products_L0", "products_L1", "products_L2"])
def test_bf16_variant_within_its_stated_tolerance(name, small_cases):
    """BASELINE.json north_star: "bf16 variant stated separately".  `layer.feature_dtype = "bf16"` gathers bfloat16 copies of
    Wh / dL/dout (fp32 accumulation); bar 2e-2 tensor-relative against the fp64 oracle for out and every gradient, and the
    returned attention (computed from the fp32 scores) stays at the fp32 bar."""
    case = small_cases[name]
    layer = make_layer(case)
    layer.feature_dtype = "bf16"
    x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    ei = torch.from_numpy(case["edge_index"]).cuda()
    out = layer(x, ei)
    go, _ = cases.upstream_grads(case, out.shape[0], out.shape[1], 1)
    (out * torch.from_numpy(go).cuda()).sum().backward()
    fw = O.forward(case["x"], case["edge_index"].astype(np.int64), case["W"], case["a"], case["nh"], case["f"], case["concat"],
                   case["add_self_loops"], case["bias"], case["const_attention"])
    gr = O.backward(fw, go, None)
    errs = dict(out=O.rel_err(out.detach().cpu().numpy(), fw["out"]), gx=O.rel_err(x.grad.cpu().numpy(), gr["x"]),
                gW=O.rel_err(layer.W.weight.grad.cpu().numpy(), gr["W"]), ga=O.rel_err(layer.a.weight.grad.cpu().numpy(), gr["a"]))
    assert all(e < 2e-2 for e in errs.values()), (name, errs)
    assert errs["out"] > 1e-6, "the bf16 path did not run (fp32-exact output)"
    _, (_, alpha) = layer(x.detach(), ei, return_attention_weights=True)
    assert O.rel_err(alpha.detach().cpu().numpy(), fw["alpha"]) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cora_L0", "pubmed_L1", "ppi_L1", "ppi_L2", "pattern_L1", "adv_wide"])
@pytest.mark.parametrize("with_alpha", [False, True])
def test_bf16_variant_on_shapes_without_a_bf16_kernel(name, with_alpha, small_cases):
    """The variant is available for EVERY shape and both backward families: where no bf16 kernel exists (more than 4 heads, narrow
    or very wide rows, shared head-mean gradients, an upstream dL/dalpha) the gathered matrices are rounded to bfloat16 and the fp32
    kernels run -- the variant's numerics, same 2e-2 bar; the returned attention stays at the fp32 bar (fp32 scores)."""
    from gat_pytorch_b200 import _lib
    case = small_cases[name]
    fp = (case["f"] + 3) // 4 * 4
    assert not (_lib.load().gat_edge_bf16_native(case["nh"], fp, int(not case["concat"] and case["nh"] > 1)) and not with_alpha)
    layer = make_layer(case)
    layer.feature_dtype = "bf16"
    x = torch.from_numpy(case["x"]).cuda().requires_grad_(True)
    ei = torch.from_numpy(case["edge_index"]).cuda()
    fw = O.forward(case["x"], case["edge_index"].astype(np.int64), case["W"], case["a"], case["nh"], case["f"], case["concat"],
                   case["add_self_loops"], case["bias"], case["const_attention"])
    if with_alpha:
        out, (_, alpha) = layer(x, ei, return_attention_weights=True)
        go, ga = cases.upstream_grads(case, out.shape[0], out.shape[1], alpha.shape[0])
        ((out * torch.from_numpy(go).cuda()).sum() + (alpha * torch.from_numpy(ga).cuda()).sum()).backward()
        gr = O.backward(fw, go, ga)
        assert O.rel_err(alpha.detach().cpu().numpy(), fw["alpha"]) < TOL
    else:
        out = layer(x, ei)
        go, _ = cases.upstream_grads(case, out.shape[0], out.shape[1], 1)
        (out * torch.from_numpy(go).cuda()).sum().backward()
        gr = O.backward(fw, go, None)
    errs = dict(out=O.rel_err(out.detach().cpu().numpy(), fw["out"]), gx=O.rel_err(x.grad.cpu().numpy(), gr["x"]),
                gW=O.rel_err(layer.W.weight.grad.cpu().numpy(), gr["W"]), ga=O.rel_err(layer.a.weight.grad.cpu().numpy(), gr["a"]))
    # da is a sum of g = alpha*(dalpha - S): the cancellation amplifies the 2^-9 rounding of the gathered rows on heads of 3-8
    # features and on PATTERN's epsilon-dominated checkpoint weights (measured: 3.1e-2 on pubmed_L1, 5.9e-2 on pattern_L1, <= 3e-3
    # elsewhere), hence its own bar
    assert all(e < (1e-1 if k == "ga" else 2e-2) for k, e in errs.items()), (name, errs)
    assert errs["out"] > 1e-6, "the bf16 rounding did not happen (fp32-exact output)"
