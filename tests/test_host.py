"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/gat_b200.h
declares, the drop-in module mirrors the reference's constructor contract, the product path refuses
to run without CUDA, and the synthetic generators are deterministic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gat_b200.h")).read()
    return sorted(set(re.findall(r"GAT_API\s+[\w\s\*]+?\b(gat_\w+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ["gat_edges_scan", "gat_csr_build", "gat_gemm", "gat_scores_fwd", "gat_edge_max", "gat_edge_fwd",
                 "gat_edge_bwd_main", "gat_edge_bwd_rowsum", "gat_edge_bwd_finish", "gat_project_fwd", "gat_last_error",
                 "gat_version"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from gat_pytorch_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s
    bound = _lib.load()
    assert set(_lib.SIGNATURES) == set(declared_symbols())
    assert bound.gat_version() >= 100
    assert bound.gat_last_error() is not None


def test_constructor_contract_matches_reference():
    """Reference: models/gat_layer.py:13-40 -- attribute names, sub-modules, state_dict keys and init RNG order."""
    from gat_pytorch_b200 import GATLayer
    torch.manual_seed(0)
    layer = GATLayer(in_features=10, out_features=3, num_heads=4, concat=True, dropout=0.5, add_self_loops=True, bias=True)
    assert list(layer.state_dict().keys()) == ["bias_param", "W.weight", "a.weight"]
    assert layer.W.weight.shape == (12, 10) and layer.a.weight.shape == (4, 24) and layer.bias_param.shape == (12,)
    assert isinstance(layer.dropout_layer, torch.nn.Dropout) and layer.normalised_attention_coeffs is None
    for attr in ["in_features", "out_features", "num_heads", "concat", "dropout", "add_self_loops", "bias", "const_attention", "device"]:
        assert hasattr(layer, attr)
    # same RNG stream as constructing Linear(W) then Linear(a) then xavier on both (gat_layer.py:27-40,142-147)
    torch.manual_seed(0)
    w = torch.nn.Linear(10, 12, bias=False)
    a = torch.nn.Linear(24, 4, bias=False)
    torch.nn.init.xavier_uniform_(w.weight)
    torch.nn.init.xavier_uniform_(a.weight)
    assert torch.equal(layer.W.weight, w.weight) and torch.equal(layer.a.weight, a.weight)
    const = GATLayer(10, 3, 4, False, const_attention=True)
    assert not hasattr(const, "a") and list(const.state_dict().keys()) == ["W.weight"]


def test_checkpoint_weights_load():
    """State-dict keys of the committed checkpoints load into the drop-in layer."""
    from gat_pytorch_b200 import GATLayer
    z = np.load(os.path.join(ROOT, "tests", "golden", "ckpt_weights.npz"))
    layer = GATLayer(1433, 8, 8, True, dropout=0.6, add_self_loops=True)
    sd = {"W.weight": torch.from_numpy(z["Cora.gat_layer_list.0.W.weight"]), "a.weight": torch.from_numpy(z["Cora.gat_layer_list.0.a.weight"])}
    layer.load_state_dict(sd, strict=True)


def test_no_cpu_fallback():
    from gat_pytorch_b200 import GATLayer, build_structure
    layer = GATLayer(4, 2, 2, True)
    with pytest.raises(RuntimeError, match="CUDA"):
        layer(torch.randn(3, 4), torch.zeros((2, 2), dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        build_structure(torch.zeros((2, 2), dtype=torch.long), 3, True)


def test_product_path_does_not_import_oracle():
    pkg = os.path.join(ROOT, "gat-pytorch_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "gat_oracle" not in text and "torch_port" not in text, fn


def test_overlay_resolves_models_gat_layer():
    import importlib
    import sys
    overlay = os.path.join(ROOT, "gat-pytorch_b200", "overlay")
    sys.path.insert(0, overlay)
    try:
        sys.modules.pop("models", None)
        sys.modules.pop("models.gat_layer", None)
        mod = importlib.import_module("models.gat_layer")
        from gat_pytorch_b200 import GATLayer
        assert mod.GATLayer is GATLayer
    finally:
        sys.path.remove(overlay)
        sys.modules.pop("models", None)
        sys.modules.pop("models.gat_layer", None)


def test_synth_is_deterministic_and_shaped():
    from gat_pytorch_b200 import synth
    x1, e1 = synth.cora()
    x2, e2 = synth.cora()
    assert np.array_equal(x1, x2) and np.array_equal(e1, e2)
    assert x1.shape == (2708, 1433) and e1.shape == (2, 10556) and not np.any(e1[0] == e1[1])
    assert set(np.unique(x1)) <= {0.0, 1.0}
    xp, ep = synth.products(scale=1 / 512)
    assert xp.shape[1] == 100 and ep.shape[0] == 2 and ep.max() < xp.shape[0]
    # symmetric: every (u,v) has its (v,u)
    fwd = set(map(tuple, ep.T[:2000].tolist()))
    allp = set(map(tuple, ep.T.tolist()))
    assert all((v, u) in allp for (u, v) in fwd)


def test_torch_port_matches_oracle(small_cases):
    """The CPU-baseline port (oracle/torch_port.py) computes what the oracle computes."""
    import gat_oracle as O
    import torch_port
    case = small_cases["adv_concat"]
    out, ei2, alpha = torch_port.layer_forward(torch.from_numpy(case["x"]), torch.from_numpy(case["edge_index"]),
                                               torch.from_numpy(case["W"]), torch.from_numpy(case["a"]), case["nh"], case["f"], True, True)
    fw = O.forward(case["x"], case["edge_index"], case["W"], case["a"], case["nh"], case["f"], True, True)
    assert np.array_equal(ei2.numpy(), fw["edge_index"])
    assert O.rel_err(out.numpy(), fw["out"]) < 1e-5 and O.rel_err(alpha.numpy(), fw["alpha"]) < 1e-5


def test_visualisation_oracle_matches_segment_formulas(small_cases):
    """The oracle restatement of the vis scripts' per-node loops (entropy_histograms.py:103-115, weight_histograms.py:74-87)
    against closed forms on the CSR segments: uniform entropy = log2(deg), entropy = log2(S) - sum(a log2 a)/S with S the
    row sum, degree-scaled weights = alpha * deg[dst] in stable target order."""
    import gat_oracle as O
    case = small_cases["adv_concat"]
    fw = O.forward(case["x"], case["edge_index"], case["W"], case["a"], case["nh"], case["f"], case["concat"], True)
    ei, alpha, n = fw["edge_index"], fw["alpha"].astype(np.float64), case["x"].shape[0]
    ent, uni = O.neighbourhood_entropy(ei, alpha, n)
    deg = np.bincount(ei[1], minlength=n)
    assert np.allclose(uni[deg > 0], np.log2(deg[deg > 0])) and np.all(uni[deg == 0] == 0)
    ssum = np.zeros((n, alpha.shape[1]))
    np.add.at(ssum, ei[1], alpha)
    alog = np.zeros_like(ssum)
    np.add.at(alog, ei[1], np.where(alpha > 0, alpha * np.log2(np.maximum(alpha, 1e-300)), 0.0))
    want = np.where(ssum > 0, np.log2(np.maximum(ssum, 1e-300)) - alog / np.maximum(ssum, 1e-300), 0.0)
    assert np.allclose(ent, want, atol=1e-9)
    order = np.argsort(ei[1], kind="stable")
    assert np.allclose(O.degree_scaled_attention(ei, alpha, n), alpha[order] * deg[ei[1][order]][:, None])


def test_c_abi_rejects_bad_arguments_with_a_message_before_touching_the_device():
    """include/gat_b200.h: every entry point returns an int status and leaves a message in gat_last_error(); argument
    validation happens before any CUDA call, so it can be exercised on a machine without a GPU (no compute is launched)."""
    from gat_pytorch_b200 import _lib
    lib = _lib.load()
    assert lib.gat_version() >= 100
    assert lib.gat_tgt_pack_stride(4) == 16 and lib.gat_tgt_pack_stride(8) == 32
    assert lib.gat_edge_fwd_workspace_bytes() == 256
    assert lib.gat_scores_fwd(None, 10, 64, None, None, 9, None, None, None) != 0
    assert b"num_heads 9" in lib.gat_last_error()
    assert lib.gat_f32_to_bf16(None, None, 8, None) != 0 and b"gat_f32_to_bf16" in lib.gat_last_error()
    assert lib.gat_attention_entropy(None, None, 5, None, 9, None, None, None) != 0 and b"gat_attention_entropy" in lib.gat_last_error()
    # the tcgen05 path needs 16-byte aligned leading dimensions: Cora's K = 1433 goes to the FFMA kernel
    assert lib.gat_gemm_tc_supported(0, 1, 100, 64, 1433, 1433, 1433, 64) == 0
    assert lib.gat_gemm_tc_supported(0, 1, 100, 64, 1024, 1024, 1024, 64) == 1
    with pytest.raises(RuntimeError, match="gat_scores_fwd"):
        _lib.call("gat_scores_fwd", None, 10, 64, None, None, 9, None, None, None)
    # round-2 entry points: the output glue, the fused regulariser, the bf16 coverage query, the device micro-F1
    assert lib.gat_out_glue_adjoint(None, None, 4, 8, 1, 0.5, 1, None, None) != 0 and b"gat_out_glue_adjoint" in lib.gat_last_error()
    assert lib.gat_head_merge_fwd_glue(None, 4, 2, 4, 4, 1, None, 0, 1, 1.5, 0, None, None) != 0 and b"dropout" in lib.gat_last_error()
    assert lib.gat_attention_norm_scores(None, None, 4, 8, None, None, None, None, 4, 0, None, None, None, 0, None) != 0
    assert b"gat_attention_norm_scores" in lib.gat_last_error()
    assert lib.gat_micro_f1_counts(None, None, 8, None, None) != 0 and b"gat_micro_f1_counts" in lib.gat_last_error()
    assert lib.gat_f32_round_bf16(None, None, 8, None) != 0 and b"gat_f32_round_bf16" in lib.gat_last_error()
    # bf16 kernels: NH <= 4, padded rows of 132..256 floats, unshared gradient (products-class shapes); everything else rounds
    assert lib.gat_edge_bf16_native(4, 64, 0) == 1 and lib.gat_edge_bf16_native(4, 48, 0) == 1
    assert lib.gat_edge_bf16_native(4, 48, 1) == 0 and lib.gat_edge_bf16_native(8, 8, 0) == 0 and lib.gat_edge_bf16_native(4, 256, 0) == 0


def test_model_forward_folds_the_glue_as_documented():
    """glue.model_forward on mock layers, no GPU: which switches each layer gets -- ELU on every layer but the last,
    the next layer's input dropout folded into a layer's output unless that next layer has a skip connection (it then needs the
    undropped tensor too), the skip rows head-averaged for a head-mean layer (GATModel.py:139-145) -- and that the layers come
    back with their own settings."""
    import types
    import torch
    from gat_pytorch_b200.glue import model_forward

    calls = []

    class MockLayer(torch.nn.Module):
        def __init__(self, out_dim):
            super().__init__()
            self.out_dim, self.output_activation, self.output_dropout, self.attention_norm = out_dim, None, 0.0, False
            self.attention_norm_value = None

        def forward(self, x, edge_index, return_attention_weights=False, skip=None):
            calls.append(dict(act=self.output_activation, drop=self.output_dropout, norm=self.attention_norm,
                              skip=None if skip is None else tuple(skip.shape), x_zero_frac=float((x == 0).float().mean())))
            self.attention_norm_value = torch.tensor(float(len(calls)))
            out = torch.ones(x.size(0), self.out_dim)
            if return_attention_weights:
                return out, (edge_index, torch.full((edge_index.size(1), 2), 0.5))
            return out

    def model(add_skip, concat, dims, heads, feats, p):
        m = torch.nn.Module()
        m.gat_layer_list = torch.nn.ModuleList([MockLayer(d) for d in dims])
        m.skip_layer_list = torch.nn.ModuleList([torch.nn.Identity() for s in add_skip if s])
        m.add_skip_connection, m.heads_concat_per_layer = add_skip, concat
        m.num_heads_per_layer, m.head_output_features_per_layer, m.dropout = heads, feats, p
        return m

    data = types.SimpleNamespace(x=torch.ones(6, 8), edge_index=torch.zeros((2, 5), dtype=torch.long))
    # three layers, skip on the middle one (PPI's pattern), dropout 0.5 in training mode
    m = model([False, True, False], [True, True, False], [8, 8, 4], [1, 2, 2, 2], [8, 4, 4, 2], 0.5)
    m.train()
    out = model_forward(m, data)
    assert tuple(out.shape) == (6, 4)
    assert [c["act"] for c in calls] == ["elu", "elu", None]
    # layer 0 may NOT fold layer 1's dropout (layer 1 has a skip and needs the undropped tensor); layer 1 folds layer 2's
    assert [c["drop"] for c in calls] == [0.0, 0.5, 0.0]
    assert [c["skip"] for c in calls] == [None, (6, 8), None]
    assert calls[0]["x_zero_frac"] > 0.2          # raw input: torch dropout in front of the first layer
    assert calls[1]["x_zero_frac"] > 0.2          # not folded -> torch dropout in front of layer 1
    assert calls[2]["x_zero_frac"] == 0.0         # folded into layer 1's output kernel: no torch dropout here
    assert all(l.output_activation is None and l.output_dropout == 0.0 and l.attention_norm is False for l in m.gat_layer_list)
    # eval mode: nothing dropped anywhere; attention requested -> the reference's triple; the norm is the mean over layers
    calls.clear()
    m.eval()
    out, ei2, att, norm = model_forward(m, data, True, attention_norm=True)
    assert [c["drop"] for c in calls] == [0.0, 0.0, 0.0] and all(c["x_zero_frac"] == 0.0 for c in calls)
    assert all(c["norm"] for c in calls) and len(att) == 3 and float(norm) == 2.0
    # a head-mean layer with a skip connection receives the head-averaged skip rows
    calls.clear()
    m2 = model([True], [False], [4], [1, 2], [8, 4], 0.0)
    model_forward(m2, types.SimpleNamespace(x=torch.arange(48.).view(6, 8), edge_index=data.edge_index))
    assert calls[0]["skip"] == (6, 4) and calls[0]["act"] is None


def test_head_groups_cover_every_head_within_the_kernel_limits():
    """gat_layer._head_groups: layers beyond the edge kernels' limits (more than 8 heads, more than 1024 floats per padded row) are
    processed in consecutive head groups that each fit."""
    from gat_pytorch_b200.gat_layer import MAX_HEADS, MAX_ROW_FLOATS, _head_groups
    for nh, fp in [(9, 8), (12, 8), (16, 72), (4, 300), (10, 128), (64, 4), (3, 1024), (8, 128), (1, 4)]:
        groups = _head_groups(nh, fp)
        assert groups[0][0] == 0 and groups[-1][1] == nh
        assert all(a[1] == b[0] for a, b in zip(groups, groups[1:]))
        assert all(0 < h1 - h0 <= MAX_HEADS and (h1 - h0) * fp <= MAX_ROW_FLOATS for h0, h1 in groups)
    assert _head_groups(8, 128) == [(0, 8)]          # exactly at the limits: one group


@pytest.mark.skipif(__import__("shutil").which("gcc") is None, reason="needs gcc")
def test_layer_descriptor_layout_matches_the_header(tmp_path):
    """struct gat_layer_desc crosses the C ABI by pointer: the ctypes mirror in _lib.py must agree with the header field by field
    (size and every offset), checked by compiling a C program against include/gat_b200.h -- a mismatch would corrupt pointers
    silently on the GPU box."""
    import ctypes
    import subprocess
    from gat_pytorch_b200 import _lib
    names = [n for n, _ in _lib.LayerDesc._fields_]
    src = ('#include <stdio.h>\n#include <stddef.h>\n#include "gat_b200.h"\nint main(void) {\n  printf("%zu\\n", sizeof(gat_layer_desc));\n'
           + "".join(f'  printf("{n} %zu\\n", offsetof(gat_layer_desc, {n}));\n' for n in names) + "  return 0;\n}\n")
    (tmp_path / "t.c").write_text(src)
    exe = str(tmp_path / "t")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(tmp_path / "t.c"), "-o", exe],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr          # the header is plain C (no C++ / CUDA types in the boundary)
    lines = subprocess.run([exe], capture_output=True, text=True).stdout.split()
    assert int(lines[0]) == ctypes.sizeof(_lib.LayerDesc)
    for name, off in zip(lines[1::2], lines[2::2]):
        assert getattr(_lib.LayerDesc, name).offset == int(off), name
    # and no field of the header is missing from the mirror: the struct's size leaves no room for one
    last = names[-1]
    assert getattr(_lib.LayerDesc, last).offset + getattr(_lib.LayerDesc, last).size == ctypes.sizeof(_lib.LayerDesc)


def test_binding_argument_counts_match_the_header():
    """Every prototype of include/gat_b200.h against its ctypes signature in _lib.SIGNATURES: same number of parameters, pointer
    parameters bound as pointers, 64-bit integers as 64-bit (a wrong count or width passes garbage across the C ABI)."""
    import ctypes
    import re
    from gat_pytorch_b200 import _lib
    text = open(os.path.join(ROOT, "include", "gat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = re.findall(r"GAT_API\s+([\w\s\*]+?)\s*\b(gat_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S)
    assert len(protos) >= 50
    seen = set()
    for _ret, name, args in protos:
        seen.add(name)
        params = [a.strip() for a in args.replace("\n", " ").split(",")]
        if params == ["void"] or params == [""]:
            params = []
        restype, argtypes = _lib.SIGNATURES[name]
        assert len(params) == len(argtypes), (name, len(params), len(argtypes))
        for p, t in zip(params, argtypes):
            if "*" in p or "gat_stream_t" in p:
                assert t is ctypes.c_void_p, (name, p, t)
            elif re.search(r"\b(int64_t|uint64_t|size_t)\b", p):
                assert ctypes.sizeof(t) == 8 and t is not ctypes.c_void_p and t is not ctypes.c_double, (name, p, t)
            elif re.search(r"\bfloat\b", p):
                assert t is ctypes.c_float, (name, p, t)
            elif re.search(r"\bint\b", p):
                assert t is ctypes.c_int, (name, p, t)
    assert seen == set(_lib.SIGNATURES), (seen ^ set(_lib.SIGNATURES))
