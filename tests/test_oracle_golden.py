"""Pins the oracle (oracle/gat_oracle.py) on outputs of the REFERENCE itself.

tests/golden/golden.npz was produced by tests/golden/make_golden.py, which imports
/root/reference/models/gat_layer.py and runs it (fp32 as shipped, and .double()) with autograd.
The reference has no tests or golden vectors of its own.
"""
import numpy as np
import pytest

import cases
import gat_oracle as O

CASE_NAMES = [c["name"] for c in cases.adversarial_cases()] + [
    f"{m}_L{i}" for m in ("cora", "pubmed", "ppi", "pattern", "products") for i in range(len(cases.synth.LAYER_SHAPES[m]))]

# fp32 tolerance: the reference's own fp32 run differs from its fp64 run by up to ~5e-6;
# the eps-dominated adversarial case amplifies rounding in the logits further.
F32_TOL = {"default": 2e-5, "adv_eps_dominated": 2e-3, "pattern_L2": 5e-5, "pattern_L3": 5e-5}


def _sampled(full, shape_rows=64, shape_cols=160):
    v2 = np.asarray(full).reshape(full.shape[0], -1)
    return v2[np.ix_(cases.sample_idx(v2.shape[0], shape_rows), cases.sample_idx(v2.shape[1], shape_cols))]


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
def test_oracle_matches_reference(name, golden, small_cases):
    case = small_cases[name]
    fw, res = run_oracle(case)
    # integer path: rewritten edge list and degree counts, bit exact
    ei = fw["edge_index"]
    assert tuple(golden[f"{name}/ei_shape"]) == ei.shape
    cols = cases.sample_idx(ei.shape[1], 512)
    assert np.array_equal(golden[f"{name}/ei_cols"], ei[:, cols])
    w = (np.arange(ei.shape[1], dtype=np.int64) % 1000003) + 1
    assert np.array_equal(golden[f"{name}/ei_checksum"], [(ei[0].astype(np.int64) * w).sum(), (ei[1].astype(np.int64) * w).sum()])
    deg = O.in_degrees(ei, case["x"].shape[0])
    assert np.array_equal(golden[f"{name}/deg_checksum"], [(deg * (np.arange(deg.size) % 1000003 + 1)).sum(), deg.max()])
    rowptr, col, eid = O.csr_by_target(ei, case["x"].shape[0])
    assert np.array_equal(np.diff(rowptr), deg)
    # floating point: fp64 run of the reference pins the restatement; fp32 run bounds its noise
    tol32 = F32_TOL.get(name, F32_TOL["default"])
    for k, v in res.items():
        if name == "adv_int32" and k.startswith("g"):
            # torch 2.11's autograd for advanced indexing with int32 indices drops duplicate contributions on
            # CPU: the reference's own int32 gradients differ from its int64 gradients by O(1) while the
            # forward is bit-identical (measured).  Only the forward is pinned for int32.
            continue
        scale = max(float(golden[f"{name}/f64/{k}_max"]), 1e-30)
        got = _sampled(v)
        err64 = np.abs(got - golden[f"{name}/f64/{k}"]).max() / scale if got.size else 0.0
        # const_attention: the reference builds float32 zeros (gat_layer.py:92), so even its .double() run
        # computes alpha in fp32
        tol64 = 3e-7 if case["const_attention"] else 1e-9
        assert err64 < tol64, (k, err64)
        assert abs(np.asarray(v, np.float64).sum() - float(golden[f"{name}/f64/{k}_sum"])) <= tol64 * max(float(golden[f"{name}/f64/{k}_abs"]), 1e-30)
        err32 = np.abs(got - golden[f"{name}/f32/{k}"]).max() / scale if got.size else 0.0
        assert err32 < tol32, (k, err32)


def test_self_loop_rewrite_semantics():
    ei = np.array([[0, 1, 1, 2, 3, 3], [1, 1, 0, 2, 0, 0]], dtype=np.int64)   # two loops, one duplicate
    out = O.add_remaining_self_loops(ei)
    assert out.tolist() == [[0, 1, 3, 3, 0, 1, 2, 3], [1, 0, 0, 0, 0, 1, 2, 3]]
    assert np.array_equal(O.add_remaining_self_loops(out), out)       # idempotent (section 9.3)


def test_csr_roundtrip():
    rng = np.random.default_rng(3)
    ei = rng.integers(0, 50, size=(2, 400))
    rowptr, col, eid = O.csr_by_target(ei, 57)
    rowptr_t, col_t, pos_t = O.csr_by_source(ei, 57, eid)
    assert np.array_equal(ei[0][eid], col) and np.all(np.diff(ei[1][eid]) >= 0)
    for i in range(57):       # stable: edge ids ascending inside a row
        assert np.all(np.diff(eid[rowptr[i]:rowptr[i + 1]]) > 0)
    # transposed slots point back at the same edges
    src_sorted = np.repeat(np.arange(57), np.diff(rowptr_t))
    assert np.array_equal(col[pos_t], src_sorted)
    dst_of_slot = np.repeat(np.arange(57), np.diff(rowptr))
    assert np.array_equal(dst_of_slot[pos_t], col_t)
