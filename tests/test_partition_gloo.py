"""world_size-2 gloo test of the destination-range partitioned layer on CPU.

The numerical backend is the oracle-based stand-in (tests/_oracle_backend.py); what is under test is the
product's host logic in gat-pytorch_b200/partition.py: the partition plan, the rank-local slice of the
rewritten edge list, and the collective choreography (all-gather of Wh/s_src, all-reduce-max of M,
all-reduce of (Gamma,|T|), reduce-scatter of dWh, gradient all-reduce)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case_name, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cases
        from _oracle_backend import OracleBackend
        from gat_pytorch_b200.partition import PartitionedGATLayer, local_edge_list, make_plan
        case = {c["name"]: c for c in cases.adversarial_cases()}[case_name]
        x, ei = torch.from_numpy(case["x"]), torch.from_numpy(case["edge_index"].astype(np.int64))
        plan = make_plan(x.size(0), world, rank)
        backend = OracleBackend()
        n_idx = int(ei.max()) + 1
        st = backend.build_structure(local_edge_list(ei, n_idx, plan.lo, plan.hi, True), plan.n)
        layer = PartitionedGATLayer(x.size(1), case["f"], case["nh"], case["concat"], backend)
        with torch.no_grad():
            layer.W.weight.copy_(torch.from_numpy(case["W"]))
            layer.a.weight.copy_(torch.from_numpy(case["a"]))
        xl = x[plan.lo:plan.hi].clone().requires_grad_(True)
        out = layer(xl, st, plan)
        go_full, _ = cases.upstream_grads(case, x.size(0), out.size(1), 1)
        (out * torch.from_numpy(go_full[plan.lo:plan.hi])).sum().backward()
        ret[rank] = dict(lo=plan.lo, hi=plan.hi, out=out.detach().numpy(), gx=xl.grad.numpy(),
                         gW=layer.W.weight.grad.numpy(), ga=layer.a.weight.grad.numpy(),
                         edges=np.stack([st["src"].numpy(), st["dst"].numpy()]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name,world", [("adv_concat", 2), ("adv_mean_oddF", 2), ("adv_ties", 2), ("adv_concat", 3)])
def test_two_rank_partition_matches_single_process_oracle(case_name, world):
    """world 3 on 97 nodes: unequal last range (33 + 33 + 31) and a padded gathered buffer."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cases
    import gat_oracle as O
    ret = mp.Manager().dict()
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(world, port, case_name, ret), nprocs=world, join=True)
    case = {c["name"]: c for c in cases.adversarial_cases()}[case_name]
    fw = O.forward(case["x"], case["edge_index"], case["W"], case["a"], case["nh"], case["f"], case["concat"], True)
    go, _ = cases.upstream_grads(case, fw["out"].shape[0], fw["out"].shape[1], 1)
    gr = O.backward(fw, go, None)
    out = np.concatenate([ret[r]["out"] for r in range(world)])
    gx = np.concatenate([ret[r]["gx"] for r in range(world)])
    assert O.rel_err(out, fw["out"]) < 1e-5
    assert O.rel_err(gx, gr["x"]) < 1e-5
    for r in range(world):                      # replicated parameters carry the GLOBAL gradient on every rank
        assert O.rel_err(ret[r]["gW"], gr["W"]) < 1e-5
        assert O.rel_err(ret[r]["ga"], gr["a"]) < 1e-5
    # the rank-local edge lists tile the rewritten list: same multiset, per-target order preserved
    ei2 = fw["edge_index"]
    for r in range(world):
        sel = (ei2[1] >= ret[r]["lo"]) & (ei2[1] < ret[r]["hi"])
        assert np.array_equal(ret[r]["edges"], ei2[:, sel])


def _replicated_worker(rank, world, port, case_name, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cases
        from _oracle_backend import OracleBackend
        from gat_pytorch_b200.partition import PartitionedGATLayer, local_edge_list, make_plan
        case = {c["name"]: c for c in cases.adversarial_cases()}[case_name]
        x, ei = torch.from_numpy(case["x"]), torch.from_numpy(case["edge_index"].astype(np.int64))
        plan = make_plan(x.size(0), world, rank)
        backend = OracleBackend()
        st = backend.build_structure(local_edge_list(ei, int(ei.max()) + 1, plan.lo, plan.hi, True), plan.n)
        layer = PartitionedGATLayer(x.size(1), case["f"], case["nh"], case["concat"], backend)
        with torch.no_grad():
            layer.W.weight.copy_(torch.from_numpy(case["W"]))
            layer.a.weight.copy_(torch.from_numpy(case["a"]))
        x_full = torch.zeros((plan.n_pad, x.size(1)))
        x_full[:x.size(0)] = x                      # every rank holds all node features (what PartitionedGAT all-gathers once per upload)
        out = layer.forward_replicated(x_full, st, plan)
        go_full, _ = cases.upstream_grads(case, x.size(0), out.size(1), 1)
        (out * torch.from_numpy(go_full[plan.lo:plan.hi])).sum().backward()
        ret[rank] = dict(out=out.detach().numpy(), gW=layer.W.weight.grad.numpy(), ga=layer.a.weight.grad.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name,world", [("adv_concat", 2), ("adv_mean_oddF", 3), ("adv_ties", 2)])
def test_replicated_input_layer_matches_single_process_oracle(case_name, world):
    """The first partitioned layer on a replicated input (no feature exchange in either direction: every rank projects all nodes,
    dW is formed from the rank's partial dWh over all nodes and summed by the parameter all-reduce)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cases
    import gat_oracle as O
    ret = mp.Manager().dict()
    port = 29900 + (os.getpid() % 40)
    mp.spawn(_replicated_worker, args=(world, port, case_name, ret), nprocs=world, join=True)
    case = {c["name"]: c for c in cases.adversarial_cases()}[case_name]
    fw = O.forward(case["x"], case["edge_index"], case["W"], case["a"], case["nh"], case["f"], case["concat"], True)
    go, _ = cases.upstream_grads(case, fw["out"].shape[0], fw["out"].shape[1], 1)
    gr = O.backward(fw, go, None)
    assert O.rel_err(np.concatenate([ret[r]["out"] for r in range(world)]), fw["out"]) < 1e-5
    for r in range(world):
        assert O.rel_err(ret[r]["gW"], gr["W"]) < 1e-5
        assert O.rel_err(ret[r]["ga"], gr["a"]) < 1e-5


def _extras_worker(rank, world, port, case_name, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cases
        from _oracle_backend import OracleBackend
        from gat_pytorch_b200.partition import PartitionedGATLayer, local_edge_list, make_plan
        case = {c["name"]: c for c in cases.adversarial_cases()}[case_name]
        x, ei = torch.from_numpy(case["x"]), torch.from_numpy(case["edge_index"].astype(np.int64))
        plan = make_plan(x.size(0), world, rank)
        backend = OracleBackend()
        st = backend.build_structure(local_edge_list(ei, int(ei.max()) + 1, plan.lo, plan.hi, True), plan.n)
        layer = PartitionedGATLayer(x.size(1), case["f"], case["nh"], case["concat"], backend, const_attention=case["const_attention"])
        with torch.no_grad():
            layer.W.weight.copy_(torch.from_numpy(case["W"]))
            if not case["const_attention"]:
                layer.a.weight.copy_(torch.from_numpy(case["a"]))
        xl = x[plan.lo:plan.hi].clone().requires_grad_(True)
        out, (edges, alpha) = layer(xl, st, plan, return_attention_weights=True)
        go_full, _ = cases.upstream_grads(case, x.size(0), out.size(1), 1)
        (out * torch.from_numpy(go_full[plan.lo:plan.hi])).sum().backward()
        ret[rank] = dict(lo=plan.lo, hi=plan.hi, out=out.detach().numpy(), gx=xl.grad.numpy(), gW=layer.W.weight.grad.numpy(),
                         ga=None if case["const_attention"] else layer.a.weight.grad.numpy(),
                         edges=edges.numpy(), alpha=alpha.numpy(), alpha_requires_grad=bool(alpha.requires_grad),
                         has_a=hasattr(layer, "a"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case_name", ["adv_const", "adv_concat"])
def test_partitioned_layer_const_attention_and_returned_attention(case_name):
    """`const_attention` (gat_layer.py:89-92) and `return_attention_weights` in the partitioned layer: outputs and gradients
    against the single-process oracle; the returned attention is the oracle's, restricted to the edges whose target the rank
    owns, in the rewritten order; it is an output only (not differentiable)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import cases
    import gat_oracle as O
    world = 2
    ret = mp.Manager().dict()
    port = 29950 + (os.getpid() % 40)
    mp.spawn(_extras_worker, args=(world, port, case_name, ret), nprocs=world, join=True)
    case = {c["name"]: c for c in cases.adversarial_cases()}[case_name]
    fw = O.forward(case["x"], case["edge_index"], case["W"], case["a"], case["nh"], case["f"], case["concat"], True, None,
                   case["const_attention"])
    go, _ = cases.upstream_grads(case, fw["out"].shape[0], fw["out"].shape[1], 1)
    gr = O.backward(fw, go, None)
    assert O.rel_err(np.concatenate([ret[r]["out"] for r in range(world)]), fw["out"]) < 1e-5
    assert O.rel_err(np.concatenate([ret[r]["gx"] for r in range(world)]), gr["x"]) < 1e-5
    ei2 = fw["edge_index"]
    for r in range(world):
        assert O.rel_err(ret[r]["gW"], gr["W"]) < 1e-5
        if not case["const_attention"]:
            assert O.rel_err(ret[r]["ga"], gr["a"]) < 1e-5
        assert ret[r]["has_a"] == (not case["const_attention"])
        sel = (ei2[1] >= ret[r]["lo"]) & (ei2[1] < ret[r]["hi"])
        assert np.array_equal(ret[r]["edges"], ei2[:, sel])
        assert O.rel_err(ret[r]["alpha"], fw["alpha"][sel]) < 1e-5
        assert not ret[r]["alpha_requires_grad"]


def _exchange_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gat_pytorch_b200.partition import edge_slice, exchange_edge_list, local_edge_list, make_plan
        g = torch.Generator().manual_seed(5)
        n = 101
        ei = torch.randint(0, n - 3, (2, 1500), generator=g)        # existing self-loops, duplicates, trailing isolated nodes
        ei[:, 7] = ei[:, 3]
        ei[1, 20:40] = ei[0, 20:40]
        plan = make_plan(n, world, rank)
        c0, c1 = edge_slice(ei.size(1), world, rank)
        got, n_idx = exchange_edge_list(ei[:, c0:c1].contiguous(), plan, None, True)
        want = local_edge_list(ei, int(ei.max()) + 1, plan.lo, plan.hi, True)
        ret[rank] = (bool(torch.equal(got, want)), n_idx == int(ei.max()) + 1, tuple(got.shape))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_edge_exchange_equals_the_local_filter(world):
    """Each rank uploads 1/P of the edge list and the buckets are exchanged (all-to-all): the rank-local rewritten list must be
    identical, edge for edge and in order, to filtering the whole list (so the CSR rows and the forward stay bit-identical)."""
    ret = mp.Manager().dict()
    port = 29950 + (os.getpid() % 40)
    mp.spawn(_exchange_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret[r][0] and ret[r][1], (r, ret[r])


def _balanced_worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from _oracle_backend import OracleBackend
        from gat_pytorch_b200.partition import (PartitionedGATLayer, edge_balanced_bounds, edge_slice, exchange_edge_list,
                                                make_balanced_plan)
        x, ei, w, a, nh, f = _skewed_case()
        n = x.size(0)
        c0, c1 = edge_slice(ei.size(1), world, rank)
        bounds = edge_balanced_bounds(ei[:, c0:c1].contiguous(), n, world)
        plan = make_balanced_plan(bounds, rank)
        backend = OracleBackend()
        local, n_idx = exchange_edge_list(ei[:, c0:c1].contiguous(), plan, None, True)
        st = backend.build_structure(local, plan.n)
        layer = PartitionedGATLayer(x.size(1), f, nh, True, backend)
        with torch.no_grad():
            layer.W.weight.copy_(w)
            layer.a.weight.copy_(a)
        xl = x[bounds[rank]:bounds[rank + 1]].clone().requires_grad_(True)
        out = layer(xl, st, plan)
        g = torch.Generator().manual_seed(11)
        go = torch.randn(n, out.size(1), generator=g)
        (out * go[bounds[rank]:bounds[rank + 1]]).sum().backward()
        ret[rank] = dict(bounds=bounds, n_edges=int(local.size(1)), out=out.detach().numpy(), gx=xl.grad.numpy(),
                         gW=layer.W.weight.grad.numpy(), ga=layer.a.weight.grad.numpy())
    finally:
        dist.destroy_process_group()


def _skewed_case():
    """Degree-sorted graph: the first nodes are hubs (what equal node ranges split badly)."""
    g = torch.Generator().manual_seed(3)
    n, e, nh, f, f_in = 120, 2400, 2, 4, 6
    dst = (torch.rand(e, generator=g) ** 3 * (n - 5)).long()            # mass concentrated on low ids
    src = torch.randint(0, n - 5, (e,), generator=g)
    ei = torch.stack([src, dst])
    x = torch.randn(n, f_in, generator=g)
    w = torch.randn(nh * f, f_in, generator=g) * 0.3
    a = torch.randn(nh, nh * 2 * f, generator=g) * 0.3
    return x, ei, w, a, nh, f


@pytest.mark.parametrize("world", [2, 3])
def test_edge_balanced_partition_matches_single_process_oracle(world):
    """destination ranges chosen to equalise EDGE counts.  The pipeline then runs on slab ids (rank r's node g is
    r*R + g - b_r); outputs and gradients must still equal the single-process oracle, and the edge counts must be balanced where
    equal node ranges are not."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gat_oracle as O
    ret = mp.Manager().dict()
    port = 29900 + (os.getpid() % 40)
    mp.spawn(_balanced_worker, args=(world, port, ret), nprocs=world, join=True)
    x, ei, w, a, nh, f = _skewed_case()
    fw = O.forward(x.numpy(), ei.numpy(), w.numpy(), a.numpy(), nh, f, True, True)
    g = torch.Generator().manual_seed(11)
    go = torch.randn(x.size(0), fw["out"].shape[1], generator=g).numpy()
    gr = O.backward(fw, go, None)
    assert O.rel_err(np.concatenate([ret[r]["out"] for r in range(world)]), fw["out"]) < 1e-5
    assert O.rel_err(np.concatenate([ret[r]["gx"] for r in range(world)]), gr["x"]) < 1e-5
    for r in range(world):
        assert O.rel_err(ret[r]["gW"], gr["W"]) < 1e-5 and O.rel_err(ret[r]["ga"], gr["a"]) < 1e-5
    counts = [ret[r]["n_edges"] for r in range(world)]
    assert sum(counts) == fw["edge_index"].shape[1]
    assert max(counts) <= 1.25 * sum(counts) / world, counts                  # balanced ...
    dst = fw["edge_index"][1]
    per = (x.size(0) + world - 1) // world
    naive = [int(((dst >= r * per) & (dst < (r + 1) * per)).sum()) for r in range(world)]
    assert max(naive) > 1.5 * sum(naive) / world, naive                      # ... where equal node ranges are not


def test_plan_covers_all_rows():
    from gat_pytorch_b200.partition import make_plan
    for n, world in [(97, 2), (10, 4), (8, 8), (5, 8)]:
        plans = [make_plan(n, world, r) for r in range(world)]
        assert plans[0].lo == 0 and plans[-1].hi == n
        assert all(a.hi == b.lo for a, b in zip(plans, plans[1:]))
        assert all(p.rows <= p.rows_per_rank and p.n_pad >= n for p in plans)


def test_cuda_backend_implements_the_backend_interface():
    """Every method the partitioned layer calls on its backend exists on the product (C-ABI) backend with the same
    parameter names as on the oracle stand-in used above (guards against the two drifting apart)."""
    import inspect
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _oracle_backend import OracleBackend
    from gat_pytorch_b200.partition import CudaBackend
    for name, fn in inspect.getmembers(OracleBackend, predicate=inspect.isfunction):
        if name.startswith("_"):
            continue
        assert hasattr(CudaBackend, name), name
        assert list(inspect.signature(fn).parameters) == list(inspect.signature(getattr(CudaBackend, name)).parameters), name


def _model_worker(rank, world, port, balance, replicate, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from _oracle_backend import OracleBackend
        from gat_pytorch_b200 import synth
        from gat_pytorch_b200.partition import PartitionedGAT
        x, ei, shapes, weights = _model_inputs(synth)
        model = PartitionedGAT(shapes, weights, torch.from_numpy(x), torch.from_numpy(ei), torch.device("cpu"), backend=OracleBackend(),
                               balance=balance, replicate_input=replicate)
        loss = model.step_resident()
        total = loss.detach().clone()
        dist.all_reduce(total)
        e2e = model.step_e2e()                     # fresh upload (edge exchange, structure, x all-gather) + the same step
        ret[rank] = dict(loss=float(total), e2e=float(e2e), bounds=model.plan.bounds, rows=model.plan.rows, n_edges=model.n_edges_global,
                         x_full=model.x_full is not None, out=model.last_out.numpy(),
                         grads=[(l.W.weight.grad.numpy().copy(), l.a.weight.grad.numpy().copy()) for l in model.layers])
    finally:
        dist.destroy_process_group()


def _model_inputs(synth):
    """A 3-layer stack of the products pattern in miniature (concat, concat, four-head head mean) on the adversarial graph, sorted
    by degree so that equal node ranges would be badly unbalanced."""
    rng = np.random.default_rng(11)
    x, ei = synth.adversarial()
    shapes = [(x.shape[1], 2, 8, True), (16, 2, 8, True), (16, 4, 3, False)]
    weights = [(synth.xavier_uniform(rng, nh * f, f_in), synth.xavier_uniform(rng, nh, 2 * nh * f)) for (f_in, nh, f, _c) in shapes]
    return x, ei.astype(np.int64), shapes, weights


@pytest.mark.parametrize("balance,replicate,world", [("edges", True, 2), ("edges", False, 2), ("nodes", True, 2), ("edges", True, 3)])
def test_partitioned_model_matches_the_single_process_stack(balance, replicate, world):
    """bench.py's multi-GPU model end to end on 2 gloo ranks with the oracle backend: distributed edge-list upload, edge-balanced
    (or equal) destination ranges, the first layer on a replicated input (or exchanged like the others), ELU between layers, the
    head-mean output layer with ONE shared gradient row per target, loss share per rank, parameter gradients summed over ranks --
    against the torch port of the reference formulation in fp64 (loss, output rows, every dW / da)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port
    from gat_pytorch_b200 import synth
    ret = mp.Manager().dict()
    port = 29800 + (os.getpid() % 50)
    mp.spawn(_model_worker, args=(world, port, balance, replicate, ret), nprocs=world, join=True)
    x, ei, shapes, weights = _model_inputs(synth)
    ws = [(torch.from_numpy(w).double().requires_grad_(True), torch.from_numpy(a).double().requires_grad_(True)) for w, a in weights]
    h = torch.from_numpy(x).double()
    eit = torch.from_numpy(ei)
    for i, ((w, a), (_fi, nh, f, concat)) in enumerate(zip(ws, shapes)):
        h, _, _ = torch_port.layer_forward(h, eit, w, a, nh, f, concat)
        if i != len(shapes) - 1:
            h = torch.nn.functional.elu(h)
    loss = h.square().mean()
    loss.backward()
    loss = loss.detach()
    rel = lambda got, want: float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))   # noqa: E731
    out = np.concatenate([ret[r]["out"] for r in range(world)])
    assert out.shape == tuple(h.shape) and rel(out, h.detach().numpy()) < 1e-5
    for r in range(world):
        assert abs(ret[r]["loss"] - float(loss)) < 1e-6 * abs(float(loss)) + 1e-12
        assert abs(ret[r]["e2e"] - float(loss)) < 1e-6 * abs(float(loss)) + 1e-12
        assert ret[r]["x_full"] == replicate
        assert (ret[r]["bounds"] is not None) == (balance == "edges")
        for (gw, ga), (w, a) in zip(ret[r]["grads"], ws):
            assert rel(gw, w.grad.numpy()) < 2e-5 and rel(ga, a.grad.numpy()) < 2e-5
    assert sum(ret[r]["rows"] for r in range(world)) == x.shape[0]
    assert ret[0]["n_edges"] == int(torch_port.rewrite_edges(eit).size(1))
