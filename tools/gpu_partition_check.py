"""torchrun --nproc-per-node=P tools/gpu_partition_check.py: P-rank partitioned model vs the single-GPU GATLayer stack
(same weights, same graph).  Forward must be bit-identical; gradients within fp32 reduction-order noise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist, torch.nn.functional as F

def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import gat_pytorch_b200 as g
    from gat_pytorch_b200.partition import PartitionedGAT
    scale = float(os.environ.get("GAT_CHECK_SCALE", 1 / 64))
    x_np, ei_np = g.synth.products(scale=scale)
    shapes, weights = g.synth.LAYER_SHAPES["products"], g.synth.seeded_weights("products")
    x_host, ei_host = torch.from_numpy(x_np), torch.from_numpy(ei_np)
    model = PartitionedGAT(shapes, weights, x_host, ei_host, dev)
    plan = model.plan
    # partitioned forward/backward, keeping the outputs
    h = model.x_local
    for i, layer in enumerate(model.layers):
        h = layer(h, model.st, plan)
        if i != len(model.layers) - 1:
            h = F.elu(h)
    out_local = h
    loss = out_local.square().sum() / (plan.n_real * out_local.size(1))
    loss.backward()
    outs = [torch.zeros((plan.rows_per_rank, out_local.size(1)), device=dev) for _ in range(world)]
    pad = torch.zeros((plan.rows_per_rank, out_local.size(1)), device=dev); pad[:plan.rows] = out_local.detach()
    dist.all_gather(outs, pad)
    ok = True
    if rank == 0:
        if plan.bounds is None:
            out_part = torch.cat(outs)[:plan.n]
        else:       # edge-balanced ranges: rank r's rows are the first b_{r+1} - b_r rows of its slab
            out_part = torch.cat([o[:plan.bounds[r + 1] - plan.bounds[r]] for r, o in enumerate(outs)])
        layers = []
        for (f_in, nh, f, concat), (w, a) in zip(shapes, weights):
            l = g.GATLayer(f_in, f, nh, concat, add_self_loops=True).to(dev)
            with torch.no_grad():
                l.W.weight.copy_(torch.from_numpy(w)); l.a.weight.copy_(torch.from_numpy(a))
            layers.append(l)
        xd, eid = x_host.to(dev), ei_host.to(dev)
        h = xd
        for i, l in enumerate(layers):
            h = l(h, eid)
            if i != len(layers) - 1:
                h = F.elu(h)
        h.square().mean().backward()
        same = torch.equal(out_part, h.detach())
        rel = ((out_part - h.detach()).abs().max() / h.detach().abs().max()).item()
        print(f"forward bit-identical: {same} (rel diff {rel:.2e}) N={plan.n_real} E'={model.n_edges_global} bounds={plan.bounds}")
        ok = ok and rel < 1e-6
        for i, (lp, ls) in enumerate(zip(model.layers, layers)):
            for nm in ("W", "a"):
                gp, gs = getattr(lp, nm).weight.grad, getattr(ls, nm).weight.grad
                r = ((gp - gs).abs().max() / gs.abs().max()).item()
                print(f"layer {i} d{nm}: rel diff {r:.2e}")
                ok = ok and r < 1e-5
    ok = extras(g, model, x_host, ei_host, dev, rank, world) and ok
    if rank == 0:
        print("PARTITION_OK" if ok else "PARTITION_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)

def extras(g, model, x_host, ei_host, dev, rank, world):
    """Returned attention, attention dropout and const_attention of the partitioned layer (reference semantics:
    gat_layer.py:89-92, :113-115, :137-138) against the single-GPU layer on the products hidden-layer shape."""
    from gat_pytorch_b200.partition import PartitionedGATLayer
    plan, st = model.plan, model.st
    f_in, nh, f, concat = g.synth.LAYER_SHAPES["products"][0]
    w, a = g.synth.seeded_weights("products")[0]
    ok = torch.ones(1, device=dev)
    single = g.GATLayer(f_in, f, nh, concat, add_self_loops=True).to(dev)
    part = PartitionedGATLayer(f_in, f, nh, concat, model.backend, dropout=0.5).to(dev)
    with torch.no_grad():
        for l in (single, part):
            l.W.weight.copy_(torch.from_numpy(w)); l.a.weight.copy_(torch.from_numpy(a))
    xd, eid = x_host.to(dev), ei_host.to(dev)
    # (1) returned attention, eval mode: bit-identical to the single-GPU attention of the same edges
    part.eval(); single.eval()
    with torch.no_grad():
        out_l, (edges_l, alpha_l) = part(model.x_local, st, plan, return_attention_weights=True)
        out_s, (ei2, alpha_s) = single(xd, eid, return_attention_weights=True)
    if plan.bounds is None:
        sel = (ei2[1] >= plan.lo) & (ei2[1] < plan.hi)
        same_edges = torch.equal(edges_l.to(ei2.dtype), ei2[:, sel])
    else:       # slab ids: compare through the relabelling
        from gat_pytorch_b200.partition import to_slab_ids
        s2 = torch.stack([to_slab_ids(ei2[0], plan), to_slab_ids(ei2[1], plan)])
        sel = (s2[1] >= plan.lo) & (s2[1] < plan.hi)
        same_edges = torch.equal(edges_l.to(s2.dtype), s2[:, sel])
    same_alpha = torch.equal(alpha_l, alpha_s[sel])
    if not (same_edges and same_alpha and not alpha_l.requires_grad):
        print(f"[rank {rank}] returned attention mismatch: edges {same_edges} alpha {same_alpha}", flush=True)
        ok.zero_()
    # (2) attention dropout, training mode: alpha unchanged (pre-dropout), output differs from eval, unbiased over draws,
    #     and the backward runs on the forward's mask (two backwards of the same forward graph are not possible; check that the
    #     gradient is finite and differs from the eval gradient)
    part.train()
    torch.manual_seed(1234 + rank)
    with torch.no_grad():
        draws = [part(model.x_local, st, plan) for _ in range(24)]
    mean = sum(draws) / len(draws)
    rel_bias = ((mean - out_l).abs().mean() / out_l.abs().mean().clamp(min=1e-30)).item()
    differs = not torch.equal(draws[0], out_l)
    xl = model.x_local.clone().requires_grad_(True)
    part(xl, st, plan).square().sum().backward()
    finite = bool(torch.isfinite(xl.grad).all() and torch.isfinite(part.W.weight.grad).all() and torch.isfinite(part.a.weight.grad).all())
    if not (differs and finite and rel_bias < 0.35):
        print(f"[rank {rank}] dropout check failed: differs {differs} finite {finite} rel_bias {rel_bias:.3f}", flush=True)
        ok.zero_()
    # (3) const_attention: forward bit-identical, gradients within reduction-order noise
    single_c = g.GATLayer(f_in, f, nh, concat, add_self_loops=True, const_attention=True).to(dev)
    part_c = PartitionedGATLayer(f_in, f, nh, concat, model.backend, const_attention=True).to(dev)
    with torch.no_grad():
        single_c.W.weight.copy_(torch.from_numpy(w)); part_c.W.weight.copy_(torch.from_numpy(w))
    xl = model.x_local.clone().requires_grad_(True)
    xs = xd.clone().requires_grad_(True)
    oc_l = part_c(xl, st, plan)
    oc_s = single_c(xs, eid)
    (oc_l.square().sum() / oc_s.numel()).backward()
    oc_s.square().mean().backward()
    g_lo, g_hi = (plan.lo, plan.hi) if plan.bounds is None else (plan.bounds[rank], plan.bounds[rank + 1])
    same_c = torch.equal(oc_l.detach(), oc_s.detach()[g_lo:g_hi])
    r_w = ((part_c.W.weight.grad - single_c.W.weight.grad).abs().max() / single_c.W.weight.grad.abs().max()).item()
    r_x = ((xl.grad - xs.grad[g_lo:g_hi]).abs().max() / xs.grad.abs().max()).item()
    if not (same_c and r_w < 1e-5 and r_x < 1e-5):
        print(f"[rank {rank}] const_attention mismatch: forward identical {same_c} dW {r_w:.2e} dx {r_x:.2e}", flush=True)
        ok.zero_()
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"extras (returned attention, dropout, const_attention): {'ok' if ok.item() else 'MISMATCH'}; dropout rel bias over 24 draws "
              f"{rel_bias:.3f}, const dW {r_w:.2e} dx {r_x:.2e}", flush=True)
    return bool(ok.item())


if __name__ == "__main__":
    main()
