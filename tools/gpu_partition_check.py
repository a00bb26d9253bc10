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
        print("PARTITION_OK" if ok else "PARTITION_MISMATCH", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)

if __name__ == "__main__":
    main()
