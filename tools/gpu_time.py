"""Quick per-kernel CUDA-event timing of one layer fwd+bwd on a products-shaped graph (scale arg)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gat_pytorch_b200 as g

def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.125
    layer_idx = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    t0 = time.time()
    x0, ei = g.synth.products(scale=scale)
    f_in, nh, f, concat = g.synth.LAYER_SHAPES["products"][layer_idx]
    n = x0.shape[0]
    print("graph", n, ei.shape, "gen %.1fs" % (time.time() - t0), flush=True)
    x = torch.randn(n, f_in, device="cuda", requires_grad=True)
    eit = torch.from_numpy(ei).cuda()
    layer = g.GATLayer(f_in, f, nh, concat, add_self_loops=True).cuda()
    t0 = time.time(); out = layer(x, eit); torch.cuda.synchronize(); print("first fwd (incl CSR build) %.3fs" % (time.time() - t0), flush=True)
    st = layer.structure_cache.get(eit, n, True)
    print("E'=", st.n_edges, "max deg", int(st.in_degrees().max()), flush=True)
    go = torch.randn_like(out)
    for it in range(3):
        out = layer(x, eit); out.backward(go)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    reps = 5
    tf = tb = 0.0
    for it in range(reps):
        ev[0].record(); out = layer(x, eit); ev[1].record(); out.backward(go); ev[2].record()
        torch.cuda.synchronize()
        tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
    tf /= reps; tb /= reps
    print("fwd %.3f ms  bwd %.3f ms  edges/s %.3e" % (tf, tb, st.n_edges / ((tf + tb) * 1e-3)), flush=True)
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        out = layer(x, eit); out.backward(go); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))

if __name__ == "__main__":
    main()
