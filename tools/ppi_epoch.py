#!/usr/bin/env python
"""PPI epoch time (the third figure of BASELINE.json's metric): a plain loop reproducing `PPI_GAT.training_step`
(models/ppi_gat.py:15-41) over one epoch of the PPI-shaped split -- 20 train graphs, batch 2 => 10 steps
(run_config.py:17-33) -- on synthetic graphs of that shape (no dataset in the image).

Per step, as the reference does: `forward_and_return_attention` glue (GATModel.py:153-187: layer -> skip -> ELU, attention
returned by every layer), BCE-with-logits, `calc_attention_norm`, attention penalty, micro-F1 with sklearn on the CPU
(a D2H sync per step, ppi_gat.py:38), backward, Adam(lr 0.005).

    python tools/ppi_epoch.py [--epochs 3] [--penalty 1.0] [--cpu-steps 2]

GPU arm: the drop-in GATLayer + gat_pytorch_b200.attention_norm.  CPU arm (a reported baseline): the torch port of the
reference layer (oracle/torch_port.py) on all host cores for --cpu-steps steps, scaled to an epoch.  One JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.nn.functional as F


def batches(n_batches, seed=0):
    import gat_pytorch_b200 as g
    rng = np.random.default_rng(seed)
    out = []
    for b in range(n_batches):
        x, ei = g.synth.ppi(seed=100 + b)
        y = (rng.random((x.shape[0], 121)) < 0.3).astype(np.float32)
        out.append((torch.from_numpy(x), torch.from_numpy(ei), torch.from_numpy(y)))
    return out


def f1_micro(out, y):
    from sklearn.metrics import f1_score
    return f1_score(y_pred=out.detach().cpu().numpy() > 0, y_true=y.detach().cpu().numpy(), average="micro")


def gpu_epoch(data, epochs, penalty, f1="sklearn", fused=False):
    """f1 = "sklearn": the reference's own call (ppi_gat.py:38); "gpu": gat_pytorch_b200.micro_f1 (same value, counted on the device).
    fused: forward through gat_pytorch_b200.model_forward with the regulariser computed by the layers (no attention tensors)."""
    import gat_pytorch_b200 as g
    import types
    f1_fn = f1_micro if f1 == "sklearn" else (lambda out, y: g.micro_f1(out, y))
    dev = torch.device("cuda", 0)
    shapes, skip = g.synth.LAYER_SHAPES["ppi"], g.synth.SKIP["ppi"]
    torch.manual_seed(42)
    layers = torch.nn.ModuleList([g.GATLayer(fi, f, nh, c, add_self_loops=True) for (fi, nh, f, c) in shapes]).to(dev)
    opt = torch.optim.Adam(layers.parameters(), lr=0.005)
    dd = [(x.to(dev), ei.to(dev), y.to(dev)) for x, ei, y in data]
    lib = g._lib.load() if hasattr(g, "_lib") else None

    model = types.SimpleNamespace(gat_layer_list=layers, skip_layer_list=torch.nn.ModuleList([torch.nn.Identity() for s_ in skip if s_]),
                                  add_skip_connection=list(skip), heads_concat_per_layer=[c for (_fi, _nh, _f, c) in shapes],
                                  num_heads_per_layer=[1] + [nh for (_fi, nh, _f, _c) in shapes],
                                  head_output_features_per_layer=[shapes[0][0]] + [f for (_fi, _nh, f, _c) in shapes],
                                  dropout=0.0, training=True)

    def step(x, ei, y):
        if fused:
            h, norm = g.model_forward(model, types.SimpleNamespace(x=x, edge_index=ei), attention_norm=True)
        else:
            att = []
            h = x
            for i, layer in enumerate(layers):
                inp = h
                h, (ei, a) = layer(h, ei, return_attention_weights=True)        # GATModel.py:166 (rewritten list feeds the next layer)
                att.append(a)
                if skip[i]:
                    h = h + inp                                                  # identity skip, GATModel.py:171-181
                if i != len(layers) - 1:
                    h = F.elu(h)
            norm = g.attention_norm(ei, att)
        loss = F.binary_cross_entropy_with_logits(h, y)
        if penalty != 0.0:
            loss = loss + penalty * norm
        f1 = f1_fn(h, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return float(loss.detach()), f1

    def breakdown(x, ei, y):
        """The same step with a device synchronisation after every part: where an epoch's time goes (ms per step)."""
        def tick(t=[None]):
            torch.cuda.synchronize()
            now = time.perf_counter()
            dt = None if t[0] is None else (now - t[0]) * 1e3
            t[0] = now
            return dt
        parts = {}
        tick()
        att, h = [], x
        for i, layer in enumerate(layers):
            inp = h
            h, (ei, a) = layer(h, ei, return_attention_weights=True)
            att.append(a)
            if skip[i]:
                h = h + inp
            if i != len(layers) - 1:
                h = F.elu(h)
        parts["forward (3 layers + glue)"] = tick()
        loss = F.binary_cross_entropy_with_logits(h, y)
        norm = g.attention_norm(ei, att)
        loss = loss + penalty * norm
        parts["loss + attention_norm"] = tick()
        f1_micro(h, y)
        parts["sklearn micro-F1 on the CPU (D2H + f1_score)"] = tick()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        parts["backward"] = tick()
        opt.step()
        parts["Adam step"] = tick()
        return parts

    for x, ei, y in dd[:3]:
        step(x, ei, y)
    torch.cuda.synchronize()
    gpu_epoch.breakdown = {}
    for x, ei, y in dd:
        for k, v in breakdown(x, ei, y).items():
            gpu_epoch.breakdown[k] = gpu_epoch.breakdown.get(k, 0.0) + v / len(dd)
    times = []
    for _ in range(epochs):
        t0 = time.perf_counter()
        for x, ei, y in dd:
            last = step(x, ei, y)
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    return min(times), last


def cpu_epoch(data, steps, penalty):
    import torch_port
    import gat_pytorch_b200 as g
    shapes, skip = g.synth.LAYER_SHAPES["ppi"], g.synth.SKIP["ppi"]
    torch.set_num_threads(os.cpu_count() or 1)
    ws = [(torch.from_numpy(w).requires_grad_(True), torch.from_numpy(a).requires_grad_(True)) for w, a in g.synth.seeded_weights("ppi")]
    opt = torch.optim.Adam([t for pair in ws for t in pair], lr=0.005)

    def step(x, ei, y):
        att, h = [], x
        for i, ((w, a), (_fi, nh, f, c)) in enumerate(zip(ws, shapes)):
            inp = h
            h, ei2, al = torch_port.layer_forward(h, ei, w, a, nh, f, c)
            att.append(al)
            if skip[i]:
                h = h + inp
            if i != len(shapes) - 1:
                h = F.elu(h)
        loss = F.binary_cross_entropy_with_logits(h, y)
        dst = ei2[1]
        deg = torch.zeros(dst.numel()).scatter_add_(0, dst, torch.ones(dst.numel())).index_select(0, dst)   # GATModel.py:196-201
        norm = sum(torch.norm(al * deg[:, None] - 1.0, p=1) / dst.numel() for al in att) / len(att)
        if penalty != 0.0:
            loss = loss + penalty * norm
        f1_micro(h, y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    step(*data[0])
    t0 = time.perf_counter()
    for i in range(steps):
        step(*data[i % len(data)])
    return (time.perf_counter() - t0) / steps * len(data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--penalty", type=float, default=1.0)
    ap.add_argument("--cpu-steps", type=int, default=2)
    args = ap.parse_args()
    data = batches(10)
    gpu_s, (loss, f1) = gpu_epoch(data, args.epochs, args.penalty)
    cpu_s = cpu_epoch(data, args.cpu_steps, args.penalty) if args.cpu_steps > 0 else None
    n, e = data[0][0].shape[0], data[0][1].shape[1]
    print(json.dumps({"metric": "ppi_epoch_time", "unit": "s", "value": gpu_s, "higher_is_better": False, "steps_per_epoch": len(data),
                      "config": {"workload": f"PPI-shaped epoch: 10 steps of 2-graph batches (N={n}, E={e} before self-loops), 3-layer GAT "
                                             "4x256 / 4x256+skip / 6x121 mean, BCE + attention penalty, sklearn micro-F1, Adam",
                                 "attention_penalty": args.penalty},
                      "last_step": {"loss": loss, "train_f1": f1},
                      "ms_per_step_breakdown_synchronised": {k: round(v, 3) for k, v in gpu_epoch.breakdown.items()},
                      "cpu_baseline": None if cpu_s is None else {"value": cpu_s, "unit": "s", "cores": torch.get_num_threads(), "kind": "port",
                                                                  "sample": f"{args.cpu_steps} steps of oracle/torch_port.py scaled to 10"}}))


if __name__ == "__main__":
    main()
