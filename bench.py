#!/usr/bin/env python
"""bench.py -- GAT layer fwd+bwd edges/s on the ogbn-products-shaped synthetic graph.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME] [--scale S]

One "step" = forward + backward of the 3-layer, 4-head GAT of BASELINE.json configs[4]
(100 -> 4x64 -> 4x64 -> 4x47 mean; hidden width per SURVEY.md section 8) over the whole graph
(2 449 029 nodes, 61 859 140 edges, E' = edges after the self-loop rewrite), with the reference's
inter-layer glue (layer -> ELU, GATModel.py:120-151).  metric = layer-edges/s = L * E' / t
(BASELINE.md section 2).  Prints ONE JSON line on rank 0.

  value    device-resident inputs, graph structure cached (steady-state training step)
  e2e      same step through the public GATLayer API starting from PINNED HOST buffers: H2D of x and
           edge_index, CSR/CSR^T build for the freshly uploaded graph, fwd+bwd, D2H of the loss
  roofline dominant kernel (the fused source-major backward pass gat_edge_bwd_fused on the hidden-layer
           shape) timed live with CUDA events inside the timed region; algorithmic bytes per SURVEY.md
           section 8-d / DESIGN.md section 4; traffic = ncu dram bytes of the same launch (profiles/traffic.json)
  cpu_baseline / --impl reference: the reference's OWN `GATLayer` (unmodified copy under oracle/_ref/, written by
           oracle/make_ref.py; kind "reference") stacked with GATModel.forward's glue, on all host cores, on a 1/64-scale
           products-shaped graph (the reference formulation cannot allocate full scale, SURVEY
           5.7); the line's config states the N / E' / scale actually timed.  Falls back to the torch port
           (oracle/torch_port.py, kind "port") only when oracle/_ref/ is absent.
  extra    (default single-GPU run only) the other figures of BASELINE.json's metric, driver-run in the same command: the four
           small named shapes at 1 GPU (`named_shapes`, same method as --workload NAME) and the PPI epoch time (`ppi_epoch`)
  checksums / parity_vs_n1: loss, per-layer sum|dW| / sum|da| (fp64) and an order-independent bit checksum of the final
           output, printed for every N; at N > 1 rank 0 re-runs the single-GPU model on the same inputs after the timed
           region and reports the differences (forward must be bit-identical, gradients within reduction-order noise).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.nn.functional as F

METRIC = "gat_layer_fwd_bwd_edges_per_s"
UNIT = "edges/s"
CPU_SCALE = 1.0 / 64     # BASELINE.md section 3 / SURVEY 8-d: the scale the reference is timed at


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="products")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the named-shape lines and the PPI epoch appended to the default run")
    ap.add_argument("--unfused-glue", action="store_true", help="run the inter-layer ELU as separate torch kernels")
    ap.add_argument("--feature-dtype", default="f32", choices=["f32", "bf16"],
                    help="bf16: the opt-in variant whose per-edge gathers read bfloat16 copies (stated separately from the fp32 headline)")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def load_synth():
    import importlib.util
    spec = importlib.util.spec_from_file_location("gat_b200_synth", os.path.join(ROOT, "gat-pytorch_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_workload(name, scale):
    synth = load_synth()
    gen = synth.GENERATORS[name]
    x, ei = gen(scale=scale) if name == "products" else gen()
    shapes = synth.LAYER_SHAPES[name]
    if name in ("products", "ppi"):
        weights = synth.seeded_weights(name)
    else:
        z = np.load(os.path.join(ROOT, "tests", "golden", "ckpt_weights.npz"))
        tag = {"cora": "Cora", "citeseer": "Citeseer", "pubmed": "Pubmed", "pattern": "PATTERN"}[name]
        weights = [(z[f"{tag}.gat_layer_list.{i}.W.weight"], z[f"{tag}.gat_layer_list.{i}.a.weight"]) for i in range(len(shapes))]
    return x, ei, shapes, weights


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons every 100 ms while the timed region runs (NVML)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self._halt = index, [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while self.nv is not None and not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------
# CPU reference arm: the reference's own GATLayer (oracle/_ref, an unmodified copy) on the host cores
# ----------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def reference_available():
    return os.path.isfile(os.path.join(REF_DIR, "models", "gat_layer.py"))


def cpu_step_fn(name, scale):
    """fwd+bwd of the stacked model on CPU tensors with GATModel.forward's glue (layer -> ELU, GATModel.py:120-151; dropout 0,
    no skip connections on this config).  Returns (step, E', layers, N, kind): kind "reference" = the reference's
    unmodified models/gat_layer.py imported from oracle/_ref, "port" = oracle/torch_port.py (only when _ref is absent)."""
    x, ei, shapes, weights = make_workload(name, scale)
    torch.set_num_threads(os.cpu_count() or 1)
    xt, eit = torch.from_numpy(x), torch.from_numpy(ei)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port
    e_prime = int(torch_port.rewrite_edges(eit).size(1))
    if reference_available():
        sys.path.insert(0, REF_DIR)
        from models.gat_layer import GATLayer as RefGATLayer      # the reference's file, byte for byte
        layers = []
        for (f_in, nh, f, concat), (w, a) in zip(shapes, weights):
            layer = RefGATLayer(in_features=f_in, out_features=f, num_heads=nh, concat=concat, dropout=0, add_self_loops=True,
                                bias=False)
            with torch.no_grad():
                layer.W.weight.copy_(torch.from_numpy(w))
                layer.a.weight.copy_(torch.from_numpy(a))
            layers.append(layer)

        def step():
            h = xt
            for i, layer in enumerate(layers):
                layer.W.weight.grad = layer.a.weight.grad = None
                h = layer(h, eit)
                if i != len(layers) - 1:
                    h = F.elu(h)
            loss = h.square().mean()
            loss.backward()
            return float(loss.detach())

        return step, e_prime, len(shapes), x.shape[0], "reference"
    ws = [(torch.from_numpy(w).requires_grad_(True), torch.from_numpy(a).requires_grad_(True)) for w, a in weights]

    def step():
        for w, a in ws:
            w.grad = a.grad = None
        return torch_port.model_step(xt, eit, ws, shapes)

    return step, e_prime, len(shapes), x.shape[0], "port"


def cpu_scale(args):
    return CPU_SCALE * args.scale if args.workload == "products" else 1.0


def cpu_sample_text(args, n, e_prime, n_layers, kind, how):
    impl = ("the reference's own models/gat_layer.py (unmodified copy in oracle/_ref)" if kind == "reference"
            else "oracle/torch_port.py (oracle/_ref absent)")
    return (f"{args.workload}-shaped graph at scale {cpu_scale(args):.6g} of full size (N={n}, E'={e_prime}), {n_layers}-layer fwd+bwd, "
            f"{how}, {impl}, torch {torch.__version__} CPU on {torch.get_num_threads()} threads")


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    step, e_prime, n_layers, n, kind = cpu_step_fn(args.workload, cpu_scale(args))
    loss = None
    for _ in range(max(args.warmup, 1)):
        loss = step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step()
    dt = (time.perf_counter() - t0) / args.steps
    val = n_layers * e_prime / dt
    sample = cpu_sample_text(args, n, e_prime, n_layers, kind, f"mean of {args.steps} steps after {max(args.warmup, 1)} warm-up")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            # same workload family as the B200 arm; n_nodes / n_edges_rewritten / scale are the bounded sample ACTUALLY timed
            # here (the reference formulation cannot allocate the full-size graph), not the B200 arm's full size
            "config": dict(workload_config(args.workload, args.gpus), scale=cpu_scale(args), n_nodes=n, n_edges_rewritten=e_prime,
                           layers=[list(s) for s in load_synth().LAYER_SHAPES[args.workload]],
                           bounded_sample=True, full_size_scale=args.scale),
            "loss": loss,
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(name, gpus):
    big = name == "products"
    return {"workload": f"{name}-shaped synthetic graph, 3-layer 4-head GAT (100->4x64->4x64->4x47 mean), fwd+bwd, "
                        "value = layers*E'/t" if big else f"{name}-shaped synthetic graph, all layers fwd+bwd, value = layers*E'/t",
            "graph": name, "partition": "single GPU" if gpus == 1 else f"destination-range over {gpus} GPUs",
            "l2_policy": "inputs larger than L2 (per-layer feature matrix 2.5 GB vs 126 MB L2); no flush" if big else
                         "working set fits in L2: a 512 MB buffer is overwritten between timed steps (flush outside the timed events)"}


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def algorithmic_bytes(kernel, e, n, nh, d, d_out):
    """SURVEY.md section 8-d per-launch byte counts (fp32, int32 indices)."""
    if kernel == "gat_edge_fwd":
        return e * (4 + 4 * nh + 4 * d) + n * (8 + 8 * nh + 4 * d_out)
    if kernel == "gat_edge_bwd_main":
        # per edge: col_t, s_tgt[dst], Z[dst], gather dOut[dst], write record {d_alpha, alpha}; per node: rowptr_t, Wh row,
        # s_src, write dWh row
        return e * (4 + 8 * nh + 4 * d_out + 8 * nh) + n * (8 + 4 * d + 4 * nh + 4 * d)
    if kernel == "gat_edge_bwd_fused":
        # per edge: col_t, s_tgt[dst], Z[dst], S[dst], gather dOut[dst]; per node: rowptr_t, Wh row, s_src, tie counts,
        # ds_tgt read+write, write ds_src, write dWh row   (no per-edge record is written)
        return e * (4 + 12 * nh + 4 * d_out) + n * (8 + 4 * d + 4 * nh + 8 * nh + 8 * nh + 4 * nh + 4 * d)
    if kernel == "gat_edge_bwd_rowsum":
        return e * (4 + 8 * nh) + n * (8 + 12 * nh)
    if kernel == "gat_edge_bwd_finish":
        return e * (4 + 8 * nh + 4 * nh) + n * (8 + 8 * d + 16 * nh)
    raise KeyError(kernel)


class SingleGpuModel:
    """The stacked drop-in GATLayers of one config on one GPU, with the reference's inter-layer glue (GATModel.py:148-149:
    F.elu on every layer's output but the last).  The ELU rides in the edge kernel's epilogue and the backward's per-node
    pass (opt-in fusion); --unfused-glue runs it as separate torch kernels."""

    def __init__(self, g, shapes, weights, dev, args):
        self.layers = []
        for (f_in, nh, f, concat), (w, a) in zip(shapes, weights):
            layer = g.GATLayer(f_in, f, nh, concat, dropout=0.0, add_self_loops=True).to(dev)
            with torch.no_grad():
                layer.W.weight.copy_(torch.from_numpy(w))
                layer.a.weight.copy_(torch.from_numpy(a))
            self.layers.append(layer)
        self.fused_glue = [not args.unfused_glue and i != len(self.layers) - 1 and bool(layer.concat)
                           for i, layer in enumerate(self.layers)]
        for layer, fz in zip(self.layers, self.fused_glue):
            layer.output_activation = "elu" if fz else None
            layer.feature_dtype = None if args.feature_dtype == "f32" else args.feature_dtype
        self.last_loss = self.last_out = None

    def fwd_bwd(self, x, ei):
        h = x
        for i, layer in enumerate(self.layers):
            layer.W.weight.grad = layer.a.weight.grad = None
            h = layer(h, ei)
            if i != len(self.layers) - 1 and not self.fused_glue[i]:
                h = F.elu(h)
        loss = h.square().mean()
        loss.backward()
        self.last_loss, self.last_out = loss.detach(), h.detach()
        return loss


def checksums(loss, out_rows, grads, world):
    """Driver-verifiable fingerprints of one step: the loss, per-layer sum|dW| / sum|da| accumulated in fp64, and an
    order-independent checksum of the final output's BIT PATTERNS (int64 sum of the fp32 words, so equal outputs give equal
    sums whatever the row partition or summation order).  At N > 1 `out_rows` is this rank's row range and `loss` its share:
    both are summed over ranks; the gradients are already the global sums on every rank."""
    bits = out_rows.contiguous().view(torch.int32).to(torch.int64).sum().reshape(1)
    fsum = torch.stack([loss.double().reshape(()), out_rows.double().abs().sum()])
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(bits)
        dist.all_reduce(fsum)
    return {"loss": float(fsum[0].item()), "out_bits_sum": int(bits.item()), "out_abs_sum": float(fsum[1].item()),
            "grad_abs_sums": [[float(gw.double().abs().sum().item()), float(ga.double().abs().sum().item())] for gw, ga in grads]}


def parity_vs_single_gpu(g, shapes, weights, x_host, ei_host, dev, args, multi, multi_grads):
    """Rank 0 of a multi-GPU run: the single-GPU model on the same inputs and weights, compared with what the partitioned
    run produced (its checksums and its all-reduced parameter gradients)."""
    ref = SingleGpuModel(g, shapes, weights, dev, args)
    ref.fwd_bwd(x_host.to(dev), ei_host.to(dev))
    one = checksums(ref.last_loss, ref.last_out, [(l.W.weight.grad, l.a.weight.grad) for l in ref.layers], 1)
    grad_rel = 0.0
    for (gw, ga), l in zip(multi_grads, ref.layers):
        for got, want in ((gw, l.W.weight.grad), (ga, l.a.weight.grad)):
            grad_rel = max(grad_rel, float(((got - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()))
    return {"n1": one, "forward_bit_identical": multi["out_bits_sum"] == one["out_bits_sum"],
            "loss_rel_diff": abs(multi["loss"] - one["loss"]) / max(abs(one["loss"]), 1e-30),
            "out_abs_sum_rel_diff": abs(multi["out_abs_sum"] - one["out_abs_sum"]) / max(abs(one["out_abs_sum"]), 1e-30),
            "grad_max_rel_diff": grad_rel}


def measure_named_shape(g, _lib, lib, name, dev, steps=20, warmup=5):
    """One of the small BASELINE shapes (Cora / Pubmed / PPI x2 / PATTERN x128) on the product path, the way `--workload NAME`
    times it: device-resident inputs, structure cached, 512 MB written between steps to flush L2, CUDA events per step."""
    x_np, ei_np, shapes, weights = make_workload(name, 1.0)

    class _A:
        unfused_glue, feature_dtype = False, "f32"
    model = SingleGpuModel(g, shapes, weights, dev, _A)
    x, ei = torch.from_numpy(x_np).to(dev), torch.from_numpy(ei_np).to(dev)
    e_prime = g.GLOBAL_CACHE.get(ei, x_np.shape[0], True).n_edges
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timed_pass():
        pairs = []
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.fwd_bwd(x, ei)
            e1.record()
            pairs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in pairs)

    for _ in range(warmup):
        model.fwd_bwd(x, ei)
    torch.cuda.synchronize()
    l0 = lib.gat_launch_count()
    ms = timed_pass() / steps
    launches = int(lib.gat_launch_count() - l0) / steps
    with _lib.KernelTimer() as kt:
        model.fwd_bwd(x, ei)
        torch.cuda.synchronize()
        kt.records.clear()
        ms_pk = timed_pass()
    kernel_ms = sum(v["ms_total"] for v in kt.summary().values())
    return {"graph": name, "n_nodes": int(x_np.shape[0]), "n_edges_rewritten": int(e_prime), "layers": [list(sh) for sh in shapes],
            "ms_per_step": ms, "value": len(shapes) * e_prime / (ms * 1e-3), "unit": UNIT, "steps": steps, "warmup": warmup,
            "gpu_launches_per_step": launches, "kernel_time_share_of_step_per_kernel_pass": kernel_ms / ms_pk if ms_pk else None}


def extras(g, _lib, lib, dev):
    """Appended to the default single-GPU line so that the other figures BASELINE.json's metric names are driver-run too: the
    four small named shapes at 1 GPU and the PPI epoch time (tools/ppi_epoch.py: `PPI_GAT.training_step` in a plain loop)."""
    out = {"named_shapes": [], "ppi_epoch": None}
    for name in ("cora", "pubmed", "ppi", "pattern"):
        try:
            out["named_shapes"].append(measure_named_shape(g, _lib, lib, name, dev))
        except Exception as exc:   # never lose the headline line over an extra
            out["named_shapes"].append({"graph": name, "error": repr(exc)[:300]})
    try:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import ppi_epoch
        data = ppi_epoch.batches(10)
        secs, (loss, f1) = ppi_epoch.gpu_epoch(data, 2, 1.0)
        out["ppi_epoch"] = {"metric": "ppi_epoch_time", "value": secs, "unit": "s", "higher_is_better": False, "steps_per_epoch": len(data),
                            "attention_penalty": 1.0, "last_step": {"loss": loss, "train_f1": f1},
                            "ms_per_step_breakdown_synchronised": {k: round(v, 3) for k, v in ppi_epoch.gpu_epoch.breakdown.items()}}
        # the same epoch with the opt-in caller-side pieces: model_forward (skip / ELU in the layers' kernels, regulariser fused: no
        # attention tensors) and the micro-F1 counted on the device instead of sklearn on host copies (same value)
        secs2, (loss2, f12) = ppi_epoch.gpu_epoch(data, 2, 1.0, f1="gpu", fused=True)
        out["ppi_epoch"]["with_model_forward_and_device_f1"] = {"value": secs2, "unit": "s", "last_step": {"loss": loss2, "train_f1": f12}}
    except Exception as exc:
        out["ppi_epoch"] = {"error": repr(exc)[:300]}
    return out


def run_b200(args):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import gat_pytorch_b200 as g
    from gat_pytorch_b200 import _lib
    lib = _lib.load()

    x_np, ei_np, shapes, weights = make_workload(args.workload, args.scale)
    n = x_np.shape[0]
    x_host = torch.from_numpy(x_np).pin_memory()
    ei_host = torch.from_numpy(ei_np).pin_memory()

    single = None
    if world > 1:
        from gat_pytorch_b200.partition import PartitionedGAT
        model = PartitionedGAT(shapes, weights, x_host, ei_host, dev, fuse_glue=not args.unfused_glue)
        step_resident, step_e2e, e_prime = model.step_resident, model.step_e2e, model.n_edges_global

        def current_state():
            return model.last_loss, model.last_out, [(l.W.weight.grad, l.a.weight.grad) for l in model.layers]
    else:
        single = SingleGpuModel(g, shapes, weights, dev, args)
        x_dev, ei_dev = x_host.to(dev), ei_host.to(dev)
        e_prime = g.GLOBAL_CACHE.get(ei_dev, n, True).n_edges

        def step_resident():
            return single.fwd_bwd(x_dev, ei_dev)

        def current_state():
            return single.last_loss, single.last_out, [(l.W.weight.grad, l.a.weight.grad) for l in single.layers]

        copy_stream = torch.cuda.Stream(device=dev)

        def step_e2e():
            g.GLOBAL_CACHE.clear()                       # a freshly uploaded graph: structure is rebuilt
            main = torch.cuda.current_stream(dev)
            copy_stream.wait_stream(main)
            with torch.cuda.stream(copy_stream):         # edge list first: the CSR build overlaps the upload of x
                eid = ei_host.to(dev, non_blocking=True)
                ev_ei = torch.cuda.Event(); ev_ei.record(copy_stream)
                xd = x_host.to(dev, non_blocking=True)
                ev_x = torch.cuda.Event(); ev_x.record(copy_stream)
            eid.record_stream(main); xd.record_stream(main)
            main.wait_event(ev_ei)
            g.GLOBAL_CACHE.get(eid, n, True)             # Kernel 1 (one host read-back of the sizes) while x is in flight
            main.wait_event(ev_x)
            return float(single.fwd_bwd(xd, eid).item())  # D2H read of the step's result

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    n_layers = len(shapes)
    flush_buf = None if args.workload == "products" else torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = lib.gat_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if flush_buf is None:
        # products: the per-kernel CUDA events that feed `roofline` are recorded inside the timed region itself (the layer then
        # runs kernel by kernel from Python, which costs nothing next to 88 ms of kernels)
        with _lib.KernelTimer() as kt:
            barrier()
            ev0.record()
            for _ in range(args.steps):
                step_resident()
            ev1.record()
            barrier()
            ms_total = ev0.elapsed_time(ev1)
        launches = int(lib.gat_launch_count() - launches0)
        kernels = kt.summary()
    else:
        # L2-resident workloads (the small named graphs): flush between steps, time each step separately, on the PRODUCT path --
        # one C-ABI call per layer and direction (gat_layer_fwd / gat_layer_bwd).  The per-kernel breakdown comes from a second
        # pass of the same steps with the per-kernel timer on (same kernels, issued one by one from Python).
        def timed_pass():
            pairs = []
            for _ in range(args.steps):
                flush_buf.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step_resident()
                e1.record()
                pairs.append((e0, e1))
            barrier()
            return sum(a.elapsed_time(b) for a, b in pairs)
        barrier()
        ms_total = timed_pass()
        launches = int(lib.gat_launch_count() - launches0)
        with _lib.KernelTimer() as kt:
            step_resident()
            barrier()
            kt.records.clear()
            timed_pass()
        kernels = kt.summary()
    clocks = sampler.stop()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_layers * e_prime / (ms_step * 1e-3)

    # ---- end-to-end from pinned host buffers
    e2e = None
    if not args.no_e2e:
        step_e2e()
        barrier()
        ev0.record()
        for _ in range(args.steps):
            step_e2e()
        ev1.record()
        barrier()
        ms_e2e = ev0.elapsed_time(ev1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        e2e = {"value": n_layers * e_prime / (ms_e2e / args.steps * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(x_host.numel() * 4 + ei_host.numel() * 8), "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e2e / args.steps}

    # ---- driver-verifiable fingerprints of the step (every N), and at N > 1 the comparison with the single-GPU model
    step_resident()
    loss_t, out_t, grads_t = current_state()
    sums = checksums(loss_t, out_t, grads_t, world)
    parity = None
    if world > 1:
        if rank == 0:
            parity = parity_vs_single_gpu(g, shapes, weights, x_host, ei_host, dev, args, sums, grads_t)
        barrier()

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (hidden-layer shape), timed live above
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    f_in, nh, f, concat = shapes[min(1, n_layers - 1)]
    fp = (f + 3) // 4 * 4
    d = nh * fp
    n_local = n if world == 1 else model.n_local
    e_local = e_prime if world == 1 else model.n_edges_local
    per_kernel = {}
    for kname in ("gat_edge_fwd", "gat_edge_bwd_fused", "gat_edge_bwd_main", "gat_edge_bwd_rowsum", "gat_edge_bwd_finish"):
        rec = kernels.get((kname, (nh, fp)))
        if rec:
            # the source-major passes of a partitioned run walk ALL n source rows (with this rank's edges)
            n_rows = n if kname in ("gat_edge_bwd_fused", "gat_edge_bwd_main", "gat_edge_bwd_finish") else n_local
            b = algorithmic_bytes(kname, e_local, n_rows, nh, d, d)
            gbs = b / (rec["ms_avg"] * 1e-3) / 1e9
            per_kernel[kname] = {"ms_avg": rec["ms_avg"], "calls": rec["calls"], "algorithmic_bytes": b, "GBps": gbs, "frac": gbs / peak}
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms_avg"]) if per_kernel else None
    roofline = None
    if dom:
        pk = per_kernel[dom]
        # DRAM bytes of the same launch from an `ncu --set full` capture (profiles/traffic.json: full-size graph on ONE GPU);
        # a partitioned run launches the kernel on 1/P of the edges, for which no capture exists -> null
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and world == 1 and args.workload == "products" and args.scale == 1.0 and args.feature_dtype == "f32":
            traffic = json.load(open(tpath)).get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": pk["GBps"], "peak": peak, "unit": "GB/s", "frac": pk["frac"],
                    "traffic": traffic, "peak_source": peak_src, "ms_avg": pk["ms_avg"], "algorithmic_bytes": pk["algorithmic_bytes"]}
    total_kernel_ms = sum(v["ms_total"] for v in kernels.values())
    breakdown = {f"{k[0]}{'' if k[1] is None else list(k[1])}": round(v["ms_total"] / args.steps, 4) for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms_total"])}

    extra = None
    if world == 1 and args.workload == "products" and not args.no_extras:
        extra = extras(g, _lib, lib, dev)

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        step, e_cpu, l_cpu, n_cpu, kind = cpu_step_fn(args.workload, cpu_scale(args))
        step()
        best = float("inf")
        for _ in range(2):
            t0 = time.perf_counter()
            step()
            best = min(best, time.perf_counter() - t0)
        cpu_baseline = {"value": l_cpu * e_cpu / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                        "sample": cpu_sample_text(args, n_cpu, e_cpu, l_cpu, kind, "best of 2 after 1 warm-up")}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32" if args.feature_dtype == "f32" else "f32 accumulate, bf16 gathered features (variant; parity bar 2e-2)",
            "data": "synthetic", "config": dict(workload_config(args.workload, world), n_nodes=n, n_edges_rewritten=e_prime,
                                                 scale=args.scale, layers=[list(s) for s in shapes]),
            "e2e": e2e, "checksums": sums, "parity_vs_n1": parity, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "kernels_ms_per_step": breakdown, "edge_kernels": per_kernel,
            "kernel_time_share_of_step": total_kernel_ms / ms_total if ms_total else None, "extra": extra}
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
    if dist_env()[2] > 1 and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
