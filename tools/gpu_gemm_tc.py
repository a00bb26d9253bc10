"""tcgen05 3xTF32 GEMM vs fp64 torch and vs the FFMA path: accuracy and timing (run on the GPU box)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gat_pytorch_b200.gat_layer import gemm
from gat_pytorch_b200 import _lib

def run(m, n, k, ta, tb, algo, reps=0, scale=1.0):
    torch.manual_seed(m * 7 + n * 3 + k)
    a = torch.randn((k, m) if ta else (m, k), device="cuda") * scale
    b = torch.randn((n, k) if tb else (k, n), device="cuda")
    c = torch.full((m, n), float("nan"), device="cuda")
    gemm(ta, tb, m, n, k, a, a.stride(0), b, b.stride(0), c, n, algo=algo)
    torch.cuda.synchronize()
    want = (a.double().T if ta else a.double()) @ (b.double().T if tb else b.double())
    err = ((c.double() - want).abs().max() / want.abs().max()).item()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gemm(ta, tb, m, n, k, a, a.stride(0), b, b.stride(0), c, n, algo=algo)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return err, ms

shapes = [(128, 64, 32), (128, 256, 256), (300, 192, 100), (1000, 72, 520), (4097, 256, 1024), (257, 128, 36), (5000, 1024, 48)]
lib = _lib.load()
for (m, n, k) in shapes:
    ok = lib.gat_gemm_tc_supported(0, 1, m, n, k, k, k, n)
    e2, _ = run(m, n, k, False, True, 2) if ok else (None, None)
    e1, _ = run(m, n, k, False, True, 1)
    print(f"NT m={m} n={n} k={k}: tc_supported={ok} err_tc={e2} err_ffma={e1}", flush=True)
for (m, n, k) in [(2449029, 256, 256), (2449029, 256, 100), (2449029, 192, 256)]:
    e2, t2 = run(m, n, k, False, True, 2, reps=5)
    e1, t1 = run(m, n, k, False, True, 1, reps=3)
    print(f"NT m={m} n={n} k={k}: tc {t2:.3f} ms err {e2:.2e} | ffma {t1:.3f} ms err {e1:.2e} | {2*m*n*k/t2/1e9:.1f} TFLOP/s(fp32-equivalent)", flush=True)

print("---- TN (MN-major operands, split-K) ----", flush=True)
for (m, n, k) in [(128, 64, 32), (256, 256, 4096), (192, 256, 10000), (256, 100, 5000), (64, 72, 3333), (1024, 1024, 4800), (260, 136, 70000)]:
    ok = lib.gat_gemm_tc_supported(1, 0, m, n, k, m, n, n)
    e2, _ = run(m, n, k, True, False, 2) if ok else (None, None)
    e1, _ = run(m, n, k, True, False, 1)
    print(f"TN m={m} n={n} k={k}: tc_supported={ok} err_tc={e2} err_ffma={e1}", flush=True)
for (m, n, k) in [(256, 256, 2449029), (192, 256, 2449029), (256, 100, 2449029)]:
    e2, t2 = run(m, n, k, True, False, 2, reps=5)
    e1, t1 = run(m, n, k, True, False, 1, reps=3)
    print(f"TN m={m} n={n} k={k}: tc {t2:.3f} ms err {e2:.2e} | ffma {t1:.3f} ms err {e1:.2e}", flush=True)
