"""CTA-pair tcgen05 GEMM (csrc/gemm_pair.cu) vs fp64 torch: accuracy of C and of the fused score columns, ELU operand /
ELU' output glue, ragged shapes; timing on the products shapes.  Run on the GPU box; GAT_GEMM_PAIR=0 times the
one-CTA-per-tile kernel for comparison."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gat_pytorch_b200 import _lib
from gat_pytorch_b200.gat_layer import gemm

lib = _lib.load()
dev = "cuda"


def rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def project(m, n, k, nh, x_act=False, reps=0, seed=0):
    torch.manual_seed(seed + m + 3 * n + 7 * k)
    x = torch.randn((m, k), device=dev)
    w = torch.randn((n, k), device=dev) / k ** 0.5
    a_src = torch.randn((nh, n), device=dev) / n ** 0.5
    a_tgt = torch.randn((nh, n), device=dev) / n ** 0.5
    wh = torch.full((m, n), float("nan"), device=dev)
    s_src = torch.full((m, nh), float("nan"), device=dev)
    s_tgt = torch.full((m, nh), float("nan"), device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        _lib.call("gat_project_fwd", x.data_ptr(), m, k, x.stride(0), int(x_act), w.data_ptr(), w.stride(0), n,
                  a_src.data_ptr(), a_tgt.data_ptr(), nh, wh.data_ptr(), s_src.data_ptr(), s_tgt.data_ptr(), 2, None, 0, st)
    run()
    torch.cuda.synchronize()
    xd = torch.nn.functional.elu(x.double()) if x_act else x.double()
    want = xd @ w.double().T
    e_wh = rel(wh, want)
    # the reference forms the scores from the fp32-rounded Wh (gat_layer.py:76-82); compare with both
    e_ss = rel(s_src, want @ a_src.double().T)
    e_st = rel(s_tgt, want @ a_tgt.double().T)
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            run()
        e0.record()
        for _ in range(reps):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return e_wh, e_ss, e_st, ms


def plain(m, n, k, act_a=False, mul=False, reps=0):
    torch.manual_seed(m + 3 * n + 7 * k + 1)
    a = torch.randn((m, k), device=dev)
    b = torch.randn((n, k), device=dev)
    c = torch.full((m, n), float("nan"), device=dev)
    msrc = torch.randn((m, n), device=dev) if mul else None
    gemm(False, True, m, n, k, a, a.stride(0), b, b.stride(0), c, n, algo=2, act_a=act_a, mul_elu_grad=msrc)
    torch.cuda.synchronize()
    ad = torch.nn.functional.elu(a.double()) if act_a else a.double()
    want = ad @ b.double().T
    if mul:
        md = msrc.double()
        want = want * torch.where(md > 0, torch.ones_like(md), md.exp())
    err = rel(c, want)
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            gemm(False, True, m, n, k, a, a.stride(0), b, b.stride(0), c, n, algo=2, act_a=act_a, mul_elu_grad=msrc)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return err, ms


print("GAT_GEMM_PAIR =", os.environ.get("GAT_GEMM_PAIR", "1"), flush=True)
for (m, n, k, nh, act) in [(16384, 256, 256, 4, False), (20001, 256, 100, 4, False), (19133, 192, 256, 4, True),
                           (33000, 64, 64, 8, False), (16500, 128, 48, 1, False), (70001, 100, 256, 2, True),
                           (16384 * 3 + 5, 200, 36, 4, False)]:
    e_wh, e_ss, e_st, _ = project(m, n, k, nh, act)
    print(f"project m={m} n={n} k={k} nh={nh} elu_in={act}: err wh {e_wh:.2e} s_src {e_ss:.2e} s_tgt {e_st:.2e}", flush=True)
for (m, n, k, act, mul) in [(16384, 256, 256, False, False), (20001, 100, 256, False, True), (50000, 256, 192, True, True),
                            (16385, 8, 16, False, False), (25000, 72, 252, False, False)]:
    err, _ = plain(m, n, k, act, mul)
    print(f"plain   m={m} n={n} k={k} elu_in={act} elu_grad_out={mul}: err {err:.2e}", flush=True)
M = 2449029
for (n, k, nh) in [(256, 256, 4), (256, 100, 4), (192, 256, 4)]:
    e_wh, e_ss, e_st, ms = project(M, n, k, nh, n == 192, reps=5)
    gb = 4.0 * M * (n + k) / 1e9
    print(f"project m={M} n={n} k={k}: {ms:.3f} ms  {gb / ms:.0f} GB/s  {3 * 2 * M * n * k / ms / 1e9:.0f} TF/s(3xTF32)  err wh {e_wh:.2e} s {max(e_ss, e_st):.2e}", flush=True)
for (n, k, mul) in [(256, 256, True), (100, 256, False), (256, 192, True)]:
    err, ms = plain(M, n, k, False, mul, reps=5)
    gb = 4.0 * M * (n + k + (n if mul else 0)) / 1e9
    print(f"dX-like m={M} n={n} k={k} elu_grad_out={mul}: {ms:.3f} ms  {gb / ms:.0f} GB/s  err {err:.2e}", flush=True)

print("---- TN (dW): CTA-pair kernel for K >= 65536 ----", flush=True)
def tn(m, n, k, act_b=False, reps=0):
    torch.manual_seed(k + m)
    a = torch.randn((k, m), device=dev); b = torch.randn((k, n), device=dev)
    c = torch.full((m, n), float("nan"), device=dev)
    gemm(True, False, m, n, k, a, m, b, n, c, n, algo=2, act_b=act_b)
    torch.cuda.synchronize()
    bd = torch.nn.functional.elu(b.double()) if act_b else b.double()
    want = a.double().T @ bd
    err = rel(c, want)
    ms = None
    if reps:
        for _ in range(2): gemm(True, False, m, n, k, a, m, b, n, c, n, algo=2, act_b=act_b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): gemm(True, False, m, n, k, a, m, b, n, c, n, algo=2, act_b=act_b)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return err, ms
for (m, n, k, act) in [(256, 256, 70000, False), (192, 256, 100003, False), (256, 100, 66000, True), (64, 72, 131072, False), (8, 8, 65536, False)]:
    err, _ = tn(m, n, k, act)
    print(f"TN m={m} n={n} k={k} elu_b={act}: err {err:.2e}", flush=True)
for (m, n) in [(256, 256), (192, 256), (256, 100)]:
    err, ms = tn(m, n, M, False, reps=5)
    print(f"TN m={m} n={n} k={M}: {ms:.3f} ms err {err:.2e}", flush=True)
