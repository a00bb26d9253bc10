"""Per-call CUDA-event timing of the GEMM entry points on the products shapes (min / median over reps, after warm-up)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gat_pytorch_b200 import _lib
from gat_pytorch_b200.gat_layer import gemm
dev = "cuda"
M = int(os.environ.get("M", 2449029))

def timeit(fn, reps=12, warm=3):
    for _ in range(warm): fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for e0, e1 in ev:
        e0.record(); fn(); e1.record()
    torch.cuda.synchronize()
    ts = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    return ts[0], ts[len(ts) // 2]

st = torch.cuda.current_stream().cuda_stream
print("GAT_GEMM_PAIR =", os.environ.get("GAT_GEMM_PAIR", "1"), "M =", M, flush=True)
for (n, k, nh, act) in [(256, 256, 4, False), (256, 100, 4, False), (192, 256, 4, True)]:
    x = torch.randn((M, k), device=dev); w = torch.randn((n, k), device=dev) / k ** 0.5
    a_src = torch.randn((nh, n), device=dev); a_tgt = torch.randn((nh, n), device=dev)
    wh = torch.empty((M, n), device=dev); s_src = torch.empty((M, nh), device=dev); s_tgt = torch.empty((M, nh), device=dev)
    fn = lambda: _lib.call("gat_project_fwd", x.data_ptr(), M, k, k, int(act), w.data_ptr(), k, n, a_src.data_ptr(), a_tgt.data_ptr(), nh,
                           wh.data_ptr(), s_src.data_ptr(), s_tgt.data_ptr(), 2, None, 0, st)
    lo, med = timeit(fn)
    gb = 4.0 * M * (n + k) / 1e6
    print(f"project  n={n} k={k} elu_in={act}: min {lo:.3f} med {med:.3f} ms   {gb / med:.0f} GB/s algorithmic   {6.0 * M * n * k / med / 1e9:.0f} TF/s (3xTF32)", flush=True)
    del x, wh
for (n, k, mul) in [(256, 256, True), (256, 256, False), (100, 256, False), (256, 192, True)]:
    a = torch.randn((M, k), device=dev); b = torch.randn((n, k), device=dev); c = torch.empty((M, n), device=dev)
    msrc = torch.randn((M, n), device=dev) if mul else None
    ws_bytes = int(_lib.load().gat_gemm_workspace_bytes(0, 1, M, n, k, 2))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
    fn = lambda: _lib.call("gat_gemm_ex", 0, 1, M, n, k, a.data_ptr(), k, b.data_ptr(), k, c.data_ptr(), n, 0, 0,
                           msrc.data_ptr() if mul else None, n if mul else 0, 2, ws.data_ptr(), ws_bytes, st)
    lo, med = timeit(fn)
    gb = 4.0 * M * (n + k + (n if mul else 0)) / 1e6
    print(f"dX-like  n={n} k={k} elu_grad_out={mul}: min {lo:.3f} med {med:.3f} ms   {gb / med:.0f} GB/s algorithmic", flush=True)
    del a, c, msrc
