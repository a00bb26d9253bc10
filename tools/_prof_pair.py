import os, sys
sys.path.insert(0, "/root/repo")
import torch
from gat_pytorch_b200 import _lib
dev = "cuda"; M = 2449029; n = k = 256; nh = 4
x = torch.randn((M, k), device=dev); w = torch.randn((n, k), device=dev) / 16
a_src = torch.randn((nh, n), device=dev); a_tgt = torch.randn((nh, n), device=dev)
wh = torch.empty((M, n), device=dev); s_src = torch.empty((M, nh), device=dev); s_tgt = torch.empty((M, nh), device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    _lib.call("gat_project_fwd", x.data_ptr(), M, k, k, 0, w.data_ptr(), k, n, a_src.data_ptr(), a_tgt.data_ptr(), nh,
              wh.data_ptr(), s_src.data_ptr(), s_tgt.data_ptr(), 2, None, 0, st)
    _lib.call("gat_gemm_ex", 0, 1, M, n, k, x.data_ptr(), k, w.data_ptr(), k, wh.data_ptr(), n, 0, 0, None, 0, 2, None, 0, st)
torch.cuda.synchronize()
print("ok")
