"""Diagnostic sweep run on the GPU box: per-case parity errors (no asserts) + quick kernel timings."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases, gat_oracle as O
from test_gpu_parity import run_cuda, run_oracle, CASE_NAMES

def main():
    print(torch.cuda.get_device_name(0), flush=True)
    allc = {c["name"]: c for c in cases.small_cases()}
    only = sys.argv[1:] or CASE_NAMES
    for name in only:
        case = allc[name]
        try:
            got, ei2 = run_cuda(case)
            torch.cuda.synchronize()
            fw, want = run_oracle(case)
            errs = {k: float("%.3g" % O.rel_err(got[k], want[k])) for k in want}
            print(name, "ei_ok=%s" % np.array_equal(ei2, fw["edge_index"]), errs, flush=True)
        except Exception as e:  # noqa
            print(name, "EXCEPTION", repr(e)[:300], flush=True)
            if "CUDA" in repr(e) or "cuda" in repr(e):
                break

if __name__ == "__main__":
    main()
