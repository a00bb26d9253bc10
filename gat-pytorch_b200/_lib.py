"""ctypes binding of libgat_b200.so (the C ABI declared in include/gat_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgat_b200.so")
LONG_ROW_EDGES = 256   # GAT_LONG_ROW_EDGES in include/gat_b200.h
CSRC = os.path.join(_HERE, "csrc")

_lock = threading.Lock()
_lib = None


class LayerDesc(ctypes.Structure):
    """struct gat_layer_desc of include/gat_b200.h (graph structure + layer configuration + parameter pointers)."""
    _fields_ = [(n, c_void_p) for n in ("rowptr", "col", "eid", "order", "rowptr_t", "col_t", "pos_t", "order_t", "tpos")] + [
        ("n_long", c_int64), ("n_long_t", c_int64), ("n", c_int64), ("n_edges", c_int64), ("f_in", c_int64),
        ("nh", ctypes.c_int32), ("f", ctypes.c_int32), ("fp", ctypes.c_int32),
        ("concat", ctypes.c_int32), ("const_attention", ctypes.c_int32), ("x_act", ctypes.c_int32), ("out_act", ctypes.c_int32),
        ("gemm_algo", ctypes.c_int32), ("p_drop", c_float), ("seed", c_uint64), ("W", c_void_p), ("a", c_void_p),
        ("skip", c_void_p), ("ld_skip", c_int64), ("out_drop_p", c_float), ("out_drop_seed", c_uint64), ("grad_skip", c_void_p),
        ("norm_out", c_void_p), ("grad_norm", c_void_p)]


# name -> (restype, argtypes); mirrors include/gat_b200.h one to one.
_P = c_void_p
SIGNATURES = {
    "gat_version": (c_int, []),
    "gat_last_error": (c_char_p, []),
    "gat_launch_count": (ctypes.c_ulonglong, []),
    "gat_edges_scan": (c_int, [_P, c_int64, c_int64, c_int, _P, _P]),
    "gat_csr_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "gat_csr_build": (c_int, [_P, c_int64, c_int64, c_int, c_int, c_int64, c_int64, c_int64,
                              _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gat_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int64, c_int64, c_int64, c_int]),
    "gat_gemm_tc_supported": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64]),
    "gat_gemm": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, _P, c_int64, _P, c_int64, _P, c_int64,
                         c_int, _P, c_size_t, _P]),
    "gat_gemm_ex": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, _P, c_int64, _P, c_int64, _P, c_int64,
                            c_int, c_int, _P, c_int64, c_int, _P, c_size_t, _P]),
    "gat_project_fwd": (c_int, [_P, c_int64, c_int64, c_int64, c_int, _P, c_int64, c_int, _P, _P, c_int, _P, _P, _P,
                                c_int, _P, c_size_t, _P]),
    "gat_project_fwd_allgather": (c_int, [_P, c_int64, c_int64, c_int64, c_int, _P, c_int64, c_int, _P, _P, c_int, _P, c_int, c_int64,
                                          _P, _P, _P]),
    "gat_scores_fwd": (c_int, [_P, c_int64, c_int, _P, _P, c_int, _P, _P, _P]),
    "gat_scores_bwd_workspace_bytes": (c_size_t, [c_int, c_int]),
    "gat_scores_bwd": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gat_edge_max": (c_int, [_P, _P, _P, c_int64, c_int64, _P, _P, c_int, _P, _P, c_size_t, _P]),
    "gat_edge_fwd_workspace_bytes": (c_size_t, []),
    "gat_edge_fwd": (c_int, [_P, _P, _P, _P, c_int64, c_int64, _P, c_int, c_int, _P, _P, _P, c_int, c_float, c_uint64, c_uint64,
                             _P, c_int, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gat_edge_fwd_glue": (c_int, [_P, _P, _P, _P, c_int64, c_int64, _P, c_int, c_int, _P, _P, _P, c_int, c_float, c_uint64, c_uint64,
                                  _P, c_int, _P, c_int64, c_float, c_uint64, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gat_edge_bwd_rowdot_glue": (c_int, [_P, c_int, _P, c_int, _P, c_int64, c_float, c_uint64, _P, _P, _P, _P, c_float,
                                         _P, c_int64, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gat_edge_bwd_fused_norm": (c_int, [_P, _P, _P, _P, c_int64, _P, c_int64, _P, c_int, c_int, _P, _P, _P, _P,
                                        c_float, c_uint64, c_uint64, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_float,
                                        _P, _P, _P, _P, c_size_t, _P]),
    "gat_attention_norm_scores": (c_int, [_P, _P, c_int64, c_int64, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "gat_head_merge_fwd_glue": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, _P, c_int64, c_int, c_float, c_uint64, _P, _P]),
    "gat_out_glue_adjoint": (c_int, [_P, _P, c_int64, c_int, c_int, c_float, c_uint64, _P, _P]),
    "gat_head_merge_fwd": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, _P, _P]),
    "gat_head_merge_bwd": (c_int, [_P, c_int64, c_int, c_int, c_int, c_int, _P, _P]),
    "gat_edge_bwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "gat_edge_bwd_main": (c_int, [_P, _P, _P, _P, c_int64, _P, c_int64, _P, c_int, c_int, _P, _P, _P, _P, c_int,
                                  c_float, c_uint64, c_uint64, _P, c_int, _P, _P, _P, _P, c_size_t, _P]),
    "gat_edge_bwd_fused": (c_int, [_P, _P, _P, _P, c_int64, _P, c_int64, _P, c_int, c_int, _P, _P, _P, _P,
                                   c_float, c_uint64, c_uint64, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64,
                                   _P, _P, _P, _P, c_int, c_int, c_int64, _P, c_size_t, _P]),
    "gat_attention_norm_workspace_bytes": (c_size_t, []),
    "gat_attention_norm_fwd": (c_int, [_P, c_int, _P, _P, c_int64, c_int, _P, _P, c_size_t, _P]),
    "gat_attention_norm_bwd": (c_int, [_P, c_int, _P, _P, c_int64, c_int, _P, _P, _P]),
    "gat_f32_to_bf16": (c_int, [_P, _P, c_int64, _P]),
    "gat_f32_round_bf16": (c_int, [_P, _P, c_int64, _P]),
    "gat_edge_bf16_native": (c_int, [c_int, c_int, c_int]),
    "gat_edge_fwd_bf16": (c_int, [_P, _P, _P, _P, c_int64, c_int64, _P, c_int, c_int, _P, _P, _P, c_int, c_float, c_uint64, c_uint64,
                             _P, c_int, _P, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gat_edge_bwd_fused_bf16": (c_int, [_P, _P, _P, _P, c_int64, _P, c_int64, _P, c_int, c_int, _P, _P, _P, _P,
                                   c_float, c_uint64, c_uint64, _P, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64,
                                   _P, _P, _P, _P, c_int, c_int, c_int64, _P, c_size_t, _P]),
    "gat_micro_f1_counts": (c_int, [_P, _P, c_int64, _P, _P]),
    "gat_attention_entropy": (c_int, [_P, _P, c_int64, _P, c_int, _P, _P, _P]),
    "gat_attention_degree_scaled": (c_int, [_P, _P, c_int64, _P, c_int, _P, _P]),
    "gat_attention_neighbourhood": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, c_int64, _P, _P, _P, _P]),
    "gat_slab_sum": (c_int, [_P, c_int, c_int64, c_int, _P, _P]),
    "gat_head_mean_bwd_shared": (c_int, [_P, c_int64, c_int, c_int, c_int, _P, _P]),
    "gat_edge_bwd_rowsum": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gat_tgt_pack_stride": (c_int, [c_int]),
    "gat_edge_bwd_rowdot": (c_int, [_P, c_int, _P, c_int, _P, _P, c_int64, c_int, c_int, _P, _P, _P, _P, _P, c_size_t, _P]),
    "gat_edge_bwd_finish": (c_int, [_P, _P, _P, c_int64, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int64,
                                    _P, _P, _P, _P, c_size_t, _P]),
    "gat_edge_bwd_gamma": (c_int, [_P, c_size_t, _P, _P]),
    "gat_pack_params": (c_int, [_P, _P, c_int, c_int, c_int, c_int64, _P, _P, _P, _P, _P]),
    "gat_unpack_param_grads": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_int64, _P, _P, _P]),
    "gat_layer_fwd_arena_bytes": (c_size_t, [_P]),
    "gat_layer_bwd_scratch_bytes": (c_size_t, [_P, c_int, c_int, c_int, c_int]),
    "gat_layer_fwd": (c_int, [_P, _P, c_int64, _P, c_size_t, _P, _P, c_int, _P]),
    "gat_layer_bwd": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, _P, c_size_t, _P, _P, _P, _P]),
}


def build(verbose: bool = False) -> str:
    """Compile libgat_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    res = subprocess.run(["make", "-j8", "-C", CSRC], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout[-4000:])
        print(res.stderr[-4000:])
    if res.returncode != 0:
        raise RuntimeError("building libgat_b200.so failed")
    return LIB_PATH


def load() -> ctypes.CDLL:
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the B200 GAT layer has no CPU or PyTorch fallback. "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or `make -C gat-pytorch_b200/csrc`.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)   # AttributeError here = header and library out of sync
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gat_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


class KernelTimer:
    """Optional per-entry-point CUDA-event timing (bench.py's live roofline measurement).  While active,
    every C-ABI call made through `call()` is bracketed by two events on the current stream; no
    synchronisation happens until `summary()`."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _timer
        _timer = self
        return self

    def __exit__(self, *exc):
        global _timer
        _timer = None

    def summary(self):
        import torch
        torch.cuda.synchronize()
        out = {}
        for name, tag, e0, e1 in self.records:
            d = out.setdefault((name, tag), [0, 0.0])
            d[0] += 1
            d[1] += e0.elapsed_time(e1)
        return {k: dict(calls=v[0], ms_total=v[1], ms_avg=v[1] / v[0]) for k, v in out.items()}


_timer = None


class timed:
    """Bracket an arbitrary region of the current stream (a collective, a torch op) with the active KernelTimer's events,
    so that bench.py's per-step breakdown also shows the time between our kernels; free when no timer is active."""

    def __init__(self, name, tag=None):
        self.name, self.tag = name, tag

    def __enter__(self):
        if _timer is not None:
            import torch
            self.e0, self.e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _timer is not None and hasattr(self, "e0"):
            self.e1.record()
            _timer.records.append((self.name, self.tag, self.e0, self.e1))


def call(name: str, *args, tag=None):
    """Invoke C-ABI entry point `name`; raise on a non-zero status."""
    fn = getattr(load(), name)
    if _timer is None:
        rc = fn(*args)
    else:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        _timer.records.append((name, tag, e0, e1))
    check(rc, name)
