"""Graph structure for the edge kernels: rewritten edge list + CSR by target + CSR by source.

Host side of Kernel 1 (include/gat_b200.h: gat_edges_scan / gat_csr_build).  Replaces
`add_remaining_self_loops` / `maybe_num_nodes` (models/utils.py:47-72) and the per-call edge
grouping that every scatter_add_/index_select in models/gat_layer.py:99-127 implies.

The structure depends only on (edge_index, N, add_self_loops), so it is cached per edge_index
tensor; the one host read-back (N_idx and the kept-edge count, which size E') happens only for a
graph not seen before, where the reference syncs on every call (utils.py:72).
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass

import torch

from . import _lib


@dataclass
class GraphStructure:
    n: int                      # rows of x
    n_idx: int                  # max(edge_index)+1 (utils.py:72)
    n_edges: int                # E' after the self-loop rewrite
    edge_index: torch.Tensor    # (2, E') in the reference's rewritten order, dtype of the input
    rowptr: torch.Tensor        # (n+1,) int32 -- diffs are the in-degrees of GATModel.py:196-201
    col: torch.Tensor           # (E',) int32 source ids, grouped by target, stable
    eid: torch.Tensor           # (E',) int32 position of each CSR slot in `edge_index`
    rowptr_t: torch.Tensor      # (n+1,) int32
    col_t: torch.Tensor         # (E',) int32 target ids, grouped by source, stable
    pos_t: torch.Tensor         # (E',) int32 CSR-by-target slot of each CSR-by-source slot
    tpos: torch.Tensor = None       # (E',) int32 CSR-by-source slot of each CSR-by-target slot (inverse of pos_t)
    order: torch.Tensor = None      # (n,) int32 scheduling permutation of the target rows (long rows first)
    order_t: torch.Tensor = None    # (n,) int32 same for the source rows
    n_long: int = -1                # rows of `order` with more than GAT_LONG_ROW_EDGES edges (-1: unknown)
    n_long_t: int = -1              # same for `order_t`

    def in_degrees(self) -> torch.Tensor:
        return self.rowptr[1:] - self.rowptr[:-1]


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def build_structure(edge_index: torch.Tensor, n_nodes: int, add_self_loops: bool) -> GraphStructure:
    if not edge_index.is_cuda:
        raise RuntimeError("gat_b200: edge_index must be a CUDA tensor (no CPU fallback)")
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError(f"edge_index must have shape (2, E), got {tuple(edge_index.shape)}")
    if edge_index.dtype not in (torch.int64, torch.int32):
        raise TypeError(f"edge_index must be int64 or int32, got {edge_index.dtype}")
    lib = _lib.load()
    dev = edge_index.device
    ei = edge_index if edge_index.stride(1) == 1 or edge_index.size(1) <= 1 else edge_index.contiguous()
    if ei.size(1) <= 1:
        ei = ei.contiguous()
    n_in = ei.size(1)
    is64 = int(ei.dtype == torch.int64)
    with torch.cuda.device(dev):
        st = _stream(dev)
        stats = torch.empty(3, dtype=torch.int64, device=dev)
        _lib.call("gat_edges_scan", ei.data_ptr(), n_in, ei.stride(0), is64, stats.data_ptr(), st)
        n_idx, n_keep, vmin = (int(v) for v in stats.tolist())   # the one host sync per new graph
        if vmin < 0:
            raise IndexError("edge_index contains negative node ids")
        if n_idx > n_nodes:
            raise IndexError(f"index {n_idx - 1} is out of bounds for dimension 0 with size {n_nodes}")
        n_out = n_keep + n_idx if add_self_loops else n_in
        i32 = dict(dtype=torch.int32, device=dev)
        rowptr = torch.empty(n_nodes + 1, **i32)
        rowptr_t = torch.empty(n_nodes + 1, **i32)
        col, eid, col_t, pos_t, tpos = (torch.empty(n_out, **i32) for _ in range(5))
        order, order_t = torch.empty(n_nodes, **i32), torch.empty(n_nodes, **i32)
        ei_out = torch.empty((2, n_out), dtype=torch.int64, device=dev) if add_self_loops else None
        ws_bytes = int(lib.gat_csr_workspace_bytes(n_in, n_out, n_nodes))
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
        n_long = torch.empty(2, dtype=torch.int64, device=dev)
        _lib.call("gat_csr_build", ei.data_ptr(), n_in, ei.stride(0), is64, int(add_self_loops), n_idx, n_out, n_nodes,
                                     ei_out.data_ptr() if ei_out is not None else None,
                                     rowptr.data_ptr(), col.data_ptr(), eid.data_ptr(),
                                     rowptr_t.data_ptr(), col_t.data_ptr(), pos_t.data_ptr(), tpos.data_ptr(),
                                     order.data_ptr(), order_t.data_ptr(), n_long.data_ptr(), ws.data_ptr(), ws_bytes, st)
        n_long, n_long_t = (int(v) for v in n_long.tolist())   # sizes the cooperative long-row launches
    if ei_out is None:
        ei_ret = edge_index
    else:
        ei_ret = ei_out if edge_index.dtype == torch.int64 else ei_out.to(edge_index.dtype)
    return GraphStructure(n_nodes, n_idx, n_out, ei_ret, rowptr, col, eid, rowptr_t, col_t, pos_t, tpos, order, order_t,
                          n_long, n_long_t)


class StructureCache:
    """LRU keyed on the identity of the edge_index tensor (storage pointer, shape, strides, version
    counter), N and the self-loop flag.  The rewritten edge list returned by a layer is registered as
    an alias of its own structure: `GATModel.forward_and_return_attention` feeds it to the next layer
    (GATModel.py:166), and rewriting a rewritten list is the identity."""

    def __init__(self, capacity: int = 16):
        self.capacity = capacity
        self._items: "OrderedDict[tuple, tuple]" = OrderedDict()

    @staticmethod
    def _key(edge_index, n_nodes, add_self_loops):
        return (edge_index.data_ptr(), tuple(edge_index.shape), tuple(edge_index.stride()), edge_index.dtype,
                edge_index._version, str(edge_index.device), int(n_nodes), bool(add_self_loops))

    def get(self, edge_index, n_nodes, add_self_loops) -> GraphStructure:
        key = self._key(edge_index, n_nodes, add_self_loops)
        hit = self._items.get(key)
        if hit is not None:
            self._items.move_to_end(key)
            return hit[1]
        s = build_structure(edge_index, n_nodes, add_self_loops)
        self._put(key, edge_index, s)
        if add_self_loops and s.edge_index is not edge_index:
            self._put(self._key(s.edge_index, n_nodes, True), s.edge_index, s)
        return s

    def find(self, edge_index):
        """Structure of an edge list some layer has already seen (any N / self-loop flag), or None."""
        probe = self._key(edge_index, 0, False)[:6]
        for key, (_, s) in reversed(self._items.items()):
            if key[:6] == probe and s.edge_index.data_ptr() == edge_index.data_ptr():
                return s
        return None

    def _put(self, key, tensor, s):
        self._items[key] = (tensor, s)   # holding `tensor` keeps its storage (and data_ptr) alive
        while len(self._items) > self.capacity:
            self._items.popitem(last=False)

    def clear(self):
        self._items.clear()


GLOBAL_CACHE = StructureCache()
