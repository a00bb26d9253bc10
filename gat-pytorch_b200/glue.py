"""Caller-side glue of the reference model stack, on the structures the layer already built.

`attention_norm(edge_index, attention_list)` mirrors `GATModel.calc_attention_norm(edge_index, attention_list)`
(models/GATModel.py:189-234): the mean over layers of  sum_{e,h} |alpha_l[e,h] * deg(dst_e) - 1| / E'.  The reference
rebuilds the per-edge degrees with a scatter_add and an index_select and runs five (E', NH) torch passes per layer; here
the degree is a `rowptr` difference of the cached CSR (the rewritten `edge_index` a layer returns is registered in the
structure cache, graph.py), the forward is one pass over alpha and the backward writes dL/dalpha in one pass
(include/gat_b200.h: gat_attention_norm_fwd / _bwd).  CUDA only, like the layer.
"""
from __future__ import annotations

import torch

from . import _lib
from .graph import GLOBAL_CACHE


class _AttentionNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, alpha, edge_dst, rowptr):
        lib = _lib.load()
        dev = alpha.device
        alpha = alpha.contiguous()
        e, nh = alpha.shape
        with torch.cuda.device(dev):
            s = torch.cuda.current_stream(dev).cuda_stream
            out = torch.empty((), dtype=torch.float32, device=dev)
            ws_bytes = int(lib.gat_attention_norm_workspace_bytes())
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.call("gat_attention_norm_fwd", edge_dst.data_ptr(), int(edge_dst.dtype == torch.int64), rowptr.data_ptr(),
                      alpha.data_ptr(), e, nh, out.data_ptr(), ws.data_ptr(), ws_bytes, s)
        ctx.save_for_backward(alpha, edge_dst, rowptr)
        return out

    @staticmethod
    def backward(ctx, grad):
        alpha, edge_dst, rowptr = ctx.saved_tensors
        dev = alpha.device
        e, nh = alpha.shape
        with torch.cuda.device(dev):
            s = torch.cuda.current_stream(dev).cuda_stream
            g = grad.to(torch.float32).contiguous()
            ga = torch.empty_like(alpha)
            _lib.call("gat_attention_norm_bwd", edge_dst.data_ptr(), int(edge_dst.dtype == torch.int64), rowptr.data_ptr(),
                      alpha.data_ptr(), e, nh, g.data_ptr(), ga.data_ptr(), s)
        return ga, None, None


def attention_norm(edge_index: torch.Tensor, attention_list, n_nodes: int | None = None) -> torch.Tensor:
    """`GATModel.calc_attention_norm` (GATModel.py:189-234) for the rewritten `edge_index` and the per-layer attention a
    stack of `GATLayer`s returned.  Differentiable w.r.t. every attention tensor."""
    if not edge_index.is_cuda:
        raise RuntimeError("gat_b200.attention_norm runs on CUDA only; there is no CPU fallback")
    if len(attention_list) == 0:
        raise ValueError("attention_list is empty")
    n = int(n_nodes) if n_nodes is not None else None
    st = GLOBAL_CACHE.find(edge_index) if n is None else GLOBAL_CACHE.get(edge_index, n, False)
    if st is None:
        raise RuntimeError("attention_norm: edge_index is not a rewritten edge list returned by a GATLayer on this graph; "
                           "pass n_nodes to build its structure")
    dst = edge_index[1]
    if dst.stride(0) != 1:
        dst = dst.contiguous()
    total = None
    for alpha in attention_list:
        if alpha.size(0) != st.n_edges:
            raise ValueError(f"attention has {alpha.size(0)} rows, the edge list {st.n_edges}")
        norm_l = _AttentionNorm.apply(alpha, dst, st.rowptr)
        total = norm_l if total is None else total + norm_l
    return total / len(attention_list)


def model_forward(model, data, return_attention_weights=None, attention_norm=False):
    """`GATModel.forward` (GATModel.py:118-151) / `GATModel.forward_and_return_attention` (:153-187) for a model whose
    `gat_layer_list` holds B200 `GATLayer`s, with the inter-layer glue folded into the layers' kernels:

      * the skip connection's rows (Identity / Linear of the layer input, head-averaged for a head-mean layer, :135-145) are
        added where the layer's output is written (`forward(..., skip=)`),
      * the ELU between layers (:148-149) is applied in the same place (`output_activation`),
      * the NEXT layer's input dropout (:130) as well (`output_dropout`, Philox mask regenerated in the backward) -- unless that
        layer has a skip connection and therefore needs the undropped tensor too; then, and for the raw input of the first
        layer, the dropout stays a torch op.

    `return_attention_weights=None` returns what `forward` returns (the output); True / False what
    `forward_and_return_attention(data, flag)` returns: `(x, edge_index, attention_weights_list)`.  Same arithmetic as the
    reference's op-by-op sequence; the layers' own attributes are restored before returning, so the unchanged callers keep
    working on the same modules.

    `attention_norm=True` appends `GATModel.calc_attention_norm(edge_index, attention_weights_list)` (GATModel.py:189-234: the
    mean over layers of sum |alpha*deg - 1| / E') to the result, computed by the layers from their score terms
    (`GATLayer.attention_norm`): the training steps of planetoid_gat.py:19-27 / ppi_gat.py:22-33 then need no attention tensor
    at all (`return_attention_weights=None`) and their backward stays the one-pass source-major kernel."""
    import torch.nn.functional as F
    x, edge_index = data.x, data.edge_index
    layers = model.gat_layer_list
    n_layers = len(layers)
    p = float(model.dropout) if model.training else 0.0
    want_attention = bool(return_attention_weights)
    attention, norms, skip_count, dropped = [], [], 0, False
    for i, layer in enumerate(layers):
        layer_input = x
        if p > 0.0 and not dropped:
            x = F.dropout(x, p=p, training=True)
        skip = None
        if model.add_skip_connection[i]:
            skip = model.skip_layer_list[skip_count](layer_input)
            skip_count += 1
            if not model.heads_concat_per_layer[i]:
                skip = skip.view(-1, model.num_heads_per_layer[i + 1], model.head_output_features_per_layer[i + 1]).mean(dim=1)
        last = i == n_layers - 1
        fold_dropout = (not last) and p > 0.0 and not model.add_skip_connection[i + 1]
        saved = (layer.output_activation, layer.output_dropout, layer.attention_norm)
        layer.output_activation, layer.output_dropout = (None if last else "elu"), (p if fold_dropout else 0.0)
        layer.attention_norm = bool(attention_norm)
        try:
            if return_attention_weights is None:
                x = layer(x, edge_index, skip=skip)
            else:
                res = layer(x, edge_index, return_attention_weights=want_attention, skip=skip)
                if want_attention:
                    x, (edge_index, alpha) = res
                    attention.append(alpha)
                else:       # the reference unpacks a tuple here and fails for False (GATModel.py:166); this path just works
                    x = res
                    attention.append(None)
        finally:
            layer.output_activation, layer.output_dropout, layer.attention_norm = saved
        if attention_norm:
            norms.append(layer.attention_norm_value)
        dropped = fold_dropout
    norm = (sum(norms) / len(norms),) if attention_norm else ()
    if return_attention_weights is None:
        return (x,) + norm if attention_norm else x
    return (x, edge_index, attention) + norm


def micro_f1(logits: torch.Tensor, y_true: torch.Tensor) -> float:
    """`sklearn.metrics.f1_score(y_pred=logits > 0, y_true=y_true, average="micro")` for multilabel indicator targets -- the metric
    `PPI_GAT` logs in every training / validation / test step (ppi_gat.py:38, :48, :56) after copying both (n, classes) matrices to
    the host -- from three integer counts taken on the device (include/gat_b200.h: gat_micro_f1_counts); one 24-byte read-back.
    A maintainer replaces the `f1_score(...)` call with `micro_f1(out, batch.y)`."""
    if not logits.is_cuda:
        raise RuntimeError("gat_b200.micro_f1 runs on CUDA only; there is no CPU fallback")
    if logits.shape != y_true.shape:
        raise ValueError(f"logits {tuple(logits.shape)} and y_true {tuple(y_true.shape)} differ in shape")
    lg = logits.detach().to(torch.float32).contiguous()
    yt = y_true.detach().to(device=lg.device, dtype=torch.float32).contiguous()
    with torch.cuda.device(lg.device):
        counts = torch.empty(3, dtype=torch.int64, device=lg.device)
        _lib.call("gat_micro_f1_counts", lg.data_ptr(), yt.data_ptr(), lg.numel(), counts.data_ptr(),
                  torch.cuda.current_stream(lg.device).cuda_stream)
    tp, fp, fn = (int(v) for v in counts.tolist())
    den = 2 * tp + fp + fn
    return 2.0 * tp / den if den else 0.0


def _structure_for(edge_index: torch.Tensor, n_nodes: int | None, who: str):
    if not edge_index.is_cuda:
        raise RuntimeError(f"gat_b200.{who} runs on CUDA only; there is no CPU fallback")
    st = GLOBAL_CACHE.find(edge_index) if n_nodes is None else GLOBAL_CACHE.get(edge_index, int(n_nodes), False)
    if st is None:
        raise RuntimeError(f"{who}: edge_index is not a rewritten edge list returned by a GATLayer on this graph; "
                           "pass n_nodes to build its structure")
    return st


def neighbourhood_entropy(edge_index: torch.Tensor, attention: torch.Tensor, n_nodes: int | None = None):
    """Visualisation feed.  For the rewritten `edge_index` and one layer's attention (E', NH) as returned by
    `GATLayer`, returns `(entropy (N, NH), uniform_entropy (N,))`:
    `entropy[i, h] == scipy.stats.entropy(attention[edge_index[1] == i, h], base=2)` and `uniform_entropy[i] ==
    scipy.stats.entropy(ones(deg_i) / deg_i, base=2)` -- the two lists `visualisation/entropy_histograms.py:103-115` builds
    with one full-edge-list mask per node, here one pass over the CSR segments."""
    st = _structure_for(edge_index, n_nodes, "neighbourhood_entropy")
    alpha = attention.detach().to(torch.float32).contiguous()
    if alpha.dim() != 2 or alpha.size(0) != st.n_edges:
        raise ValueError(f"attention must be (E', NH) with E' = {st.n_edges}")
    dev, nh = alpha.device, alpha.size(1)
    with torch.cuda.device(dev):
        ent = torch.empty((st.n, nh), dtype=torch.float32, device=dev)
        uni = torch.empty((st.n,), dtype=torch.float32, device=dev)
        _lib.call("gat_attention_entropy", st.rowptr.data_ptr(), st.eid.data_ptr(), st.n, alpha.data_ptr(), nh,
                  ent.data_ptr(), uni.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    return ent, uni


def degree_scaled_attention(edge_index: torch.Tensor, attention: torch.Tensor, n_nodes: int | None = None) -> torch.Tensor:
    """Visualisation feed: `attention[e, h] * in_degree(dst_e)` for every edge, ordered node by node and,
    inside a node, in edge-list order -- the concatenation `visualisation/weight_histograms.py:74-87` builds (before its
    `weight < 5` filter) with one full-edge-list mask per node.  Shape (E', NH)."""
    st = _structure_for(edge_index, n_nodes, "degree_scaled_attention")
    alpha = attention.detach().to(torch.float32).contiguous()
    if alpha.dim() != 2 or alpha.size(0) != st.n_edges:
        raise ValueError(f"attention must be (E', NH) with E' = {st.n_edges}")
    dev, nh = alpha.device, alpha.size(1)
    with torch.cuda.device(dev):
        out = torch.empty((st.n_edges, nh), dtype=torch.float32, device=dev)
        _lib.call("gat_attention_degree_scaled", st.rowptr.data_ptr(), st.eid.data_ptr(), st.n, alpha.data_ptr(), nh,
                  out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    return out


def neighbourhood_attention(edge_index: torch.Tensor, attention: torch.Tensor, node_ids, head: int, n_nodes: int | None = None):
    """Visualisation feed of `visualisation/neighbourhood_attention_weights.py:45-60`: for every node in
    `node_ids`, `(neighbour_ids, edge_widths)` where `neighbour_ids == edge_index[0][edge_index[1] == node]` and
    `edge_widths == attention[edge_index[1] == node, head] / max(...) * 60 / len(neighbour_ids)` -- what the script feeds to
    igraph, built from the node's CSR segment instead of a mask over the whole edge list.  Returns a list of (int64, float32)
    tensor pairs, one per requested node, in request order."""
    st = _structure_for(edge_index, n_nodes, "neighbourhood_attention")
    alpha = attention.detach().to(torch.float32).contiguous()
    if alpha.dim() != 2 or alpha.size(0) != st.n_edges:
        raise ValueError(f"attention must be (E', NH) with E' = {st.n_edges}")
    dev, nh = alpha.device, alpha.size(1)
    if not 0 <= int(head) < nh:
        raise IndexError(f"head {head} out of range for {nh} heads")
    nodes = torch.as_tensor(list(node_ids), dtype=torch.int64, device=dev)
    if nodes.numel() and (int(nodes.min()) < 0 or int(nodes.max()) >= st.n):
        raise IndexError("node id out of range")
    with torch.cuda.device(dev):
        deg = (st.rowptr[nodes + 1] - st.rowptr[nodes]).to(torch.int64)
        off = torch.cumsum(deg, 0) - deg
        sizes = deg.tolist()
        total = int(sum(sizes))
        src = torch.empty(total, dtype=torch.int64, device=dev)
        w = torch.empty(total, dtype=torch.float32, device=dev)
        _lib.call("gat_attention_neighbourhood", st.rowptr.data_ptr(), st.col.data_ptr(), st.eid.data_ptr(), alpha.data_ptr(), nh, int(head),
                  nodes.data_ptr(), nodes.numel(), off.data_ptr(), src.data_ptr(), w.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    return list(zip(torch.split(src, sizes), torch.split(w, sizes)))
