"""Torch CPU port of the reference layer's formulation -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Used by `bench.py` for the `cpu_baseline` leg and `--impl reference` (the reference itself is
Python and `/root/reference` does not exist on the GPU box, so the reference arm times this port:
`cpu_baseline.kind = "port"`).  It performs the same ATen op sequence as
`/root/reference/models/gat_layer.py:53-140` (index, cat, mm, max, leaky_relu, exp, scatter_add_,
index_select, div, mul, scatter_add_, view/mean) with autograd providing the backward, so its
timing is representative of the reference's CPU path; `tests/test_oracle_golden.py` checks it
against the golden vectors minted from the reference.  Never imported by the product path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def rewrite_edges(edge_index: torch.Tensor) -> torch.Tensor:
    """utils.py:47-72: drop every loop, keep order, append loops 0..max."""
    n_idx = int(edge_index.max()) + 1
    keep = edge_index[0] != edge_index[1]
    loops = torch.arange(n_idx, dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index[:, keep], torch.stack([loops, loops])], dim=1)


def _segment_sum(values: torch.Tensor, index: torch.Tensor, n: int) -> torch.Tensor:
    """utils.py:6-27 (scatter_add_ into zeros with an explicitly broadcast index)."""
    out = values.new_zeros((n,) + tuple(values.shape[1:]))
    idx = index.view((-1,) + (1,) * (values.dim() - 1)).expand_as(values)
    return out.scatter_add_(0, idx, values)


def layer_forward(x, edge_index, w, a, nh, f, concat=True, add_self_loops=True, const_attention=False):
    if add_self_loops:
        edge_index = rewrite_edges(edge_index)
    n, e = x.size(0), edge_index.size(1)
    src, dst = edge_index[0], edge_index[1]
    wh = F.linear(x, w).view(n, nh, f)                                   # :64-65
    wh_src = wh[src]                                                     # :70
    if const_attention:
        act = x.new_zeros((e, nh))                                       # :89-92
    else:
        pairs = torch.cat([wh_src, wh[dst]], dim=-1).view(e, nh * 2 * f)  # :71-81
        logits = F.linear(pairs, a)                                      # :82
        act = F.leaky_relu(logits - logits.max())                        # :85-87
    p = act.exp()                                                        # :96
    z = _segment_sum(p, dst, n)                                          # :99-103
    alpha = p / (torch.index_select(z, 0, dst) + 1e-8)                   # :106-109
    o = _segment_sum(alpha.view(e, nh, 1) * wh_src, dst, n)              # :119-127
    out = o.view(n, nh * f) if concat else o.mean(dim=1)                 # :129-132
    return out, edge_index, alpha


def model_step(x, edge_index, weights, shapes):
    """fwd+bwd of the stacked model with GATModel.forward's glue (layer -> ELU, GATModel.py:120-151);
    returns the loss value.  `weights` = [(W, a)] leaf tensors with requires_grad."""
    h = x
    for i, ((w, a), (_f_in, nh, f, concat)) in enumerate(zip(weights, shapes)):
        h, _, _ = layer_forward(h, edge_index, w, a, nh, f, concat)
        if i != len(shapes) - 1:
            h = F.elu(h)
    loss = h.square().mean()
    loss.backward()
    return float(loss.detach())
