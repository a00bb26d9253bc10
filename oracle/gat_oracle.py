"""CPU oracle for the GAT hot path -- TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the shipped layer (`gat-pytorch_b200/gat_layer.py`) never does
and fails loudly when its CUDA library is missing.

It restates, in numpy, what `/root/reference/models/gat_layer.py:42-140` and
`/root/reference/models/utils.py:6-72` compute, in the reference's own formulation (gather
both endpoints, concatenate, multiply by the full cross-head `a`, ONE global max, LeakyReLU
slope 0.01, exp, scatter-add, +1e-8, divide, scatter-add), plus the backward pass that the
reference leaves to autograd, written out explicitly.

Pinning: the reference has no tests or golden vectors for this path, so
parity is pinned on outputs of the reference itself: `tests/golden/make_golden.py` imports
`/root/reference/models/gat_layer.py` in the build container, runs it in fp32 and fp64 with
autograd and commits the results under `tests/golden/`; `tests/test_oracle_golden.py` checks
this file against them (fp64 run: <= 1e-9; fp32 run: within the reference's own fp32 noise).

Arithmetic is done in `dtype` (float64 = ground truth used by the GPU parity tests).
"""
from __future__ import annotations

import numpy as np

LEAKY_SLOPE = 0.01   # nn.LeakyReLU() default, gat_layer.py:87
SOFTMAX_EPS = 1e-8   # gat_layer.py:109


# --------------------------------------------------------------------------------------
# integer path  (utils.py:47-72)
# --------------------------------------------------------------------------------------
def add_remaining_self_loops(edge_index: np.ndarray) -> np.ndarray:
    """utils.py:47-67 with num_nodes=None: N_idx = max+1 (utils.py:72); drop every loop keeping
    input order (utils.py:61,65); append (k,k) for k in 0..N_idx-1 (utils.py:63-65)."""
    edge_index = np.asarray(edge_index)
    n_idx = int(edge_index.max()) + 1
    keep = edge_index[0] != edge_index[1]
    loops = np.arange(n_idx, dtype=edge_index.dtype)
    return np.concatenate([edge_index[:, keep], np.stack([loops, loops])], axis=1)


def in_degrees(edge_index: np.ndarray, n: int) -> np.ndarray:
    """Degree counts the reference derives by scatter-adding ones over targets
    (GATModel.py:196-201)."""
    return np.bincount(edge_index[1], minlength=n).astype(np.int64)


def csr_by_target(edge_index: np.ndarray, n: int):
    """Destination-sorted CSR with the reference's edge order inside each row: stable argsort
    of the target row.  Returns rowptr (n+1), col (source ids), eid
    (position of each CSR slot in the rewritten edge list)."""
    dst = edge_index[1]
    eid = np.argsort(dst, kind="stable")
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(dst, minlength=n), out=rowptr[1:])
    return rowptr, edge_index[0][eid], eid


def csr_by_source(edge_index: np.ndarray, n: int, eid: np.ndarray):
    """Transposed CSR: stable argsort of the source row; colT = target ids; posT = CSR slot
    (in the target-sorted CSR) of each transposed slot."""
    src = edge_index[0]
    perm_t = np.argsort(src, kind="stable")
    rowptr_t = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(src, minlength=n), out=rowptr_t[1:])
    slot_of_edge = np.empty_like(eid)
    slot_of_edge[eid] = np.arange(eid.shape[0], dtype=eid.dtype)
    return rowptr_t, edge_index[1][perm_t], slot_of_edge[perm_t]


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _scatter_rows(values: np.ndarray, index: np.ndarray, n: int) -> np.ndarray:
    """sum_over_neighbourhood (utils.py:6-27): out[index[e]] += values[e]."""
    flat = values.reshape(values.shape[0], -1)
    order = np.argsort(index, kind="stable")
    sorted_idx = index[order]
    out = np.zeros((n, flat.shape[1]), dtype=values.dtype)
    if flat.shape[0]:
        starts = np.flatnonzero(np.r_[True, sorted_idx[1:] != sorted_idx[:-1]])
        out[sorted_idx[starts]] = np.add.reduceat(flat[order], starts, axis=0)
    return out.reshape((n,) + values.shape[1:])


def split_attention(a: np.ndarray, nh: int, f: int):
    """`a.weight` (NH, NH*2F) acts on cat([Wh_src, Wh_dst], -1).view(E, NH*2F)
    (gat_layer.py:76-82): column h'*2F+j multiplies Wh_src[h', j], column h'*2F+F+j multiplies
    Wh_dst[h', j].  Returns A_src, A_tgt of shape (NH, NH*F)."""
    a3 = a.reshape(nh, nh, 2 * f)
    return a3[:, :, :f].reshape(nh, nh * f), a3[:, :, f:].reshape(nh, nh * f)


# --------------------------------------------------------------------------------------
# forward  (gat_layer.py:42-140)
# --------------------------------------------------------------------------------------
def forward(x, edge_index, W, a, nh, f, concat=True, add_self_loops=False, bias=None,
            const_attention=False, drop_mask=None, dtype=np.float64):
    """Returns dict(out, edge_index, alpha, ...intermediates).  `drop_mask` (E', NH), already
    scaled by 1/(1-p), stands in for nn.Dropout (gat_layer.py:113-115); None = eval / p=0."""
    x = np.asarray(x, dtype=dtype)
    W = np.asarray(W, dtype=dtype)
    if add_self_loops:                                              # :53-54
        edge_index = add_remaining_self_loops(edge_index)
    n, e = x.shape[0], edge_index.shape[1]
    src, dst = edge_index[0], edge_index[1]
    wh = (x @ W.T).reshape(n, nh, f)                                # :64-65
    wh_src = wh[src]                                                # :70
    if not const_attention:
        a = np.asarray(a, dtype=dtype)
        pairs = np.concatenate([wh_src, wh[dst]], axis=-1).reshape(e, nh * 2 * f)   # :76,:81
        logits = pairs @ a.T                                        # :82
        gmax = logits.max() if e else dtype(0)                      # :85
        shifted = logits - gmax
        act = np.where(shifted >= 0, shifted, shifted * dtype(LEAKY_SLOPE))         # :87
    else:
        logits = np.zeros((e, nh), dtype=dtype)                     # :89-92
        gmax = dtype(0)
        act = logits
    p = np.exp(act)                                                 # :96
    z = _scatter_rows(p, dst, n)                                    # :99-103
    alpha = p / (z[dst] + dtype(SOFTMAX_EPS))                       # :106-109
    alpha_drop = alpha if drop_mask is None else alpha * np.asarray(drop_mask, dtype=dtype)
    weighted = alpha_drop[:, :, None] * wh_src                      # :119
    o = _scatter_rows(weighted, dst, n)                             # :123-127
    out = o.reshape(n, nh * f) if concat else o.mean(axis=1)        # :129-132
    if bias is not None:
        out = out + np.asarray(bias, dtype=dtype)                   # :134-135
    return dict(out=out, edge_index=edge_index, alpha=alpha, wh=wh, logits=logits, gmax=gmax,
                z=z, alpha_drop=alpha_drop, x=x, W=W, a=a, nh=nh, f=f, concat=concat,
                const_attention=const_attention, drop_mask=drop_mask, dtype=dtype)


# --------------------------------------------------------------------------------------
# backward  (autograd of the above)
# --------------------------------------------------------------------------------------
def backward(fw: dict, grad_out, grad_alpha=None):
    """Gradients of sum(out*grad_out) + sum(alpha*grad_alpha) w.r.t. x, W, a, bias.
    Includes the gradient through the un-detached global max (gat_layer.py:85), split evenly
    over the arg-max set as torch.max() backward does."""
    dtype = fw["dtype"]
    x, W, wh, alpha = fw["x"], fw["W"], fw["wh"], fw["alpha"]
    nh, f, concat = fw["nh"], fw["f"], fw["concat"]
    src, dst = fw["edge_index"]
    n, e = x.shape[0], src.shape[0]
    g = np.asarray(grad_out, dtype=dtype)
    go = g.reshape(n, nh, f) if concat else np.broadcast_to(g[:, None, :] / dtype(nh), (n, nh, f))
    mask = np.ones((e, nh), dtype=dtype) if fw["drop_mask"] is None else np.asarray(fw["drop_mask"], dtype=dtype)
    # value path: d weighted -> d Wh[src]
    d_wh = _scatter_rows((mask * alpha)[:, :, None] * go[dst], src, n)
    # attention path
    d_alpha = mask * np.einsum("ehf,ehf->eh", go[dst], wh[src])
    if grad_alpha is not None:
        d_alpha = d_alpha + np.asarray(grad_alpha, dtype=dtype)
    grads = dict(bias=g.sum(axis=0) if concat else None)
    if fw["const_attention"]:
        d_wh_flat = d_wh.reshape(n, nh * f)
        grads.update(a=None)
    else:
        a = fw["a"]
        s = _scatter_rows(alpha * d_alpha, dst, n)
        shifted = fw["logits"] - fw["gmax"]
        slope = np.where(shifted > 0, dtype(1), dtype(LEAKY_SLOPE))  # torch: grad at 0 is the slope
        d_shift = alpha * (d_alpha - s[dst]) * slope
        gamma = d_shift.sum()
        ties = fw["logits"] == fw["gmax"]
        d_logit = d_shift - ties * (gamma / max(int(ties.sum()), 1))
        a_src, a_tgt = split_attention(a, nh, f)
        ds_src = _scatter_rows(d_logit, src, n)
        ds_tgt = _scatter_rows(d_logit, dst, n)
        wh_flat = wh.reshape(n, nh * f)
        d_wh_flat = d_wh.reshape(n, nh * f) + ds_src @ a_src + ds_tgt @ a_tgt
        da_src = (ds_src.T @ wh_flat).reshape(nh, nh, f)
        da_tgt = (ds_tgt.T @ wh_flat).reshape(nh, nh, f)
        grads.update(a=np.concatenate([da_src, da_tgt], axis=-1).reshape(nh, 2 * nh * f),
                     ds_src=ds_src, ds_tgt=ds_tgt, gamma=gamma, n_ties=int(ties.sum()))
    grads.update(W=d_wh_flat.T @ x, x=d_wh_flat @ W, wh=d_wh_flat)
    return grads


def attention_norm(edge_index, alphas, n):
    """calc_attention_norm (GATModel.py:189-234): mean over layers of ||alpha*deg - 1||_1 / E'."""
    deg = in_degrees(edge_index, n)[edge_index[1]].astype(np.float64)
    total = 0.0
    for al in alphas:
        total += np.abs(np.asarray(al, np.float64) * deg[:, None] - 1.0).sum() / edge_index.shape[1]
    return total / len(alphas)


def neighbourhood_entropy(edge_index, alpha, n):
    """visualisation/entropy_histograms.py:103-115, restated loop for loop: for every node, mask the edge list with
    `target_nodes == node_id` and take scipy.stats.entropy(weights, base=2) of the node's incoming attention (one head at a
    time, :95-97), next to the entropy of the uniform distribution over the same neighbourhood (:115).
    Returns (entropy (n, NH), uniform (n,)).  Nodes without incoming edges (never visited with self-loops) give 0."""
    from scipy.stats import entropy
    alpha = np.asarray(alpha, np.float64)
    target_nodes = np.asarray(edge_index[1])
    ent = np.zeros((n, alpha.shape[1]))
    uni = np.zeros(n)
    for node_id in range(n):
        sel = target_nodes == node_id
        k = int(sel.sum())
        if k == 0:
            continue
        for head in range(alpha.shape[1]):
            ent[node_id, head] = entropy(alpha[sel, head], base=2)
        uni[node_id] = entropy(np.ones(k) / k, base=2)
    return ent, uni


def degree_scaled_attention(edge_index, alpha, n):
    """visualisation/weight_histograms.py:74-87: per node, the incoming attention weights times the neighbourhood size,
    concatenated node by node (edge-list order inside a node); the reference's `weight < 5` filter is the consumer's."""
    alpha = np.asarray(alpha, np.float64)
    target_nodes = np.asarray(edge_index[1])
    out = []
    for node_id in range(n):
        sel = target_nodes == node_id
        out.append(alpha[sel] * int(sel.sum()))
    return np.concatenate(out, axis=0) if out else np.zeros((0, alpha.shape[1]))


def rel_err(got, want) -> float:
    """Tensor-relative error used by every parity test: max|got-want| / max(|want|, tiny)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    if want.size == 0:
        return 0.0
    return float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))


def elementwise_report(got, want, rtol: float = 1e-5) -> dict:
    """Element-wise companions of `rel_err` (which divides by the LARGEST reference entry, so a few large entries could hide
    errors in small-magnitude rows):
      atol_needed_over_mean  smallest atol with |got-want| <= rtol*|want| + atol for EVERY element, as a multiple of mean|want|
      row_rel_{median,p99,max}  per row r (node / edge / weight row): max_j|got-want| / max_j|want|, rows whose reference is
                             exactly zero excluded -- a row is judged against its own magnitude
      frac_within_rtol       fraction of elements with |got-want| <= rtol*|want| + 1e-7*mean|want|"""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    if want.size == 0:
        return dict(tensor_rel=0.0, atol_needed_over_mean=0.0, row_rel_median=0.0, row_rel_p99=0.0, row_rel_max=0.0, frac_within_rtol=1.0)
    d = np.abs(got - want)
    aw = np.abs(want)
    mean = max(float(aw.mean()), 1e-300)
    d2, w2 = d.reshape(d.shape[0], -1), aw.reshape(aw.shape[0], -1)
    row_ref = w2.max(axis=1)
    ok = row_ref > 0
    row_rel = d2.max(axis=1)[ok] / row_ref[ok] if ok.any() else np.zeros(1)
    return dict(tensor_rel=float(d.max() / max(aw.max(), 1e-30)),
                atol_needed_over_mean=float(np.maximum(d - rtol * aw, 0.0).max() / mean),
                row_rel_median=float(np.median(row_rel)), row_rel_p99=float(np.percentile(row_rel, 99)),
                row_rel_max=float(row_rel.max()),
                frac_within_rtol=float((d <= rtol * aw + 1e-7 * mean).mean()))


def neighbourhood_attention(edge_index, alpha, node_ids, head):
    """visualisation/neighbourhood_attention_weights.py:41-60, statement for statement: per requested node the mask over the
    whole edge list, the neighbours' ids, the head's weights over them divided by their maximum, times 60 / size."""
    source_nodes, target_nodes = edge_index[0], edge_index[1]
    out = []
    for node_id in node_ids:
        neighbour_node_indices = target_nodes == node_id                                   # :46
        neighbour_nodes_ids = source_nodes[neighbour_node_indices]                         # :49
        size_of_neighborhood = len(neighbour_nodes_ids)                                    # :50
        w = np.array(alpha[neighbour_node_indices, head], dtype=np.float32)                # :53
        w /= np.max(w)                                                                     # :56
        w *= (60 / size_of_neighborhood)                                                   # :58
        out.append((neighbour_nodes_ids.astype(np.int64), w))
    return out
