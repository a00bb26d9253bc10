"""Seeded synthetic graphs of the five BASELINE shapes.

The reference ships no datasets (they are downloaded by torch_geometric at run time,
`models/planetoid_gat.py:56-59`, `models/ppi_gat.py:61-64`, `models/pattern_gat.py`), and the
build box has no network, so every parity test and every bench line runs on graphs generated
here.  Only numpy's `default_rng` (PCG64, stream-stable across numpy versions) is used, so the
same seed yields the same graph in this container, on the GPU box and in the golden-vector
generator.

Every generator returns `(x, edge_index)` as numpy arrays: `x` float32 `(N, F_in)`,
`edge_index` int64 `(2, E)` with edges pointing row 0 -> row 1 (`models/gat_layer.py:47-48`).
Shapes follow `run_config.py:17-98` and `BASELINE.json`.
"""
from __future__ import annotations

import numpy as np

# (in_features, heads, per-head width, concat) per layer, from run_config.py:17-98.
LAYER_SHAPES = {
    "cora": [(1433, 8, 8, True), (64, 1, 7, False)],
    "citeseer": [(3703, 8, 8, True), (64, 1, 6, False)],
    "pubmed": [(500, 8, 8, True), (64, 8, 3, False)],
    "ppi": [(50, 4, 256, True), (1024, 4, 256, True), (1024, 6, 121, False)],
    "pattern": [(3, 4, 12, True), (48, 4, 24, True), (96, 4, 12, True), (48, 1, 1, False)],
    # BASELINE.json gives "100-dim in, 3-layer 4-head GAT"; the hidden width is not stated.
    # SURVEY.md section 8 assumes H=64 per head (D=256); 47 classes as in ogbn-products.
    "products": [(100, 4, 64, True), (256, 4, 64, True), (256, 4, 47, False)],
}

# run_config.py: add_skip_connection per dataset (GATModel.py:97-112 builds Identity/Linear).
SKIP = {
    "cora": [False, False], "citeseer": [False, False], "pubmed": [False, False],
    "ppi": [False, True, False], "pattern": [True, True, True, True],
    "products": [False, False, False],
}
DROPOUT = {"cora": 0.6, "citeseer": 0.6, "pubmed": 0.6, "ppi": 0.0, "pattern": 0.0, "products": 0.0}


def _powerlaw_ranks(rng, n, size, exponent, offset):
    """Ranks in [0, n) with density ~ (rank+offset)^-exponent: analytic inverse CDF of the continuous
    power law on [offset, n+offset), floored (no O(n) table, so 3e7 draws take well under a second)."""
    a, b, q = float(offset), float(n + offset), 1.0 - exponent
    t = (rng.random(size) * (b ** q - a ** q) + a ** q) ** (1.0 / q)
    return np.clip((t - a).astype(np.int64), 0, n - 1)


def _symmetric_powerlaw_pairs(rng, n, n_pairs, exponent, offset):
    """Chung-Lu style undirected pairs: both endpoints drawn with probability ~ (rank+offset)^-exponent."""
    perm = rng.permutation(n)  # decouple degree from node id
    u = perm[_powerlaw_ranks(rng, n, n_pairs, exponent, offset)]
    v = perm[_powerlaw_ranks(rng, n, n_pairs, exponent, offset)]
    clash = u == v
    v[clash] = (v[clash] + 1 + rng.integers(0, n - 1, size=int(clash.sum()))) % n
    return u.astype(np.int64), v.astype(np.int64)


def _directed_both_ways(u, v, n, sort=True):
    src = np.concatenate([u, v])
    dst = np.concatenate([v, u])
    if sort:  # PyG datasets are coalesced: ordered by (src, dst); duplicates are kept
        key = src * np.int64(n) + dst
        key.sort()
        src, dst = key // np.int64(n), key % np.int64(n)
    return np.stack([src, dst]).astype(np.int64)


def cora(seed=0, n=2708, e=10556, f_in=1433, density=0.0127):
    """Cora-shaped: 5278 undirected pairs, bag-of-words {0,1} features."""
    rng = np.random.default_rng(seed)
    u, v = _symmetric_powerlaw_pairs(rng, n, e // 2, exponent=0.55, offset=3.0)
    x = (rng.random((n, f_in)) < density).astype(np.float32)
    return x, _directed_both_ways(u, v, n)


def citeseer(seed=0):
    return cora(seed=seed, n=3327, e=9104, f_in=3703, density=0.0086)


def pubmed(seed=0, n=19717, e=88648, f_in=500):
    """Pubmed-shaped: TF-IDF-like 10 % dense features in (0, 0.2)."""
    rng = np.random.default_rng(seed)
    u, v = _symmetric_powerlaw_pairs(rng, n, e // 2, exponent=0.6, offset=5.0)
    mask = rng.random((n, f_in)) < 0.10
    x = (rng.random((n, f_in)) * 0.2 * mask).astype(np.float32)
    return x, _directed_both_ways(u, v, n)


def ppi(seed=0, graphs=2, nodes_per_graph=2400, edges_per_graph=34000, f_in=50):
    """PPI-shaped batch: block-diagonal union of `graphs` graphs (PyG Batch semantics)."""
    rng = np.random.default_rng(seed)
    xs, eis = [], []
    for g in range(graphs):
        u, v = _symmetric_powerlaw_pairs(rng, nodes_per_graph, edges_per_graph // 2, 0.5, 8.0)
        ei = _directed_both_ways(u, v, nodes_per_graph) + g * nodes_per_graph
        eis.append(ei)
        xs.append(rng.standard_normal((nodes_per_graph, f_in)).astype(np.float32))
    return np.concatenate(xs), np.concatenate(eis, axis=1)


def pattern(seed=0, graphs=128, p_in=0.5, p_out=0.35, lo=44, hi=188):
    """PATTERN-shaped batch: SBM graphs with 5 communities, one-hot 3-valued features."""
    rng = np.random.default_rng(seed)
    xs, eis, base = [], [], 0
    for _ in range(graphs):
        n = int(rng.integers(lo, hi + 1))
        comm = rng.integers(0, 5, size=n)
        prob = np.where(comm[:, None] == comm[None, :], p_in, p_out)
        upper = np.triu(rng.random((n, n)) < prob, k=1)
        u, v = np.nonzero(upper)
        eis.append(_directed_both_ways(u.astype(np.int64), v.astype(np.int64), n) + base)
        xs.append(np.eye(3, dtype=np.float32)[rng.integers(0, 3, size=n)])
        base += n
    return np.concatenate(xs), np.concatenate(eis, axis=1)


def products(seed=0, scale=1.0, n=2449029, e=61859140, f_in=100, max_degree=17000.0, sort=True):
    """ogbn-products-shaped: symmetric, truncated power-law degrees (mean ~25, max ~17k at scale 1).

    `scale` < 1 shrinks nodes and edges together (mean degree kept); it is what the CPU
    baseline uses, because the reference cannot allocate its (E', NH, F) intermediates at full
    size.
    """
    rng = np.random.default_rng(seed)
    n = max(int(round(n * scale)), 64)
    n_pairs = max(int(round(e * scale)) // 2, 64)
    exponent = 0.62
    target = min(max_degree * max(scale, 1e-3) ** 0.5, n / 4)
    lo_off, hi_off = 0.5, float(n)  # bisection on the rank offset so the hub hits `target`
    for _ in range(60):
        off = 0.5 * (lo_off + hi_off)
        w = (np.arange(n, dtype=np.float64) + off) ** (-exponent)
        hub = 2.0 * n_pairs * w[0] / w.sum()
        lo_off, hi_off = (off, hi_off) if hub > target else (lo_off, off)
    u, v = _symmetric_powerlaw_pairs(rng, n, n_pairs, exponent, 0.5 * (lo_off + hi_off))
    x = rng.standard_normal((n, f_in), dtype=np.float32)
    return x, _directed_both_ways(u, v, n, sort=sort)


def adversarial(seed=0, n=97, e=900, f_in=11):
    """Small graph with every edge case the reference exhibits:
    existing self-loops (dropped by the rewrite), duplicate edges, a hub of in-degree >> 32,
    trailing isolated nodes that never appear in edge_index (no self-loop: utils.py:59,72),
    nodes with out-edges only."""
    rng = np.random.default_rng(seed)
    n_idx = n - 5  # last 5 nodes never appear
    src = rng.integers(0, n_idx, size=e)
    dst = rng.integers(0, n_idx, size=e)
    dst[: e // 4] = 7  # hub
    src[e // 4: e // 4 + 20] = dst[e // 4: e // 4 + 20]  # explicit self-loops
    src[-30:] = src[-60:-30]  # duplicates
    dst[-30:] = dst[-60:-30]
    order = rng.permutation(e)
    ei = np.stack([src[order], dst[order]]).astype(np.int64)
    ei[0, 0], ei[1, 0] = n_idx - 1, 3  # make sure max id is present
    x = rng.standard_normal((n, f_in)).astype(np.float32)
    return x, ei


GENERATORS = {"cora": cora, "citeseer": citeseer, "pubmed": pubmed, "ppi": ppi,
              "pattern": pattern, "products": products, "adversarial": adversarial}


def xavier_uniform(rng, out_f, in_f):
    """Same distribution as nn.init.xavier_uniform_ (gat_layer.py:142-145); numpy stream."""
    bound = float(np.sqrt(6.0 / (in_f + out_f)))
    return rng.uniform(-bound, bound, size=(out_f, in_f)).astype(np.float32)


def seeded_weights(name, seed=42):
    """Per-layer (W, a) for configs whose checkpoints are absent (PPI, products)."""
    rng = np.random.default_rng(seed)
    out = []
    for (f_in, nh, f, _c) in LAYER_SHAPES[name]:
        out.append((xavier_uniform(rng, nh * f, f_in), xavier_uniform(rng, nh, 2 * nh * f)))
    return out
