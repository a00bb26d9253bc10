"""Synthetic datasets with the shapes of the ones the reference names (planetoid_gat.py:57-61, ppi_gat.py:61-64,
pattern_gat.py:72-75), from the generators of gat-pytorch_b200/synth.py (node / edge counts, feature widths,
split sizes).  Labels are PLANTED, not random: a fixed teacher -- two rounds of mean aggregation over the in-neighbourhood
(a constant-attention GAT, gat_layer.py:89-92) of a seeded random projection of the features -- so that a GAT can learn
them and loss / accuracy / F1 curves of two layer implementations can be compared (training-parity tests).  Class count and
prevalence follow each dataset; the absolute numbers say nothing about the real datasets."""
import importlib.util
import os

import numpy as np
import torch

from .data import Data

_SYNTH = None


def _synth():
    global _SYNTH
    if _SYNTH is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "synth.py")
        spec = importlib.util.spec_from_file_location("gat_b200_synth_for_shims", os.path.normpath(path))
        _SYNTH = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_SYNTH)
    return _SYNTH


def _teacher_scores(x, ei, n_out, seed):
    """Â²(x P): P a seeded (F_in, n_out) projection, Â = mean over {in-neighbours, self}; columns standardised."""
    rng = np.random.default_rng(seed)
    n = x.shape[0]
    h = x.astype(np.float64) @ rng.standard_normal((x.shape[1], n_out))
    src, dst = ei[0], ei[1]
    deg = np.bincount(dst, minlength=n).astype(np.float64) + 1.0
    for _ in range(2):
        agg = h.copy()
        np.add.at(agg, dst, h[src])
        h = agg / deg[:, None]
    return (h - h.mean(axis=0)) / (h.std(axis=0) + 1e-12)


def planted_classes(x, ei, n_classes, seed=1234):
    return np.argmax(_teacher_scores(x, ei, n_classes, seed), axis=1).astype(np.int64)


def planted_multilabel(x, ei, n_labels, positive_rate, seed=1234):
    z = _teacher_scores(x, ei, n_labels, seed)
    return z > np.quantile(z, 1.0 - positive_rate, axis=0, keepdims=True)


class _ListDataset:
    def __init__(self, graphs):
        self._graphs = graphs

    def __len__(self):
        return len(self._graphs)

    def __getitem__(self, i):
        return self._graphs[i]

    def __iter__(self):
        return iter(self._graphs)


class Planetoid(_ListDataset):
    """One graph with train / val / test masks (140-ish / 500 / 1000 nodes, the public Planetoid split sizes)."""
    _CLASSES = {"Cora": 7, "Citeseer": 6, "Pubmed": 3}

    def __init__(self, root=None, name="Cora", **kwargs):
        gen = {"Cora": _synth().cora, "Citeseer": _synth().citeseer, "Pubmed": _synth().pubmed}[name]
        x, ei = gen()
        n, c = x.shape[0], self._CLASSES[name]
        rng = np.random.default_rng(7)
        y = torch.from_numpy(planted_classes(x, ei, c))
        idx = torch.from_numpy(rng.permutation(n))
        masks = [torch.zeros(n, dtype=torch.bool) for _ in range(3)]
        masks[0][idx[:20 * c]] = True
        masks[1][idx[20 * c:20 * c + 500]] = True
        masks[2][idx[-1000:]] = True
        super().__init__([Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(ei), y=y, train_mask=masks[0], val_mask=masks[1],
                               test_mask=masks[2])])
        self.num_classes, self.num_node_features = c, x.shape[1]


class PPI(_ListDataset):
    """20 / 2 / 2 graphs (ppi_gat.py:61-64), ~2.4k nodes and ~34k edges each, 50 features, 121 binary labels."""

    def __init__(self, root=None, split="train", **kwargs):
        count, base = {"train": (20, 0), "val": (2, 100), "test": (2, 200)}[split]
        graphs = []
        for g in range(count):
            x, ei = _synth().ppi(seed=base + g, graphs=1)
            y = planted_multilabel(x, ei, 121, 0.3).astype(np.float32)
            graphs.append(Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(ei), y=torch.from_numpy(y)))
        super().__init__(graphs)


class GNNBenchmarkDataset(_ListDataset):
    """PATTERN-shaped SBM graphs (44-188 nodes), 3 one-hot features, binary node labels with 17.65 % positives
    (pattern_gat.py:11).  The stand-in keeps the split small (64 / 16 / 16 graphs; the real one has 10k / 2k / 2k)."""

    def __init__(self, root=None, name="PATTERN", split="train", **kwargs):
        if name != "PATTERN":
            raise NotImplementedError(name)
        count, base = {"train": (64, 0), "val": (16, 1000), "test": (16, 2000)}[split]
        graphs = []
        for g in range(count):
            x, ei = _synth().pattern(seed=base + g, graphs=1)
            y = planted_multilabel(x, ei, 1, 0.1765)[:, 0].astype(np.int64)
            graphs.append(Data(x=torch.from_numpy(x), edge_index=torch.from_numpy(ei), y=torch.from_numpy(y)))
        super().__init__(graphs)
