"""Namespace-package overlay for the reference checkout.

The reference's `models/` directory has no `__init__.py`, so `models` is a namespace package: with
`sys.path = [<this repo>/gat-pytorch_b200/overlay, <this repo>, <reference checkout>, ...]`
`models.gat_layer` resolves to this file while `models.utils`, `models.GATModel`, ... still resolve to
the reference, and `planetoid_gat.py`, `ppi_gat.py`, `pattern_gat.py`, `train.py`, `vis.py` run unchanged
(`models/GATModel.py:13` does `from .gat_layer import GATLayer`).  See INTEGRATION.md.
"""
from gat_pytorch_b200.gat_layer import GATLayer  # noqa: F401

__all__ = ["GATLayer"]
