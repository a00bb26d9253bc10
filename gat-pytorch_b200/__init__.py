"""gat-pytorch_b200: B200-native (sm_100a) replacement for the hot path of loodvn/gat-pytorch.

Only what the path needs lives here:
  csrc/            CUDA kernels + the C ABI (include/gat_b200.h) -> libgat_b200.so
  _lib.py          ctypes binding (no fallback)
  graph.py         host side of the CSR / transposed-CSR builder + per-graph cache
  gat_layer.py     drop-in `GATLayer` (reference: models/gat_layer.py)
  glue.py          caller-side glue on the cached structure: GATModel.calc_attention_norm, the visualisation
                   feed (per-node attention entropy, degree-scaled weights), GATModel.forward with the skip add /
                   ELU / input dropout between layers folded into the layers' kernels (model_forward)
  partition.py     destination-range partitioned layer for graphs spanning several GPUs
  synth.py         seeded synthetic graphs of the BASELINE shapes
  overlay/models/  namespace-package overlay so `from models.gat_layer import GATLayer` resolves here

The directory name contains a hyphen (it mirrors the reference's repository name), so it is imported
through the `gat_pytorch_b200` shim module at the repository root.
"""
from .gat_layer import GATLayer  # noqa: F401
from .graph import GLOBAL_CACHE, GraphStructure, StructureCache, build_structure  # noqa: F401
from .glue import attention_norm, degree_scaled_attention, micro_f1, model_forward, neighbourhood_attention, neighbourhood_entropy  # noqa: F401
from . import synth  # noqa: F401

__all__ = ["GATLayer", "GraphStructure", "StructureCache", "build_structure", "GLOBAL_CACHE", "attention_norm", "neighbourhood_entropy",
           "degree_scaled_attention", "neighbourhood_attention", "model_forward", "micro_f1", "synth"]
