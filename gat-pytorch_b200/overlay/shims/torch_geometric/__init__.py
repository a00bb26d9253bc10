"""Offline stand-in for the slice of `torch_geometric` (pinned 1.6.3 by the reference, env/gat_req_mac_version.yml:180) that
loodvn/gat-pytorch uses -- SURVEY.md section 8-f2.  NOT the product path and not a graph library: `data.Data` /
`data.DataLoader` (block-diagonal batching), the three dataset classes the reference names -- generating SYNTHETIC graphs of
each dataset's shape (there is no network and no dataset in the image; SURVEY.md 8-d gives the shapes) -- and an import-only
`nn.GATConv`.  Put this directory on PYTHONPATH only when the real package is absent.
"""
from . import data, datasets, nn  # noqa: F401

__version__ = "0.0-gat-b200-shim"
