"""Import-only stand-in for matplotlib (absent from the offline image): the reference's visualisation modules draw with
`matplotlib.pyplot`; every call is accepted and recorded, nothing is rendered.  See ../pytorch_lightning/__init__.py."""
__version__ = "0.0-gat-b200-shim"
