"""Module path under which Lightning 1.2.x pickled `EarlyStopping` into the reference's committed checkpoints."""
from . import EarlyStopping  # noqa: F401
