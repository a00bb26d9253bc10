"""Module path under which Lightning 1.2.x pickled `ModelCheckpoint` into the reference's committed checkpoints."""
from . import ModelCheckpoint  # noqa: F401
