"""Test infrastructure.  Empty stand-in: the reference's tensor/layers.py imports matplotlib.pyplot for a plotting helper that the
sweep never calls; the image has no matplotlib."""
from . import pyplot  # noqa: F401
