"""blues_b200 — B200-native NCMC engine behind the BLUES Python API.

Importing the package never touches the GPU; creating a ``Context`` / ``Simulation`` needs a CUDA device and
the in-tree ``libblues_b200.so`` (no CPU fallback).
"""
__version__ = '0.1.0'

from . import unit  # noqa: F401
from . import system as app  # noqa: F401  (enum names: app.PME, app.HBonds, ...)
