"""Device-backed mirror of ``pyxrd.calculations`` (same module and function names,
same arguments, same mutation / error conventions).

Every function here packs its DataObject arguments, runs the CUDA kernels through
the C-ABI (``pyxrd_b200.runtime``) and hands back host numpy arrays, exactly where
the reference would have put them.  Nothing in this package computes on the CPU
and nothing imports ``oracle/``.
"""
from . import atoms, components, CSDS, goniometer, math_tools, phases, specimen, mixture, statistics, exceptions  # noqa: F401
from .. import data_objects  # noqa: F401
