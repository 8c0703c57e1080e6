"""Loader of ``oracle/_ref`` (made by ``oracle/make_ref.py``): the unmodified ``pyxrd.calculations`` of the reference,
for bench.py's CPU arm (``--impl reference``, ``kind: "reference"``).  Test / baseline infrastructure only.

The reference targets older numpy / scipy; four removed spellings are restored on the LIBRARY modules before the
import -- ``np.complex_``, ``np.complex``, ``np.float_`` and the ``iprint`` keyword of
``scipy.optimize.fmin_l_bfgs_b`` -- no reference file is edited.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "pyxrd", "calculations", "mixture.py"))


def load():
    """-> the reference's ``pyxrd.calculations.mixture`` module (its functions take the same DataObject graphs the
    product takes), or None when oracle/_ref is absent"""
    if not available():
        return None
    import numpy as np
    for name, value in (("complex_", np.complex128), ("complex", complex), ("float_", np.float64)):
        if not hasattr(np, name):
            setattr(np, name, value)
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pyxrd.calculations.mixture as mixture
    if not os.path.abspath(mixture.__file__).startswith(REF_DIR):
        return None                 # another pyxrd on the path: not the vendored, unmodified copy
    if not getattr(mixture, "_iprint_shim", False):
        import scipy.optimize
        real = scipy.optimize.fmin_l_bfgs_b

        def fmin_l_bfgs_b(*args, **kwargs):
            kwargs.pop("iprint", None)
            return real(*args, **kwargs)

        mixture.fmin_l_bfgs_b = fmin_l_bfgs_b
        mixture._iprint_shim = True
    return mixture
