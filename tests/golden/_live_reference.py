"""Load the UNMODIFIED reference ``pyxrd.calculations`` from /root/reference.

Only usable in the build container (the GPU box has no /root/reference): used by
``make_golden.py`` to produce the committed fixtures and by nothing else.

The reference targets older numpy/scipy.  Four removed spellings are restored on
the *library modules* before import; no reference file is edited:
``np.complex_``, ``np.complex``, ``np.float_`` and the ``iprint`` kwarg of
``scipy.optimize.fmin_l_bfgs_b``.
"""
import sys
import warnings

import numpy as np

REFERENCE_ROOT = "/root/reference"


def load():
    if not hasattr(np, "complex_"):
        np.complex_ = np.complex128
    if not hasattr(np, "complex"):
        np.complex = complex
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pyxrd.calculations.atoms as atoms
        import pyxrd.calculations.components as components
        import pyxrd.calculations.CSDS as CSDS
        import pyxrd.calculations.goniometer as goniometer
        import pyxrd.calculations.phases as phases
        import pyxrd.calculations.specimen as specimen
        import pyxrd.calculations.statistics as statistics
        import pyxrd.calculations.mixture as mixture
    import scipy.optimize

    if not getattr(mixture, "_iprint_shim", False):
        real = scipy.optimize.fmin_l_bfgs_b

        def fmin_l_bfgs_b(*args, **kwargs):
            kwargs.pop("iprint", None)
            return real(*args, **kwargs)

        mixture.fmin_l_bfgs_b = fmin_l_bfgs_b
        mixture._iprint_shim = True

    class Ref(object):
        pass

    ref = Ref()
    for name, mod in dict(atoms=atoms, components=components, CSDS=CSDS, goniometer=goniometer,
                          phases=phases, specimen=specimen, statistics=statistics,
                          mixture=mixture).items():
        setattr(ref, name, mod)
    return ref
