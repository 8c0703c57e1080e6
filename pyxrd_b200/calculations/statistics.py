"""Device versions of ``pyxrd/calculations/statistics.py`` (the pattern statistics the GUI shows,
specimen/models/statistics.py:16,115-118, and the residual functions ``settings.RESIDUAL_METHOD`` names,
mixture.py:22-27).  Inside ``calculate_mixture`` the residual is formed by the batch kernels; these are the same
quantities for plain host arrays."""
import numpy as np

from ..runtime import default_context


def _call(kind, exp, calc=None, phase=None, extra=0.0):
    ctx = default_context()
    with ctx.lock:
        return ctx.leaf_statistic(kind, exp, calc, phase, extra)


def R_squared(exp, calc, *args):
    """statistics.py:12-20."""
    return _call(2, exp, calc)


def Rp(exp, calc, *args):
    """statistics.py:22-27."""
    return _call(0, exp, calc)


def smooth_pattern(pattern):
    """statistics.py:29-33: ``smooth(pattern, 15)`` (math_tools.py:46-99, Blackman window of 31 points)."""
    ctx = default_context()
    with ctx.lock:
        return ctx.leaf_derive(np.asarray(pattern, dtype=float), smooth_only=True)


def derive(pattern):
    """statistics.py:35-41: ``np.gradient(smooth_pattern(pattern))``."""
    ctx = default_context()
    with ctx.lock:
        return ctx.leaf_derive(np.asarray(pattern, dtype=float))


def Rpder(exp, calc):
    """statistics.py:43-48."""
    return Rp(derive(exp), derive(calc))


def Rpw(exp, calc):
    """statistics.py:50-68 (1-D patterns; for the 2-D clipped patterns of mixture.py:88-89 the reference raises, and so
    does ``Context.set_residual_method``)."""
    exp = np.asarray(exp)
    if exp.ndim != 1:
        raise ValueError("The truth value of an array with more than one element is ambiguous. Use a.any() or a.all()")
    return _call(1, exp, calc)


def Rpe(exp, calc, num_params):
    """statistics.py:70-79."""
    return _call(3, exp, None, None, float(num_params))


def Rphase(exp, calc, phase):
    """statistics.py:81-100."""
    return _call(4, exp, calc, phase)
