"""Device version of ``pyxrd/calculations/CSDS.py``."""
from ..runtime import default_context


def calculate_distribution(CSDS):
    """CSDS.py:13-46 -> (q[0..maximum] normalised, real mean)."""
    ctx = default_context()
    with ctx.lock:
        arr, mean, _ = ctx.leaf_csds(CSDS.average, CSDS.alpha_scale, CSDS.alpha_offset, CSDS.beta_scale,
                                     CSDS.beta_offset, int(CSDS.minimum), int(CSDS.maximum))
    return arr, mean


def progression_factors(CSDS):
    """coefficient of Q^n in the series, n = 0..maximum+1 (phases.py:147-151, index = power)."""
    ctx = default_context()
    with ctx.lock:
        _, _, coef = ctx.leaf_csds(CSDS.average, CSDS.alpha_scale, CSDS.alpha_offset, CSDS.beta_scale,
                                   CSDS.beta_offset, int(CSDS.minimum), int(CSDS.maximum))
    return coef
