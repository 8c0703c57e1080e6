"""Device versions of the array helpers of ``pyxrd/calculations/math_tools.py`` the hot path uses."""
from ..runtime import default_context


def mmult(A, B):
    """math_tools.py:16-17: batched complex matrix product of [X, r, r] stacks (``k_leaf_mmult``: every product and sum
    rounded as numpy rounds the reference's broadcast-and-sum)"""
    ctx = default_context()
    with ctx.lock:
        return ctx.leaf_mmult(A, B)
