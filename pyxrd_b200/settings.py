"""The three flags the hot path reads at call time (reference pyxrd/data/settings.py:15,17,32;
read in calculations/mixture.py:89,161,246).  When this package runs inside a PyXRD process
the reference's own ``pyxrd.data.settings`` module is consulted, so a user changing a flag
there changes it here too."""
import sys

RESIDUAL_METHOD = "Rp"
BGSHIFT = True
DEBUG = False


def get(name):
    ref = sys.modules.get("pyxrd.data.settings")
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    return globals()[name]
