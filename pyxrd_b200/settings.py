"""The flags the hot path reads at call time (reference pyxrd/data/settings.py:15,17,26,32;
read in calculations/mixture.py:89,161,246 and phases/models/CSDS.py:127,140).  When this package runs inside a PyXRD process
the reference's own ``pyxrd.data.settings`` module is consulted, so a user changing a flag
there changes it here too."""
import sys

RESIDUAL_METHOD = "Rp"
BGSHIFT = True
DEBUG = False
LOG_NORMAL_MAX_CSDS_FACTOR = 2.5


def get(name):
    ref = sys.modules.get("pyxrd.data.settings")
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    return globals()[name]
