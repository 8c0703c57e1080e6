"""The flags the hot path reads at call time (reference pyxrd/data/settings.py:15,17,26,32;
read in calculations/mixture.py:89,161,246 and phases/models/CSDS.py:127,140).  When this package runs inside a PyXRD process
the reference's own ``pyxrd.data.settings`` module is consulted, so a user changing a flag
there changes it here too."""
import sys

RESIDUAL_METHOD = "Rp"
BGSHIFT = True
DEBUG = False
LOG_NORMAL_MAX_CSDS_FACTOR = 2.5
# not a reference flag: how optimize_mixture minimises.  "device" = the device-resident optimiser (pyxrd_optimize; end
# point <= the reference's L-BFGS-B end point * (1 + 5e-3)) whenever the mixture is in its scope,
# "scipy" = the reference's own fmin_l_bfgs_b call with the device evaluating the objective (same path as the
# reference, ~700 device round trips per trial)
MIXTURE_OPTIMIZER = "device"


def get(name):
    ref = sys.modules.get("pyxrd.data.settings")
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    return globals()[name]
