"""Load the UNMODIFIED reference MODEL layer (pyxrd.probabilities.models, pyxrd.phases.models ...)
headless, for golden vectors of the rows SURVEY.md §8(f) marks "next".  Build container only.

On top of ``_live_reference.load()``: ``inspect.getargspec`` / ``scipy.integrate.trapz`` aliases, a stub
``imp`` module and MagicMock stand-ins for the GUI / plotting / Pyro packages; the
reference's ``settings._parse_args`` is replaced as its own test utilities do (test/utils.py:39-42)."""
import inspect
import sys
import types
from unittest.mock import MagicMock

import _live_reference


def load():
    ref = _live_reference.load()
    if not hasattr(inspect, "getargspec"):
        inspect.getargspec = inspect.getfullargspec
    import scipy.integrate
    if not hasattr(scipy.integrate, "trapz"):
        scipy.integrate.trapz = scipy.integrate.trapezoid
    if "imp" not in sys.modules:
        imp = types.ModuleType("imp")
        imp.find_module = lambda *a, **k: None           # imported, never called (refinement/methods/__init__.py:12)
        sys.modules["imp"] = imp
    for name in ("gi", "gi.repository", "cairocffi", "matplotlib", "matplotlib.pyplot", "matplotlib.figure",
                 "matplotlib.backends", "matplotlib.backends.backend_gtk3", "matplotlib.backends.backend_gtk3cairo",
                 "matplotlib.font_manager", "matplotlib.transforms", "matplotlib.text", "matplotlib.patches",
                 "matplotlib.offsetbox", "matplotlib.ticker", "matplotlib.gridspec", "matplotlib.lines",
                 "matplotlib.mathtext", "matplotlib.cbook", "matplotlib.widgets", "matplotlib.image",
                 "matplotlib.artist", "matplotlib.axes", "matplotlib.colors",
                 "mpl_toolkits", "mpl_toolkits.axisartist", "mpl_toolkits.axes_grid1", "mpl_toolkits.mplot3d",
                 "Pyro4", "Pyro4.errors", "Pyro4.naming", "Pyro4.util"):
        sys.modules.setdefault(name, MagicMock())
    from pyxrd.data import settings
    settings._parse_args = lambda: types.SimpleNamespace(script=None, filename="", debug=False, clear_cache=False)
    try:
        settings.initialize()
    except Exception:
        pass
    import pyxrd.probabilities.models as pm
    from pyxrd.probabilities.models.base_models import _AbstractProbability
    # change-monitoring helpers (base_models.py:289-309) stack level matrices of different sizes with
    # np.asanyarray, which numpy >= 1.24 refuses; they do no arithmetic: always report "changed"
    _AbstractProbability._stash_matrices = lambda self: None
    _AbstractProbability._compare_stashed_matrices = lambda self: False
    ref.probabilities = pm
    return ref
