"""Golden vectors from the reference's own integration fixture, ``test/test_mixture/test refinement.pyxrd``
(3 specimens x 3 phases -- two R0 G1 and one three-component phase -- with the project's real atom
types, goniometer and observed patterns).  Build container only: the project file is parsed by the UNMODIFIED reference
model layer (headless, ``_live_models``), ``Mixture.data_object`` (mixture/models/mixture.py:53-85) is what the reference
hands to ``pyxrd.calculations``; that graph is copied object by object into ``pyxrd_b200.data_objects`` instances (same
class names, same attributes; shared objects stay shared) so that it can be unpickled without the reference, and the
reference's outputs for it are stored next to it.

    python tests/golden/make_fixture.py   ->   tests/golden/fixture.pkl, tests/golden/fixture.npz
"""
import copy
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import _live_models  # noqa: E402

FIXTURE = "/root/reference/test/test_mixture/test refinement.pyxrd"


def convert(obj, memo):
    """reference DataObject graph -> pyxrd_b200.data_objects graph (layout only)"""
    from pyxrd_b200 import data_objects as mine
    if id(obj) in memo:
        return memo[id(obj)]
    cls = type(obj).__name__
    if hasattr(mine, cls) and cls.endswith("Data"):
        out = getattr(mine, cls)()
        memo[id(obj)] = out
        for key, value in vars(obj).items():
            setattr(out, key, convert(value, memo))
        return out
    if isinstance(obj, np.ndarray):
        return np.array(obj)
    if isinstance(obj, (list, tuple)):
        return type(obj)(convert(v, memo) for v in obj)
    if isinstance(obj, dict):
        return {k: convert(v, memo) for k, v in obj.items()}
    if isinstance(obj, (np.floating, np.integer, np.bool_)):
        return obj.item()
    if obj is None or isinstance(obj, (bool, int, float, str)):
        return obj
    raise TypeError("cannot carry %r (%s) into the fixture" % (obj, type(obj)))


def main():
    ref = _live_models.load()
    import pyxrd.project.models  # noqa: F401  (registers the storables)
    from pyxrd.file_parsers.json_parser import JSONParser
    project = JSONParser.parse(FIXTURE)
    data = project.mixtures[0].data_object
    mine = convert(data, {})
    with open(os.path.join(HERE, "fixture.pkl"), "wb") as f:
        pickle.dump(mine, f, protocol=4)
    out = {"source": np.array("reference test/test_mixture/test refinement.pyxrd through Mixture.data_object; numpy %s"
                              % np.__version__)}
    work = copy.deepcopy(data)
    ref.mixture.calculate_mixture(work)
    for s, spec in enumerate(work.specimens):
        out["corr_s%d" % s] = np.array(spec.correction)
        out["pi_s%d" % s] = np.array(spec.phase_intensities)
        out["total_s%d" % s] = np.array(spec.total_intensity)
    out["residuals"] = np.array(work.residuals, dtype=float)
    work = copy.deepcopy(data)
    r = ref.mixture.get_optimized_residual(work)
    out["opt_residual"] = np.array(r, dtype=float)
    out["opt_fractions"] = np.array(work.fractions, dtype=float)
    out["opt_scales"] = np.array(work.scales, dtype=float)
    out["opt_bgshifts"] = np.array(work.bgshifts, dtype=float)
    out["opt_residuals"] = np.array(work.residuals, dtype=float)
    np.savez_compressed(os.path.join(HERE, "fixture.npz"), **out)
    print("fixture: residuals", out["residuals"], "optimised", out["opt_residual"], "fractions", out["opt_fractions"])


if __name__ == "__main__":
    main()
