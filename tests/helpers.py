"""Shared helpers for the parity tests."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pyxrd_b200 import synthetic as syn  # noqa: E402

SEED = 2024          # population seed used by tests/golden/make_golden.py
N_GOLDEN = {"c1": 3, "c2": 3, "c3": 1, "c4": 3, "c5": 2}


def rel_err(a, b):
    """max |a-b| / |b| per point (0/0 -> 0)."""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    den = np.abs(b)
    num = np.abs(a - b)
    out = np.zeros_like(num)
    nz = den > 0
    out[nz] = num[nz] / den[nz]
    out[~nz] = num[~nz]
    return float(out.max()) if out.size else 0.0


def fingerprint(mix):
    """Must match tests/golden/make_golden.py:fingerprint."""
    acc = []
    for spec in mix.specimens:
        acc += [float(np.sum(spec.range_theta)), float(spec.goniometer.wavelength)]
        for phase in spec.phases[0]:
            acc += [float(np.sum(phase.P)), float(np.trace(phase.W)), float(phase.sigma_star),
                    float(phase.CSDS.average), float(phase.CSDS.maximum)]
            for comp in phase.components:
                acc += [comp.d001, comp.weight, comp.volume, comp.delta_c]
                acc += [a.pn for a in comp.layer_atoms + comp.interlayer_atoms]
                acc += [a.default_z for a in comp.layer_atoms + comp.interlayer_atoms]
    return np.array(acc, dtype=float)


def golden_mixtures(name, g, n=None):
    """[(tag, MixtureData)] rebuilt from the synthetic builders with the golden 'observed' attached."""
    cfg = syn.CONFIGS[name]
    n = N_GOLDEN[name] if n is None else n
    out = [("truth", cfg.truth())] + [("k%d" % i, cfg.candidate(SEED, i)) for i in range(n)]
    for tag, mix in out:
        assert np.array_equal(fingerprint(mix), g["%s_%s_fp" % (name, tag)]), \
            "synthetic builder drifted from the golden inputs (%s %s)" % (name, tag)
        for s, spec in enumerate(mix.specimens):
            spec.observed_intensity = np.array(g["%s_observed_s%d" % (name, s)])
    return out


def reference_fixture():
    """(MixtureData, golden outputs) of the reference's own integration fixture ``test/test_mixture/test refinement.pyxrd``
    (tests/golden/make_fixture.py): the DataObject graph the reference's model layer builds for it, carried into
    pyxrd_b200.data_objects instances"""
    import pickle
    golden = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(golden, "fixture.pkl"), "rb") as f:
        mix = pickle.load(f)
    return mix, np.load(os.path.join(golden, "fixture.npz"))
