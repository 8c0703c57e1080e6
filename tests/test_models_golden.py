"""Pin the probability-model oracle (oracle/pyxrd_oracle_models.py) against the live reference
model classes (tests/golden/models.npz, made by ``tests/golden/make_golden.py models``)."""
import numpy as np
import pytest

from oracle import pyxrd_oracle_models as om

CASES = [(1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (2, 2), (3, 2), (4, 2), (5, 3), (6, 4), (7, 3)]


@pytest.mark.parametrize("model,G", CASES)
def test_oracle_models_bit_identical_to_live_reference(golden, model, G):
    g = golden("models")
    key = "m%d_g%d" % (model, G)
    params, Ws, Ps, valid = g[key + "_params"], g[key + "_W"], g[key + "_P"], g[key + "_valid"]
    assert len(params) >= 100
    seen = set()
    for p, W, P, v in zip(params, Ws, Ps, valid):
        gW, gP, gv = om.probabilities(model, G, p)
        assert np.array_equal(gW, W, equal_nan=True) and np.array_equal(gP, P, equal_nan=True), (key, p)
        assert gv == bool(v), (key, p)
        seen.add(bool(v))
    if model not in (om.R0, om.R1G2, om.R3G2):        # R0 / R1G2 cannot leave [0, 1], R3G2 clamps
        assert seen == {True, False}, "golden set must exercise valid and invalid parameter sets"


def test_r1g2_matches_the_closed_form_used_by_the_synthetic_config():
    # synthetic._c2 writes the R1G2 relations out by hand; the model must give the same matrices
    W, P, valid = om.probabilities(om.R1G2, 2, [0.7, 0.8])
    assert valid
    assert np.array_equal(np.diag(W), [0.7, 1.0 - 0.7])
    p22 = 0.8
    p21 = 1.0 - p22
    p12 = (1.0 - 0.7) * p21 / 0.7
    assert np.array_equal(P, [[1.0 - p12, p12], [p21, p22]])
