"""Binding tables discovered from a model layer (population.discover_template).

``binding.pkl / binding.npz`` were made in the build container by ``tests/golden/make_binding_golden.py``: the REFERENCE's
running model layer (its integration fixture, 40 refinables switched on, a reference ``Refiner``) was probed for the table;
the golden vectors are what the reference itself makes of 12 random solutions."""
import copy
import os
import pickle

import numpy as np
import pytest

from helpers import ROOT, rel_err
from pyxrd_b200 import synthetic as syn
from pyxrd_b200.population import PopulationTemplate, Refinable, discover_template

GOLDEN = os.path.join(ROOT, "tests", "golden")


def reference_binding():
    with open(os.path.join(GOLDEN, "binding.pkl"), "rb") as f:
        mix = pickle.load(f)
    g = np.load(os.path.join(GOLDEN, "binding.npz"))
    tmpl = PopulationTemplate.from_table(mix, g["bindings"], g["bounds"], names=[str(n) for n in g["names"]],
                                         csds_auto=g["csds_auto"] if g["csds_auto"].size else None,
                                         csds_factor=float(g["csds_factor"]))
    return tmpl, g


def test_discovered_table_reproduces_the_reference_model_layer():
    """expansion of the stored table = packing of the graphs the reference's MVC built for the same solutions"""
    tmpl, g = reference_binding()
    assert tmpl.n_params == 40 and len(tmpl.bindings) == 60
    got = tmpl.expand(g["x"])
    K = len(g["x"])
    for key in ("phase_d", "comp_d", "atom_d", "prob_params", "fractions", "scales", "bgshifts"):
        want = g["packed_" + key].reshape((K,) + tuple(g["packed_" + key].shape[1:]))
        have = np.asarray(got[key]).reshape(want.shape)
        assert np.allclose(have, want, rtol=1e-12, atol=0.0), key
    want = g["packed_phase_i"]
    have = got["phase_i"].reshape(want.shape)
    assert np.array_equal(have[:, :, :5], want[:, :, :5])               # G, rank, CSDS range, flags (offsets move per candidate)
    # among them the weight fractions F1, F2 of the three R0 G3 phases (one per specimen): probability parameters
    names = [str(n) for n in g["names"]]
    assert sum("W1/Sum" in n or "W2/Sum" in n for n in names) == 6
    assert int(np.sum(tmpl.bindings["target"] == 4)) == 6
    # and the cell length b of a component whose d001 is refined as well: its volume a * b * d001
    # (phases/models/unit_cell_prop.py:199) is set by b and multiplied by d001 / d001_0
    b = names.index("Cell length b [nm]")
    row = tmpl.bindings[tmpl.bindings["param"] == b]
    assert len(row) == 1 and int(row["target"][0]) == 1 and int(row["index"][0]) % 8 == 4 and int(row["mode"][0]) == 0
    shared = tmpl.bindings[(tmpl.bindings["target"] == 1) & (tmpl.bindings["index"] == row["index"][0])]
    assert len(shared) == 2 and int(shared["mode"][1]) == 1 and shared["offset"][1] == 0.0
    # three interlayer atom contents of one smectite component (atom_relations.py:416: every content drives the pn of
    # its atoms): the component weight is set by the first and the other two add to it
    contents = [i for i, n in enumerate(names) if n in ("Glycol content", "Water content", "Ca content")]
    assert len(contents) == 3
    weight = tmpl.bindings[np.isin(tmpl.bindings["param"], contents) & (tmpl.bindings["target"] == 1)]
    assert len(weight) == 3 and len(set(weight["index"])) == 1 and int(weight["index"][0]) % 8 == 5
    assert [int(m) for m in weight["mode"]] == [0, 2, 2]
    assert int(np.sum(np.isin(tmpl.bindings["param"], contents) & (tmpl.bindings["target"] == 2))) == 7     # the pn slots


def _toy_model_layer(nonlinear=False, coupled=False):
    """a stand-in for an MVC: x -> MixtureData with inheritance (two phases share sigma*), a derived value
    (volume = 0.468 d001), a probability model and the automatic CSDS maximum"""
    base, _ = syn.template("c4")
    ad, eg = base.specimens[0].phases[0], base.specimens[1].phases[0]

    def build(x):
        mix = copy.deepcopy(base)
        a, e = mix.specimens[0].phases[0], mix.specimens[1].phases[0]
        for row in (a, e):
            row[3].sigma_star = x[0]
            row[3].prob_params = [x[1], row[3].prob_params[1]]
            row[0].CSDS.average = x[2]
            row[0].CSDS.maximum = int(2.5 * x[2])
        a[3].components[1].d001 = x[3]
        a[3].components[1].volume = 0.468 * (x[3] * (x[0] + 2.0) if coupled == "product" else
                                             x[3] * x[0] + x[3] + x[0] if coupled == "mixed" else x[3])
        if coupled == "product":                      # a sum as well: a component weight over two refined contents
            a[3].components[1].weight = 300.0 + 16.0 * x[1] + 27.0 * x[4]
        if nonlinear:
            e[3].components[1].delta_c = 0.01 * x[3] ** 3
        if coupled == "product":                      # quadratic in one refinable: distinct roots, and a square
            e[3].components[1].delta_c = 0.01 * (x[3] + 0.5) * (2.0 * x[3] - 1.0)
            e[3].components[1].volume = 0.3 * (0.2 * x[3] + 0.9) ** 2
        mix.fractions = np.array([x[4], 0.1, 0.1, 0.25, 0.1, 0.45 - x[4]])
        return mix

    def phase_model(s, z, m):
        ph = (ad, eg)[s][m]
        return (ph.prob_model, ["p%d" % i for i in range(len(ph.prob_params))], None) if getattr(ph, "prob_model", None) else None

    return build, base


def test_discovery_on_a_toy_model_layer():
    build, base = _toy_model_layer()
    x0 = np.array([6.0, 0.6, 10.0, 1.25, 0.3])
    bounds = np.array([[2.0, 12.0], [0.3, 0.9], [8.0, 14.0], [1.22, 1.28], [0.05, 0.4]])

    def get(x):
        mix = build(x)
        return mix

    # phases keep the prob_model attributes synthetic.template gave them: no phase_model callback needed
    tmpl = discover_template(get, x0, bounds, names=["sigma", "W1", "T", "d001", "f0"], verify=5)
    table = {(int(r["param"]), int(r["target"]), int(r["index"])): (float(r["scale"]), float(r["offset"])) for r in tmpl.bindings}
    by_param = {i: [k for k in table if k[0] == i] for i in range(5)}
    assert len(by_param[0]) == 2 and len(by_param[1]) == 2          # both copies of the phase
    assert len(by_param[2]) == 2 and tmpl.csds_auto is not None and tmpl.csds_auto.sum() == 2
    scales = sorted(table[k][0] for k in by_param[3])
    assert len(scales) == 2 and abs(scales[0] - 0.468) < 1e-9 and scales[1] == 1.0
    got4 = sorted(table[k] for k in by_param[4])
    assert len(got4) == 2 and np.allclose(got4, sorted([(1.0, 0.0), (-1.0, 0.45)]), rtol=0, atol=1e-12)
    x = bounds[:, 0] + np.random.default_rng(1).uniform(0, 1, (6, 5)) * (bounds[:, 1] - bounds[:, 0])
    got = tmpl.expand(x)
    from pyxrd_b200 import packing
    want = packing.BatchPack([build(row) for row in x], tmpl.static)
    for key in ("phase_d", "comp_d", "atom_d", "prob_params", "fractions"):
        assert np.allclose(got[key], getattr(want, key), rtol=1e-12, atol=0.0), key
    assert np.array_equal(got["phase_i"][:, 3], want.phase_i[:, 3])


def test_discovery_of_slots_driven_by_two_refinables():
    """a cell volume a * b * d001 with two of the factors refined (phases/models/unit_cell_prop.py:199) is a product
    of affine functions, a component weight over two refined atom contents (atom_relations.py:244) a sum: both are
    found with one joint probe and expanded by multiply / add rows"""
    from pyxrd_b200 import packing
    from pyxrd_b200.population import BIND_MODE_ADD, BIND_MODE_MULTIPLY, BIND_MODE_SET
    build, _ = _toy_model_layer(coupled="product")
    x0 = np.array([6.0, 0.6, 10.0, 1.25, 0.3])
    bounds = np.array([[2.0, 12.0], [0.3, 0.9], [8.0, 14.0], [1.22, 1.28], [0.05, 0.4]])
    tmpl = discover_template(build, x0, bounds, verify=5)
    modes = {(int(r["param"]), int(r["mode"])) for r in tmpl.bindings if int(r["target"]) == 1}
    assert (0, BIND_MODE_SET) in modes and (3, BIND_MODE_MULTIPLY) in modes          # volume: sigma first, then d001
    assert (1, BIND_MODE_SET) in modes and (4, BIND_MODE_ADD) in modes               # weight: W1 first, then f0
    # quadratic in d001 alone (cell lengths a and b both following one atom ratio make a volume like this): two rows
    quad = {}
    for r in tmpl.bindings[(tmpl.bindings["param"] == 3) & (tmpl.bindings["target"] == 1)]:
        quad.setdefault(int(r["index"]), []).append(int(r["mode"]))
    assert sorted(quad.values()) == [[0], [0, 1], [0, 1], [1]]      # d001 itself, delta_c, the square, the product above
    x = bounds[:, 0] + np.random.default_rng(3).uniform(0, 1, (7, 5)) * (bounds[:, 1] - bounds[:, 0])
    got = tmpl.expand(x)
    want = packing.BatchPack([build(row) for row in x], tmpl.static)
    for key in ("phase_d", "comp_d", "atom_d", "prob_params", "fractions"):
        assert np.allclose(got[key], getattr(want, key), rtol=1e-12, atol=0.0), key
    # the stored table round-trips through rows with and without the mode column
    again = PopulationTemplate.from_table(tmpl.mixture, [tuple(r) for r in tmpl.bindings[["param", "target", "index", "scale", "offset", "mode"]]],
                                          bounds, csds_auto=tmpl.csds_auto, csds_factor=tmpl.csds_factor)
    assert np.array_equal(again.bindings, tmpl.bindings)


def test_discovery_refuses_what_a_table_cannot_express():
    x0 = np.array([6.0, 0.6, 10.0, 1.25, 0.3])
    bounds = np.array([[2.0, 12.0], [0.3, 0.9], [8.0, 14.0], [1.22, 1.28], [0.05, 0.4]])
    build, _ = _toy_model_layer(nonlinear=True)
    with pytest.raises(NotImplementedError, match="non-affinely"):
        discover_template(build, x0, bounds)
    build, _ = _toy_model_layer(coupled="mixed")
    with pytest.raises(NotImplementedError, match="together|interact"):
        discover_template(build, x0, bounds)

    def fixed_matrices(x):                                  # a probability matrix driven without a named model
        mix = _toy_model_layer()[0](x)
        for spec in mix.specimens:
            ph = spec.phases[0][3]
            ph.prob_model = None
            ph.W = np.diag([x[1], 1.0 - x[1]])
            ph.P = np.array([[x[1] ** 2, 1.0 - x[1] ** 2], [0.5, 0.5]])
        return mix
    with pytest.raises(NotImplementedError, match="probability matrix"):
        discover_template(fixed_matrices, x0, bounds)
