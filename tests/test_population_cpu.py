"""Host half of the population path, no GPU: the native expansion of a
template + binding table (``pyxrd_expand_population``) must produce exactly the arrays that packing K
separately built DataObject graphs produces."""
import copy

import numpy as np
import pytest

from oracle import pyxrd_oracle_models as om
from pyxrd_b200 import packing, runtime
from pyxrd_b200 import synthetic as syn
from pyxrd_b200.population import PopulationTemplate, Refinable


def _ensure_built():
    import os
    if not os.path.exists(runtime.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def built_candidates(tmpl, x):
    """K independent DataObject graphs: the template with solution x[k] written into it"""
    out = []
    for row in x:
        tmpl.apply(row, probabilities=om.probabilities)
        out.append(copy.deepcopy(tmpl.mixture))
    return out


@pytest.mark.parametrize("name,K", [("c1", 5), ("c2", 7), ("c4", 4), ("c5", 3), ("models", 6)])
def test_expansion_equals_packing_of_rebuilt_graphs(name, K):
    _ensure_built()
    mix, refs = syn.template(name)
    tmpl = PopulationTemplate(mix, refs)
    x = syn.sample_solutions(refs, 11, K)
    got = tmpl.expand(x)
    want = packing.BatchPack(built_candidates(tmpl, x), tmpl.static)
    assert got["K"] == K == want.K and got["M"] == want.M
    for key in ("phase_i", "phase_d", "phase_comp", "comp_d", "comp_i", "atom_d", "atom_type", "job_phase",
                "fractions", "scales", "bgshifts", "prob_model", "prob_params"):
        assert np.array_equal(got[key], getattr(want, key)), key
    # fixed (not model-driven) matrices travel unchanged
    for p, row in enumerate(want.phase_i):
        if (want.prob_model is None or want.prob_model[p] == 0) and row[1] > 0:
            n = 2 * row[1] * row[1]
            assert np.array_equal(got["matrices"][row[6]:row[6] + n], want.matrices[row[6]:row[6] + n])
    # the automatic CSDS maximum follows the bound average (phases/models/CSDS.py:127,140)
    assert np.array_equal(got["phase_i"][:, 3], (2.5 * got["phase_d"][:, 1]).astype(np.int32))


def test_mixture_level_and_matrix_bindings():
    _ensure_built()
    mix = syn.CONFIGS["c4"].truth()
    ph = mix.specimens[0].phases[0][3]
    refs = [Refinable("f0", [(mix, "fractions:0"), (mix, "fractions:2", -1.0, 0.5)]),
            Refinable("s1", [(mix, "scales:1", 2.0, 0.0)]),
            Refinable("b0", [(mix, "bgshifts:0")]),
            Refinable("p01", [(ph, "P:0:1"), (ph, "P:0:0", -1.0, 1.0)]),
            Refinable("pn", [(ph.components[0].layer_atoms[2], "pn")])]
    tmpl = PopulationTemplate(mix, refs, csds_auto_maximum=False)
    x = np.array([[0.2, 3.0, 1.5, 0.25, 7.0], [0.4, 5.0, -2.0, 0.5, 9.0]])
    got = tmpl.expand(x)
    base = tmpl.batch
    assert np.array_equal(got["fractions"][:, 0], x[:, 0]) and np.array_equal(got["fractions"][:, 2], 0.5 - x[:, 0])
    assert np.array_equal(got["fractions"][:, 1], np.repeat(base.fractions[0, 1], 2))
    assert np.array_equal(got["scales"][:, 1], 2.0 * x[:, 1]) and np.array_equal(got["bgshifts"][:, 0], x[:, 2])
    pid = base.phase_index[id(ph)]
    off, NM = base.phase_i[pid, 6], base.matrices.size
    for k in range(2):
        P = got["matrices"][k * NM + off + 4:k * NM + off + 8].reshape(2, 2)
        assert P[0, 1] == x[k, 3] and P[0, 0] == 1.0 - x[k, 3]
        assert np.array_equal(P[1], np.asarray(ph.P)[1])
    aidx = base.atom_index[id(ph.components[0].layer_atoms[2])]
    NA = base.atom_type.size
    assert got["atom_d"][aidx, 0] == 7.0 and got["atom_d"][NA + aidx, 0] == 9.0
    # indices of the second candidate point into its own tables
    NP = base.phase_i.shape[0]
    assert np.array_equal(got["job_phase"][base.job_phase.size:], base.job_phase + NP)
    assert np.array_equal(got["phase_i"][NP:, 6], base.phase_i[:, 6] + NM)
    assert np.array_equal(got["phase_i"][:NP, 3], base.phase_i[:, 3])        # maximum untouched


def test_binding_errors():
    _ensure_built()
    mix, refs = syn.template("c2")
    tmpl = PopulationTemplate(mix, refs)
    with pytest.raises(ValueError):
        tmpl.expand(np.zeros((2, 3)))                       # wrong number of columns
    bad = tmpl.bindings.copy()
    bad["index"][0] = 10 ** 6
    with pytest.raises(ValueError, match="out of range"):
        runtime.expand_population(tmpl.batch, np.zeros((1, len(refs))), bad)
    other = syn.CONFIGS["c1"].truth()
    with pytest.raises(KeyError):
        PopulationTemplate(mix, [Refinable("x", [(other.specimens[0].phases[0][0], "sigma_star")])])
    with pytest.raises(KeyError):
        PopulationTemplate(mix, [Refinable("x", [(mix.specimens[0].phases[0][0], "W:0:0")])])   # model-driven
    with pytest.raises(ValueError):
        ph = syn.CONFIGS["c2"].truth().specimens[0].phases[0][0]
        ph.prob_model, ph.prob_params = "R2G2", [0.5]
        packing.prob_model_of(ph)


def test_packing_caches_are_keyed_by_value():
    """the serial drop-in calls repack a graph per trial: atom-type keys are memoised per packing call (by object),
    wavelength tables cached across calls (by value) -- neither may change what is packed"""
    mix, _ = syn.template("c2")
    spec = mix.specimens[0]
    g = spec.goniometer
    a = packing.wavelength_tables(spec.range_theta, g.wavelength, g.wavelength_distribution)
    b = packing.wavelength_tables(np.array(spec.range_theta), g.wavelength, list(g.wavelength_distribution))
    assert all(u is v for u, v in zip(a, b)) and not a[1].flags.writeable            # one read-only entry
    fresh = packing._wavelength_tables(spec.range_theta, g.wavelength, g.wavelength_distribution)
    assert all(np.array_equal(u, v) for u, v in zip(a, fresh))
    c = packing.wavelength_tables(spec.range_theta, g.wavelength * 1.001, g.wavelength_distribution)
    assert not np.array_equal(a[2], c[2])
    # equal-valued atom types held by different objects land on the same table row, a changed one on a new row
    other = copy.deepcopy(mix)
    atoms = [at for ph in packing.iter_phases(other.specimens[0]) for comp in ph.components
             for at in list(comp.layer_atoms) + list(comp.interlayer_atoms) if at is not None and at.atom_type is not None]
    atoms[0].atom_type = copy.deepcopy(atoms[0].atom_type)
    types = packing.collect_atom_types(other.specimens)
    assert types == packing.collect_atom_types(mix.specimens)
    one = packing.BatchPack([other], packing.StaticPack(other.specimens, types))
    two = packing.BatchPack([mix], packing.StaticPack(mix.specimens, types))
    assert np.array_equal(one.atom_type, two.atom_type)
    atoms[0].atom_type.debye = float(atoms[0].atom_type.debye) + 0.25
    assert len(packing.collect_atom_types(other.specimens)) == len(types) + 1


def test_optimiser_scope_follows_the_mixture_masks(monkeypatch):
    """the device optimiser frees every fraction and scale and minimises Rp.  A template whose mixture fixes
    a fraction, switches auto_scales off (scales_mask zero, mixture/models/mixture.py:64-68), frees only part of the
    background shifts or asks for another residual must refuse ``optimized_residuals`` -- the refinement then stays on the
    per-trial path -- instead of silently minimising another objective."""
    from pyxrd_b200 import settings
    mix, refs = syn.template("c4")
    rng = np.random.default_rng(3)
    for spec in mix.specimens:
        spec.observed_intensity = rng.uniform(20.0, 60.0, (spec.range_theta.size, 1))
    tmpl = PopulationTemplate(mix, refs)
    assert tmpl.optimiser_scope() is None and tmpl.mixture_refines_background()
    x = syn.sample_solutions(refs, 1, 2)

    def refused(fragment):
        reason = tmpl.optimiser_scope()
        assert reason and fragment in reason, reason
        with pytest.raises(ValueError):          # raised before any device call
            tmpl.optimized_residuals(x, ctx=object())

    mix.fractions_mask = np.array([1, 1, 0, 1, 1, 1.0])
    refused("fractions_mask")
    mix.fractions_mask = np.ones(6)
    mix.scales_mask = np.zeros(2)
    refused("scales_mask")
    mix.scales_mask = np.ones(2)
    mix.bgshifts_mask = np.array([1.0, 0.0])
    refused("bgshifts_mask")
    assert not tmpl.mixture_refines_background()         # a partially fixed background is not "free"
    mix.bgshifts_mask = np.zeros(2)
    assert tmpl.optimiser_scope() is None and not tmpl.mixture_refines_background()
    mix.bgshifts_mask = np.ones(2)
    real_get = settings.get          # (inside a PyXRD process the reference's own settings module is consulted)
    monkeypatch.setattr(settings, "get", lambda name: "Rpder" if name == "RESIDUAL_METHOD" else real_get(name))
    refused("Rpder")
    monkeypatch.setattr(settings, "get", real_get)
    mix.specimens[1].observed_intensity = np.zeros((0, 1))
    refused("no observations")
