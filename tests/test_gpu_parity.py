"""Parity of the CUDA path (through the C-ABI) against the live-reference golden
vectors and the CPU oracle.  Tolerance for intensities / totals / residuals at a given
solution: relative 1e-9 per point (north_star); measured agreement is ~1e-12."""
import copy

import numpy as np
import pytest

from helpers import SEED, golden_mixtures, rel_err
from oracle import pyxrd_oracle as orc
from pyxrd_b200 import synthetic as syn
from pyxrd_b200.data_objects import AtomData, AtomTypeData, ComponentData, CSDSData

pytestmark = pytest.mark.gpu

RTOL = 1e-9          # north_star tolerance, fp64
OPT_RTOL = 5e-3      # optimiser end point: only reproducible to ~1e-3 (see test_oracle_golden.py)

KAT_THETA = np.asanyarray([2.2551711385, 2.478038901, 2.7001518288, 2.9214422642, 3.1418428,
                           3.3612863, 3.5797059197, 3.7970351263])


@pytest.fixture(scope="module")
def calc():
    import pyxrd_b200.calculations as c
    return c


# ---- the reference's own known-answer vectors, through the device leaf functions ---------
def test_kat_lpf(calc):
    got = calc.goniometer.get_lorentz_polarisation_factor(KAT_THETA, 12, 2.3, 2.3, 0)
    want = [3.00643375e-03, 4.83799816e-03, 1.33173586e-02, 6.56714627e-02, 5.11941694e+02,
            6.59637604e-02, 1.35695934e-02, 4.97673826e-03]
    assert np.allclose(got, want)                     # test_goniometer.py:74-86
    assert rel_err(got, orc.get_lorentz_polarisation_factor(KAT_THETA, 12, 2.3, 2.3, 0)) < 1e-11


def test_kat_ads(calc):
    got = calc.goniometer.get_fixed_to_ads_correction_range(KAT_THETA, None)
    want = [7.74814435e-01, 6.15920411e-01, 4.27242611e-01, 2.18376385e-01, -2.50146408e-04,
            -2.17930643e-01, -4.24231682e-01, -6.09510086e-01]
    assert np.allclose(got, want)                     # test_goniometer.py:63-72


def test_kat_asf(calc):
    at = AtomTypeData(par_a=np.full(5, 10.2), par_b=np.full(5, 10.2), par_c=20.2, debye=0.0)
    stl = np.array([0.1152, 0.5756, 1.1469])
    got = calc.atoms.get_atomic_scattering_factor((stl * 0.05) ** 2, at)
    want = [71.1827439324707, 70.77093939463964, 69.51772020477823]   # test_atom_type.py:58-75
    assert rel_err(got, want) < 1e-14
    assert np.array_equal(calc.atoms.get_atomic_scattering_factor(stl, None), np.zeros(3))


def test_leaf_component_and_csds(calc, golden):
    g = golden("kat")
    h = AtomTypeData(par_a=np.asanyarray([0.413048, 0.294953, 0.187491, 0.080701, 0.023736]),
                     par_b=np.asanyarray([15.569946, 32.398468, 5.711404, 61.889874, 1.334118]),
                     par_c=0.000049, debye=0)
    comp = ComponentData(layer_atoms=[AtomData(atom_type=h, pn=1, default_z=0, z=0)],
                         interlayer_atoms=[AtomData(atom_type=h, pn=1, default_z=0.55, z=0.55)],
                         volume=1.0, weight=1.0, d001=0.55, default_c=0.6, delta_c=0.02, lattice_d=0.45)
    sf, phi = calc.components.get_factors(g["kat_theta"], comp)
    assert np.max(np.abs(sf - g["kat_comp_sf"])) < 1e-13
    assert np.max(np.abs(phi - g["kat_comp_phi"])) < 1e-13
    comp2 = copy.deepcopy(comp)
    orc.get_factors(g["kat_theta"], comp2)
    assert comp.interlayer_atoms[0].z == comp2.interlayer_atoms[0].z        # the z write-back
    sfa = calc.atoms.get_structure_factor(g["kat_theta"], comp.interlayer_atoms[0])
    assert np.max(np.abs(sfa - orc.get_structure_factor(g["kat_theta"], comp2.interlayer_atoms[0]))) < 1e-13
    csds = CSDSData(average=10, maximum=50, minimum=1, alpha_scale=0.9485, alpha_offset=0.017,
                    beta_scale=0.1032, beta_offset=0.0034)
    arr, mean = calc.CSDS.calculate_distribution(csds)
    assert rel_err(arr, g["kat_csds_arr"]) < 1e-13 and abs(mean - float(g["kat_csds_mean"])) < 1e-12
    coef = calc.CSDS.progression_factors(csds)
    want = orc.progression_factors(g["kat_csds_arr"], 1, 50)
    assert rel_err(coef[1:51], [want[n] for n in range(1, 51)]) < 1e-12


# ---- the BASELINE configs against live-reference outputs ----------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c4", "c5", "c3"])
def test_config_matches_live_reference(calc, golden, name):
    g = golden(name)
    for tag, mix in golden_mixtures(name, g):
        calc.mixture.calculate_mixture(mix)
        for s, spec in enumerate(mix.specimens):
            assert rel_err(spec.correction, g["%s_%s_corr_s%d" % (name, tag, s)]) < 1e-14
            assert rel_err(spec.phase_intensities, g["%s_%s_pi_s%d" % (name, tag, s)]) < RTOL, (name, tag, s)
            assert rel_err(spec.total_intensity, g["%s_%s_total_s%d" % (name, tag, s)]) < RTOL
        assert rel_err(mix.residuals, g["%s_%s_residuals" % (name, tag)]) < RTOL
        x = np.array(g["%s_%s_x" % (name, tag)])
        mix.calculated = False
        res = calc.mixture._get_residuals(x, mix)
        assert rel_err(res, g["%s_%s_x_residuals" % (name, tag)]) < RTOL
        for s, spec in enumerate(mix.specimens):
            assert rel_err(spec.total_intensity, g["%s_%s_x_total_s%d" % (name, tag, s)]) < RTOL


@pytest.mark.parametrize("name", ["c1", "c4"])
@pytest.mark.parametrize("mode", ["scipy", "device"])
def test_optimize_mixture_against_live_reference(calc, golden, name, mode):
    """"scipy": the reference's own L-BFGS-B call with the device objective ends where the reference ends;
    "device" (default): never worse than that end point (one-sided)"""
    from pyxrd_b200 import settings
    g = golden(name)
    settings.MIXTURE_OPTIMIZER = mode
    try:
        for tag, mix in golden_mixtures(name, g)[:2]:
            r = calc.mixture.get_optimized_residual(mix)
            want = float(g["%s_%s_opt_residual" % (name, tag)][0])
            assert r.shape == (1,)
            if mode == "scipy":
                assert abs(float(r[0]) - want) <= OPT_RTOL * want, (name, tag, float(r[0]), want)
            else:
                assert float(r[0]) <= want * (1.0 + OPT_RTOL), (name, tag, float(r[0]), want)
            assert mix.optimized and mix.calculated
            assert abs(np.sum(mix.fractions) - 1.0) < 1e-5
            # the reported residual is the objective at the solution handed back (up to the 6-digit rounding)
            assert abs(mix.residuals[0] - float(r[0])) <= 2e-3 * float(r[0])
    finally:
        settings.MIXTURE_OPTIMIZER = "device"


def test_population_batch_equals_single_calls(calc, golden):
    from pyxrd_b200.batched import population_phase_intensities, population_residuals
    g = golden("c4")
    mixes = [m for _, m in golden_mixtures("c4", g)]
    res = population_residuals(mixes)
    pis = population_phase_intensities(mixes)
    for k, (tag, mix) in enumerate(golden_mixtures("c4", g)):
        assert rel_err(res[k], g["c4_%s_residuals" % tag]) < RTOL
        for s in range(2):
            assert rel_err(pis[k][s], g["c4_%s_pi_s%d" % (tag, s)]) < RTOL


# ---- edge cases the reference handles -------------------------------------------------
def _oracle_vs_device(calc, mix, tol=RTOL):
    ref = copy.deepcopy(mix)
    orc.calculate_mixture(ref)
    calc.mixture.calculate_mixture(mix)
    for a, b in zip(mix.specimens, ref.specimens):
        if a is None:
            continue
        assert rel_err(a.phase_intensities, b.phase_intensities) < tol
        assert rel_err(a.total_intensity, b.total_intensity) < tol
    assert rel_err(mix.residuals, ref.residuals) < tol
    return mix, ref


def _with_observed(mix, level=50.0):
    rng = np.random.default_rng(3)
    for spec in mix.specimens:
        spec.observed_intensity = rng.uniform(0.5, 1.5, (spec.range_theta.size, 1)) * level
    return mix


def test_none_phase_and_invalid_probabilities_give_zeros(calc):
    mix = _with_observed(syn.CONFIGS["c4"].truth())
    mix.specimens[0].phases[0][1] = None
    mix.specimens[1].phases[0][3].valid_probs = False
    mix, _ = _oracle_vs_device(calc, mix)
    assert not mix.specimens[0].phase_intensities[1].any()
    assert not mix.specimens[1].phase_intensities[3].any()


def test_flags_lpf_correction_and_machine_setups(calc):
    mix = _with_observed(syn.CONFIGS["c4"].truth())
    mix.specimens[0].phases[0][0].apply_lpf = False
    mix.specimens[0].phases[0][3].apply_correction = False
    mix.specimens[0].goniometer = syn.make_goniometer("cu_pair", divergence_mode="AUTOMATIC", absorption=True)
    mix.specimens[1].goniometer = syn.make_goniometer("single", absorption=True)
    mix.specimens[1].goniometer.mcr_2theta = 26.6
    _oracle_vs_device(calc, mix)


def test_ragged_specimens_and_exclusions(calc):
    cfg = syn.CONFIGS["c4"]
    mix = cfg.truth()
    th = syn.theta_grid(4.0, 30.0, 0.05)
    old = mix.specimens[1]
    mix.specimens[1] = syn.make_specimen(th, old.phases[0], old.goniometer, exclusion=(10.0, 12.5))
    _with_observed(mix)
    mix, ref = _oracle_vs_device(calc, mix)
    assert mix.specimens[1].phase_intensities.shape == (6, 1, th.size)


def test_no_observations_and_empty_mixture(calc):
    mix = syn.CONFIGS["c1"].truth()
    mix.specimens[0].observed_intensity = np.array([], dtype=float)
    calc.mixture.calculate_mixture(mix)
    assert mix.residuals == [0.0, 0.0]
    empty = syn.make_mixture([], 0, fractions=[])
    out = calc.mixture.calculate_mixture(empty)           # AssertionError swallowed (mixture.py:229-236)
    assert out is empty and not empty.calculated
    assert calc.mixture.optimize_mixture(empty) is empty and not empty.optimized


def test_generic_rank_and_csds_quirks(calc):
    types = syn.atom_types()
    rng = np.random.default_rng(11)
    theta = syn.theta_grid(3.0, 20.0, 0.05)
    phases = []
    # rank 25 (G=5, R2): generic kernel; rank 16 (G=4, R2) and (G=2, R4): staged kernels; rank 6 R1
    for G, R in ((5, 2), (4, 2), (2, 4), (6, 1), (3, 2)):
        comps = [syn.make_component(types, k, rng) for k in
                 (["illite", "smectite_ad", "chlorite", "smectite_eg", "kaolinite", "illite"][:G])]
        W, P = syn.make_probabilities(rng, G, R)
        phases.append(syn.make_phase(comps, W, P, syn.make_csds(6.0 + G), sigma_star=4.0))
    phases[1].CSDS.minimum = 3                      # n < minimum terms vanish
    phases[2].CSDS = syn.make_csds(1.2, maximum=1)  # empty series: S = mean * Id
    phases[3].CSDS.minimum = 0                      # python wrap-around Qn[-1] = Q^(N+1)
    spec = syn.make_specimen(theta, phases, syn.make_goniometer("cu_pair"))
    mix = _with_observed(syn.make_mixture([spec], len(phases)))
    _oracle_vs_device(calc, mix)


def test_single_phase_helpers(calc):
    mix = syn.CONFIGS["c2"].truth()
    spec = mix.specimens[0]
    phase = spec.phases[0][0]
    theta = spec.range_theta[::7]
    stl = 2 * np.sin(theta) / spec.goniometer.wavelength
    assert rel_err(calc.phases.get_diffracted_intensity(theta, stl, phase),
                   orc.get_diffracted_intensity(theta, stl, phase)) < RTOL
    assert rel_err(calc.phases.get_intensity(theta, stl, 2.3, 2.3, 0.0, phase),
                   orc.get_intensity(theta, stl, 2.3, 2.3, 0.0, phase)) < RTOL
    corr, pi = calc.specimen.calculate_phase_intensities(spec)
    corr_o, pi_o = orc.calculate_phase_intensities(spec)
    assert rel_err(corr, corr_o) < 1e-14 and rel_err(pi, pi_o) < RTOL
    assert rel_err(calc.specimen.get_machine_correction_range(spec), corr_o) < 1e-14
    spec.correction, spec.phase_intensities = corr_o, pi_o
    ref = copy.deepcopy(spec)
    calc.specimen.calculate_scaled_intensities(spec, 2.5, np.array([0.7]), 3.0)
    orc.calculate_scaled_intensities(ref, 2.5, np.array([0.7]), 3.0)
    assert rel_err(spec.total_intensity, ref.total_intensity) < 1e-14
    assert rel_err(spec.scaled_phase_intensities, ref.scaled_phase_intensities) < 1e-14
    assert rel_err(spec.background_intensity, ref.background_intensity) < 1e-14


# ---- size-independent properties at full BASELINE sizes -----------------------------------
def test_properties_full_size(calc):
    from pyxrd_b200.batched import population_phase_intensities, population_residuals
    cfg = syn.CONFIGS["c2"]
    mixes = cfg.population(SEED, 64)
    for m in mixes:
        _with_observed(m)
    pis = population_phase_intensities(mixes)
    again = population_phase_intensities(mixes[::-1])[::-1]
    for a, b in zip(pis, again):                      # order independence, bit for bit
        assert np.array_equal(a[0], b[0])
    assert all(np.isfinite(p[0]).all() and (p[0] > 0).all() for p in pis)
    # Rp is homogeneous of degree 0 under a joint rescale of observed and (scale, bgshift)
    r1 = population_residuals(mixes)
    for m in mixes:
        m.specimens[0].observed_intensity = m.specimens[0].observed_intensity * 4.0
        m.scales = np.asarray(m.scales) * 4.0
        m.bgshifts = np.asarray(m.bgshifts) * 4.0
    r2 = population_residuals(mixes)
    assert rel_err(r2, r1) < 1e-12
    # a wavelength distribution equal to the single tube wavelength doubles the pattern
    # (sequential in-place accumulation, specimen.py:71-74) except where the shifted axis
    # rounds outside the range at the two end points
    one = syn.CONFIGS["c1"].truth()
    _with_observed(one)
    base = population_phase_intensities([one])[0][0]
    one.specimens[0].goniometer.wavelength_distribution = []
    bare = population_phase_intensities([one])[0][0]
    assert rel_err(base[0, 0, 1:-1], 2.0 * bare[0, 0, 1:-1]) < 1e-9


# ---- device-resident optimiser (pyxrd_optimize) ---------------------------------------------
@pytest.mark.parametrize("name", ["c1", "c2", "c4", "c5"])
def test_device_optimizer_reaches_reference_optimum(calc, golden, name):
    """one-sided contract: the device optimum is never worse than the reference's L-BFGS-B end
    point by more than OPT_RTOL (it is often lower: the reference stalls on the L1 objective)"""
    from pyxrd_b200.batched import get_optimized_residuals
    g = golden(name)
    tagged = golden_mixtures(name, g)
    mixes = [m for _, m in tagged]
    out = get_optimized_residuals(mixes)
    for (tag, mix), r in zip(tagged, out):
        want = float(g["%s_%s_opt_residual" % (name, tag)][0])
        assert r.shape == (1,) and np.isfinite(r[0])
        assert r[0] <= want * (1.0 + OPT_RTOL), (name, tag, float(r[0]), want)
        assert mix.optimized and abs(np.sum(mix.fractions) - 1.0) < 1e-5
        assert (np.asarray(mix.fractions) >= 0).all() and (np.asarray(mix.scales) > 0).all()
        assert (np.asarray(mix.bgshifts) >= 0).all()


@pytest.mark.parametrize("name", ["c1", "c2", "c4"])
def test_device_optimizer_reports_the_true_objective(calc, golden, name):
    """the residual the optimiser returns is the reference objective at the solution it returns (c1 / c2: the
    single-lane solver for at most five unknowns, c4: the register team solver)"""
    from pyxrd_b200.batched import pack_population
    from pyxrd_b200.runtime import default_context
    g = golden(name)
    mixes = [m for _, m in golden_mixtures(name, g)]
    for m in mixes:
        calc.mixture._parse_bookkeeping(m)
    static, batch = pack_population(mixes)
    ctx = default_context()
    with ctx.lock:
        ctx.set_static(static)
        ctx.set_batch(batch)
        ctx.compute_intensities()
        ctx.optimize(refine_bg=True)
        sol, res = ctx.read_solutions(), ctx.read_opt_residuals()
        M, S = batch.M, batch.S
        ctx.set_solution(sol[:, :M], sol[:, M:M + S], sol[:, M + S:])
        ctx.compute_residuals()
        check = ctx.read_residuals()
    assert rel_err(res, check) < 1e-9
    assert np.all(np.abs(sol[:, :M].sum(axis=1) - 1.0) < 1e-12)


def test_device_optimizer_fixed_background_and_fallback(calc, golden):
    from pyxrd_b200.batched import get_optimized_residuals
    g = golden("c4")
    mixes = [m for _, m in golden_mixtures("c4", g)][:2]
    for m in mixes:
        m.bgshifts_mask = np.zeros(2)
        m.bgshifts = np.array([15.0, 18.0])
    free = get_optimized_residuals([m for _, m in golden_mixtures("c4", g)][:2])
    fixed = get_optimized_residuals(mixes)
    for m, a, b in zip(mixes, fixed, free):
        assert np.array_equal(m.bgshifts, [15.0, 18.0])
        assert a[0] >= b[0] * (1 - 1e-6)              # fewer free variables cannot fit better
    # the same with one specimen and one phase (single-lane solver, background row fixed)
    g2 = golden("c2")
    one = [m for _, m in golden_mixtures("c2", g2)][:1]
    one[0].bgshifts_mask = np.zeros(1)
    one[0].bgshifts = np.array([7.0])
    a = get_optimized_residuals(one)[0]
    b = get_optimized_residuals([m for _, m in golden_mixtures("c2", g2)][:1])[0]
    assert np.array_equal(one[0].bgshifts, [7.0]) and np.isfinite(a[0]) and a[0] >= b[0] * (1 - 1e-6)
    # a partially fixed fraction vector is outside the device optimiser: scipy-driven path
    part = [m for _, m in golden_mixtures("c4", g)][0]
    part.fractions_mask = np.array([1, 1, 1, 0, 1, 1])
    r = get_optimized_residuals([part])[0]
    assert np.isfinite(r[0]) and part.optimized and abs(part.fractions[3] - 0.25) < 1e-6


# ---- RawPatternPhase and the alternative residual methods -------------------------------------
def test_raw_pattern_phase_matches_live_reference(calc, golden):
    """phases.py:97-104 through k_phase_raw: unsorted raw abscissa, fill 0 outside, LPF / correction flags"""
    g = golden("extras")
    mix = syn.raw_case()
    calc.mixture.calculate_mixture(mix)
    spec = mix.specimens[0]
    assert rel_err(spec.phase_intensities, g["raw_pi"]) < RTOL
    assert np.array_equal(spec.phase_intensities[1, 0] == 0.0, g["raw_pi"][1, 0] == 0.0)
    assert rel_err(spec.total_intensity, g["raw_total"]) < RTOL
    assert rel_err(mix.residuals, g["raw_residuals"]) < RTOL
    theta = spec.range_theta
    stl = 2 * np.sin(theta) / spec.goniometer.wavelength
    assert rel_err(calc.phases.get_diffracted_intensity(theta, stl, spec.phases[0][1]), g["raw_single"]) < RTOL
    assert rel_err(calc.phases.get_intensity(theta, stl, 2.3, 2.3, 0.0, spec.phases[0][2]), g["raw_single_lpf"]) < RTOL


def test_residual_methods(calc, golden):
    """settings.RESIDUAL_METHOD: Rpder on the device (statistics.py:29-48); Rpw / Rphase raise what the
    reference raises for them (statistics.py:58-61,81 cannot take the 2-D clipped patterns)."""
    from pyxrd_b200 import settings
    from pyxrd_b200.batched import get_optimized_residuals, population_residuals
    g = golden("extras")
    try:
        settings.RESIDUAL_METHOD = "Rpder"
        mix = syn.raw_case()
        calc.mixture.calculate_mixture(mix)
        assert rel_err(mix.residuals, g["rpder_raw_residuals"]) < RTOL
        mix = syn.rpder_case()
        calc.mixture.calculate_mixture(mix)
        assert rel_err(mix.residuals, g["rpder_c4_residuals"]) < RTOL
        assert rel_err(population_residuals([syn.rpder_case()])[0], g["rpder_c4_residuals"]) < RTOL
        # the optimiser minimises whatever settings select: scipy-driven path with the device objective
        want = float(g["rpder_c4_opt_residual"][0])
        got = float(get_optimized_residuals([syn.rpder_case()])[0][0])
        assert got <= want * (1.0 + 5e-2), (got, want)
        settings.RESIDUAL_METHOD = "Rpw"
        with pytest.raises(ValueError):
            calc.mixture.calculate_mixture(syn.raw_case())
        settings.RESIDUAL_METHOD = "Rphase"
        with pytest.raises(TypeError):
            calc.mixture.calculate_mixture(syn.raw_case())
        settings.RESIDUAL_METHOD = "nonsense"
        with pytest.raises(KeyError):
            calc.mixture.calculate_mixture(syn.raw_case())
    finally:
        settings.RESIDUAL_METHOD = "Rp"
    mix = syn.raw_case()
    calc.mixture.calculate_mixture(mix)
    assert rel_err(mix.residuals, g["raw_residuals"]) < RTOL


def test_two_patterns_per_specimen(calc, golden):
    """Z = 2 (z_list of length 2): per-pattern phase sets, a None phase in one pattern only, observed [X, 2]"""
    from pyxrd_b200.batched import get_optimized_residuals, population_residuals
    g = golden("extras")
    mix = syn.z2_case()
    calc.mixture.calculate_mixture(mix)
    for s, spec in enumerate(mix.specimens):
        assert spec.phase_intensities.shape == (6, 2, 2100) and spec.total_intensity.shape == (2, 2100)
        assert rel_err(spec.phase_intensities, g["z2_pi_s%d" % s]) < RTOL
        assert rel_err(spec.total_intensity, g["z2_total_s%d" % s]) < RTOL
    assert not mix.specimens[1].phase_intensities[2, 1].any()
    assert rel_err(mix.residuals, g["z2_residuals"]) < RTOL
    assert rel_err(population_residuals([syn.z2_case(), syn.z2_case()])[1], g["z2_residuals"]) < RTOL
    want = float(g["z2_opt_residual"][0])
    got = float(get_optimized_residuals([syn.z2_case()])[0][0])
    assert got <= want * (1.0 + OPT_RTOL), (got, want)
    got = float(calc.mixture.get_optimized_residual(syn.z2_case())[0])
    assert got <= want * (1.0 + OPT_RTOL)


def test_compute_all_equals_the_two_calls(calc, golden):
    """pyxrd_compute_all = pyxrd_compute_intensities + pyxrd_compute_residuals; a None phase stays zero through the
    wavelength passes; specimens with wavelength lists of different length share the launches"""
    from pyxrd_b200.batched import pack_population
    from pyxrd_b200.runtime import default_context
    g = golden("c4")
    mixes = [m for _, m in golden_mixtures("c4", g)]
    mixes[1].specimens[0].phases[0][2] = None
    for m in mixes:           # the specimens (goniometer, observations) are shared by the candidates of a batch
        m.specimens[1].goniometer.wavelength_distribution = [(syn.CU_KA1, 1.0)]
    static, batch = pack_population(mixes)
    ctx = default_context()
    with ctx.lock:
        ctx.set_static(static, force=True)
        ctx.set_batch(batch)
        ctx.compute_intensities()
        ctx.compute_residuals(True)
        want = [ctx.read_intensities_flat(), ctx.read_totals_flat(), ctx.read_scaled_flat(), ctx.read_residuals()]
        ctx.set_batch(batch)
        ctx.compute_all(True)
        got = [ctx.read_intensities_flat(), ctx.read_totals_flat(), ctx.read_scaled_flat(), ctx.read_residuals()]
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    for k, mix in enumerate(mixes[:3]):
        ref = copy.deepcopy(mix)
        orc.calculate_mixture(ref)
        assert rel_err(got[3][k], ref.residuals) < RTOL


def _random_mixture(rng, types):
    """a random small mixture over the shapes the reference's model layer can produce (and a few it cannot)"""
    kinds = ["illite", "smectite_ad", "chlorite", "smectite_eg", "kaolinite", "illite"]
    shapes = [(1, 0), (2, 0), (3, 0), (4, 0), (2, 1), (3, 1), (4, 1), (2, 2), (3, 2), (2, 3), (5, 1), (6, 0)]
    S, Z, M = int(rng.integers(1, 4)), int(rng.integers(1, 3)), int(rng.integers(1, 5))
    specs = []
    for s in range(S):
        theta = syn.theta_grid(float(rng.uniform(2.0, 6.0)), float(rng.uniform(25.0, 60.0)), float(rng.uniform(0.11, 0.3)))
        rows = []
        for z in range(Z):
            row = []
            for m in range(M):
                if rng.uniform() < 0.12:
                    row.append(None)
                    continue
                G, R = shapes[int(rng.integers(len(shapes)))]
                comps = [syn.make_component(types, kinds[(m + c) % len(kinds)], rng, delta_c=float(rng.choice([0.0, 0.02, 0.05])))
                         for c in range(G)]
                W, P = syn.make_probabilities(rng, G, R)
                avg = float(rng.uniform(3.0, 24.0))
                csds = syn.make_csds(avg, minimum=int(rng.choice([0, 1, 1, 1, 2])))
                ph = syn.make_phase(comps, W, P, csds, sigma_star=float(rng.uniform(1.0, 20.0)),
                                    apply_lpf=bool(rng.uniform() < 0.85), apply_correction=bool(rng.uniform() < 0.85),
                                    valid=bool(rng.uniform() < 0.93))
                row.append(ph)
            rows.append(row)
        gon = syn.make_goniometer(str(rng.choice(["single", "cu_pair"])), divergence_mode=str(rng.choice(["FIXED", "AUTOMATIC"])),
                                  absorption=bool(rng.uniform() < 0.4))
        gon.mcr_2theta = float(rng.choice([0.0, 26.6]))
        if rng.uniform() < 0.2:
            gon.wavelength_distribution = []
        excl = (12.0, 14.0) if rng.uniform() < 0.4 else None
        sp = syn.make_specimen(theta, rows[0], gon, exclusion=excl)
        sp.phases = rows
        sp.z_list = [float(z) for z in range(Z)]
        sp.observed_intensity = rng.uniform(10.0, 90.0, (theta.size, Z))
        specs.append(sp)
    fr = rng.dirichlet(np.ones(M))
    return syn.make_mixture(specs, M, fractions=fr, scales=rng.uniform(0.3, 3.0, S), bgshifts=rng.uniform(0.0, 6.0, S))


def test_random_mixtures_against_the_oracle(calc):
    """32 random mixtures: 1-3 ragged specimens, 1-2 patterns each, 1-4 phases of random (G, R) / CSDS / flags, None phases,
    invalid probabilities, both divergence modes, absorption, monochromator, empty / single / double wavelength lists"""
    import os
    rng = np.random.default_rng(int(os.environ.get("PYXRD_FUZZ_SEED", "20261020")))
    types = syn.atom_types()
    worst = 0.0
    for _ in range(int(os.environ.get("PYXRD_FUZZ_COUNT", "32"))):
        mix = _random_mixture(rng, types)
        ref = copy.deepcopy(mix)
        orc.calculate_mixture(ref)
        calc.mixture.calculate_mixture(mix)
        for a, b in zip(mix.specimens, ref.specimens):
            assert a.phase_intensities.shape == b.phase_intensities.shape
            worst = max(worst, rel_err(a.phase_intensities, b.phase_intensities), rel_err(a.total_intensity, b.total_intensity))
            assert np.array_equal(a.phase_intensities == 0.0, b.phase_intensities == 0.0)
        worst = max(worst, rel_err(mix.residuals, ref.residuals))
    assert worst < RTOL, worst


def test_reference_integration_fixture(calc):
    """the reference's own project file test/test_mixture/test refinement.pyxrd, as its model layer hands it over
    (Mixture.data_object): patterns / totals / residuals at 1e-9 against the live reference, optimised residual
    (live reference: 0.05744468) within the optimiser contract in both modes"""
    from helpers import reference_fixture
    from pyxrd_b200 import settings
    from pyxrd_b200.batched import get_optimized_residuals
    mix, g = reference_fixture()
    work = copy.deepcopy(mix)
    calc.mixture.calculate_mixture(work)
    for s, spec in enumerate(work.specimens):
        assert rel_err(spec.correction, g["corr_s%d" % s]) < 1e-14
        assert rel_err(spec.phase_intensities, g["pi_s%d" % s]) < RTOL
        assert rel_err(spec.total_intensity, g["total_s%d" % s]) < RTOL
    assert rel_err(work.residuals, g["residuals"]) < RTOL
    want = float(g["opt_residual"][0])
    r = calc.mixture.get_optimized_residual(copy.deepcopy(mix))
    assert float(r[0]) <= want * (1.0 + OPT_RTOL) + 1e-9, (float(r[0]), want)
    r = get_optimized_residuals([copy.deepcopy(mix), copy.deepcopy(mix)])
    assert float(r[1][0]) <= want * (1.0 + OPT_RTOL) + 1e-9
    settings.MIXTURE_OPTIMIZER = "scipy"
    try:
        work = copy.deepcopy(mix)
        r = calc.mixture.get_optimized_residual(work)
        assert abs(float(r[0]) - want) <= OPT_RTOL * want, (float(r[0]), want)
        assert rel_err(work.fractions, g["opt_fractions"]) < 1e-3
    finally:
        settings.MIXTURE_OPTIMIZER = "device"


def test_dense_and_structured_matrices_of_the_same_rank(calc):
    """ranks 4 / 8 / 9 / 16 / 27: a P with the Reichweite >= 2 transition structure takes the structured kernels
    (k_phase_small<.., STRUCT = 1>, k_phase_staged<.., STRUCT = 1> for 16 / 27), a dense P of the same rank the dense
    ones (vector Horner, DMMA for 8 / 16 / 27);
    Reichweite 0 with 3 - 6 components takes the scalar series, a P of that rank with unequal rows the dense kernel"""
    types = syn.atom_types()
    rng = np.random.default_rng(77)
    theta = syn.theta_grid(3.0, 30.0, 0.05)
    kinds = ["illite", "smectite_ad", "chlorite", "smectite_eg", "kaolinite", "illite"]
    phases = []
    for G, R in ((2, 2), (2, 3), (3, 2), (2, 4), (4, 2), (3, 3)):
        for dense in (False, True):
            comps = [syn.make_component(types, kinds[c], rng) for c in range(G)]
            W, P = syn.make_probabilities(rng, G, R, dense=dense)
            phases.append(syn.make_phase(comps, W, P, syn.make_csds(7.0 + G + R), sigma_star=5.0))
    for G in (3, 4, 5, 6):
        for equal_rows in (True, False):
            comps = [syn.make_component(types, kinds[c], rng) for c in range(G)]
            W, P = syn.make_probabilities(rng, G, 0) if equal_rows else syn.make_probabilities(rng, G, 1)
            phases.append(syn.make_phase(comps, W, P, syn.make_csds(9.0), sigma_star=4.0))
    spec = syn.make_specimen(theta, phases, syn.make_goniometer("cu_pair"))
    mix = _with_observed(syn.make_mixture([spec], len(phases)))
    _oracle_vs_device(calc, mix)


def test_statistics_on_plain_arrays(calc, golden):
    """pyxrd_b200.calculations.statistics (k_leaf_statistic, k_leaf_derive) against the live reference"""
    g = golden("extras")
    st = calc.statistics
    e, c, ph = g["stat_exp"], g["stat_calc"], g["stat_phase"]
    got = [st.Rp(e, c), st.Rpw(e, c), st.R_squared(e, c), st.Rpe(e, c, 7), st.Rphase(e, c, ph), st.Rpder(e, c)]
    assert rel_err(got, g["stat_values"]) < 1e-12
    assert np.max(np.abs(st.derive(e) - g["stat_derive"])) < 1e-10 * np.max(np.abs(g["stat_derive"]))
    assert np.max(np.abs(st.smooth_pattern(e) - g["stat_smooth"])) < 1e-12 * np.max(np.abs(g["stat_smooth"]))
    assert st.Rpw(np.zeros(40), np.ones(40)) == 0
    with pytest.raises(ValueError, match="window size"):
        st.derive(np.ones(20))
    with pytest.raises(ValueError):
        st.Rpw(np.ones((2, 40)), np.ones((2, 40)))
    with pytest.raises(ValueError):
        st.Rp(np.ones(10), np.ones(11))


def test_remaining_leaf_helpers(calc, golden):
    """get_T (goniometer.py:20-25), get_structure_factors (phases.py:20-39), get_absolute_scale (phases.py:48-66)"""
    g = golden("kat")
    theta = g["kat_theta"]
    T = calc.goniometer.get_T(theta, 12, 2.3, 2.3)
    lpf = calc.goniometer.get_lorentz_polarisation_factor(theta, 12, 2.3, 2.3, 0)
    assert rel_err(T * (1.0 + np.cos(2.0 * theta) ** 2) / np.sin(theta), lpf) < 1e-14
    assert rel_err(T, orc.get_T(theta, 12, 2.3, 2.3)) < 1e-13
    mix = syn.CONFIGS["c2"].truth()
    ph = mix.specimens[0].phases[0][0]
    stl = 2 * np.sin(mix.specimens[0].range_theta[::50]) / mix.specimens[0].goniometer.wavelength
    SF, PF = calc.phases.get_structure_factors(stl, ph.G, ph.components)
    SFo, PFo = orc.get_structure_factors(stl, ph.G, copy.deepcopy(ph.components))
    assert SF.shape == (stl.size, 2) and np.max(np.abs(SF - SFo)) < 1e-11 * np.max(np.abs(SFo))
    assert np.max(np.abs(PF - PFo)) < 1e-13
    assert calc.phases.get_absolute_scale(ph.components, 9.7, ph.W) == orc.get_absolute_scale(ph.components, 9.7, ph.W)
    assert calc.phases.get_absolute_scale([None, None], 9.7, ph.W) == 0.0


def test_rank_two_series_with_nearly_degenerate_eigenvalues(calc):
    """the rank-2 series runs as a Clenshaw recurrence on (trace, determinant) of Q; segregated (P11, P22 -> 1) and
    nearly periodic (P11, P22 -> 0) stackings put both eigenvalues of Q close to the unit circle / to each other, long
    CSDS distributions (N = 120) give the recurrence time to amplify rounding errors"""
    types = syn.atom_types()
    theta = syn.theta_grid(3.0, 50.0, 0.02)
    phases = []
    for w1, p11 in ((0.5, 0.999), (0.5, 0.9999999), (0.5, 1e-4), (0.9, 0.999), (0.7, 0.71), (0.5, 0.5), (0.999, 0.9995)):
        w2 = 1.0 - w1
        p12 = 1.0 - p11
        p21 = w1 * p12 / w2
        P = np.array([[p11, p12], [p21, 1.0 - p21]])
        comps = [syn.make_component(types, "illite"), syn.make_component(types, "smectite_ad", delta_c=0.0)]
        phases.append(syn.make_phase(comps, np.diag([w1, w2]), P, syn.make_csds(48.0, maximum=120), sigma_star=6.0))
    # the same with identical components: phi_0 = phi_1, Q has the eigenvalues phi and phi (p11 + p22 - 1)
    comps = [syn.make_component(types, "illite"), syn.make_component(types, "illite")]
    phases.append(syn.make_phase(comps, np.diag([0.5, 0.5]), np.array([[0.999, 0.001], [0.001, 0.999]]),
                                 syn.make_csds(48.0, maximum=120), sigma_star=6.0))
    spec = syn.make_specimen(theta, phases, syn.make_goniometer("single"))
    mix = _with_observed(syn.make_mixture([spec], len(phases)))
    _oracle_vs_device(calc, mix)
