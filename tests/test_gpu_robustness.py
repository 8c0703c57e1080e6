"""State, threading and process conventions of the drop-in functions on the device: residency of a
parsed mixture across interleaved calls, two threads on one context, a child forked after CUDA was initialised, a
specimen without observations in the batched calls, stale static tables."""
import copy
import multiprocessing as mp
import threading

import numpy as np
import pytest

from helpers import rel_err
from oracle import pyxrd_oracle as orc
from pyxrd_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def calc():
    import pyxrd_b200.calculations as c
    return c


def _observed(mix, seed=1):
    rng = np.random.default_rng(seed)
    for spec in mix.specimens:
        spec.observed_intensity = rng.uniform(20.0, 60.0, (spec.range_theta.size, 1))
    return mix


def test_interleaved_calls_do_not_reuse_foreign_device_state(calc):
    """parse_mixture(mix) leaves the patterns resident; another entry point (get_intensity,
    calculate_phase_intensities, a population call, another mixture) then replaces the context's batch.  The later
    calculate_mixture / optimize_mixture of the first mixture must notice and recompute, not use foreign patterns."""
    from pyxrd_b200.batched import population_residuals
    mix = _observed(syn.CONFIGS["c4"].candidate(2024, 3))
    ref = copy.deepcopy(mix)
    orc.calculate_mixture(ref)
    calc.mixture.parse_mixture(mix)
    other = _observed(syn.CONFIGS["c4"].candidate(2024, 4), seed=1)        # same shapes (K, M, S agree), other patterns
    spec = other.specimens[0]
    phase = spec.phases[0][3]
    stl = 2.0 * np.sin(spec.range_theta) / spec.goniometer.wavelength
    calc.phases.get_intensity(spec.range_theta, stl, 2.3, 2.3, 0.0, phase)
    calc.mixture.calculate_mixture(mix)
    assert rel_err(mix.residuals, ref.residuals) < RTOL
    for a, b in zip(mix.specimens, ref.specimens):
        assert rel_err(a.total_intensity, b.total_intensity) < RTOL
    # a whole other batch of the same shape in between
    mix.calculated = False
    population_residuals([other])
    calc.mixture.calculate_mixture(mix)
    assert rel_err(mix.residuals, ref.residuals) < RTOL
    # specimen-level entry point in between, then the optimiser
    mix.calculated = False
    calc.specimen.calculate_phase_intensities(other.specimens[1])
    calc.mixture.calculate_mixture(mix)
    assert rel_err(mix.residuals, ref.residuals) < RTOL
    # a parsed mixture that went through pickle (a worker of another process) never matches a resident token
    import pickle
    clone = pickle.loads(pickle.dumps(mix))
    clone.calculated = False
    calc.mixture.calculate_mixture(clone)
    assert rel_err(clone.residuals, ref.residuals) < RTOL


def test_two_threads_share_one_context(calc):
    """the functions are called from the GUI thread and a refinement thread at once; the context
    serialises them (per-context lock) and each caller gets its own results"""
    mixes = [_observed(syn.CONFIGS[name].candidate(2024, i), seed=i) for i, name in enumerate(("c1", "c4", "c1", "c4"))]
    refs = []
    for m in mixes:
        r = copy.deepcopy(m)
        orc.calculate_mixture(r)
        refs.append(r)
    errors, worst = [], [0.0] * len(mixes)

    def worker(i):
        try:
            for _ in range(6):
                m = copy.deepcopy(mixes[i])
                calc.mixture.calculate_mixture(m)
                worst[i] = max(worst[i], rel_err(m.residuals, refs[i].residuals),
                               max(rel_err(a.total_intensity, b.total_intensity) for a, b in zip(m.specimens, refs[i].specimens)))
                calc.mixture.get_optimized_residual(copy.deepcopy(mixes[i]))
        except Exception as err:        # pragma: no cover
            errors.append(err)

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(len(mixes))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert max(worst) < RTOL, worst


def _child(queue):
    import pyxrd_b200.calculations as c
    from pyxrd_b200.runtime import DeviceError
    try:
        c.mixture.calculate_mixture(_observed(syn.CONFIGS["c1"].truth()))
        queue.put(("computed", ""))
    except DeviceError as err:
        queue.put(("DeviceError", str(err)))
    except Exception as err:            # pragma: no cover
        queue.put((type(err).__name__, str(err)))


def test_fork_after_cuda_initialisation(calc):
    """the reference's Pool workers are forked (pyxrd_server.py:57-64): a child forked after the parent touched CUDA
    gets a DeviceError that names the provider / spawn, and the parent keeps working"""
    calc.mixture.calculate_mixture(_observed(syn.CONFIGS["c1"].truth()))          # the parent has a context now
    ctx = mp.get_context("fork")
    queue = ctx.Queue()
    proc = ctx.Process(target=_child, args=(queue,))
    proc.start()
    kind, text = queue.get(timeout=120)
    proc.join(timeout=120)
    assert kind == "DeviceError" and "forked" in text and "B200AsyncServerProvider" in text, (kind, text)
    mix = _observed(syn.CONFIGS["c1"].truth())
    ref = copy.deepcopy(mix)
    orc.calculate_mixture(ref)
    calc.mixture.calculate_mixture(mix)
    assert rel_err(mix.residuals, ref.residuals) < RTOL


def test_spawned_worker_computes(calc):
    """... while a SPAWNED worker (fresh interpreter) creates its own context and computes"""
    ctx = mp.get_context("spawn")
    queue = ctx.Queue()
    proc = ctx.Process(target=_child, args=(queue,))
    proc.start()
    kind, text = queue.get(timeout=300)
    proc.join(timeout=120)
    assert kind == "computed", (kind, text)


def test_specimen_without_observations_in_the_batched_calls(calc):
    """a specimen without observed data reports 0.0 (mixture.py:247-251) in every call that has
    calculate_mixture semantics, the batched ones included (the device knows the specimen has no observations)"""
    from pyxrd_b200.batched import population_residuals
    from pyxrd_b200.population import PopulationTemplate
    mixes = [_observed(syn.CONFIGS["c4"].candidate(2024, i)) for i in range(3)]
    for m in mixes:
        m.specimens[1].observed_intensity = np.zeros((0, 1))
    res = population_residuals(mixes)
    assert np.isfinite(res).all() and np.array_equal(res[:, 2], np.zeros(3))
    for k, m in enumerate(mixes):
        ref = copy.deepcopy(m)
        orc.calculate_mixture(ref)
        assert rel_err(res[k], ref.residuals) < RTOL
    tmix, refs = syn.template("c4")
    _observed(tmix)
    tmix.specimens[0].observed_intensity = np.zeros((0, 1))
    tmpl = PopulationTemplate(tmix, refs)
    res = tmpl.residuals(syn.sample_solutions(refs, 5, 4))
    assert np.isfinite(res).all() and np.array_equal(res[:, 1], np.zeros(4)) and np.all(res[:, 2] > 0.0)
    assert np.allclose(res[:, 0], 0.5 * res[:, 2], rtol=1e-15)


def test_static_tables_follow_a_changed_secondary_wavelength(calc):
    """two goniometers that differ only in a SECONDARY wavelength of the distribution (same
    primary wavelength, same fractions) must not share resident interpolation tables"""
    mix = _observed(syn.CONFIGS["c2"].candidate(2024, 0))
    calc.mixture.calculate_mixture(mix)
    changed = copy.deepcopy(mix)
    changed.parsed = changed.calculated = False
    (w1, f1), (w2, f2) = changed.specimens[0].goniometer.wavelength_distribution
    changed.specimens[0].goniometer.wavelength_distribution = [(w1, f1), (w2 * 1.004, f2)]
    ref = copy.deepcopy(changed)
    orc.calculate_mixture(ref)
    calc.mixture.calculate_mixture(changed)
    assert rel_err(changed.specimens[0].phase_intensities, ref.specimens[0].phase_intensities) < RTOL
    assert rel_err(changed.residuals, ref.residuals) < RTOL
    assert rel_err(changed.specimens[0].phase_intensities, mix.specimens[0].phase_intensities) > 1e-6


def test_candidates_with_other_observations_are_refused(calc):
    """... and candidates whose observed data / selected range differ from the batch's first candidate are refused
    (same end points and sizes are not enough)"""
    from pyxrd_b200.batched import pack_population
    mixes = [_observed(syn.CONFIGS["c1"].candidate(2024, i)) for i in range(2)]
    pack_population(mixes)
    mixes[1].specimens[0].observed_intensity = mixes[1].specimens[0].observed_intensity * 1.01
    with pytest.raises(ValueError):
        pack_population(mixes)
    mixes[1].specimens[0].observed_intensity = mixes[0].specimens[0].observed_intensity.copy()
    sel = np.ones(mixes[1].specimens[0].range_theta.size, dtype=bool)
    sel[100:200] = False
    mixes[1].specimens[0].selected_range = sel
    with pytest.raises(ValueError):
        pack_population(mixes)


def test_layer_tables_on_and_off_agree(calc):
    """candidate-invariant layer stacks (components.py:15-16 only moves interlayer atoms) come from a table computed once
    per batch; switched off ("layer_tables" 0) every layer is evaluated inside the phase kernels.  Both against the oracle
    at 1e-9, against each other at 1e-12; a population whose refinable drives a LAYER atom keeps that component out of
    the table and still matches graphs rebuilt per candidate."""
    from oracle import pyxrd_oracle_models as om
    from pyxrd_b200.batched import population_phase_intensities
    from pyxrd_b200.population import PopulationTemplate, Refinable
    from pyxrd_b200.runtime import default_context
    ctx = default_context()
    mixes = [_observed(syn.CONFIGS["c4"].candidate(2024, i)) for i in range(4)]
    got = {}
    for flag in (1, 0):
        ctx.set_option("layer_tables", flag)
        try:
            got[flag] = population_phase_intensities([copy.deepcopy(m) for m in mixes])
        finally:
            ctx.set_option("layer_tables", 1)
    for k, m in enumerate(mixes):
        ref = copy.deepcopy(m)
        orc.calculate_mixture(ref)
        for s, spec in enumerate(ref.specimens):
            assert rel_err(got[1][k][s], spec.phase_intensities) < RTOL
            assert rel_err(got[0][k][s], spec.phase_intensities) < RTOL
            assert rel_err(got[1][k][s], got[0][k][s]) < 1e-12
    # a refinable on the content of a LAYER atom of one component, the other components keep their tables
    tmix, refs = syn.template("c2")
    _observed(tmix)
    phase = tmix.specimens[0].phases[0][0]
    refs = list(refs) + [Refinable("illite.Fe", [(phase.components[0].layer_atoms[7], "pn")], 0.1, 0.6)]
    tmpl = PopulationTemplate(tmix, refs)
    x = syn.sample_solutions(refs, 9, 5)
    res = tmpl.residuals(x)
    pats = tmpl.phase_intensities(x)
    for k in range(len(x)):
        tmpl.apply(x[k], probabilities=om.probabilities)
        ref = copy.deepcopy(tmpl.mixture)
        orc.calculate_mixture(ref)
        assert rel_err(pats[k][0], ref.specimens[0].phase_intensities) < RTOL
        assert rel_err(res[k], ref.residuals) < RTOL


def test_mmult_and_q_matrices_match_numpy(calc):
    """a7: ``mmult`` (math_tools.py:16-17) and ``get_Q_matrices`` (phases.py:41-46) as device leaves, against the
    reference's own numpy expressions (restated here: broadcast-and-sum over the contraction index).  1e-13 of the
    largest entry: numpy's complex product is not bit-portable (its SIMD loops fuse one of the two products)"""
    rng = np.random.default_rng(8)

    def np_mmult(A, B):
        return np.sum(np.transpose(A, (0, 2, 1))[:, :, :, np.newaxis] * B[:, :, np.newaxis, :], -3)

    for X, r in ((37, 1), (50, 2), (21, 3), (9, 8), (3, 27)):
        A = rng.standard_normal((X, r, r)) + 1j * rng.standard_normal((X, r, r))
        B = rng.standard_normal((X, r, r)) + 1j * rng.standard_normal((X, r, r))
        want = np_mmult(A, B)
        assert np.max(np.abs(calc.math_tools.mmult(A, B) - want)) <= 1e-13 * np.max(np.abs(want)), (X, r)
        Q = 0.6 * A / r
        want = np.zeros((6,) + Q.shape, dtype=complex)
        want[0] = Q
        for n in range(1, 6):
            want[n] = np_mmult(want[n - 1], Q)
        got = calc.phases.get_Q_matrices(Q, 5)
        assert got.shape == want.shape, (X, r)
        for n in range(6):
            assert np.max(np.abs(got[n] - want[n])) <= 1e-13 * np.max(np.abs(want[n])), (X, r, n)
    with pytest.raises(ValueError):
        calc.math_tools.mmult(np.zeros((2, 3, 3)), np.zeros((2, 2, 2)))


def test_generations_of_one_population_reuse_the_resident_template(calc):
    """``pyxrd_update_population``: later generations of the same template only upload x.  Same results, bit for bit, as
    a fresh upload; a changed population size, another template or any other call in between falls back to the full
    upload; CSDS maxima beyond the first generation's (longer series, larger coefficient tables) are handled."""
    from pyxrd_b200.population import PopulationTemplate
    from pyxrd_b200.runtime import default_context
    ctx = default_context()
    tmix, refs = syn.template("c4")
    _observed(tmix)
    tmpl = PopulationTemplate(tmix, refs)
    omix, orefs = syn.template("c4")
    other = PopulationTemplate(_observed(omix), orefs)                  # same content, another template object
    lo = np.array([r.minimum for r in refs])
    hi = np.array([r.maximum for r in refs])
    rng = np.random.default_rng(12)

    def fresh(x, optimised):
        ctx.invalidate_residency()
        return tmpl.optimized_residuals(x) if optimised else tmpl.residuals(x)

    x1 = lo + rng.uniform(0.0, 0.5, (6, lo.size)) * (hi - lo)          # small CSDS averages first
    x2 = lo + rng.uniform(0.5, 1.0, (6, lo.size)) * (hi - lo)          # then larger ones: longer series than resident
    first = tmpl.residuals(x1)
    assert ctx.population_is_resident(tmpl._key(), x1.shape)
    second = tmpl.residuals(x2)                                         # update path
    assert ctx.population_is_resident(tmpl._key(), x2.shape)
    assert np.array_equal(second, fresh(x2, False)) and np.array_equal(first, fresh(x1, False))
    opt_a = tmpl.optimized_residuals(x1)                                # update path again (after the fresh uploads)
    assert np.array_equal(opt_a, fresh(x1, True))
    x3 = np.concatenate([x1, x2])                                       # another K: full upload
    assert np.array_equal(tmpl.residuals(x3)[:6], first)
    calc.mixture.calculate_mixture(_observed(syn.CONFIGS["c1"].truth()))   # a drop-in call in between drops the population
    assert not ctx.population_is_resident(tmpl._key(), x3.shape)
    assert np.array_equal(tmpl.residuals(x3)[6:], second)
    assert np.array_equal(other.residuals(x3), tmpl.residuals(x3))      # templates do not share residency


def test_layer_tables_with_none_atom_types(calc):
    """a None atom type contributes nothing (atoms.py:65-67), also inside a tabulated layer stack"""
    from pyxrd_b200.batched import population_phase_intensities
    mixes = [_observed(syn.CONFIGS["c2"].candidate(2024, i)) for i in range(3)]
    for m in mixes:
        for comp in m.specimens[0].phases[0][0].components:
            comp.layer_atoms[4].atom_type = None                        # the same layer position in every component
    got = population_phase_intensities([copy.deepcopy(m) for m in mixes])
    for k, m in enumerate(mixes):
        ref = copy.deepcopy(m)
        orc.calculate_mixture(ref)
        assert rel_err(got[k][0], ref.specimens[0].phase_intensities) < RTOL


def test_all_wavelength_passes_inside_the_residual_kernel(calc):
    """pyxrd_compute_all takes every pass of the wavelength distribution inside the residual kernel where the gathers
    stay local (k_patterns_residual): same patterns / totals / residuals, bit for bit, as the pass kernels
    ("fuse_patterns" 0); a residuals-only call does not store the final patterns and says so; a distribution with a
    far-away line (Cu K-beta: the gathers reach thousands of points) keeps the pass kernels and still matches the oracle
    (specimen.py:64-74)."""
    from pyxrd_b200.batched import pack_population
    from pyxrd_b200.runtime import DeviceError, default_context
    ctx = default_context()
    mixes = [_observed(syn.CONFIGS["c4"].candidate(2024, i)) for i in range(11)]      # 11: a block's last group is ragged
    mixes[3].specimens[1].phases[0][1] = None
    static, batch = pack_population(mixes)
    got = {}
    with ctx.lock:
        for flag in (1, 0):
            ctx.set_option("fuse_patterns", flag)
            try:
                ctx.set_static(static, force=True)
                ctx.set_batch(batch)
                ctx.compute_all(True)
                got[flag] = [ctx.read_intensities_flat(), ctx.read_totals_flat(), ctx.read_scaled_flat(), ctx.read_residuals()]
                ctx.set_batch(batch)
                ctx.compute_all(residuals_only=True)
                got[flag].append(ctx.read_residuals())
                if flag:
                    with pytest.raises(DeviceError):
                        ctx.read_intensities_flat()          # the final patterns were not stored
            finally:
                ctx.set_option("fuse_patterns", 1)
    for a, b in zip(got[1], got[0]):
        assert np.array_equal(a, b)
    assert np.array_equal(got[1][3], got[1][4])
    for k in (0, 3, 10):
        ref = copy.deepcopy(mixes[k])
        orc.calculate_mixture(ref)
        assert rel_err(got[1][3][k], ref.residuals) < RTOL
    # K-beta in the list: no halo holds that gather
    wide = [_observed(syn.CONFIGS["c2"].candidate(7, i)) for i in range(3)]
    for m in wide:
        for spec in m.specimens:
            spec.goniometer.wavelength_distribution = [(syn.CU_KA1, 0.9), (0.1392218, 0.1)]
    res = calc.mixture.calculate_mixture
    for m in wide:
        ref = copy.deepcopy(m)
        orc.calculate_mixture(ref)
        res(m)
        assert rel_err(m.residuals, ref.residuals) < RTOL
        assert rel_err(m.specimens[0].phase_intensities, ref.specimens[0].phase_intensities) < RTOL


@pytest.mark.parametrize("lines", [1, 3, 4, 5])
def test_fused_passes_for_every_list_length(calc, lines):
    """k_patterns_residual is instantiated per number of passes (1 .. 4; five entries keep the pass kernels): a list of
    ``lines`` close wavelengths, fused against the pass kernels bit for bit (compute_all and compute_intensities) and against
    the oracle (specimen.py:64-74)."""
    from pyxrd_b200.batched import pack_population
    from pyxrd_b200.runtime import default_context
    ctx = default_context()
    dist = [(0.1544426, 0.60), (0.153475, 0.25), (0.15406, 0.08), (0.15411, 0.04), (0.15390, 0.03)][:lines]
    mixes = [_observed(syn.CONFIGS["c2"].candidate(11, i)) for i in range(5)]
    for m in mixes:
        for spec in m.specimens:
            spec.goniometer.wavelength_distribution = list(dist)
    static, batch = pack_population(mixes)
    got = {}
    with ctx.lock:
        for flag in (1, 0):
            ctx.set_option("fuse_patterns", flag)
            try:
                ctx.set_static(static, force=True)
                ctx.set_batch(batch)
                ctx.compute_all()
                got[flag] = [ctx.read_intensities_flat(), ctx.read_totals_flat(), ctx.read_residuals()]
                ctx.set_batch(batch)
                ctx.compute_intensities()
                got[flag].append(ctx.read_intensities_flat())
            finally:
                ctx.set_option("fuse_patterns", 1)
    for a, b in zip(got[1], got[0]):
        assert np.array_equal(a, b)
    assert np.array_equal(got[1][0], got[1][3])
    for k in (0, 4):
        ref = copy.deepcopy(mixes[k])
        orc.calculate_mixture(ref)
        assert rel_err(got[1][2][k], ref.residuals) < RTOL


def test_fused_passes_on_ragged_shapes(calc):
    """the tile + halo bookkeeping of k_patterns_residual on awkward shapes: pattern lengths around the tile size and below
    the halo, a coarse grid (gathers that reach further), an exclusion window, two patterns per specimen (Z = 2) with a
    None phase in one of them -- fused against the pass kernels bit for bit, and against the oracle"""
    from pyxrd_b200.batched import pack_population
    from pyxrd_b200.runtime import default_context
    ctx = default_context()

    def reshaped(npoints, lo, hi, seed):
        out = []
        for i in range(3):
            m = syn.CONFIGS["c4"].candidate(seed, i)
            for si, spec in enumerate(m.specimens):
                theta = np.radians(np.linspace(lo, hi, npoints + 7 * si) / 2.0)
                new = syn.make_specimen(theta, spec.phases[0], spec.goniometer, exclusion=(12.0, 12.6) if si == 0 else None)
                new.goniometer.wavelength_distribution = list(syn.CU_PAIR)
                m.specimens[si] = new
            out.append(_observed(m, seed=seed))
        return out

    cases = [reshaped(511, 3.0, 45.0, 3), reshaped(513, 3.0, 45.0, 4), reshaped(1030, 20.0, 110.0, 5), reshaped(40, 5.0, 30.0, 6),
             [syn.z2_case(), syn.z2_case()]]
    for mixes in cases:
        static, batch = pack_population(mixes)
        got = {}
        with ctx.lock:
            for flag in (1, 0):
                ctx.set_option("fuse_patterns", flag)
                try:
                    ctx.set_static(static, force=True)
                    ctx.set_batch(batch)
                    ctx.compute_all()
                    got[flag] = [ctx.read_intensities_flat(), ctx.read_totals_flat(), ctx.read_residuals()]
                finally:
                    ctx.set_option("fuse_patterns", 1)
        for a, b in zip(got[1], got[0]):
            assert np.array_equal(a, b)
        ref = copy.deepcopy(mixes[0])
        orc.calculate_mixture(ref)
        assert rel_err(got[1][2][0], ref.residuals) < RTOL
