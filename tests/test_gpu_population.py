"""Population path on the device through the C-ABI:
probability models (``k_probabilities`` / ``pyxrd_leaf_probabilities``) against the live reference's model
classes, and template + binding-table populations against the oracle run on separately rebuilt DataObject
graphs.  Tolerances: W / P bit-identical; intensities / residuals relative 1e-9 (north_star)."""
import copy

import numpy as np
import pytest

from helpers import rel_err
from oracle import pyxrd_oracle as orc
from oracle import pyxrd_oracle_models as om
from pyxrd_b200 import synthetic as syn
from pyxrd_b200.population import PopulationTemplate, Refinable

pytestmark = pytest.mark.gpu

RTOL = 1e-9
OPT_RTOL = 5e-3
CASES = [(1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (1, 6), (2, 2), (3, 2), (4, 2), (5, 3), (6, 4), (7, 3)]


@pytest.fixture(scope="module")
def ctx():
    from pyxrd_b200.runtime import default_context
    return default_context()


@pytest.mark.parametrize("model,G", CASES)
def test_device_models_bit_identical_to_live_reference(ctx, golden, model, G):
    g = golden("models")
    key = "m%d_g%d" % (model, G)
    for p, W, P, v in zip(g[key + "_params"], g[key + "_W"], g[key + "_P"], g[key + "_valid"]):
        gW, gP, gv = ctx.leaf_probabilities(model, G, p)
        assert np.array_equal(gW, W, equal_nan=True), (key, p)
        assert np.array_equal(gP, P, equal_nan=True), (key, p)
        assert gv == bool(v), (key, p)


def test_leaf_probabilities_rejects_unknown_models(ctx):
    from pyxrd_b200.runtime import DeviceError
    with pytest.raises(DeviceError):
        ctx.leaf_probabilities(2, 3, [0.5, 0.5])           # R1 exists for G = 2 only here
    with pytest.raises(DeviceError):
        ctx.leaf_probabilities(9, 2, [0.5, 0.5])


def _observe(mix, level=50.0):
    rng = np.random.default_rng(3)
    for spec in mix.specimens:
        spec.observed_intensity = rng.uniform(0.5, 1.5, (spec.range_theta.size, 1)) * level
    return mix


def _template(name):
    mix, refs = syn.template(name)
    _observe(mix)
    return PopulationTemplate(mix, refs), refs


@pytest.mark.parametrize("name,K", [("c1", 6), ("c2", 6), ("c4", 5), ("c5", 3), ("models", 12)])
def test_population_matches_oracle_on_rebuilt_graphs(ctx, name, K):
    import os
    K = int(os.environ.get("PYXRD_FUZZ_COUNT", K))
    tmpl, refs = _template(name)
    x = syn.sample_solutions(refs, int(os.environ.get("PYXRD_FUZZ_SEED", "5")), K)
    res = tmpl.residuals(x)
    pis = tmpl.phase_intensities(x)
    assert res.shape == (K, 1 + tmpl.static.S)
    for k in range(K):
        tmpl.apply(x[k], probabilities=om.probabilities)
        ref = copy.deepcopy(tmpl.mixture)
        orc.calculate_mixture(ref)
        for s, spec in enumerate(ref.specimens):
            assert rel_err(pis[k][s], spec.phase_intensities) < RTOL, (name, k, s)
        assert rel_err(res[k], ref.residuals) < RTOL, (name, k)
    if name == "models":         # both verdicts of validate() occur in the sample
        valid = [[om.probabilities(om.PROB_CODES[p.prob_model], p.G, p.prob_params)[2] for p in
                  tmpl.apply(row, probabilities=om.probabilities).specimens[0].phases[0]] for row in x]
        assert np.any(valid) and not np.all(valid)


def test_population_equals_the_batch_path_bit_for_bit(ctx):
    """same kernels, same packed values: a template population and K packed DataObject graphs agree exactly"""
    from pyxrd_b200.batched import population_residuals
    tmpl, refs = _template("c4")
    x = syn.sample_solutions(refs, 6, 8)
    res = tmpl.residuals(x)
    graphs = []
    for row in x:
        tmpl.apply(row, ctx=ctx)                  # W / P from the device leaf
        graphs.append(copy.deepcopy(tmpl.mixture))
    for g in graphs:                              # hand the matrices over as fixed ones
        for spec in g.specimens:
            for ph in spec.phases[0]:
                ph.prob_model = None
    assert np.array_equal(res, population_residuals(graphs))


def test_invalid_probabilities_inside_a_population(ctx):
    tmpl, refs = _template("c4")
    names = [r.name for r in refs]
    x = syn.sample_solutions(refs, 8, 4)
    # R2G2 with W1 = 0.5, P21 = 0.1, P112 = 0.5: P211 = P112 * W11 / W21 = 4.5 leaves [0, 1]
    x[1, names.index("IS_R2.prob0")] = 0.5
    x[1, names.index("IS_R2.prob1")] = 0.5
    x[1, names.index("IS_R2.prob2")] = 0.1
    x[1, names.index("IS_R2.prob3")] = 0.9
    assert not om.probabilities(om.R2G2, 2, x[1, names.index("IS_R2.prob0"):names.index("IS_R2.prob0") + 4])[2]
    pis = tmpl.phase_intensities(x)
    res = tmpl.residuals(x)
    assert not pis[1][0][5].any() and not pis[1][1][5].any()        # zeros (phases.py:109-111)
    assert pis[0][0][5].any() and pis[1][0][4].any()
    tmpl.apply(x[1], probabilities=om.probabilities)
    assert not tmpl.mixture.specimens[0].phases[0][5].valid_probs
    ref = copy.deepcopy(tmpl.mixture)
    orc.calculate_mixture(ref)
    assert rel_err(res[1], ref.residuals) < RTOL


def test_out_of_range_parameters_are_clamped_like_the_setters(ctx):
    tmpl, refs = _template("c2")
    names = [r.name for r in refs]
    x = syn.sample_solutions(refs, 9, 2)
    x[0, names.index("IS.prob1")] = 1.3           # clamped to 1.0 (cast_property.py:44-47)
    x[1, names.index("IS.prob1")] = 1.0
    x[1, [i for i in range(len(names)) if i != names.index("IS.prob1")]] = \
        x[0, [i for i in range(len(names)) if i != names.index("IS.prob1")]]
    res = tmpl.residuals(x)
    assert np.array_equal(res[0], res[1])


def test_population_optimiser(ctx):
    """get_optimized_residual semantics for a whole population: equal to the batch path's device optimiser"""
    from pyxrd_b200.batched import get_optimized_residuals
    tmpl, refs = _template("c4")
    x = syn.sample_solutions(refs, 10, 6)
    got, sol = tmpl.optimized_residuals(x, want_solutions=True)
    graphs = []
    for row in x:
        tmpl.apply(row, ctx=ctx)
        g = copy.deepcopy(tmpl.mixture)
        for spec in g.specimens:
            for ph in spec.phases[0]:
                ph.prob_model = None
        graphs.append(g)
    want = np.array([r[0] for r in get_optimized_residuals(graphs)])
    assert np.array_equal(got, want)
    M = tmpl.batch.M
    assert np.all(np.abs(sol[:, :M].sum(axis=1) - 1.0) < 1e-12) and np.all(sol[:, :M] >= 0)
    # against the oracle's scipy-driven optimiser on one rebuilt graph (one-sided contract of DESIGN.md §4)
    ref = copy.deepcopy(graphs[0])
    ref.optimized = False
    want0 = float(orc.get_optimized_residual(ref)[0])
    assert got[0] <= want0 * (1.0 + OPT_RTOL), (got[0], want0)


def test_full_size_population_properties(ctx):
    """c2 at bench size (2048 candidates x 6700 points): permutation of the solution rows permutes the
    residuals bit for bit; duplicated rows give duplicated residuals; everything finite"""
    tmpl, refs = _template("c2")
    x = syn.sample_solutions(refs, 12, 2048)
    x[1000] = x[3]
    res = tmpl.residuals(x)
    assert np.isfinite(res).all() and (res >= 0).all()
    assert np.array_equal(res[1000], res[3])
    perm = np.random.default_rng(0).permutation(len(x))
    assert np.array_equal(tmpl.residuals(x[perm]), res[perm])


def test_binding_to_fixed_matrices_and_mixture_values(ctx):
    mix = _observe(syn.CONFIGS["c4"].truth())
    ph_ad, ph_eg = mix.specimens[0].phases[0][3], mix.specimens[1].phases[0][3]
    refs = [Refinable("f0", [(mix, "fractions:0")]), Refinable("s1", [(mix, "scales:1")]),
            Refinable("b0", [(mix, "bgshifts:0")]),
            Refinable("p01", [(ph_ad, "P:0:1"), (ph_ad, "P:0:0", -1.0, 1.0), (ph_eg, "P:0:1"), (ph_eg, "P:0:0", -1.0, 1.0)])]
    tmpl = PopulationTemplate(mix, refs, csds_auto_maximum=False)
    x = np.array([[0.2, 3.0, 1.5, 0.25], [0.4, 0.5, 0.0, 0.5]])
    res = tmpl.residuals(x)
    for k in range(2):
        tmpl.apply(x[k], probabilities=om.probabilities)
        ref = copy.deepcopy(tmpl.mixture)
        orc.calculate_mixture(ref)
        assert rel_err(res[k], ref.residuals) < RTOL


@pytest.mark.parametrize("name,K", [("c4", 7), ("models", 9), ("c5", 4)])
def test_device_expansion_equals_host_expansion(ctx, name, K):
    """pyxrd_set_population expands on the device (k_expand_population); pyxrd_expand_population is the same
    expansion on the host (checked against packing on the CPU): identical batches, identical bits out"""
    import types
    tmpl, refs = _template(name)
    x = syn.sample_solutions(refs, 21, K)
    with ctx.lock:
        ctx.set_static(tmpl.static)
        ctx.set_population(tmpl.batch, x, tmpl.bindings, tmpl.csds_auto, tmpl.csds_factor)
        ctx.compute_intensities()
        ctx.compute_residuals()
        dev_int, dev_res = ctx.read_intensities_flat(), ctx.read_residuals()
        host = types.SimpleNamespace(**tmpl.expand(x))
        host.S, host.Z = tmpl.batch.S, tmpl.batch.Z
        ctx.set_batch(host)
        ctx.compute_intensities()
        ctx.compute_residuals()
        assert np.array_equal(dev_int, ctx.read_intensities_flat())
        assert np.array_equal(dev_res, ctx.read_residuals())


def test_multiply_and_add_rows_on_the_device(ctx):
    """slots driven by two refinables (a product: cell volume; a sum: component weight): the device expansion applies
    the rows in table order like the host expansion, and the result is what the model layer packs"""
    import types
    from pyxrd_b200 import packing
    from pyxrd_b200.population import discover_template
    from test_binding_golden import _toy_model_layer
    build, _ = _toy_model_layer(coupled="product")
    x0 = np.array([6.0, 0.6, 10.0, 1.25, 0.3])
    bounds = np.array([[2.0, 12.0], [0.3, 0.9], [8.0, 14.0], [1.22, 1.28], [0.05, 0.4]])
    tmpl = discover_template(build, x0, bounds, verify=2)
    assert {1, 2} <= set(int(m) for m in tmpl.bindings["mode"])
    x = bounds[:, 0] + np.random.default_rng(5).uniform(0, 1, (6, 5)) * (bounds[:, 1] - bounds[:, 0])
    with ctx.lock:
        ctx.set_static(tmpl.static)
        ctx.set_population(tmpl.batch, x, tmpl.bindings, tmpl.csds_auto, tmpl.csds_factor)
        ctx.compute_intensities()
        dev_int = ctx.read_intensities_flat()
        host = types.SimpleNamespace(**tmpl.expand(x))
        host.S, host.Z = tmpl.batch.S, tmpl.batch.Z
        ctx.set_batch(host)
        ctx.compute_intensities()
        assert np.array_equal(dev_int, ctx.read_intensities_flat())
        want = packing.BatchPack([build(row) for row in x], tmpl.static)
        ctx.set_batch(want)
        ctx.compute_intensities()
        assert rel_err(dev_int, ctx.read_intensities_flat()) < 1e-12
    bad = tmpl.bindings.copy()
    bad["mode"][0] = 3
    from pyxrd_b200.runtime import DeviceError
    with pytest.raises(DeviceError, match="out of range"):
        tmpl2 = copy.copy(tmpl)
        tmpl2.bindings = bad
        tmpl2.residuals(x)


def test_population_errors(ctx):
    from pyxrd_b200.runtime import DeviceError
    tmpl, refs = _template("c2")
    x = syn.sample_solutions(refs, 3, 4)
    bad = tmpl.bindings.copy()
    bad["index"][0] = 10 ** 6
    with ctx.lock:
        ctx.set_static(tmpl.static)
        with pytest.raises(DeviceError, match="out of range"):
            ctx.evaluate_population(tmpl.batch, x, bad, tmpl.csds_auto, tmpl.csds_factor)
        xs = x.copy()
        xs[2, [r.name for r in refs].index("IS.csds_average")] = 0.1      # maximum = int(0.25) = 0 < minimum = 1
        with pytest.raises(DeviceError, match="CSDS range"):
            ctx.evaluate_population(tmpl.batch, xs, tmpl.bindings, tmpl.csds_auto, tmpl.csds_factor)
        assert np.isfinite(ctx.evaluate_population(tmpl.batch, x, tmpl.bindings, tmpl.csds_auto, tmpl.csds_factor)).all()


def test_batched_lbfgsb_equals_trial_by_trial_lbfgsb(ctx):
    """the L BFGS B run (scipy_runs.py:39-48) with every gradient's d + 1 probes in one device batch takes the path
    of the trial-by-trial run: the device results do not depend on how trials are grouped"""
    from scipy.optimize import fmin_l_bfgs_b
    from pyxrd_b200 import drivers
    tmpl, refs = _template("c2")
    x0 = syn.sample_solutions(refs, 31, 1)[0]
    sol, res, info = drivers.lbfgsb(tmpl, x0, maxiter=4)
    one = lambda x: float(tmpl.optimized_residuals(np.asarray(x)[None, :])[0])          # noqa: E731
    want_sol, want_res, want_info = fmin_l_bfgs_b(one, x0, approx_grad=True, epsilon=1e-4, maxiter=4,
                                                  bounds=[tuple(b) for b in tmpl.bounds()])
    assert np.array_equal(sol, want_sol) and res == want_res and info["nit"] == want_info["nit"]
    assert info["trials"] == want_info["funcalls"] and info["batches"] * (len(refs) + 1) == info["trials"]
    assert res < one(x0)


def test_brute_force_generation(ctx):
    from pyxrd_b200 import drivers
    tmpl, refs = _template("c1")
    best, value, grid, values = drivers.brute_force(tmpl, num_samples=5)
    assert grid.shape == (3 * 25, 3) and np.isfinite(values).all()
    k = int(np.argmin(values))
    tmpl.apply(grid[k], probabilities=om.probabilities)
    ref = copy.deepcopy(tmpl.mixture)
    want = float(orc.get_optimized_residual(ref)[0])
    assert value <= want * (1.0 + OPT_RTOL)
    assert np.array_equal(drivers.evaluate_generation(tmpl, grid[:7], optimize=False), tmpl.residuals(grid[:7])[:, 0])


def test_more_than_65535_patterns_in_one_batch(ctx):
    """grid.y stops at 65535: candidates / patterns ride on grid.x in every kernel (70 000 one-phase candidates, and a
    6-phase x 2-specimen mixture with 6 000 candidates = 72 000 patterns)"""
    mix, refs = syn.template("c1")
    spec = mix.specimens[0]
    theta = spec.range_theta[::30]                                   # 70 points
    mix.specimens[0] = syn.make_specimen(theta, spec.phases[0], spec.goniometer)
    mix.specimens[0].goniometer.wavelength_distribution = list(syn.CU_PAIR)
    _observe(mix)
    tmpl = PopulationTemplate(mix, refs)
    x = syn.sample_solutions(refs, 41, 70000)
    res = tmpl.residuals(x)
    assert res.shape == (70000, 2) and np.isfinite(res).all()
    pick = np.array([0, 1, 65534, 65535, 65536, 69999])
    assert np.array_equal(res[pick], tmpl.residuals(x[pick]))
    opt = tmpl.optimized_residuals(x)
    assert np.isfinite(opt).all() and np.array_equal(opt[pick], tmpl.optimized_residuals(x[pick]))
    mix, refs = syn.template("c4")
    for s, spec in enumerate(mix.specimens):
        mix.specimens[s] = syn.make_specimen(spec.range_theta[::30], spec.phases[0], spec.goniometer)
    _observe(mix)
    tmpl = PopulationTemplate(mix, refs)
    x = syn.sample_solutions(refs, 42, 6000)
    res = tmpl.residuals(x)
    pick = np.array([0, 5461, 5462, 5999])
    assert np.isfinite(res).all() and np.array_equal(res[pick], tmpl.residuals(x[pick]))


def test_table_discovered_from_the_reference_model_layer(ctx):
    """tests/golden/binding.*: the table was discovered by probing the reference's RUNNING model layer (its integration
    fixture, 40 refinables, one slot bound as a product, one as a sum of three); here the device evaluates the 12 stored solutions from the table
    alone and must agree with
    what the reference computed for the graphs its MVC built: calculate_mixture residuals at 1e-9, get_optimized_residual
    (= Refiner.get_residual) within the optimiser contract"""
    from test_binding_golden import reference_binding
    tmpl, g = reference_binding()
    x = g["x"]
    res = tmpl.residuals(x)
    assert rel_err(res, g["residuals"]) < RTOL
    opt = tmpl.optimized_residuals(x)
    assert np.all(opt <= g["optimised"] * (1.0 + OPT_RTOL) + 1e-9), (opt, g["optimised"])
    # W, P, valid_probs of the model-driven phases as the model layer had them
    pis = tmpl.phase_intensities(x[:3])
    for k in range(3):
        for s in range(len(pis[k])):
            for m in range(pis[k][s].shape[0]):
                assert bool(pis[k][s][m].any()) == bool(g["valid_probs"][k][s][m])


def test_deap_shaped_generation_end_to_end():
    """A generation of a population method, evaluated the way the reference's DEAP runs ask for it
    (refinement/methods/deap_cma.py:345-360 ``_ask`` -> ``RefineAsyncHelper.do_async_evaluation(toolbox.generate,
    result_func=...)``, refine_async_helper.py:16-23, async_evaluatable.py:38-50) -- through the installed hook with a
    REAL device evaluation of a 256-row population, and through the batching provider with the default evaluator.
    No ``deap`` needed: the hook only sees an ``iter_func`` that yields individuals."""
    import functools
    import pyxrd_b200.calculations as calc
    from pyxrd_b200 import install
    from pyxrd_b200.provider import B200AsyncServer
    tmix, refs = syn.template("c4")
    truth = syn.CONFIGS["c4"].truth()

    def calc_total(mix):
        calc.mixture.calculate_mixture(mix)
        return [np.array(s.total_intensity) for s in mix.specimens]

    syn.attach_observed([tmix], truth, calc_total)
    tmpl = PopulationTemplate(tmix, refs)
    x = syn.sample_solutions(refs, 31, 256)

    class Individual(list):                       # what deap's creator.Individual is: a list with a fitness slot
        fitness = None

    class Refiner(object):
        def __init__(self):
            self._b200_template = tmpl             # what install._refiner_template caches after discovering the table
            self.seen = []

        def update(self, solution, residual=None):
            self.seen.append((list(solution), residual))

    class Run(object):                            # the part of RefineCMAESRun the hook touches
        def __init__(self):
            self.refiner = Refiner()

        def _user_cancelled(self):
            return False

    def must_not_run(*args, **kwargs):            # pragma: no cover
        raise AssertionError("the reference's per-trial loop must not run")

    hook = install._generation_evaluator(must_not_run)
    run = Run()
    reported = []

    def result_f(*args):                          # deap_cma.py:349-351
        run.refiner.update(*args)
        reported.append(args)

    population = hook(run, lambda: (Individual(row) for row in x), result_func=result_f)
    assert len(population) == 256 and all(isinstance(ind, Individual) for ind in population)
    assert len(reported) == 256 and len(run.refiner.seen) == 256
    values = np.array([float(r[1][0]) for r in reported])
    assert all(np.shape(r[1]) == (1,) for r in reported)                  # get_optimized_residual returns a 1-element array
    assert np.array_equal(values, tmpl.optimized_residuals(x))            # the same launch sequence, deterministic
    # a few rows against the oracle's get_optimized_residual on graphs rebuilt per candidate (optimiser contract)
    for k in (0, 97, 255):
        tmpl.apply(x[k], probabilities=om.probabilities)
        ref = copy.deepcopy(tmpl.mixture)
        want = float(orc.get_optimized_residual(ref)[0])
        assert values[k] <= want * (1.0 + 5e-3), (k, values[k], want)
        assert values[k] >= want * 0.5
    # the provider route (settings.ASYNC_SERER_PROVIDERS): submit all trials, fetch all -> one device batch
    server = B200AsyncServer()
    graphs = []
    for k in range(24):
        tmpl.apply(x[k], probabilities=om.probabilities)
        graphs.append(copy.deepcopy(tmpl.mixture))
    results = [server.submit(functools.partial(calc.mixture.get_optimized_residual, g)) for g in graphs]
    fetched = np.array([float(r.get()[0]) for r in results])
    assert np.allclose(fetched, values[:24], rtol=1e-9, atol=0.0)         # packed graphs == template expansion
    assert all(g.optimized for g in graphs)
