"""Host-side logic of the drop-in boundary: rebinding of pyxrd.calculations.* and the
batching async provider (CPU; no compute)."""
import functools
import os
import sys

import pytest

from pyxrd_b200.provider import B200AsyncServer, B200AsyncServerProvider

HAVE_REFERENCE = os.path.isdir("/root/reference/pyxrd")


@pytest.fixture(scope="module")
def reference_mixture():
    """the mixture model of the reference's integration fixture inside its running (headless) model layer, with the
    first two refinables left switched on; parsed once per module (a second parse of the same file in one process
    collides in the reference's object pool)"""
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _live_models
    _live_models.load()
    import pyxrd.project.models  # noqa: F401
    from pyxrd.file_parsers.json_parser import JSONParser
    project = JSONParser.parse("/root/reference/test/test_mixture/test refinement.pyxrd")
    mixture = project.mixtures[0]
    n = 0
    for node in mixture.refinement.refinables.iter_children():
        r = node.object
        if r.prop_descr is not None and r.refine:
            n += 1
            if n > 2:
                r.refine = False
    return mixture


def get_optimized_residual(mixture):          # a stand-in with the reference's function name
    return ("single", mixture)


def test_provider_batches_trials_and_runs_other_callables():
    seen = []

    def evaluator(mixtures):
        seen.append(list(mixtures))
        return [("batched", m) for m in mixtures]

    server = B200AsyncServer(batch_evaluator=evaluator)
    results = [server.submit(functools.partial(get_optimized_residual, i)) for i in range(5)]
    other = server.submit(lambda: 42)
    assert seen == []                                  # nothing ran yet: submit-all ...
    assert results[3].get() == ("batched", 3)          # ... then the first fetch runs one batch
    assert seen == [[0, 1, 2, 3, 4]]
    assert [r.get() for r in results] == [("batched", i) for i in range(5)]
    assert other.get() == 42
    later = server.submit(functools.partial(get_optimized_residual, 9))
    assert later.get() == ("batched", 9) and len(seen) == 2


def test_provider_propagates_errors_and_has_reference_interface():
    def evaluator(mixtures):
        raise RuntimeError("device lost")

    server = B200AsyncServer(batch_evaluator=evaluator)
    r = server.submit(functools.partial(get_optimized_residual, 1))
    with pytest.raises(RuntimeError):
        r.get()
    for name in ("get_status", "get_server", "launch_server", "stop_server"):
        assert callable(getattr(B200AsyncServerProvider, name))
    status = B200AsyncServerProvider.get_status()
    assert len(status) == 3
    assert B200AsyncServerProvider.get_server() is B200AsyncServerProvider.get_server()
    B200AsyncServerProvider.stop_server()


@pytest.mark.skipif(not HAVE_REFERENCE, reason="needs the reference tree (build container only)")
def test_install_rebinds_reference_modules():
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _live_reference
    ref = _live_reference.load()
    import pyxrd_b200.calculations as ours
    import pyxrd_b200.install as inst
    original = ref.mixture.get_optimized_residual
    try:
        replaced = inst.install()
        assert ("pyxrd.calculations.mixture", "get_optimized_residual") in replaced
        assert ref.mixture.get_optimized_residual is ours.mixture.get_optimized_residual
        assert ref.phases.get_diffracted_intensity is ours.phases.get_diffracted_intensity
        # import sites inside the reference package itself follow (specimen.py:13 imported get_intensity)
        assert ref.specimen.get_intensity is ours.phases.get_intensity
        assert ref.components.get_structure_factor is ours.atoms.get_structure_factor
    finally:
        inst.uninstall()
    assert ref.mixture.get_optimized_residual is original


@pytest.mark.filterwarnings("ignore")
@pytest.mark.skipif(not HAVE_REFERENCE, reason="needs the reference tree (build container only)")
def test_generations_of_a_reference_refinement_go_through_the_binding_table(monkeypatch, reference_mixture):
    """install(generations=True) inside the live reference: its brute-force run (custom_brute.py:36-56) hands the whole
    grid to ONE PopulationTemplate call (table discovered from its own Refiner); the reference's bookkeeping
    (Refiner.update -> history) receives the values.  The device evaluation itself is replaced by a stand-in here (no
    GPU in the build container; its numerics are pinned by tests/test_gpu_population.py on binding.npz)."""
    import threading
    import numpy as np
    import pyxrd_b200.install as inst
    from pyxrd_b200 import population
    mixture = reference_mixture
    calls, steepness = [], [1e-6]

    def stand_in(self, x, ctx=None, **kwargs):
        calls.append(np.array(x))
        return np.array([1e-3 + float(np.sum((row - np.array([7.0, 18.0])) ** 2)) * steepness[0] for row in np.atleast_2d(x)])

    monkeypatch.setattr(population.PopulationTemplate, "optimized_residuals", stand_in)
    try:
        inst.install(calculations=False)
        mixture.refinement.refine_method_index = 3          # brute force
        refiner = mixture.refinement.get_refiner()
        assert [r.text_title for r in refiner.refinables] == ["Average CSDS", "Average CSDS"]
        refiner.refine(stop=threading.Event())
    finally:
        inst.uninstall()
    assert len(calls) == 1 and calls[0].shape == (121, 2)                      # 11 x 11 grid, one call
    tmpl = refiner._b200_template
    assert isinstance(tmpl, population.PopulationTemplate) and tmpl.clip_to_bounds
    assert tmpl.csds_auto is not None and tmpl.csds_auto.sum() >= 2            # maximum follows the bound average
    grid = calls[0]
    best = grid[np.argmin(np.sum((grid - np.array([7.0, 18.0])) ** 2, axis=1))]
    assert np.allclose(refiner.history.best_solution, best) and refiner.history.best_residual < 2e-3
    # the L BFGS B run (scipy_runs.py:35-52): every gradient = one call with d + 1 rows
    del calls[:]
    steepness[0] = 1e-2
    monkeypatch.setattr(type(refiner), "update", lambda self, solution, residual=None, iteration=0: None)
    try:
        inst.install(calculations=False)
        mixture.refinement.refine_method_index = 0
        refiner = mixture.refinement.get_refiner()
        refiner.refine(stop=threading.Event())
    finally:
        inst.uninstall()
    assert len(calls) >= 2 and all(c.shape == (3, 2) for c in calls)
    b = tmpl.bounds()
    assert np.allclose(calls[-1][0], np.clip([7.0, 18.0], b[:, 0], b[:, 1]), atol=1e-2)     # the stand-in's minimum inside the ranges


@pytest.mark.filterwarnings("ignore")
@pytest.mark.skipif(not HAVE_REFERENCE, reason="needs the reference tree (build container only)")
def test_refinements_outside_the_optimiser_scope_keep_the_reference_loop(reference_mixture):
    """with ``auto_scales`` off the reference zeroes scales_mask (mixture/models/mixture.py:64-68) and
    get_optimized_residual keeps the scales fixed; the device optimiser would free them.  The install hooks must then
    hand the generation back to the reference's own loop (``_refiner_template`` -> False)."""
    import pyxrd_b200.install as inst
    from pyxrd_b200 import population
    mixture = reference_mixture
    refiner = mixture.refinement.get_refiner()
    template = inst._refiner_template(refiner)
    assert isinstance(template, population.PopulationTemplate) and template.optimiser_scope() is None
    mixture.auto_scales = False
    refiner = mixture.refinement.get_refiner()
    assert inst._refiner_template(refiner) is False
    mixture.auto_scales = True
    mixture.auto_bg = False                     # background fixed for every specimen: in scope, not refined
    refiner = mixture.refinement.get_refiner()
    template = inst._refiner_template(refiner)
    assert isinstance(template, population.PopulationTemplate) and not template.mixture_refines_background()
