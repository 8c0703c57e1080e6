"""Host-side logic of the drop-in boundary: rebinding of pyxrd.calculations.* and the
batching async provider (CPU; no compute)."""
import functools
import os
import sys

import pytest

from pyxrd_b200.provider import B200AsyncServer, B200AsyncServerProvider

HAVE_REFERENCE = os.path.isdir("/root/reference/pyxrd")


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
