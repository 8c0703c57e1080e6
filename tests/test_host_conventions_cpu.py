"""Host-side conventions of the drop-in functions that need no GPU: the solution-vector bookkeeping of
``calculations/mixture.py`` against the oracle's restatement of reference mixture.py:32-77,107-138 (random masks,
bit for bit), the ``wrap_exceptions`` convention inside a daemon worker (exceptions.py:23-37) and the
``"AbtractPhase"`` -> NotImplementedError convention (phases.py:74-75)."""
import copy
import multiprocessing as mp
import warnings

import numpy as np
import pytest

from helpers import syn
from oracle import pyxrd_oracle as orc
from pyxrd_b200.calculations import mixture as cm
from pyxrd_b200.calculations.exceptions import WrapException, wrap_exceptions
from pyxrd_b200.data_objects import MixtureData, SpecimenData


def _random_mixture(rng):
    M, S = int(rng.integers(1, 6)), int(rng.integers(1, 4))
    specs = [SpecimenData(z_list=[0.0]) if rng.uniform() > 0.1 else None for _ in range(S)]
    fr = rng.uniform(0.0, 1.0, M)
    mode = int(rng.integers(0, 4))
    if mode == 0:
        fr = fr / fr.sum()
    elif mode == 1:
        fr = fr * 3.0
    return dict(specimens=specs, fractions=fr, scales=rng.uniform(0.5, 2.0, S), bgshifts=rng.uniform(0.0, 2.0, S),
                fractions_mask=(rng.uniform(size=M) > 0.3).astype(float),
                scales_mask=(rng.uniform(size=S) > 0.3).astype(float),
                bgshifts_mask=(rng.uniform(size=S) > 0.5).astype(float) if rng.uniform() > 0.3 else np.zeros(S),
                n=S, m=M, parsed=False, calculated=False, optimized=False)


def test_solution_bookkeeping_matches_the_oracle_bit_for_bit(monkeypatch):
    monkeypatch.setattr(orc, "calculate_phase_intensities", lambda specimen: (None, None))
    rng = np.random.default_rng(5)
    seen_fixed, seen_assert, seen_renorm = 0, 0, 0
    for _ in range(1500):
        kw = _random_mixture(rng)
        a, b = MixtureData(**copy.deepcopy(kw)), MixtureData(**copy.deepcopy(kw))
        raised = [False, False]
        for i, (fn, mix) in enumerate(((cm._parse_bookkeeping, a), (orc.parse_mixture, b))):
            try:
                fn(mix)
            except AssertionError:
                raised[i] = True
        assert raised[0] == raised[1]
        if raised[0]:
            seen_assert += 1
            continue
        for name in ("m", "n", "o", "real_m", "real_n", "real_o", "z"):
            assert getattr(a, name) == getattr(b, name), name
        for name in ("selected_fractions", "selected_scales", "selected_bgshifts", "fractions"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), name
        assert a.sum_static_fractions == b.sum_static_fractions
        seen_fixed += a.m < a.real_m
        seen_renorm += not np.array_equal(a.fractions, kw["fractions"])
        if a.m == 0:
            continue
        with np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert np.array_equal(cm.get_zero_solution(a), orc.get_zero_solution(b), equal_nan=True)
            assert cm.get_bounds(a) == orc.get_bounds(b)
            assert np.array_equal(cm.get_solution(a), orc.get_solution(b))
            x = rng.uniform(0.1, 2.0, a.m + a.n + a.o)
            if rng.uniform() < 0.05:
                x[:a.m] = 0.0                     # 0 / 0 -> nan fractions in both
            held = a.fractions                    # the arrays are written in place, like np.put does
            cm.set_solution(x.copy(), a)
            orc.set_solution(x.copy(), b)
            assert held is a.fractions
            for name in ("fractions", "scales", "bgshifts"):
                assert np.array_equal(getattr(a, name), getattr(b, name), equal_nan=True), name
    assert seen_fixed > 100 and seen_assert > 10 and seen_renorm > 10


def test_zero_solution_quirk_without_free_background():
    """x0[-0:] = 0 zeroes the whole vector in the reference: scales start at 0 when o == 0"""
    mix = MixtureData(m=2, n=2, o=0, sum_static_fractions=0.0)
    assert np.array_equal(cm.get_zero_solution(mix), [1.0, 1.0, 0.0, 0.0])
    mix.o = 2
    assert np.array_equal(cm.get_zero_solution(mix), [1.0, 1.0, 1.0, 1.0, 0.0, 0.0])


@wrap_exceptions
def _failing(value):
    raise KeyError(value)


def _daemon_body(queue):
    try:
        _failing("inside a daemon")
    except WrapException as err:
        queue.put(("wrapped", type(err.exception).__name__, "KeyError" in err.formatted and "_failing" in err.formatted))
    except Exception as err:       # pragma: no cover
        queue.put(("plain", type(err).__name__, False))


def test_wrap_exceptions_in_a_daemon_worker_and_outside():
    with pytest.raises(KeyError):                 # not a daemon: re-raised as is (exceptions.py:35-36)
        _failing("main process")
    ctx = mp.get_context("fork")
    queue = ctx.Queue()
    proc = ctx.Process(target=_daemon_body, args=(queue,), daemon=True)
    proc.start()
    kind, name, has_traceback = queue.get(timeout=60)
    proc.join(timeout=60)
    assert (kind, name, has_traceback) == ("wrapped", "KeyError", True)


def test_abstract_phase_type_raises_not_implemented_before_any_device_work(monkeypatch):
    """phases.py:74-75: the (misspelt) "AbtractPhase" type raises NotImplementedError.  The check sits in the packing
    step, so it fires without a device."""
    from pyxrd_b200 import runtime
    from pyxrd_b200.calculations import phases
    from pyxrd_b200.packing import BatchPack, StaticPack, collect_atom_types

    def no_device(*args, **kwargs):               # pragma: no cover
        raise AssertionError("the type check must come before any device work")

    mix = syn.CONFIGS["c1"].truth()
    spec = mix.specimens[0]
    spec.phases[0][0].type = "AbtractPhase"
    static = StaticPack(mix.specimens, collect_atom_types(mix.specimens), n_patterns=1)
    with pytest.raises(NotImplementedError):
        BatchPack([mix], static)
    monkeypatch.setattr(runtime, "default_context", no_device)
    monkeypatch.setattr(phases, "default_context", no_device, raising=False)
    theta = spec.range_theta
    with pytest.raises(NotImplementedError):
        phases.get_diffracted_intensity(theta, 2 * np.sin(theta) / 0.154056, spec.phases[0][0])


def _forked_child_body(queue):
    from pyxrd_b200 import runtime
    try:
        runtime.default_context(0)
        queue.put(("created", ""))
    except runtime.DeviceError as err:
        queue.put(("DeviceError", str(err)))
    except Exception as err:       # pragma: no cover
        queue.put((type(err).__name__, str(err)))


def test_child_forked_after_cuda_initialisation_gets_a_clear_error():
    """SURVEY.md 8(b) threading model (iii): the reference evaluates trials in forked Pool workers
    (pyxrd_server.py:57-64).  CUDA does not survive fork(): a child of a process that holds a context must get a
    DeviceError that names the alternatives, must not inherit the parent's handles and must not call into the CUDA
    runtime with them.  (The parent's context is a stand-in here: the hook only looks at the registry.)"""
    from pyxrd_b200 import runtime

    class Held(object):
        _h = object()
        lock = None

    key = (-1, 0)
    runtime._contexts[key] = Held()
    try:
        ctx = mp.get_context("fork")
        queue = ctx.Queue()
        proc = ctx.Process(target=_forked_child_body, args=(queue,))
        proc.start()
        kind, text = queue.get(timeout=60)
        proc.join(timeout=60)
    finally:
        del runtime._contexts[key]
    assert kind == "DeviceError", (kind, text)
    assert "forked" in text and "B200AsyncServerProvider" in text and "spawn" in text
    assert not runtime._forked_after_cuda            # the parent is unaffected


def test_resident_token_is_cleared_by_every_call_that_replaces_device_state():
    """the residency of a parsed mixture is remembered by a token on the CONTEXT; every entry point that
    replaces the tables, the batch or its intensities must clear it (library calls are stubbed: host logic only)"""
    from pyxrd_b200 import runtime

    class Lib(object):
        def __getattr__(self, name):
            return lambda *a, **k: 0

    class Pack(object):
        K = M = S = Z = 1
        sumX = 4

        def signature(self):
            return id(self)

    ctx = runtime.Context.__new__(runtime.Context)
    ctx._lib, ctx._h, ctx.device = Lib(), None, 0
    ctx._static = ctx._static_sig = ctx._batch = None
    ctx._resident_token = None
    ctx.batch_desc = lambda b: runtime.BatchDesc()
    ctx._population_desc_orig = ctx._population_desc
    token = ctx.new_resident_token()
    assert token[0] == __import__("os").getpid() and ctx._resident_token == token
    assert ctx.new_resident_token() != token                      # unique per upload
    import pickle
    assert pickle.loads(pickle.dumps(token)) == token             # plain ints: survives the Pool / Pyro boundary
    static = Pack()
    static.S = static.T = static.Z = 1
    for attr in ("npts", "theta", "gonio", "nwl", "wl_fraction", "wl_index", "wl_w_hi", "wl_w_lo", "observed", "selected",
                 "atype_table"):
        setattr(static, attr, np.zeros(4, dtype=np.int32 if attr in ("npts", "nwl", "wl_index") else
                                       np.uint8 if attr == "selected" else float))
    static.stl_override = None
    calls = {
        "set_static": lambda: ctx.set_static(static, force=True),
        "set_batch": lambda: ctx.set_batch(Pack()),
        "set_intensities": lambda: ctx.set_intensities(np.zeros(4)),
        "evaluate_host": lambda: ctx.evaluate_host(Pack()),
        "optimize_host": lambda: ctx.optimize_host(Pack()),
        "evaluate_population": lambda: ctx.evaluate_population(Pack(), np.zeros((1, 1)), np.zeros(0, dtype=runtime.BINDING_DTYPE)),
        "set_population": lambda: ctx.set_population(Pack(), np.zeros((1, 1)), np.zeros(0, dtype=runtime.BINDING_DTYPE)),
        "optimize_population": lambda: ctx.optimize_population(Pack(), np.zeros((1, 1)), np.zeros(0, dtype=runtime.BINDING_DTYPE)),
    }
    for name, call in calls.items():
        ctx.new_resident_token()
        call()
        assert ctx._resident_token is None, name
    ctx.new_resident_token()
    ctx._batch = Pack()
    ctx.set_solution(np.ones(1), np.ones(1), np.zeros(1))         # the solution is not part of the residency
    ctx.compute_residuals()
    assert ctx._resident_token is not None
