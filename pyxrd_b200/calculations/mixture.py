"""Device versions of ``pyxrd/calculations/mixture.py``: same names, same arguments,
same in-place mutation of the MixtureData / SpecimenData objects.

Phase intensities, scaling / summing and the Rp residual run on the device.  The bounded
minimisation of ``optimize_mixture`` is driven by the very same scipy L-BFGS-B call as the
reference (mixture.py:178-184) with the device evaluating the objective; the batched,
fully device-resident optimiser for whole populations lives in ``pyxrd_b200.batched``.
"""
import logging
import time

import numpy as np

from .. import settings
from ..packing import BatchPack, StaticPack, collect_atom_types
from ..runtime import default_context
from .exceptions import wrap_exceptions

logger = logging.getLogger(__name__)


def _inspect_signature_has(func, name):
    import inspect
    try:
        return name in inspect.signature(func).parameters
    except (TypeError, ValueError):
        return False


# ---- solution vector handling (mixture.py:32-77): host bookkeeping on tiny vectors ------------
# x = [free fractions | free scales | free bgshifts]; which entries of the mixture's arrays are free was decided by
# parse_mixture (selected_*), m / n / o count them.
def _split_solution(x, mixture):
    """the three segments of x.  The reference addresses the background segment as x[-o:], which is the WHOLE vector
    when o == 0 -- harmless there because no index receives it; here the empty case is an empty slice."""
    x = np.asarray(x)
    nf, ns, nb = int(mixture.m), int(mixture.n), int(mixture.o)
    return x[:nf], x[nf:nf + ns], (x[x.size - nb:] if nb else x[:0])


def _store(array, where, values):
    """array[where[i]] = values[i], in place (the caller may hold a reference to the array, as with np.put)"""
    where = np.asarray(where, dtype=np.intp).ravel()
    if where.size:
        array[where] = np.asarray(values)[:where.size]
    return array


def parse_solution(x, mixture):
    """mixture.py:32-48: x written into the mixture's own arrays (which are returned); the free fractions are rescaled
    so that free + static fractions sum to one (0 / 0 -> nan for an all-zero proposal, as in the reference)."""
    free_f, free_s, free_b = _split_solution(x, mixture)
    # multiply first, divide second: the rounding of the reference's expression (mixture.py:40)
    _store(mixture.fractions, mixture.selected_fractions, free_f * (1.0 - mixture.sum_static_fractions) / np.sum(free_f))
    _store(mixture.scales, mixture.selected_scales, free_s)
    _store(mixture.bgshifts, mixture.selected_bgshifts, free_b)
    return mixture.fractions, mixture.scales, mixture.bgshifts


def get_solution(mixture):
    """mixture.py:50-55"""
    parts = [np.asarray(arr)[np.asarray(sel, dtype=np.intp)] for arr, sel in (
        (mixture.fractions, mixture.selected_fractions), (mixture.scales, mixture.selected_scales),
        (mixture.bgshifts, mixture.selected_bgshifts))]
    return np.concatenate(parts)


def set_solution(x, mixture):
    """mixture.py:57-58"""
    mixture.fractions, mixture.scales, mixture.bgshifts = parse_solution(x, mixture)


def get_zero_solution(mixture):
    """mixture.py:60-68: starting point of the reference's L-BFGS-B call.  Quirk kept on purpose: the
    reference zeroes the background segment with x0[-o:] = 0, which for o == 0 zeroes the whole vector, so without
    free background shifts the scales start at 0 (and are lifted to their lower bound 1e-3 by the optimiser)."""
    nf, ns, nb = int(mixture.m), int(mixture.n), int(mixture.o)
    x0 = np.empty(nf + ns + nb)
    x0[nf:nf + ns] = 1.0 if nb else 0.0
    x0[nf + ns:] = 0.0
    x0[:nf] = 1.0 / (1.0 - mixture.sum_static_fractions)
    return x0


def get_bounds(mixture):
    """mixture.py:73-77: fractions in [0, 1], scales >= 1e-3, background shifts >= 0"""
    return [(0., 1.)] * int(mixture.m) + [(1e-3, None)] * int(mixture.n) + [(0., None)] * int(mixture.o)


# ---- device residency -------------------------------------------------------------------
def _make_resident(mixture, ctx, download):
    """Pack + upload the mixture and compute its phase intensities on ``ctx``.

    ``mixture._b200_epoch`` (a tuple of plain ints (pid, context, counter): survives pickling, is no device handle, and
    never matches a context of another process) remembers which upload the context holds, so that later calls on
    the same parsed mixture reuse the device-resident intensities instead of recomputing them.  The context forgets
    the token whenever ANY call replaces its tables, batch or intensities (runtime.Context.invalidate_residency)."""
    specimens = mixture.specimens
    static = StaticPack(specimens, collect_atom_types(specimens), n_patterns=mixture.z)
    ctx.set_static(static)
    batch = BatchPack([mixture], static, bgshift_enabled=bool(settings.get("BGSHIFT")))
    ctx.set_batch(batch)
    ctx.compute_intensities()
    mixture._b200_epoch = ctx.new_resident_token()
    if download:
        corr = ctx.read_correction_flat()
        per_spec = ctx.split_per_specimen(ctx.read_intensities_flat(), (batch.M, static.Z))[0]
        for s, specimen in enumerate(specimens):
            if specimen is not None:
                lo, hi = int(static.offsets[s]), int(static.offsets[s + 1])
                specimen.correction = corr[lo:hi].copy()
                specimen.phase_intensities = per_spec[s].copy()


def _ensure_resident(mixture, ctx):
    token = getattr(mixture, "_b200_epoch", None)
    if token is None or ctx._resident_token != token:
        _make_resident(mixture, ctx, download=False)


def _device_residuals(mixture, ctx, want_patterns):
    """residuals [1+S] of the mixture's current fractions / scales / bgshifts"""
    _ensure_resident(mixture, ctx)
    S = len(mixture.specimens)
    bg = np.asarray(mixture.bgshifts, dtype=float) if settings.get("BGSHIFT") else np.zeros(S)
    ctx.set_solution(np.asarray(mixture.fractions, dtype=float), np.asarray(mixture.scales, dtype=float)[:S], bg[:S])
    has_obs = any(sp is not None and np.size(sp.observed_intensity) > 0 for sp in mixture.specimens)
    ctx.set_residual_method(settings.get("RESIDUAL_METHOD") if has_obs else "Rp")
    ctx.compute_residuals(want_scaled=want_patterns)
    return ctx.read_residuals()[0]


# ---- the public functions ----------------------------------------------------------------
def _free_indices(mask):
    """positions of the non-zero entries of a mask, as a flat index array"""
    return np.flatnonzero(np.asarray(mask))


def _parse_bookkeeping(mixture):
    """the host half of parse_mixture (mixture.py:107-138): sanity checks (AssertionError, which the callers swallow),
    the index sets of the free variables, the static fraction sum, the pattern count.  Conventions kept: m / n / o are
    OVERWRITTEN with the numbers of free variables while real_m / real_n keep the sizes -- and real_o is set to n, not
    o, as the reference does (mixture.py:119)."""
    mixture.parsed = False
    n_specimens, n_phases = mixture.n, mixture.m
    assert n_specimens > 0, "a mixture needs at least one specimen"
    assert n_specimens == len(mixture.specimens), "mixture.n does not match the number of specimens"
    assert n_phases > 0, "a mixture needs at least one phase"
    mixture.real_m, mixture.real_n, mixture.real_o = n_phases, n_specimens, n_specimens
    for name, mask in (("fractions", mixture.fractions_mask), ("scales", mixture.scales_mask),
                       ("bgshifts", mixture.bgshifts_mask)):
        setattr(mixture, "selected_" + name, _free_indices(mask))
    mixture.m = len(mixture.selected_fractions)
    mixture.n = len(mixture.selected_scales)
    mixture.o = len(mixture.selected_bgshifts)

    def static_share(total):
        return total - np.sum(np.asarray(mixture.fractions)[mixture.selected_fractions])

    total = np.sum(mixture.fractions)
    static = static_share(total)
    # fractions a user typed in need not sum to one: renormalise when the fixed ones would leave a negative share
    # (and the total is not one) or claim more than everything -- the reference's condition, mixture.py:126
    if (total != 1.0 and static < 0.0) or static > 1.0:
        mixture.fractions = mixture.fractions / total
        total = 1.0
    mixture.sum_static_fractions = static_share(total)
    present = [sp for sp in mixture.specimens if sp is not None]
    mixture.z = len(present[0].z_list) if present else 0
    assert mixture.z > 0, "a mixture needs at least one pattern in one of its specimens"
    for _ in range(len(mixture.specimens) - len(present)):
        logger.warning("parse_mixture reports: 'None found!'")


def parse_mixture(mixture, force=False):
    """mixture.py:103-148: bookkeeping of the selected variables + all phase intensities."""
    if mixture.parsed and not force:
        return
    _parse_bookkeeping(mixture)
    ctx = default_context()
    with ctx.lock:
        _make_resident(mixture, ctx, download=True)
    mixture.parsed = True


@wrap_exceptions
def calculate_mixture(mixture, force=False):
    """mixture.py:221-257."""
    if mixture.calculated and not force:
        return mixture
    try:
        parse_mixture(mixture)
    except AssertionError:
        for specimen in mixture.specimens:
            if specimen is not None:
                specimen.total_intensity = None
                specimen.scaled_phase_intensities = None
        return mixture
    ctx = default_context()
    with ctx.lock:
        res = _device_residuals(mixture, ctx, want_patterns=True)
        static, batch = ctx._static, ctx._batch
        totals = ctx.split_per_specimen(ctx.read_totals_flat(), (static.Z,))[0]
        scaled = ctx.split_per_specimen(ctx.read_scaled_flat(), (batch.M, static.Z))[0]
    bgs = np.asarray(mixture.bgshifts, dtype=float)
    residuals = [0.0]
    for s, specimen in enumerate(mixture.specimens):
        value = 0.0
        if specimen is not None:
            bg = float(bgs[s]) if settings.get("BGSHIFT") else 0.0
            specimen.background_intensity = bg * specimen.correction
            specimen.scaled_phase_intensities = scaled[s].copy()
            specimen.total_intensity = totals[s].copy()
            if np.size(specimen.observed_intensity) > 0:
                value = float(res[1 + s])
            else:
                logger.warning("calculate_mixture reports: 'Zero observations found!'")
        residuals.append(value)
    residuals[0] = np.average(residuals[1:])
    mixture.residuals = residuals
    mixture.calculated = True
    return mixture


def _get_residuals(x, mixture):
    """mixture.py:91-94."""
    set_solution(x, mixture)
    mixture = calculate_mixture(mixture)
    return mixture.residuals


def get_residual(mixture):
    """mixture.py:96-97."""
    return _get_residuals(get_solution(mixture), mixture)


@wrap_exceptions
def optimize_mixture(mixture, force=False):
    """mixture.py:150-219.  ``settings.MIXTURE_OPTIMIZER == "device"`` (default): the device-resident optimiser when the
    mixture is in its scope (all fractions and scales free, Rp); otherwise / ``"scipy"``: the reference's own call --
    scipy L-BFGS-B (finite-difference gradient, bounds) over [fractions | scales | bgshifts] with the objective evaluated
    on the device."""
    from scipy.optimize import fmin_l_bfgs_b
    if mixture.optimized and not force:
        return mixture
    try:
        parse_mixture(mixture)
    except AssertionError:
        if settings.get("DEBUG"):
            raise
        return mixture
    ctx = default_context()
    if settings.get("MIXTURE_OPTIMIZER") == "device":
        from ..batched import _device_optimisable
        if _device_optimisable(mixture):
            # the whole minimisation in one kernel on the resident patterns (k_mixture_optimize)
            with ctx.lock:
                _ensure_resident(mixture, ctx)
                refine_bg = mixture.o > 0 and bool(settings.get("BGSHIFT"))
                ctx.optimize(refine_bg=refine_bg)
                sol = ctx.read_solutions()[0]
                res = ctx.read_opt_residuals()[0]
            M, S = len(np.asarray(mixture.fractions).ravel()), len(mixture.specimens)
            mixture.fractions = np.around(sol[:M] / np.sum(sol[:M]), 6)
            mixture.scales = np.array(sol[M:M + S]).round(6)
            if mixture.o > 0:
                mixture.bgshifts = np.array(sol[M + S:M + 2 * S])
            mixture.residual = np.array([res[0]])
            mixture.calculated = False
            mixture.optimized = True
            return mixture
    x0 = get_zero_solution(mixture)
    bounds = get_bounds(mixture)
    no_obs = [s for s, sp in enumerate(mixture.specimens) if sp is None or np.size(sp.observed_intensity) == 0]

    best = [np.inf, None]

    def objective(x):
        mixture.calculated = False
        set_solution(x, mixture)
        with ctx.lock:
            res = _device_residuals(mixture, ctx, want_patterns=False)
        if no_obs:
            per = [0.0 if s in no_obs else res[1 + s] for s in range(len(mixture.specimens))]
            value = np.average(per)
        else:
            value = res[0]
        if np.isfinite(value) and value < best[0]:
            best[0], best[1] = float(value), np.array(x, dtype=float)
        return value

    t1 = time.time()
    kwargs = dict(approx_grad=True, bounds=bounds)
    if _inspect_signature_has(fmin_l_bfgs_b, "iprint"):
        kwargs["iprint"] = -1
    lastx, residual, info = fmin_l_bfgs_b(objective, x0, **kwargs)
    if not np.all(np.isfinite(residual)) and best[1] is not None:
        # L-BFGS-B may end on a proposal with every free fraction at its lower bound 0: parse_solution then divides
        # 0 / 0 and the objective is NaN (the reference hands that NaN back: whether a run gets there depends on the last
        # bits of the objective).  The best point the run has seen is returned instead.
        lastx, residual = best[1], best[0]
    if np.isscalar(residual):
        residual = np.array([residual])
    logger.debug('%s took %0.3f ms' % ("optimize_mixture", (time.time() - t1) * 1000.0))
    fractions, scales, bgshifts = parse_solution(lastx, mixture)
    fractions = fractions.flatten()
    sum_frac = np.sum(fractions)
    if sum_frac == 0.0 and len(fractions) > 0:
        fractions[0] = 1.0
        sum_frac = 1.0
    fractions = np.around((fractions / sum_frac), 6)
    scales *= sum_frac
    scales = scales.round(6)
    mixture.fractions = fractions
    mixture.scales = scales
    mixture.bgshifts = bgshifts
    mixture.residual = residual
    mixture.calculated = False
    mixture.optimized = True
    return mixture


@wrap_exceptions
def calculate_and_optimize_mixture(mixture):
    """mixture.py:259-266."""
    return calculate_mixture(optimize_mixture(mixture))


@wrap_exceptions
def get_optimized_residual(mixture):
    """mixture.py:268-276: the optimiser's final objective value (1-element array)."""
    mixture = calculate_and_optimize_mixture(mixture)
    return mixture.residual
