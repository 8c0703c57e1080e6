"""CPU ORACLE for the PyXRD matrix-method hot path -- TEST INFRASTRUCTURE ONLY.

A numpy restatement, in this repository's own words, of the algorithm of
``pyxrd/calculations/{atoms,components,CSDS,phases,goniometer,specimen,
statistics,mixture}.py`` of the reference (v0.8.4).  Every function cites the
reference lines it follows.  It is the *checker* for the CUDA path: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import it.  The product (``pyxrd_b200``) never imports this module and has no
CPU fallback.

Pinning: the reference is pure Python and runs in the build container, so this
restatement is pinned against (i) the reference's own known-answer tests
(``test/test_calculations/test_goniometer.py:63-95``,
``test/test_atoms/test_atom_type.py:58-75``) and (ii) outputs of the live
reference on the synthetic BASELINE configs, committed under ``tests/golden/``
together with ``tests/golden/make_golden.py`` that generated them
(numpy 2.3.5 / scipy 1.18.1).  See ``tests/test_oracle_golden.py``.

Third-party arithmetic the reference relies on and that is restated here:
``scipy.interpolate.interp1d(kind='linear', bounds_error=False, fill_value=0)``
(scipy 1.18.1 ``_interpolate.py`` ``_call_linear`` / ``_check_bounds``: two-weight
form ``w_hi*y_hi + w_lo*y_lo``); ``scipy.special.erf`` (math.erf here, vectorised);
``scipy.optimize.fmin_l_bfgs_b`` is *called*, not restated (it is a library on
both sides).

Cost structure: like the reference this evaluates the full r x r complex matrix
powers Q^n for n = 1..N (O(N r^3) per 2-theta point), so that timing it is a fair
stand-in for the reference CPU path; the powers are consumed as they are formed
instead of being stored (the reference keeps (N+1) X r^2 complex128 alive).
"""
from __future__ import annotations

import math

import numpy as np

try:  # erf: a library function on both sides (reference: scipy.special.erf, goniometer.py:10)
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[float])

SQRT2PI = math.sqrt(2.0 * math.pi)
SQRT8 = math.sqrt(8.0)

# the three reference settings the path reads at call time (data/settings.py:15,17,32)
SETTINGS = dict(RESIDUAL_METHOD="Rp", BGSHIFT=True, DEBUG=False)


# ---------------------------------------------------------------------------
# atoms.py
# ---------------------------------------------------------------------------
def get_atomic_scattering_factor(angstrom_range, atom_type):
    """atoms.py:13-38 -- (sum_i a_i exp(-b_i s) + c) * exp(-B s); None type -> zeros."""
    s = np.asarray(angstrom_range, dtype=float)
    if atom_type is None:
        return np.zeros_like(s)
    gauss = np.asarray(atom_type.par_a) * np.exp(-np.asarray(atom_type.par_b) * s[..., np.newaxis])
    asf = np.sum(gauss, axis=1) + atom_type.par_c
    return asf * np.exp(-atom_type.debye * s)


def get_structure_factor(range_stl, atom):
    """atoms.py:40-67 -- asf * pn * exp(2 pi i z stl), with s = (0.05 stl)^2."""
    stl = np.asarray(range_stl, dtype=float)
    if atom is None or atom.atom_type is None:
        return np.zeros_like(stl)
    asf = get_atomic_scattering_factor((stl * 0.05) ** 2, atom.atom_type)
    return asf * atom.pn * np.exp(2 * math.pi * atom.z * stl * 1j)


# ---------------------------------------------------------------------------
# components.py
# ---------------------------------------------------------------------------
def calculate_z(default_z, lattice_d, z_factor):
    """components.py:15-16 -- stretch an interlayer position."""
    return lattice_d + z_factor * (default_z - lattice_d)


def get_factors(range_stl, component):
    """components.py:18-54 -- (SF, phi) of one layer type; writes ``atom.z``."""
    stl = np.asarray(range_stl, dtype=float)
    z_factor = (component.d001 - component.lattice_d) / (component.default_c - component.lattice_d)
    total = 0.0 + 0.0j
    for atom in component.layer_atoms:
        atom.z = atom.default_z
        total = total + get_structure_factor(stl, atom)
    for atom in component.interlayer_atoms:
        atom.z = calculate_z(atom.default_z, component.lattice_d, z_factor)
        total = total + get_structure_factor(stl, atom)
    phi = np.exp(2.0 * math.pi * stl * (component.d001 * 1j - math.pi * component.delta_c * stl))
    return total, phi


# ---------------------------------------------------------------------------
# math_tools.py / CSDS.py
# ---------------------------------------------------------------------------
def lognormal(T, a, b):
    """math_tools.py:36-37."""
    return math.exp(-(math.log(T) - a) ** 2 / (2.0 * (b ** 2))) / (SQRT2PI * abs(b) * T)


def calculate_distribution(CSDS):
    """CSDS.py:13-46 -- q_T for T = min..max (normalised), and the real mean."""
    a = CSDS.alpha_scale * math.log(CSDS.average) + CSDS.alpha_offset
    b = math.sqrt(CSDS.beta_scale * math.log(CSDS.average) + CSDS.beta_offset)
    count = int(CSDS.maximum - CSDS.minimum) + 1
    raw, total, last = {}, 0, 0
    for step in range(count):
        T = max(CSDS.minimum + step, 1e-50)
        q = lognormal(T, a, b)
        total += q
        raw[int(T)] = q
        last = T
    arr = np.zeros(shape=(last + 1,), dtype=float)
    mean = 0
    for T, q in raw.items():
        arr[T] = q / total
        mean += T * q
    return arr, mean / total


# ---------------------------------------------------------------------------
# goniometer.py
# ---------------------------------------------------------------------------
def get_S(soller1, soller2):
    """goniometer.py:15-18."""
    return math.sqrt((soller1 * 0.5) ** 2 + (soller2 * 0.5) ** 2), soller1 * soller2


def get_T(range_theta, sigma_star, soller1, soller2):
    """goniometer.py:20-25."""
    sigma_star = float(max(sigma_star, 1e-18))
    S, _ = get_S(soller1, soller2)
    st = np.sin(range_theta)
    Q = S / (SQRT8 * st * sigma_star)
    return _erf(Q) * SQRT2PI / (2.0 * sigma_star * S) - 2.0 * st * (1.0 - np.exp(-(Q ** 2.0))) / (S ** 2.0)


def get_lorentz_polarisation_factor(range_theta, sigma_star, soller1, soller2, mcr_2theta):
    """goniometer.py:27-34."""
    theta = np.asarray(range_theta, dtype=float)
    T = get_T(theta, sigma_star, soller1, soller2)
    pol = math.cos(math.radians(mcr_2theta)) ** 2
    return T * (1.0 + pol * (np.cos(2.0 * theta) ** 2)) / np.sin(theta)


def get_fixed_to_ads_correction_range(range_theta, goniometer):
    """goniometer.py:36-37."""
    return np.sin(range_theta)


# ---------------------------------------------------------------------------
# phases.py
# ---------------------------------------------------------------------------
def _bmm(A, B):
    """math_tools.py:16-17 -- batched (X, r, r) complex product."""
    return np.matmul(A, B)


def get_absolute_scale(components, CSDS_real_mean, W):
    """phases.py:48-66 -- uses the FIRST G diagonal entries of W, as written."""
    w = np.diag(W)
    volume = d001 = density = 0.0
    for i, comp in enumerate(components):
        if comp is not None:
            volume += comp.volume * w[i]
            d001 += comp.d001 * w[i]
            density += comp.weight * w[i] / comp.volume
    mass = CSDS_real_mean * volume ** 2 * density
    return d001 / mass if mass != 0.0 else 0.0


def get_structure_factors(range_stl, G, comp_list):
    """phases.py:20-39."""
    stl = np.asarray(range_stl, dtype=float)
    SF = np.zeros(stl.shape + (G,), dtype=complex)
    PF = np.zeros(stl.shape + (G,), dtype=complex)
    for i, comp in enumerate(comp_list):
        SF[:, i], PF[:, i] = get_factors(stl, comp)
    return SF, PF


def progression_factors(CSDS_arr, minimum, maximum):
    """phases.py:147-151 -- coef_n = 2 * sum_{m>n} (m-n) q_m for n = min..max."""
    out = {}
    for n in range(minimum, maximum + 1):
        pf = 0
        for m in range(n + 1, maximum + 1):
            pf += (m - n) * CSDS_arr[m]
        out[n] = 2 * pf
    return out


def get_diffracted_intensity(range_theta, range_stl, phase):
    """phases.py:68-79,106-158 -- matrix-method intensity of one phase (no LPF)."""
    if phase.type == "AbtractPhase":
        raise NotImplementedError
    if phase.type == "RawPatternPhase":
        return _raw_intensity(range_theta, phase)
    stl = np.asarray(range_stl, dtype=float)
    if not phase.valid_probs:
        return np.zeros_like(stl)
    dist, mean = calculate_distribution(phase.CSDS)
    scale = get_absolute_scale(phase.components, mean, phase.W)
    X = stl.shape[0]
    P = np.asarray(phase.P, dtype=float)
    W = np.asarray(phase.W, dtype=float)
    rank = P.shape[1]
    reps = int(rank // phase.G)
    SF, PF = get_structure_factors(stl, phase.G, phase.components)
    u = np.repeat(SF, reps, axis=1)                       # u[x, I] = SF[x, comp(I)]
    F = np.conjugate(u)[:, :, np.newaxis] * u[:, np.newaxis, :]     # F[I,J] = conj(u_I) u_J
    Q = np.repeat(PF, reps, axis=1)[:, np.newaxis, :] * P[np.newaxis, :, :]   # Q[I,J] = phi_J P[I,J]
    N, Nmin = int(phase.CSDS.maximum), int(phase.CSDS.minimum)
    coef = progression_factors(dist, Nmin, N)
    # powers[k] = Q^(k+1), k = 0..N (phases.py:41-46); term n uses powers[n-1], python-style index
    want = {}
    for n in range(Nmin, N + 1):
        want.setdefault((n - 1) % (N + 1), []).append(coef[n])
    series = np.zeros((X, rank, rank), dtype=complex)
    power = Q.astype(complex)
    for k in range(0, N + 1):
        if k > 0:
            power = _bmm(power, Q)
        for c in want.get(k, ()):
            series += c * power
    series = series + np.identity(rank, dtype=complex)[np.newaxis] * mean
    FW = _bmm(F, np.broadcast_to(W.astype(complex), (X, rank, rank)))
    total = _bmm(FW, series)
    return np.real(np.trace(total, axis1=2, axis2=1)) * scale


def _interp_linear_fill0(x, y, x_new):
    """scipy 1.18.1 interp1d(kind='linear', bounds_error=False, fill_value=0):
    argsort x; searchsorted; clip to [1, n-1]; w_hi*y_hi + w_lo*y_lo; 0 outside [x0, x_last]."""
    x = np.asarray(x, dtype=float)
    y = np.asarray(y, dtype=float)
    order = np.argsort(x, kind="mergesort")
    x, y = x[order], y[order]
    x_new = np.asarray(x_new, dtype=float)
    idx = np.searchsorted(x, x_new).clip(1, len(x) - 1).astype(int)
    lo, hi = idx - 1, idx
    x_lo, x_hi = x[lo], x[hi]
    with np.errstate(invalid="ignore", divide="ignore"):
        out = ((x_new - x_lo) / (x_hi - x_lo)) * y[hi] + ((x_hi - x_new) / (x_hi - x_lo)) * y[lo]
    out[x_new < x[0]] = 0.0
    out[x_new > x[-1]] = 0.0
    return out


def _raw_intensity(range_theta, phase):
    """phases.py:97-104."""
    return _interp_linear_fill0(phase.raw_pattern_x, phase.raw_pattern_y, np.rad2deg(2 * range_theta))


def get_intensity(range_theta, range_stl, soller1, soller2, mcr_2theta, phase):
    """phases.py:81-95."""
    intensity = get_diffracted_intensity(range_theta, range_stl, phase)
    if phase.apply_lpf:
        return intensity * get_lorentz_polarisation_factor(range_theta, phase.sigma_star,
                                                           soller1, soller2, mcr_2theta)
    return intensity


# ---------------------------------------------------------------------------
# specimen.py
# ---------------------------------------------------------------------------
def get_machine_correction_range(specimen):
    """specimen.py:21-43."""
    gon = specimen.goniometer
    st = np.sin(specimen.range_theta)
    corr = np.ones_like(specimen.range_theta)
    if gon.divergence_mode == "AUTOMATIC":
        corr *= get_fixed_to_ads_correction_range(specimen.range_theta, gon)
    if gon.has_absorption_correction > 0.0:
        mu = gon.absorption * gon.sample_surf_density * 1e-3
        if mu > 0.0:
            corr *= np.minimum(1.0 - np.exp(-2.0 * mu / st), 1.0)
    if gon.divergence_mode == "FIXED" and gon.divergence > 0:
        L_Rta = gon.sample_length / (gon.radius * math.tan(math.radians(gon.divergence)))
        corr *= np.minimum(st * L_Rta, 1)
    return corr


def calculate_phase_intensities(specimen):
    """specimen.py:45-82 -> (correction f64[X], intensities f64[M, Z, X])."""
    gon = specimen.goniometer
    theta = np.asarray(specimen.range_theta, dtype=float)
    stl = 2 * np.sin(theta) / gon.wavelength
    corr = get_machine_correction_range(specimen)
    rows = []
    for z_index in range(len(specimen.z_list)):
        row = []
        for phase in specimen.phases[z_index]:
            if phase is None:
                I = np.zeros_like(stl)
            else:
                I = get_intensity(theta, stl, gon.soller1, gon.soller2, gon.mcr_2theta, phase) \
                    * (corr if phase.apply_correction else 1.0)
            # sequential, in-place accumulation over the wavelength list (specimen.py:66-74)
            for wavelength, fraction in gon.wavelength_distribution:
                shifted = np.arcsin(stl * wavelength * 0.5)
                I = I + _interp_linear_fill0(shifted, I * fraction, theta)
            row.append(I)
        rows.append(row)
    return corr, np.swapaxes(np.array(rows, dtype=float), 0, 1)


def calculate_scaled_intensities(specimen, scale, fractions, bgshift):
    """specimen.py:84-88."""
    specimen.background_intensity = bgshift * specimen.correction
    specimen.scaled_phase_intensities = (fractions * specimen.phase_intensities.transpose()).transpose() * scale
    specimen.total_intensity = np.sum(specimen.scaled_phase_intensities, axis=0) + specimen.background_intensity
    return specimen


def get_clipped_intensities(specimen):
    """specimen.py:16-19."""
    exp = specimen.observed_intensity.transpose()[..., specimen.selected_range]
    cal = specimen.total_intensity[..., specimen.selected_range]
    return exp, cal


# ---------------------------------------------------------------------------
# statistics.py
# ---------------------------------------------------------------------------
def Rp(exp, calc, *args):
    """statistics.py:22-27."""
    return np.sum(np.abs(exp - calc)) / np.sum(np.abs(exp)) * 100


def smooth_blackman(x, half_window_len):
    """math_tools.py:46-99 with window='blackman'."""
    n = 2 * half_window_len + 1
    x = np.ravel(x)
    if x.size < n:
        raise ValueError("Input vector needs to be bigger than window size.")
    padded = np.r_[x[n - 1:0:-1], x, x[-1:-n:-1]]
    w = np.blackman(n)
    y = np.convolve(w / w.sum(), padded, mode="valid")
    return y[half_window_len:-half_window_len]


def Rpder(exp, calc, *args):
    """statistics.py:35-48 -- Rp of smoothed first derivatives."""
    return Rp(np.gradient(smooth_blackman(exp, 15)), np.gradient(smooth_blackman(calc, 15)))


def Rpw(exp, calc, *args):
    """statistics.py:50-68."""
    exp, calc = np.ravel(exp), np.ravel(calc)
    sm1 = sm2 = 0
    for e, c in zip(exp, calc):
        with np.errstate(all="ignore"):
            t = (e - c) ** 2 / e
        if not (np.isnan(t) or np.isinf(t)):
            sm1 += t
            sm2 += abs(e)
    try:
        return math.sqrt(sm1 / sm2) * 100
    except Exception:
        return 0


def R_squared(exp, calc, *args):
    """statistics.py:12-20."""
    avg = sum(exp) / exp.size
    return 1 - (np.sum((exp - calc) ** 2) / np.sum((exp - avg) ** 2))


def derive(pattern):
    """statistics.py:35-41."""
    return np.gradient(smooth_blackman(pattern, 15))


def Rpe(exp, calc, num_params):
    """statistics.py:70-79."""
    return np.sqrt((exp.size - num_params) / np.sum(exp ** 2)) * 100


def Rphase(exp, calc, phase):
    """statistics.py:81-100."""
    exp, calc, phase = np.ravel(exp), np.ravel(calc), np.ravel(phase)
    sm1 = sm2 = 0
    for e, c, p in zip(exp, calc, phase):
        with np.errstate(all="ignore"):
            t = (e - c) ** 2 * p / (e ** 2)
        if not (np.isnan(t) or np.isinf(t)):
            sm1 += t
            sm2 += abs(p)
    try:
        return math.sqrt(sm1 / sm2) * 100
    except Exception:
        return 0


_RESIDUALS = {"Rp": Rp, "Rpw": Rpw, "Rpder": Rpder}


# ---------------------------------------------------------------------------
# mixture.py
# ---------------------------------------------------------------------------
def parse_solution(x, mixture):
    """mixture.py:32-48."""
    fractions, scales, bgshifts = mixture.fractions, mixture.scales, mixture.bgshifts
    scaled = x[:mixture.m] * (1.0 - mixture.sum_static_fractions) / np.sum(x[:mixture.m])
    np.put(fractions, mixture.selected_fractions, scaled)
    np.put(scales, mixture.selected_scales, x[mixture.m:mixture.m + mixture.n])
    np.put(bgshifts, mixture.selected_bgshifts, x[-mixture.o:])
    return fractions, scales, bgshifts


def get_solution(mixture):
    """mixture.py:50-55."""
    return np.concatenate((np.take(mixture.fractions, mixture.selected_fractions),
                           np.take(mixture.scales, mixture.selected_scales),
                           np.take(mixture.bgshifts, mixture.selected_bgshifts)))


def set_solution(x, mixture):
    """mixture.py:57-58."""
    mixture.fractions, mixture.scales, mixture.bgshifts = parse_solution(x, mixture)


def get_zero_solution(mixture):
    """mixture.py:60-68 (when o == 0 the slice [-0:] is the whole vector: all zeros first)."""
    x0 = np.ones(shape=(mixture.m + mixture.n + mixture.o,))
    x0[-mixture.o:] = 0.0
    x0[:mixture.m] = 1.0 / (1.0 - mixture.sum_static_fractions)
    return x0


def get_bounds(mixture):
    """mixture.py:73-77."""
    return ([(0., 1.)] * mixture.m) + ([(1e-3, None)] * mixture.n) + ([(0., None)] * mixture.o)


def _specimen_residual(specimen):
    """mixture.py:82-89."""
    exp, cal = get_clipped_intensities(specimen)
    return _RESIDUALS[SETTINGS["RESIDUAL_METHOD"]](exp, cal)


def parse_mixture(mixture, force=False):
    """mixture.py:103-148."""
    if mixture.parsed and not force:
        return
    mixture.parsed = False
    assert mixture.n > 0
    assert mixture.n == len(mixture.specimens)
    assert mixture.m > 0
    mixture.selected_fractions = np.asanyarray(np.nonzero(mixture.fractions_mask)).flatten()
    mixture.selected_scales = np.asanyarray(np.nonzero(mixture.scales_mask)).flatten()
    mixture.selected_bgshifts = np.asanyarray(np.nonzero(mixture.bgshifts_mask)).flatten()
    mixture.real_m, mixture.real_n, mixture.real_o = mixture.m, mixture.n, mixture.n
    mixture.m = len(mixture.selected_fractions)
    mixture.n = len(mixture.selected_scales)
    mixture.o = len(mixture.selected_bgshifts)
    sm = np.sum(mixture.fractions)
    mixture.sum_static_fractions = sm - np.sum(np.take(mixture.fractions, mixture.selected_fractions))
    if sm != 1.0 and mixture.sum_static_fractions < 0.0 or mixture.sum_static_fractions > 1.0:
        mixture.fractions = mixture.fractions / sm
        sm = 1.0
    mixture.sum_static_fractions = sm - np.sum(np.take(mixture.fractions, mixture.selected_fractions))
    mixture.z = 0
    for specimen in mixture.specimens:
        if specimen is not None:
            mixture.z = len(specimen.z_list)
            break
    assert mixture.z > 0
    for specimen in mixture.specimens:
        if specimen is not None:
            specimen.correction, specimen.phase_intensities = calculate_phase_intensities(specimen)
    mixture.parsed = True


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
    fractions = np.asanyarray(mixture.fractions)
    mixture.residuals = [0.0]
    for scale, bgshift, specimen in zip(mixture.scales, mixture.bgshifts, mixture.specimens):
        res = 0.0
        if specimen is not None:
            bgshift = bgshift if SETTINGS["BGSHIFT"] else 0.0
            calculate_scaled_intensities(specimen, scale, fractions, bgshift)
            if specimen.observed_intensity.size > 0:
                res = _specimen_residual(specimen)
        mixture.residuals.append(res)
    mixture.residuals[0] = np.average(mixture.residuals[1:])
    mixture.calculated = True
    return mixture


def _get_residuals(x, mixture):
    """mixture.py:91-94."""
    set_solution(x, mixture)
    return calculate_mixture(mixture).residuals


def get_residual(mixture):
    """mixture.py:96-97."""
    return _get_residuals(get_solution(mixture), mixture)


def optimize_mixture(mixture, force=False, record=None):
    """mixture.py:150-219 -- bounded L-BFGS-B (finite-difference gradient) over [fractions|scales|bg]."""
    from scipy.optimize import fmin_l_bfgs_b
    if mixture.optimized and not force:
        return mixture
    try:
        parse_mixture(mixture)
    except AssertionError:
        if SETTINGS["DEBUG"]:
            raise
        return mixture
    x0 = get_zero_solution(mixture)
    bounds = get_bounds(mixture)

    def objective(x):
        mixture.calculated = False
        return _get_residuals(x, mixture)[0]

    lastx, residual, info = fmin_l_bfgs_b(objective, x0, approx_grad=True, bounds=bounds)
    if record is not None:
        record.update(x=np.array(lastx), info=info)
    if np.isscalar(residual):
        residual = np.array([residual])
    fractions, scales, bgshifts = parse_solution(lastx, mixture)
    fractions = fractions.flatten()
    total = np.sum(fractions)
    if total == 0.0 and len(fractions) > 0:
        fractions[0] = 1.0
        total = 1.0
    fractions = np.around(fractions / total, 6)
    scales *= total
    scales = scales.round(6)
    mixture.fractions, mixture.scales, mixture.bgshifts = fractions, scales, bgshifts
    mixture.residual = residual
    mixture.calculated = False
    mixture.optimized = True
    return mixture


def calculate_and_optimize_mixture(mixture):
    """mixture.py:259-266."""
    return calculate_mixture(optimize_mixture(mixture))


def get_optimized_residual(mixture):
    """mixture.py:268-276 -- returns the optimiser's final objective (1-element array)."""
    return calculate_and_optimize_mixture(mixture).residual
