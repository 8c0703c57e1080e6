"""Deterministic synthetic phase / specimen / mixture definitions.

Builds the shapes named in BASELINE.json ``configs`` (SURVEY.md §8(d)
"Synthetic inputs") as plain DataObject graphs, exactly what the reference's
model layer would hand to ``pyxrd.calculations`` (data_objects.py:22-271).
Nothing here computes a diffraction pattern: the ``observed`` pattern of a
specimen is produced by whatever calculator the caller plugs in (the CUDA path
in ``bench.py``, the oracle in the CPU tests).

Everything is a pure function of ``(config name, seed, candidate index)`` so
the GPU box, this container and the golden-vector generator build bit-identical
inputs (``numpy.random.default_rng`` / PCG64 streams are platform independent).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from .data_objects import (AtomData, AtomTypeData, ComponentData, CSDSData,
                           GonioData, MixtureData, PhaseData, SpecimenData)

# Five-Gaussian scattering-factor coefficients (Waasmaier & Kirfel, Acta Cryst.
# A51 (1995) 416) for the species a clay layer needs: (weight, debye, c, a1-5, b1-5)
_WK = {
    "O2-":  (15.9994, 2.0, 0.025429, (3.990247, 2.300563, 0.6072, 1.907882, 1.16708),
             (16.639956, 5.636819, 0.108493, 47.299709, 0.379984)),
    "O1-":  (15.9994, 2.0, 0.046136, (3.106934, 3.235142, 1.148886, 0.783981, 0.676953),
             (19.86808, 6.960252, 0.170043, 65.693512, 0.630757)),
    "Si4+": (28.0855, 1.5, 0.097266, (3.676722, 3.828496, 1.258033, 0.419024, 0.720421),
             (1.446851, 3.013144, 0.064397, 0.206254, 5.970222)),
    "Al3+": (26.9815386, 1.5, 0.019397, (4.132015, 0.912049, 1.102425, 0.614876, 3.219136),
             (3.528641, 7.378344, 0.133708, 0.039065, 1.644728)),
    "Mg2+": (24.305, 1.5, 0.058851, (3.062918, 4.135106, 0.853742, 1.036792, 0.85252),
             (2.015803, 4.417941, 0.065307, 9.66971, 0.187818)),
    "Fe2+": (55.845, 1.5, -9.676919, (11.776765, 11.165097, 3.533495, 0.165345, 7.036932),
             (4.912232, 0.001748, 14.166556, 42.381958, 0.341324)),
    "K1+":  (39.0983, 2.0, 0.257164, (-17.609339, 1.494873, 7.150305, 10.899569, 15.808228),
             (18.840979, 0.053453, 0.81294, 22.264105, 14.351593)),
    "Ca2+": (40.078, 2.0, -21.013187, (8.501441, 12.880483, 9.765095, 7.156669, 0.71116),
             (10.525848, -0.004033, 0.010692, 0.684443, 27.231771)),
    "Na1+": (22.98976928, 1.5, 0.0453, (3.14869, 4.073989, 0.767888, 0.995612, 0.968249),
             (2.594987, 6.046925, 0.070139, 14.122657, 0.217037)),
    "H2O":  (18.01528, 2.0, 0.025429, (3.990247, 2.300563, 0.6072, 1.907882, 1.16708),
             (16.639956, 5.636819, 0.108493, 47.299709, 0.379984)),
    "H":    (1.00794, 0.0, 0.000049, (0.413048, 0.294953, 0.187491, 0.080701, 0.023736),
             (15.569946, 32.398468, 5.711404, 61.889874, 1.334118)),
}

DRITS = dict(alpha_scale=0.9485, alpha_offset=0.017, beta_scale=0.1032, beta_offset=0.0034)
CU_KA1 = 0.154056
CU_PAIR = [(0.1544426, 0.955148885), (0.153475, 0.044851115)]


def atom_types() -> dict:
    """One shared AtomTypeData per species (shared objects, like a Project's)."""
    out = {}
    for name, (weight, debye, c, a, b) in _WK.items():
        out[name] = AtomTypeData(par_a=np.array(a, dtype=float), par_b=np.array(b, dtype=float),
                                 par_c=float(c), debye=float(debye), weight=float(weight),
                                 charge=0.0, name=name)
    return out


# (species, pn, z as a fraction of the silicate-layer height) of a 2:1 layer, 11 planes
_LAYER_PLANES = [
    ("O2-", 6.0, 0.00), ("Si4+", 3.2, 0.09), ("Al3+", 0.8, 0.09), ("O2-", 4.0, 0.335),
    ("O1-", 2.0, 0.335), ("Al3+", 3.4, 0.50), ("Mg2+", 0.4, 0.50), ("Fe2+", 0.2, 0.50),
    ("O2-", 4.0, 0.665), ("Si4+", 4.0, 0.91), ("O2-", 6.0, 1.00),
]


def make_component(types: dict, kind: str, rng: Optional[np.random.Generator] = None,
                   d001: Optional[float] = None, delta_c: float = 0.0) -> ComponentData:
    """A layer type: 11 silicate-layer planes + 0/3/4/7 interlayer planes.

    ``kind``: illite | smectite_ad | smectite_eg | chlorite | kaolinite.
    ``rng`` (optional) jitters occupancies a little so components differ.
    """
    lattice_d = 0.66
    spec = {
        "illite":      (0.998, [("K1+", 1.5, 0.5)]),
        "smectite_ad": (1.25,  [("H2O", 1.6, 0.32), ("Ca2+", 0.35, 0.5), ("H2O", 1.6, 0.68)]),
        "smectite_eg": (1.69,  [("H2O", 0.8, 0.2), ("O1-", 1.7, 0.33), ("Ca2+", 0.35, 0.5),
                                ("O1-", 1.7, 0.67), ("H2O", 0.8, 0.8), ("H", 3.4, 0.4), ("H", 3.4, 0.6)]),
        "chlorite":    (1.42,  [("O1-", 6.0, 0.27), ("Mg2+", 4.0, 0.5), ("Al3+", 2.0, 0.5), ("O1-", 6.0, 0.73)]),
        "kaolinite":   (0.715, []),
    }[kind]
    default_c, inter = spec
    if kind == "kaolinite":
        lattice_d = 0.44
    jit = (lambda: 1.0) if rng is None else (lambda: float(rng.uniform(0.97, 1.03)))
    layer_atoms, weight = [], 0.0
    planes = _LAYER_PLANES if kind != "kaolinite" else _LAYER_PLANES[:7]
    zmax = max(p[2] for p in planes)
    for name, pn, zf in planes:
        pn_j = pn * jit()
        layer_atoms.append(AtomData(atom_type=types[name], pn=pn_j, default_z=lattice_d * zf / zmax,
                                    z=lattice_d * zf / zmax, stretch_z=False))
        weight += pn_j * types[name].weight
    interlayer_atoms = []
    for name, pn, f in inter:
        pn_j = pn * jit()
        z0 = lattice_d + f * (default_c - lattice_d)
        interlayer_atoms.append(AtomData(atom_type=types[name], pn=pn_j, default_z=z0, z=z0, stretch_z=True))
        weight += pn_j * types[name].weight
    d = default_c if d001 is None else float(d001)
    return ComponentData(layer_atoms=layer_atoms, interlayer_atoms=interlayer_atoms,
                         volume=0.52 * 0.90 * d, weight=weight, d001=d, default_c=default_c,
                         delta_c=float(delta_c), lattice_d=lattice_d, name=kind)


def make_csds(average: float, maximum: Optional[int] = None, minimum: int = 1) -> CSDSData:
    """Drits log-normal CSDS; maximum = int(2.5*average) as the model layer sets it
    (phases/models/CSDS.py:127,140 + settings.py:26)."""
    if maximum is None:
        maximum = int(2.5 * average)
    return CSDSData(average=float(average), maximum=int(maximum), minimum=int(minimum), **DRITS)


def _stationary(P: np.ndarray) -> np.ndarray:
    """Stationary row vector of a row-stochastic matrix (deterministic power iteration)."""
    w = np.full(P.shape[0], 1.0 / P.shape[0])
    for _ in range(20000):
        w2 = w @ P
        w2 /= w2.sum()
        if np.max(np.abs(w2 - w)) < 1e-16:
            w = w2
            break
        w = w2
    return w


def make_probabilities(rng: np.random.Generator, G: int, R: int, dense: bool = False):
    """(W, P) of rank r = G**max(R,1): W diagonal, P row-stochastic.

    R0: P[i,j] = W_j.  R>=1: Markov chain over R-tuples (i1..iR) -> (i2..iR,j);
    ``dense=True`` instead draws a dense random row-stochastic P (config 3 in
    SURVEY.md §8(d): an r=27 phase the reference's model layer cannot itself make,
    but which its calculations accept, phases.py:125-135).
    """
    r = G ** max(R, 1)
    if R == 0:
        w = rng.dirichlet(np.full(G, 4.0))
        P = np.repeat(w[np.newaxis, :], G, axis=0)
        return np.diag(w), P
    if dense:
        P = rng.dirichlet(np.full(r, 0.6), size=r)
    else:
        P = np.zeros((r, r))
        sub = G ** (R - 1)
        for I in range(r):
            tail = I % sub                    # (i2..iR)
            probs = rng.dirichlet(np.full(G, 2.5))
            for j in range(G):
                P[I, tail * G + j] = probs[j]
    w = _stationary(P)
    return np.diag(w), P


def make_goniometer(wavelengths: str = "single", divergence_mode: str = "FIXED",
                    absorption: bool = False) -> GonioData:
    if wavelengths == "single":
        wl, dist = CU_KA1, [(CU_KA1, 1.0)]
    elif wavelengths == "cu_pair":
        wl, dist = CU_PAIR[0][0], list(CU_PAIR)
    else:
        raise ValueError(wavelengths)
    return GonioData(min_2theta=3.0, max_2theta=45.0, steps=2500, soller1=2.3, soller2=2.3,
                     has_soller1=True, has_soller2=True, divergence_mode=divergence_mode,
                     divergence=0.5, mcr_2theta=0.0, has_absorption_correction=bool(absorption),
                     absorption=45.0, sample_surf_density=20.0, radius=24.0, wavelength=wl,
                     wavelength_distribution=dist, sample_length=1.25)


def theta_grid(lo: float, hi: float, step: float) -> np.ndarray:
    """theta (radians) for a 2-theta range in degrees, as SURVEY.md §8(d) defines it."""
    return np.radians(np.arange(lo, hi, step) / 2.0)


def make_phase(components: Sequence[ComponentData], W, P, csds: CSDSData, sigma_star: float = 3.0,
               apply_lpf: bool = True, apply_correction: bool = True, valid: bool = True) -> PhaseData:
    return PhaseData(type="Phase", valid_probs=bool(valid), apply_lpf=apply_lpf,
                     apply_correction=apply_correction, components=list(components),
                     sigma_star=float(sigma_star), CSDS=csds, G=len(components),
                     W=np.asarray(W, dtype=float), P=np.asarray(P, dtype=float))


def make_specimen(theta: np.ndarray, phases: Sequence[Optional[PhaseData]], goniometer: GonioData,
                  observed: Optional[np.ndarray] = None, exclusion: Optional[tuple] = None) -> SpecimenData:
    """One pattern per specimen (Z = 1).  ``exclusion`` = (lo, hi) 2-theta degrees to mask out."""
    sel = np.ones(theta.shape, dtype=bool)
    if exclusion is not None:
        tt = np.degrees(2.0 * theta)
        sel &= ~((tt >= exclusion[0]) & (tt <= exclusion[1]))
    if observed is None:
        observed = np.zeros((theta.size, 1))
    return SpecimenData(goniometer=goniometer, range_theta=np.array(theta, dtype=float),
                        selected_range=sel, z_list=[0.0], phases=[list(phases)],
                        observed_intensity=np.asarray(observed, dtype=float))


def make_mixture(specimens: Sequence[SpecimenData], n_phases: int, fractions=None, scales=None,
                 bgshifts=None, refine_bg: bool = True) -> MixtureData:
    n = len(specimens)
    fractions = np.full(n_phases, 1.0 / n_phases) if fractions is None else np.asarray(fractions, float)
    scales = np.ones(n) if scales is None else np.asarray(scales, float)
    bgshifts = np.zeros(n) if bgshifts is None else np.asarray(bgshifts, float)
    return MixtureData(specimens=list(specimens), fractions=fractions, scales=scales, bgshifts=bgshifts,
                       fractions_mask=np.ones(n_phases), scales_mask=np.ones(n),
                       bgshifts_mask=np.ones(n) if refine_bg else np.zeros(n),
                       n=n, m=n_phases, parsed=False, calculated=False, optimized=False)


# ---------------------------------------------------------------------------
# The five BASELINE.json configurations
# ---------------------------------------------------------------------------
class Config(object):
    """name, theta grids, builder(candidate_rng | None) -> MixtureData."""

    def __init__(self, name, description, build, n_specimens, n_phases):
        self.name, self.description, self.build = name, description, build
        self.n_specimens, self.n_phases = n_specimens, n_phases

    def truth(self) -> MixtureData:
        """The hidden 'true' parameter set (no perturbation)."""
        return self.build(None)

    def candidate(self, seed: int, index: int) -> MixtureData:
        """Candidate ``index`` of the population drawn with ``seed`` (independent streams)."""
        return self.build(np.random.default_rng([int(seed), int(index)]))

    def population(self, seed: int, count: int, start: int = 0) -> List[MixtureData]:
        return [self.candidate(seed, start + i) for i in range(count)]


def _u(rng, lo, hi, centre):
    return centre if rng is None else float(rng.uniform(lo, hi))


def _c1(rng):
    """Single-phase R0 illite, 1 component, CSDS mean 10, 2theta 3-45 step 0.02 (X=2100)."""
    types = atom_types()
    comp = make_component(types, "illite", d001=_u(rng, 0.99, 1.01, 0.998))
    W, P = np.array([[1.0]]), np.array([[1.0]])
    phase = make_phase([comp], W, P, make_csds(_u(rng, 8.0, 12.9, 10.0)), sigma_star=_u(rng, 1.0, 15.0, 3.0))
    spec = make_specimen(theta_grid(3.0, 45.0, 0.02), [phase], make_goniometer("single"))
    return make_mixture([spec], 1)


def _c2(rng):
    """Two-component illite-smectite R1, air-dry, 2theta 3-70 step 0.01 (X=6700)."""
    types = atom_types()
    base = np.random.default_rng(1002)
    illite = make_component(types, "illite", d001=_u(rng, 0.99, 1.005, 0.998))
    smect = make_component(types, "smectite_ad", d001=_u(rng, 1.20, 1.30, 1.25), delta_c=0.02)
    prng = base if rng is None else rng
    w1 = float(prng.uniform(0.5, 0.9))
    p_min = max(0.0, (2.0 * w1 - 1.0) / w1)   # keeps P21 = W1 P12 / W2 <= 1
    p11 = float(prng.uniform(p_min + 1e-3, 1.0 - 1e-3))
    w2 = 1.0 - w1
    p12 = 1.0 - p11
    p21 = w1 * p12 / w2
    P = np.array([[p11, p12], [p21, 1.0 - p21]])
    W = np.diag([w1, w2])
    phase = make_phase([illite, smect], W, P, make_csds(_u(rng, 8.0, 12.9, 10.0)),
                       sigma_star=_u(rng, 1.0, 15.0, 3.0))
    spec = make_specimen(theta_grid(3.0, 70.0, 0.01), [phase], make_goniometer("cu_pair"))
    return make_mixture([spec], 1)


def _c3(rng):
    """Three-component 'R3' illite-chlorite-smectite, g^R = 27 dense, CSDS mean 30 / max 120."""
    types = atom_types()
    base = np.random.default_rng(1003)
    comps = [make_component(types, "illite", d001=_u(rng, 0.99, 1.005, 0.998)),
             make_component(types, "chlorite", d001=_u(rng, 1.41, 1.43, 1.42)),
             make_component(types, "smectite_eg", d001=_u(rng, 1.66, 1.72, 1.69), delta_c=0.02)]
    W, P = make_probabilities(base if rng is None else rng, 3, 3, dense=True)
    phase = make_phase(comps, W, P, make_csds(_u(rng, 28.0, 32.0, 30.0), maximum=120),
                       sigma_star=_u(rng, 1.0, 15.0, 3.0))
    spec = make_specimen(theta_grid(3.0, 45.0, 0.02), [phase], make_goniometer("single"))
    return make_mixture([spec], 1)


def _c3s(rng):
    """c3 with the transition structure a real Reichweite-3 chain of three components has: state (i1, i2, i3) only
    reaches (i2, i3, j) -- 3 non-zero entries per row of the 27 x 27 matrix instead of 27."""
    types = atom_types()
    base = np.random.default_rng(1003)
    comps = [make_component(types, "illite", d001=_u(rng, 0.99, 1.005, 0.998)),
             make_component(types, "chlorite", d001=_u(rng, 1.41, 1.43, 1.42)),
             make_component(types, "smectite_eg", d001=_u(rng, 1.66, 1.72, 1.69), delta_c=0.02)]
    W, P = make_probabilities(base if rng is None else rng, 3, 3, dense=False)
    phase = make_phase(comps, W, P, make_csds(_u(rng, 28.0, 32.0, 30.0), maximum=120),
                       sigma_star=_u(rng, 1.0, 15.0, 3.0))
    spec = make_specimen(theta_grid(3.0, 45.0, 0.02), [phase], make_goniometer("single"))
    return make_mixture([spec], 1)


def _mixed_layer_set(rng, base_seed, states):
    """Shared-parameter phase set for an (AD, EG) specimen pair: each phase is built
    twice with the same CSDS / sigma* / W / P and different smectite hydration state."""
    types = atom_types()
    base = np.random.default_rng(base_seed)
    prng = base if rng is None else rng
    out = {s: [] for s in states}

    def both(fn):
        args = fn()
        for s in states:
            out[s].append(args(s))

    def smec(s, **kw):
        return make_component(types, "smectite_ad" if s == "AD" else "smectite_eg", **kw)

    # 1 illite R0 G1
    def p1():
        d, avg, ss = _u(rng, 0.995, 1.002, 0.998), _u(rng, 14.0, 22.0, 18.0), _u(rng, 2.0, 10.0, 4.0)
        return lambda s: make_phase([make_component(types, "illite", d001=d)], [[1.0]], [[1.0]],
                                    make_csds(avg), sigma_star=ss)
    # 2 kaolinite R0 G1
    def p2():
        avg, ss = _u(rng, 20.0, 30.0, 25.0), _u(rng, 2.0, 10.0, 5.0)
        return lambda s: make_phase([make_component(types, "kaolinite")], [[1.0]], [[1.0]],
                                    make_csds(avg), sigma_star=ss)
    # 3 chlorite R0 G1
    def p3():
        avg, ss = _u(rng, 10.0, 18.0, 14.0), _u(rng, 2.0, 10.0, 6.0)
        return lambda s: make_phase([make_component(types, "chlorite")], [[1.0]], [[1.0]],
                                    make_csds(avg), sigma_star=ss)
    # 4 illite-smectite R1 G2
    def p4():
        W, P = make_probabilities(prng, 2, 1)
        avg, ss = _u(rng, 8.0, 14.0, 10.0), _u(rng, 2.0, 12.0, 8.0)
        dad, deg = _u(rng, 1.22, 1.28, 1.25), _u(rng, 1.66, 1.72, 1.69)
        return lambda s: make_phase([make_component(types, "illite"),
                                     smec(s, d001=dad if s == "AD" else deg, delta_c=0.02)],
                                    W, P, make_csds(avg), sigma_star=ss)
    # 5 smectite-rich illite-smectite R0 G2
    def p5():
        W, P = make_probabilities(prng, 2, 0)
        avg, ss = _u(rng, 5.0, 9.0, 7.0), _u(rng, 4.0, 14.0, 10.0)
        return lambda s: make_phase([make_component(types, "illite"), smec(s, delta_c=0.03)],
                                    W, P, make_csds(avg), sigma_star=ss)
    # 6 illite-smectite R2 G2 (rank 4)
    def p6():
        W, P = make_probabilities(prng, 2, 2)
        avg, ss = _u(rng, 8.0, 16.0, 12.0), _u(rng, 2.0, 12.0, 7.0)
        return lambda s: make_phase([make_component(types, "illite"), smec(s, delta_c=0.02)],
                                    W, P, make_csds(avg), sigma_star=ss)
    for fn in (p1, p2, p3, p4, p5, p6):
        both(fn)
    return out


def _c4(rng):
    """Two specimens (air-dry + glycolated) x 6 shared-parameter phases, 2theta 3-45 step 0.02."""
    sets = _mixed_layer_set(rng, 1004, ("AD", "EG"))
    theta = theta_grid(3.0, 45.0, 0.02)
    specs = [make_specimen(theta, sets["AD"], make_goniometer("cu_pair")),
             make_specimen(theta, sets["EG"], make_goniometer("cu_pair"), exclusion=(26.0, 28.0))]
    return make_mixture(specs, 6, fractions=[0.3, 0.1, 0.1, 0.25, 0.1, 0.15])


def _c5(rng):
    """One specimen x 4 phases, each an R3 G2 illite-smectite (rank 8), CSDS ~30 (max 75)."""
    types = atom_types()
    base = np.random.default_rng(1005)
    prng = base if rng is None else rng
    phases = []
    for i in range(4):
        W, P = make_probabilities(prng, 2, 3)
        comps = [make_component(types, "illite", d001=_u(rng, 0.995, 1.002, 0.998)),
                 make_component(types, "smectite_eg", d001=_u(rng, 1.66, 1.72, 1.69), delta_c=0.02)]
        phases.append(make_phase(comps, W, P, make_csds(_u(rng, 28.0, 32.0, 30.0)),
                                 sigma_star=_u(rng, 2.0, 12.0, 3.0 + 2 * i)))
    spec = make_specimen(theta_grid(3.0, 45.0, 0.02), phases, make_goniometer("cu_pair"))
    return make_mixture([spec], 4)


def make_raw_phase(x, y, apply_lpf: bool = False, apply_correction: bool = False, sigma_star: float = 3.0) -> PhaseData:
    """RawPatternPhase data object as phases/models/raw_pattern_phase.py:36-43 builds it."""
    return PhaseData(type="RawPatternPhase", valid_probs=True, apply_lpf=apply_lpf, apply_correction=apply_correction,
                     components=[], sigma_star=float(sigma_star), raw_pattern_x=np.asarray(x, dtype=float),
                     raw_pattern_y=np.asarray(y, dtype=float))


def raw_case() -> MixtureData:
    """c1's specimen with two RawPatternPhases next to the calculated illite: an unsorted, coarser raw
    abscissa that covers only part of the 2-theta range (fill 0 outside), the model's defaults
    (no LPF, no correction) for one and both switched on for the other."""
    mix = _c1(None)
    spec = mix.specimens[0]
    rng = np.random.default_rng(4242)
    x = np.linspace(5.013, 39.77, 811)
    y = 40.0 + 25.0 * np.sin(x * 0.9) ** 2 + rng.uniform(0.0, 3.0, x.size)
    shuffle = rng.permutation(x.size)
    raw_a = make_raw_phase(x[shuffle], y[shuffle])
    x2 = np.sort(rng.uniform(2.0, 60.0, 300))
    raw_b = make_raw_phase(x2, rng.uniform(0.0, 80.0, x2.size), apply_lpf=True, apply_correction=True, sigma_star=7.5)
    spec.phases = [[spec.phases[0][0], raw_a, raw_b]]
    out = make_mixture([spec], 3, fractions=[0.5, 0.3, 0.2], scales=[1.7], bgshifts=[2.5])
    spec.observed_intensity = rng.uniform(20.0, 90.0, (spec.range_theta.size, 1))
    return out


def z2_case() -> MixtureData:
    """two patterns per specimen (``z_list`` of length 2: specimen.py:52-63 loops over it, ``observed_intensity`` is
    [X, 2], ``phases`` is [Z][M]): c4's air-dry specimen carries its own phase set as pattern 0 and the glycolated
    set as pattern 1; the second specimen keeps one ... two patterns as well (both of its own set, one phase None)"""
    mix = _c4(None)
    ad, eg = mix.specimens
    theta = ad.range_theta
    rng = np.random.default_rng(777)
    row_ad, row_eg = list(ad.phases[0]), list(eg.phases[0])
    row_b = list(row_eg)
    row_b[2] = None
    specs = []
    for rows, gon, excl in (([row_ad, row_eg], ad.goniometer, None), ([row_eg, row_b], eg.goniometer, (26.0, 28.0))):
        sp = make_specimen(theta, rows[0], gon, exclusion=excl)
        sp.phases = [list(r) for r in rows]
        sp.z_list = [0.0, 1.0]
        tt = np.degrees(2.0 * theta)
        sp.observed_intensity = np.stack([40.0 + 25.0 * np.sin(tt * 0.8) ** 2 + rng.uniform(0.0, 5.0, tt.size),
                                          55.0 + 20.0 * np.cos(tt * 0.5) ** 2 + rng.uniform(0.0, 5.0, tt.size)], axis=1)
        specs.append(sp)
    return make_mixture(specs, 6, fractions=[0.3, 0.1, 0.1, 0.25, 0.1, 0.15], scales=[1.3, 0.8], bgshifts=[2.0, 4.0])


CONFIGS = {
    "c1": Config("c1", "single-phase R0 illite, G=1, CSDS mean 10, X=2100", _c1, 1, 1),
    "c2": Config("c2", "illite-smectite R1 G=2, air-dry, X=6700", _c2, 1, 1),
    "c3": Config("c3", "illite-chlorite-smectite rank 27, CSDS mean 30 max 120, X=2100", _c3, 1, 1),
    "c3s": Config("c3s", "c3 with the Reichweite-3 transition structure (3 non-zeros per row of the 27 x 27 matrix)", _c3s, 1, 1),
    "c4": Config("c4", "AD+EG specimens x 6 shared-parameter phases (CMA-ES population)", _c4, 2, 6),
    "c5": Config("c5", "4-phase R3 (rank 8) mixture population sweep", _c5, 1, 4),
}


def attach_observed(candidates: Sequence[MixtureData], truth: MixtureData,
                    calc_total: Callable[[MixtureData], List[np.ndarray]],
                    noise_seed: int = 1234, noise: float = 0.01, scale: float = 1000.0) -> None:
    """Give every candidate the same 'observed' patterns.

    ``calc_total(truth)`` must return, per specimen, the total intensity f64[Z, X]
    of the true parameter set (with its own fractions / unit scales); the pattern is
    normalised to ``scale`` counts at its maximum, a flat background of 2 % and
    ``noise`` relative Gaussian noise (seed 1234) are added.
    """
    totals = calc_total(truth)
    rng = np.random.default_rng(noise_seed)
    observed = []
    for tot in totals:
        tot = np.asarray(tot, dtype=float)
        y = tot / max(float(np.max(tot)), 1e-300) * scale
        y = y + 0.02 * scale
        y = y + rng.standard_normal(y.shape) * noise * scale
        observed.append(np.ascontiguousarray(np.abs(y).T))  # stored as [X, Z] like the model layer
    for mix in list(candidates) + [truth]:
        for spec, obs in zip(mix.specimens, observed):
            spec.observed_intensity = obs.copy()


def rpder_case() -> MixtureData:
    """c4 candidate 1 with a pseudo-observed pattern and a 2-degree exclusion window on specimen 0."""
    mix = CONFIGS["c4"].candidate(2024, 1)
    rng = np.random.default_rng(515)
    for s, spec in enumerate(mix.specimens):
        tt = np.degrees(2.0 * spec.range_theta)
        spec.observed_intensity = (50.0 + 30.0 * np.sin(tt * 0.7) ** 2 + rng.uniform(0.0, 4.0, tt.size)).reshape(-1, 1)
        if s == 0:
            spec.selected_range = ~((tt >= 20.0) & (tt <= 22.0))
    return mix


# ---------------------------------------------------------------------------
# Templates + refinables of the configurations (population path)
# ---------------------------------------------------------------------------
def _phase_refinables(tag, phases, avg_range, ss_range=(1.0, 15.0)):
    """CSDS average (+ automatic maximum) and sigma* of a phase; ``phases`` = the DataObjects that share
    the values (the AD / EG copies of one model phase)."""
    from .population import Refinable
    return [Refinable(tag + ".csds_average", [(p.CSDS, "average") for p in phases], *avg_range),
            Refinable(tag + ".sigma_star", [(p, "sigma_star") for p in phases], *ss_range)]


def _d001_refinable(tag, comps, lo, hi):
    """d001 of a component; the cell volume follows it as make_component defines it (0.52 * 0.90 * d001)"""
    from .population import Refinable
    targets = []
    for c in comps:
        targets += [(c, "d001"), (c, "volume", 0.52 * 0.90, 0.0)]
    return Refinable(tag + ".d001", targets, lo, hi)


def _prob_refinables(tag, phases, model, params, bounds):
    from .population import Refinable
    for p in phases:
        p.prob_model, p.prob_params = model, list(params)
    return [Refinable("%s.prob%d" % (tag, i), [(p, "prob:%d" % i) for p in phases], lo, hi)
            for i, (lo, hi) in enumerate(bounds)]


def template(name: str):
    """(template MixtureData, [Refinable]) of a configuration: the truth candidate with its probability
    matrices handed over to the device models (``prob_model`` / ``prob_params`` on the PhaseData) and the
    refinable parameters a refinement of that configuration would vary."""
    if name == "models":
        return _models_template()
    mix = CONFIGS[name].truth()
    refs = []
    if name == "c1":
        ph = mix.specimens[0].phases[0][0]
        refs += [_d001_refinable("illite", [ph.components[0]], 0.99, 1.01)]
        refs += _phase_refinables("illite", [ph], (8.0, 12.9))
    elif name == "c2":
        ph = mix.specimens[0].phases[0][0]
        refs += [_d001_refinable("illite", [ph.components[0]], 0.99, 1.005),
                 _d001_refinable("smectite", [ph.components[1]], 1.20, 1.30)]
        refs += _prob_refinables("IS", [ph], "R1G2", [0.7, 0.7], [(0.5, 0.9), (0.05, 0.95)])
        refs += _phase_refinables("IS", [ph], (8.0, 12.9))
    elif name == "c4":
        ad, eg = mix.specimens[0].phases[0], mix.specimens[1].phases[0]
        pairs = [[a, e] for a, e in zip(ad, eg)]
        refs += [_d001_refinable("illite", [p.components[0] for p in pairs[0]], 0.995, 1.002)]
        refs += _phase_refinables("illite", pairs[0], (14.0, 22.0), (2.0, 10.0))
        refs += _phase_refinables("kaolinite", pairs[1], (20.0, 30.0), (2.0, 10.0))
        refs += _phase_refinables("chlorite", pairs[2], (10.0, 18.0), (2.0, 10.0))
        refs += _prob_refinables("IS_R1", pairs[3], "R1G2", [0.6, 0.7], [(0.3, 0.9), (0.05, 0.95)])
        refs += [_d001_refinable("IS_R1.smectite_ad", [pairs[3][0].components[1]], 1.22, 1.28),
                 _d001_refinable("IS_R1.smectite_eg", [pairs[3][1].components[1]], 1.66, 1.72)]
        refs += _phase_refinables("IS_R1", pairs[3], (8.0, 14.0), (2.0, 12.0))
        refs += _prob_refinables("IS_R0", pairs[4], "R0", [0.4], [(0.05, 0.95)])
        refs += _phase_refinables("IS_R0", pairs[4], (5.0, 9.0), (4.0, 14.0))
        refs += _prob_refinables("IS_R2", pairs[5], "R2G2", [0.7, 0.6, 0.5, 0.4],
                                 [(0.5, 0.95), (0.05, 0.95), (0.05, 0.95), (0.05, 0.95)])
        refs += _phase_refinables("IS_R2", pairs[5], (8.0, 16.0), (2.0, 12.0))
    elif name == "c5":
        for i, ph in enumerate(mix.specimens[0].phases[0]):
            tag = "IS_R3_%d" % i
            refs += _prob_refinables(tag, [ph], "R3G2", [0.8, 0.6], [(0.67, 0.95), (0.05, 0.95)])
            refs += [_d001_refinable(tag + ".illite", [ph.components[0]], 0.995, 1.002),
                     _d001_refinable(tag + ".smectite", [ph.components[1]], 1.66, 1.72)]
            refs += _phase_refinables(tag, [ph], (28.0, 32.0), (2.0, 12.0))
    else:
        raise KeyError("configuration %r has no template (its W / P are not produced by a probability model)" % name)
    return mix, refs


def _models_template():
    """one specimen with a phase of each multi-component probability model the reference ships beyond the
    BASELINE configurations: R1G3 (rank 3), R1G4 (rank 4), R2G3 (rank 9); every probability parameter refinable.
    The ranges keep part of the samples inside [0, 1]; the rest exercise valid_probs = False."""
    types = atom_types()
    kinds = ["illite", "chlorite", "smectite_ad", "kaolinite"]
    refs, phases = [], []
    for tag, model, G, r, start, bounds in (
            ("R1G3", "R1G3", 3, 3, [0.6, 0.7, 0.6, 0.5, 0.5, 0.5], [(0.3, 0.9)] + [(0.2, 0.9)] * 5),
            ("R1G4", "R1G4", 4, 4, [0.6, 0.5, 0.5, 0.5, 0.5, 0.4, 0.5, 0.5, 0.8, 0.75, 0.7, 0.5], [(0.3, 0.9)] + [(0.2, 0.9)] * 11),
            ("R2G3", "R2G3", 3, 9, [0.8, 0.9, 0.9, 0.9, 0.9, 0.9], [(0.55, 0.95)] + [(0.5, 0.98)] * 5)):
        comps = [make_component(types, k, delta_c=0.02 if "smectite" in k else 0.0) for k in kinds[:G]]
        ph = make_phase(comps, np.eye(r) / r, np.full((r, r), 1.0 / r), make_csds(9.0), sigma_star=5.0)
        refs += _prob_refinables(tag, [ph], model, start, bounds)
        refs += _phase_refinables(tag, [ph], (7.0, 12.0), (2.0, 12.0))
        phases.append(ph)
    spec = make_specimen(theta_grid(3.0, 40.0, 0.05), phases, make_goniometer("cu_pair"))
    return make_mixture([spec], 3, fractions=[0.4, 0.3, 0.3]), refs


def sample_solutions(refinables, seed: int, count: int) -> np.ndarray:
    """x f64[count, d] uniform inside the refinables' bounds (refinement_generator.py:26-33 style)"""
    rng = np.random.default_rng([int(seed), 77])
    lo = np.array([r.minimum for r in refinables])
    hi = np.array([r.maximum for r in refinables])
    return lo + rng.uniform(0.0, 1.0, (int(count), lo.size)) * (hi - lo)
