#!/usr/bin/env python
"""Generate the committed golden vectors by running the LIVE reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

Inputs are the deterministic synthetic configs of ``pyxrd_b200.synthetic``
(rebuilt identically by the tests); outputs are what the unmodified reference
``pyxrd.calculations`` (v0.8.4, behind the 4-alias compatibility shim of
``_live_reference.py``) returns for them with the numpy / scipy of this image.
"""
import os
import sys
import time

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

import _live_reference  # noqa: E402
from pyxrd_b200 import synthetic as syn  # noqa: E402

N_CANDIDATES = {"c1": 3, "c2": 3, "c3": 1, "c4": 3, "c5": 2}
SEED = 2024


def fingerprint(mix):
    """A few numbers that change if the synthetic builders drift."""
    acc = []
    for spec in mix.specimens:
        acc += [float(np.sum(spec.range_theta)), float(spec.goniometer.wavelength)]
        for phase in spec.phases[0]:
            acc += [float(np.sum(phase.P)), float(np.trace(phase.W)), float(phase.sigma_star),
                    float(phase.CSDS.average), float(phase.CSDS.maximum)]
            for comp in phase.components:
                acc += [comp.d001, comp.weight, comp.volume, comp.delta_c]
                acc += [a.pn for a in comp.layer_atoms + comp.interlayer_atoms]
                acc += [a.default_z for a in comp.layer_atoms + comp.interlayer_atoms]
    return np.array(acc, dtype=float)


def kats(ref, out):
    """The reference's own known-answer vectors, re-evaluated live."""
    theta = np.asanyarray([2.2551711385, 2.478038901, 2.7001518288, 2.9214422642, 3.1418428,
                           3.3612863, 3.5797059197, 3.7970351263])
    out["kat_theta"] = theta
    out["kat_lpf"] = ref.goniometer.get_lorentz_polarisation_factor(theta, 12, 2.3, 2.3, 0)
    out["kat_ads"] = ref.goniometer.get_fixed_to_ads_correction_range(theta, None)
    out["kat_S"] = np.array(ref.goniometer.get_S(2.5, 2.5))
    # test_atom_type.py:58-75 goes through AtomType.get_atomic_scattering_factors, which squares
    # (stl * 0.05); here the calculation-layer function is called on the same abscissa
    from pyxrd_b200.data_objects import AtomTypeData, AtomData, ComponentData, CSDSData
    at = AtomTypeData(par_a=np.full(5, 10.2), par_b=np.full(5, 10.2), par_c=20.2, debye=0.0)
    stl = np.array([0.1152, 0.5756, 1.1469])
    out["kat_asf_stl"] = stl
    out["kat_asf"] = ref.atoms.get_atomic_scattering_factor((stl * 0.05) ** 2, at)
    # test_components.py:24-62 component (H atoms) -- values unpinned upstream, pinned here
    h = AtomTypeData(par_a=np.asanyarray([0.413048, 0.294953, 0.187491, 0.080701, 0.023736]),
                     par_b=np.asanyarray([15.569946, 32.398468, 5.711404, 61.889874, 1.334118]),
                     par_c=0.000049, debye=0)
    comp = ComponentData(layer_atoms=[AtomData(atom_type=h, pn=1, default_z=0, z=0)],
                         interlayer_atoms=[AtomData(atom_type=h, pn=1, default_z=0.55, z=0.55)],
                         volume=1.0, weight=1.0, d001=0.55, default_c=0.6, delta_c=0.02, lattice_d=0.45)
    sf, phi = ref.components.get_factors(theta, comp)
    out["kat_comp_sf"], out["kat_comp_phi"] = sf, phi
    # test_CSDS.py:26-35 distribution
    csds = CSDSData(average=10, maximum=50, minimum=1, alpha_scale=0.9485, alpha_offset=0.017,
                    beta_scale=0.1032, beta_offset=0.0034)
    arr, mean = ref.CSDS.calculate_distribution(csds)
    out["kat_csds_arr"], out["kat_csds_mean"] = arr, np.array(mean)


def run_config(ref, name, out):
    cfg = syn.CONFIGS[name]
    truth = cfg.truth()
    cands = cfg.population(SEED, N_CANDIDATES[name])

    def calc_total(mix):
        ref.mixture.calculate_mixture(mix)
        return [np.array(s.total_intensity) for s in mix.specimens]

    t0 = time.time()
    syn.attach_observed(cands, truth, calc_total)
    for s, spec in enumerate(truth.specimens):
        out["%s_observed_s%d" % (name, s)] = spec.observed_intensity
    allmix = [("truth", truth)] + [("k%d" % i, m) for i, m in enumerate(cands)]
    rng = np.random.default_rng(77)
    for tag, mix in allmix:
        out["%s_%s_fp" % (name, tag)] = fingerprint(mix)
        mix.parsed = mix.calculated = mix.optimized = False
        ref.mixture.calculate_mixture(mix)
        for s, spec in enumerate(mix.specimens):
            out["%s_%s_corr_s%d" % (name, tag, s)] = np.array(spec.correction)
            out["%s_%s_pi_s%d" % (name, tag, s)] = np.array(spec.phase_intensities)
            out["%s_%s_total_s%d" % (name, tag, s)] = np.array(spec.total_intensity)
        out["%s_%s_residuals" % (name, tag)] = np.array(mix.residuals, dtype=float)
        # a second, random solution vector through the reference's own unpack path
        x = np.concatenate([rng.uniform(0.05, 1.0, mix.m), rng.uniform(0.2, 3.0, mix.n),
                            rng.uniform(0.0, 5.0, mix.o)])
        out["%s_%s_x" % (name, tag)] = x
        mix.calculated = False
        res = ref.mixture._get_residuals(x.copy(), mix)
        out["%s_%s_x_residuals" % (name, tag)] = np.array(res, dtype=float)
        for s, spec in enumerate(mix.specimens):
            out["%s_%s_x_total_s%d" % (name, tag, s)] = np.array(spec.total_intensity)
        # the optimiser (fresh flags, as get_optimized_residual is called per trial)
        if name != "c3" or tag == "truth":
            mix.calculated = mix.optimized = False
            r = ref.mixture.get_optimized_residual(mix)
            out["%s_%s_opt_residual" % (name, tag)] = np.array(r, dtype=float)
            out["%s_%s_opt_fractions" % (name, tag)] = np.array(mix.fractions, dtype=float)
            out["%s_%s_opt_scales" % (name, tag)] = np.array(mix.scales, dtype=float)
            out["%s_%s_opt_bgshifts" % (name, tag)] = np.array(mix.bgshifts, dtype=float)
            out["%s_%s_opt_residuals" % (name, tag)] = np.array(mix.residuals, dtype=float)
        print("  %s %s done (%.1f s)" % (name, tag, time.time() - t0), flush=True)


def extras(ref, out):
    """cases beyond the BASELINE configs: RawPatternPhase (phases.py:97-104)"""
    mix = syn.raw_case()
    ref.mixture.calculate_mixture(mix)
    spec = mix.specimens[0]
    out["raw_pi"] = np.array(spec.phase_intensities)
    out["raw_total"] = np.array(spec.total_intensity)
    out["raw_residuals"] = np.array(mix.residuals, dtype=float)
    theta = spec.range_theta
    stl = 2 * np.sin(theta) / spec.goniometer.wavelength
    out["raw_single"] = ref.phases.get_diffracted_intensity(theta, stl, spec.phases[0][1])
    out["raw_single_lpf"] = ref.phases.get_intensity(theta, stl, 2.3, 2.3, 0.0, spec.phases[0][2])
    # statistics.py on plain arrays (GUI pattern statistics): zeros in exp make Rpw / Rphase skip terms
    from pyxrd.calculations import statistics as rs
    rng = np.random.default_rng(31)
    e = rng.uniform(5.0, 500.0, 1500)
    e[[3, 77, 900]] = 0.0
    c = e * rng.uniform(0.9, 1.1, e.size) + rng.uniform(-2.0, 2.0, e.size)
    ph = rng.uniform(0.0, 80.0, e.size)
    out["stat_exp"], out["stat_calc"], out["stat_phase"] = e, c, ph
    with np.errstate(all="ignore"):
        out["stat_values"] = np.array([rs.Rp(e, c), rs.Rpw(e, c), rs.R_squared(e, c), rs.Rpe(e, c, 7), rs.Rphase(e, c, ph),
                                       rs.Rpder(e, c)], dtype=float)
    out["stat_derive"] = rs.derive(e)
    out["stat_smooth"] = rs.smooth_pattern(e)
    out["stat_rpw_all_zero"] = np.array(rs.Rpw(np.zeros(40), np.ones(40)), dtype=float)
    # two patterns per specimen (z_list of length 2, observed [X, 2], phases [Z][M])
    mix = syn.z2_case()
    ref.mixture.calculate_mixture(mix)
    for s, spec in enumerate(mix.specimens):
        out["z2_pi_s%d" % s] = np.array(spec.phase_intensities)
        out["z2_total_s%d" % s] = np.array(spec.total_intensity)
    out["z2_residuals"] = np.array(mix.residuals, dtype=float)
    mix = syn.z2_case()
    out["z2_opt_residual"] = np.array(ref.mixture.get_optimized_residual(mix), dtype=float)
    # settings.RESIDUAL_METHOD = "Rpder" (statistics.py:29-48) on the raw case (one specimen, no
    # exclusion) and on c4 (two specimens) with an exclusion window on the first specimen
    from pyxrd.data import settings
    settings.RESIDUAL_METHOD = "Rpder"
    try:
        mix = syn.raw_case()
        ref.mixture.calculate_mixture(mix)
        out["rpder_raw_residuals"] = np.array(mix.residuals, dtype=float)
        mix = syn.rpder_case()
        ref.mixture.calculate_mixture(mix)
        out["rpder_c4_residuals"] = np.array(mix.residuals, dtype=float)
        r = ref.mixture.get_optimized_residual(syn.rpder_case())
        out["rpder_c4_opt_residual"] = np.array(r, dtype=float)
    finally:
        settings.RESIDUAL_METHOD = "Rp"


MODEL_CASES = [(1, 1, 0), (1, 2, 1), (1, 3, 2), (1, 4, 3), (1, 5, 4), (1, 6, 5), (2, 2, 2), (3, 2, 4), (4, 2, 2),
               (5, 3, 6), (6, 4, 12), (7, 3, 6)]


def model_parameter_sets(npar, n=120, seed=99):
    """deterministic parameter vectors: special values, in-range, and out-of-range (clamped by the setters)"""
    rng = np.random.default_rng([seed, npar])
    out = []
    for t in range(n):
        if npar and t < 16:
            out.append(rng.choice([0.0, 0.5, 1.0, 2.0 / 3.0, 0.75, 0.25], npar))
        elif t % 5 == 0:
            out.append(rng.uniform(-0.2, 1.2, npar))
        else:
            out.append(rng.uniform(0.0, 1.0, npar))
    return np.array(out, dtype=float).reshape(n, npar)


def models(out):
    """probability models through the live reference MODEL layer:
    W, P, valid_probs (phases/models/phase.py:45) after a second update() (converged level matrices)"""
    import warnings
    import _live_models
    pm = _live_models.load().probabilities
    for model, G, npar in MODEL_CASES:
        params = model_parameter_sets(npar)
        r = G ** {1: 1, 2: 1, 3: 2, 4: 3, 5: 1, 6: 1, 7: 2}[model]
        Ws, Ps, valid = np.zeros((len(params), r, r)), np.zeros((len(params), r, r)), np.zeros(len(params), dtype=bool)
        for t, p in enumerate(params):
            if model == 1:
                m = getattr(pm, "R0G%dModel" % G)(**{"F%d" % (i + 1): p[i] for i in range(G - 1)})
            elif model == 2:
                m = pm.R1G2Model(W1=p[0], P11_or_P22=p[1])
            elif model == 3:
                m = pm.R2G2Model(W1=p[0], P112_or_P211=p[1], P21=p[2], P122_or_P221=p[3])
            elif model == 4:
                m = pm.R3G2Model(W1=p[0], P1111_or_P2112=p[1])
            elif model == 5:
                m = pm.R1G3Model(*p)
            elif model == 6:
                m = pm.R1G4Model(*p)
            else:
                m = pm.R2G3Model(*p)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                m.update()
            Ws[t], Ps[t] = m.get_distribution_matrix(), m.get_probability_matrix()
            valid[t] = bool(all(m.P_valid) and all(m.W_valid))
        key = "m%d_g%d" % (model, G)
        out[key + "_params"], out[key + "_W"], out[key + "_P"], out[key + "_valid"] = params, Ws, Ps, valid


def main():
    ref = _live_reference.load()
    names = sys.argv[1:] or ["kat", "extras", "models", "c1", "c2", "c4", "c5", "c3"]
    for name in names:
        out = {"versions": np.array("numpy %s scipy %s python %s reference 0.8.4"
                                    % (np.__version__, scipy.__version__, sys.version.split()[0]))}
        if name == "kat":
            kats(ref, out)
        elif name == "extras":
            extras(ref, out)
        elif name == "models":
            models(out)
        else:
            run_config(ref, name, out)
        path = os.path.join(HERE, "%s.npz" % name)
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path), "bytes", flush=True)


if __name__ == "__main__":
    main()
