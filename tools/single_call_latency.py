"""Latency of the drop-in calls on ONE mixture (the way the reference's GUI / serial refinement calls them)."""
import sys, os, time, copy
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyxrd_b200 import synthetic as syn
import pyxrd_b200.calculations as calc
from pyxrd_b200.batched import get_optimized_residuals

for name in (sys.argv[1:] or ["c1", "c2", "c4", "c3", "c5"]):
    cfg = syn.CONFIGS[name]
    mix = cfg.candidate(2024, 0)
    rng = np.random.default_rng(1)
    for spec in mix.specimens:
        spec.observed_intensity = rng.uniform(20.0, 60.0, (spec.range_theta.size, 1))
    def fresh():
        m = copy.deepcopy(mix)
        m.parsed = m.calculated = m.optimized = False
        return m
    calc.mixture.calculate_mixture(fresh())
    calc.mixture.get_optimized_residual(fresh())          # first launch of an optimiser kernel loads its module
    reps = 20
    ms = [fresh() for _ in range(reps)]
    t0 = time.perf_counter()
    for m in ms:
        calc.mixture.calculate_mixture(m)
    t_calc = (time.perf_counter() - t0) / reps
    ms = [fresh() for _ in range(reps)]
    t0 = time.perf_counter()
    for m in ms:
        calc.mixture.get_optimized_residual(m)
    t_opt = (time.perf_counter() - t0) / reps
    from pyxrd_b200 import settings
    settings.MIXTURE_OPTIMIZER = "scipy"
    ms = [fresh() for _ in range(3)]
    t0 = time.perf_counter()
    for m in ms:
        calc.mixture.get_optimized_residual(m)
    t_scipy = (time.perf_counter() - t0) / 3
    settings.MIXTURE_OPTIMIZER = "device"
    ms = [fresh() for _ in range(reps)]
    t0 = time.perf_counter()
    for m in ms:
        get_optimized_residuals([m])
    t_dev = (time.perf_counter() - t0) / reps
    print("%s: calculate_mixture %.2f ms, get_optimized_residual %.2f ms (MIXTURE_OPTIMIZER = scipy: %.1f ms), "
          "get_optimized_residuals([m]) %.2f ms" % (name, t_calc * 1e3, t_opt * 1e3, t_scipy * 1e3, t_dev * 1e3))
