"""Trial evaluations per second through the template path (solution vectors x[K, d] on the host -> optimised residuals on
the host, pyxrd_optimize_population) for a sweep of population sizes: BASELINE configs[4] asks for 1k-64k trials over the
4-phase R3 mixture.  usage: python tools/population_sweep.py [c5] [1024 4096 16384 65536]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyxrd_b200 import synthetic as syn
from pyxrd_b200.population import PopulationTemplate
from pyxrd_b200.runtime import default_context

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
sizes = [int(v) for v in sys.argv[2:]] or [1024, 4096, 16384, 65536]
mix, refs = syn.template(name)
rng = np.random.default_rng(3)
for spec in mix.specimens:
    tt = np.degrees(2.0 * spec.range_theta)
    spec.observed_intensity = (200.0 + 800.0 * np.exp(-((tt - 8.8) / 0.6) ** 2) + rng.uniform(0.0, 20.0, tt.size)).reshape(-1, 1)
tmpl = PopulationTemplate(mix, refs)
ctx = default_context(0)
for K in sizes:
    x = syn.sample_solutions(refs, 50, K)
    tmpl.optimized_residuals(x, ctx=ctx)              # allocations, module loading
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        r = tmpl.optimized_residuals(x, ctx=ctx)
    dt = (time.perf_counter() - t0) / reps
    st = ctx.stage_times()
    print("%s K=%6d: %8.2f ms per generation, %9.0f trials/s end to end (prologue %.2f ms, phase %.2f ms, wavelength %.2f ms, optimiser %.2f ms), "
          "residuals finite: %s, best %.4f" % (name, K, dt * 1e3, K / dt, st["prologue"], st["phase"], st["wavelength"], st["residual"],
                                              bool(np.isfinite(r).all()), float(np.min(r))))
