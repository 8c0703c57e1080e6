"""Sweep of the weight-floor schedule (delta0, shrink, delta_min) of the device mixture optimiser on the c4 bench
population: time and excess over the best residual found by a long run (64 x 8)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyxrd_b200 import synthetic as syn
import pyxrd_b200.calculations as calc
from pyxrd_b200.batched import pack_population
from pyxrd_b200.runtime import default_context
K = 256
cfg = syn.CONFIGS["c4"]
truth, mixes = cfg.truth(), cfg.population(2024, K)
def calc_total(mix):
    calc.mixture.calculate_mixture(mix)
    return [np.array(s.total_intensity) for s in mix.specimens]
syn.attach_observed(mixes, truth, calc_total)
static, batch = pack_population(mixes)
ctx = default_context(0)
def run(outer, newton, **env):
    schedule = dict(DELTA0=1e-2, SHRINK=0.5, DELTA_MIN=1e-7)
    schedule.update(env)
    for k, v in schedule.items():
        ctx.set_option("optimizer_" + k.lower(), v)
    ctx.optimize(refine_bg=True, outer=outer, newton=newton); ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.optimize(refine_bg=True, outer=outer, newton=newton)
    ctx.synchronize()
    return np.array(ctx.read_opt_residuals()).reshape(K, -1)[:, 0].copy(), (time.perf_counter() - t0) / 3
with ctx.lock:
    ctx.set_static(static, force=True)
    ctx.set_batch(batch)
    ctx.compute_intensities()
    best, _ = run(96, 8)
    rows = []
    for outer in (16, 20, 24, 32):
        for newton in (1, 2):
            for shrink in (0.5, 0.4, 0.3, 0.2):
                for d0 in (1e-2, 3e-3):
                    r, dt = run(outer, newton, SHRINK=shrink, DELTA0=d0)
                    ex = r / np.minimum(best, r) - 1.0
                    rows.append((dt, outer, newton, shrink, d0, ex.max(), ex.mean(), int((ex > 1e-3).sum())))
    rows.sort()
    for dt, outer, newton, shrink, d0, mx, mean, n3 in rows:
        if mx < 2.5e-3:
            print("%.3f ms outer %2d newton %d shrink %.2f delta0 %.0e: max %.2e mean %.2e #>1e-3 %d" % (dt * 1e3, outer, newton, shrink, d0, mx, mean, n3))
