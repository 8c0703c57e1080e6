"""Sweep of the majoriser index from which the device mixture optimiser takes ONE Newton iteration per majoriser instead
of two ("optimizer_single_from"), on the c4 bench population: time and excess over a long run (64 x 8)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyxrd_b200 import synthetic as syn
import pyxrd_b200.calculations as calc
from pyxrd_b200.batched import pack_population
from pyxrd_b200.runtime import default_context
name = sys.argv[1] if len(sys.argv) > 1 else "c4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cfg = syn.CONFIGS[name]
truth, mixes = cfg.truth(), cfg.population(2024, K)
def calc_total(mix):
    calc.mixture.calculate_mixture(mix)
    return [np.array(s.total_intensity) for s in mix.specimens]
syn.attach_observed(mixes, truth, calc_total)
static, batch = pack_population(mixes)
ctx = default_context(0)
def run(outer, newton, single_from, tol=1e-5, coarse=(0, 1)):
    ctx.set_option("optimizer_single_from", single_from)
    ctx.set_option("optimizer_newton_tol", tol)
    ctx.set_option("optimizer_coarse_until", coarse[0])
    ctx.set_option("optimizer_coarse_stride", coarse[1])
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
    best, _ = run(96, 8, 0)
    for outer, coarse in [(32, (0, 1)), (32, (8, 2)), (32, (12, 2)), (32, (16, 2)), (32, (8, 4)), (32, (12, 4)), (32, (16, 4)), (32, (20, 4)),
                          (32, (12, 8)), (32, (16, 8)), (36, (16, 4))]:
        r, dt = run(outer, 2, 0, 1e-5, coarse)
        ex = r / np.minimum(best, r) - 1.0
        print("outer %2d coarse until %2d stride %d: %.3f ms  excess max %.2e p99 %.2e mean %.2e (#>1e-3: %d)" % (outer, coarse[0], coarse[1], dt * 1e3, ex.max(), np.quantile(ex, 0.99), ex.mean(), int((ex > 1e-3).sum())))
    ctx.set_option("optimizer_coarse_until", 0)
    ctx.set_option("optimizer_coarse_stride", 1)
    ctx.set_option("optimizer_single_from", 0)
    ctx.set_option("optimizer_newton_tol", 1e-5)
