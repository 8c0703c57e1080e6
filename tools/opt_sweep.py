"""Quality / time sweep of the device mixture optimiser (outer reweightings x Newton steps) on the c4 bench population."""
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
truth = cfg.truth()
mixes = cfg.population(2024, K)
def calc_total(mix):
    calc.mixture.calculate_mixture(mix)
    return [np.array(s.total_intensity) for s in mix.specimens]
syn.attach_observed(mixes, truth, calc_total)
static, batch = pack_population(mixes)
ctx = default_context(0)
res = {}
with ctx.lock:
    ctx.set_static(static, force=True)
    ctx.set_batch(batch)
    ctx.compute_intensities()
    for outer, newton in [(64, 8), (32, 4), (32, 2), (32, 1), (24, 2), (24, 1), (16, 4), (16, 2), (16, 1), (48, 1), (12, 2), (8, 2)]:
        ctx.optimize(refine_bg=True, outer=outer, newton=newton)
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.optimize(refine_bg=True, outer=outer, newton=newton)
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / 3
        res[(outer, newton)] = (np.array(ctx.read_opt_residuals()).reshape(K, -1)[:, 0].copy(), dt)
best = np.min(np.stack([v[0] for v in res.values()]), axis=0)
for k, (r, dt) in res.items():
    ex = r / best - 1.0
    print("outer %3d newton %d: %.3f ms  excess over best: max %.2e  p99 %.2e  mean %.2e  (#>1e-3: %d, #>5e-3: %d)" % (k[0], k[1], dt * 1e3, ex.max(), np.quantile(ex, 0.99), ex.mean(), int((ex > 1e-3).sum()), int((ex > 5e-3).sum())))
