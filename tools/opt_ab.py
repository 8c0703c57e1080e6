"""A/B of device mixture-optimiser variants on a bench population: time per launch (CUDA-synchronised wall clock over a
few launches) AND excess of the optimised residual over a long run (96 x 8).  Usage:
    python tools/opt_ab.py c4 256 '{"optimizer_lane_solver": 0}' '{"optimizer_lane_solver": 4}' ...
Every variant is a JSON dict of pyxrd_set_option keys applied on top of the defaults."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyxrd_b200 import synthetic as syn  # noqa: E402
import pyxrd_b200.calculations as calc  # noqa: E402
from pyxrd_b200.batched import pack_population  # noqa: E402
from pyxrd_b200.runtime import default_context  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 256
variants = [json.loads(a) for a in sys.argv[3:]] or [{}]
DEFAULTS = {"optimizer_lane_solver": 4, "optimizer_threads": 0, "optimizer_newton_tol": 1e-5, "optimizer_single_from": 0}
cfg = syn.CONFIGS[name]
truth, mixes = cfg.truth(), cfg.population(2024, K)


def calc_total(mix):
    calc.mixture.calculate_mixture(mix)
    return [np.array(s.total_intensity) for s in mix.specimens]


syn.attach_observed(mixes, truth, calc_total)
static, batch = pack_population(mixes)
ctx = default_context(0)


def run(outer, newton, options, reps=5):
    for key, value in {**DEFAULTS, **options}.items():
        try:
            ctx.set_option(key, value)
        except Exception:
            if key in options:
                raise
    ctx.optimize(refine_bg=True, outer=outer, newton=newton)
    ctx.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.optimize(refine_bg=True, outer=outer, newton=newton)
    ctx.synchronize()
    return np.array(ctx.read_opt_residuals()).reshape(K, -1)[:, 0].copy(), (time.perf_counter() - t0) / reps


with ctx.lock:
    ctx.set_static(static, force=True)
    ctx.set_batch(batch)
    ctx.compute_intensities()
    best, _ = run(96, 8, {"optimizer_lane_solver": 0, "optimizer_newton_tol": 1e-13}, reps=1)
    for options in variants:
        r, dt = run(32, 2, options)
        best = np.minimum(best, r)
    for options in variants:
        r, dt = run(32, 2, options)
        ex = r / best - 1.0
        print("%s K=%d %s: %.3f ms  excess max %.2e p99 %.2e mean %.2e (#>1e-3: %d) finite %s"
              % (name, K, json.dumps(options), dt * 1e3, ex.max(), np.quantile(ex, 0.99), ex.mean(), int((ex > 1e-3).sum()),
                 bool(np.isfinite(r).all())))
    run(32, 2, {})
