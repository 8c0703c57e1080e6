"""one optimiser call on a bench population (argv: candidates, configuration name, JSON dict of pyxrd_set_option keys; default 256 c4) (profiling aid: build with -DPX_OPT_PROFILE for cycle counters)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyxrd_b200 import synthetic as syn
import pyxrd_b200.calculations as calc
from pyxrd_b200.batched import pack_population
from pyxrd_b200.runtime import default_context
K = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = syn.CONFIGS[sys.argv[2] if len(sys.argv) > 2 else "c4"]
truth, mixes = cfg.truth(), cfg.population(2024, K)
def calc_total(mix):
    calc.mixture.calculate_mixture(mix)
    return [np.array(s.total_intensity) for s in mix.specimens]
syn.attach_observed(mixes, truth, calc_total)
static, batch = pack_population(mixes)
ctx = default_context(0)
if len(sys.argv) > 3:
    import json
    for key, value in json.loads(sys.argv[3]).items():
        ctx.set_option(key, value)
with ctx.lock:
    ctx.set_static(static, force=True)
    ctx.set_batch(batch)
    ctx.compute_intensities()
    for _ in range(2):
        ctx.optimize(refine_bg=True)
        ctx.synchronize()
    print("optimiser ms", ctx.stage_times()["residual"])
