"""Host-side latency of one refinement generation (solution vectors on the host in, residual table on the host out),
piece by piece, on the bench's c2 template: where the end-to-end call spends the time the device does not.
    python tools/generation_latency.py [candidates] [config]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
from pyxrd_b200 import distributed, settings, synthetic as syn  # noqa: E402
import pyxrd_b200.calculations as calc  # noqa: E402
from pyxrd_b200.runtime import default_context  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
name = sys.argv[2] if len(sys.argv) > 2 else "c2"
ctx = default_context(0)
tmpl, refs = bench.make_template(name, calc)
x = syn.sample_solutions(refs, 2024, K)
stream = torch.cuda.Stream()
ctx.set_stream(stream.cuda_stream)
acc = {}


def tick(key, t0):
    t1 = time.perf_counter()
    acc[key] = acc.get(key, 0.0) + (t1 - t0)
    return t1


with torch.cuda.stream(stream), ctx.lock:
    for rep in range(220):
        if rep == 20:
            acc.clear()
        torch.cuda.synchronize()
        t = t_begin = time.perf_counter()
        ctx.set_residual_method(settings.get("RESIDUAL_METHOD")); t = tick("set_residual_method", t)
        xs = tmpl._x(x); t = tick("_x (validation)", t)
        ctx.set_static(tmpl.static); t = tick("set_static (identity check)", t)
        ctx.load_population(tmpl.batch, xs, tmpl.bindings, tmpl.csds_auto, tmpl.csds_factor, key=tmpl._key()); t = tick("load_population (x upload + prologue launches)", t)
        ctx.compute_all(residuals_only=True); t = tick("compute_all (launches)", t)
        view = ctx.residuals_tensor(); t = tick("residuals_tensor", t)
        rows = distributed.gather_rows(view, K, 1); t = tick("gather_rows (world 1)", t)
        out = distributed._to_host(rows); t = tick("_to_host (waits for the device)", t)
        acc["total"] = acc.get("total", 0.0) + (t - t_begin)
n = 200
for key, v in acc.items():
    print("%-50s %8.1f us" % (key, v / n * 1e6))
print("device step (events): see bench.py; residual table", out.shape)
