"""Population drivers over a :class:`~pyxrd_b200.population.PopulationTemplate`.

The reference's refinement methods call ``refiner.get_residual(solution)`` one solution at a time
(refinement/refiner.py:111-121), or submit a generation to the async server one task per individual
(refinement/async_evaluatable.py:38-50).  These drivers keep the methods' own sampling / stepping rules and hand
every set of solutions that can be evaluated together to ONE device call:

* :func:`lbfgsb` -- the L BFGS B run (refinement/methods/scipy_runs.py:21-52: ``fmin_l_bfgs_b(approx_grad=True,
  epsilon=1e-4, bounds=ranges)``) with the d + 1 probes of every forward-difference gradient in one batch;
* :func:`brute_force` -- the brute force run's grid (refinement/methods/custom_brute.py:36-56);
* :func:`evaluate_generation` -- what a DEAP generation needs (deap_cma.py:345-360, deap_swarm.py:183-203,
  deap_pso_cma.py:171-201): fitness of every individual of a population matrix.

The objective is ``get_optimized_residual`` semantics (mixture.py:268-276) by default, or the plain
``calculate_mixture`` residual with ``optimize=False``.
"""
from __future__ import annotations

from itertools import combinations, product
from typing import Callable, Optional

import numpy as np


def evaluate_generation(template, population, optimize: bool = True, ctx=None) -> np.ndarray:
    """f64[K]: fitness (average residual) of every row of ``population`` f64[K, d], one device batch"""
    x = np.ascontiguousarray(np.atleast_2d(np.asarray(population, dtype=float)))
    if optimize:
        return np.asarray(template.optimized_residuals(x, ctx=ctx))
    return np.asarray(template.residuals(x, ctx=ctx))[:, 0]


def forward_difference_probes(x, epsilon, bounds):
    """The d + 1 points scipy's ``approx_grad`` evaluates (2-point scheme with an absolute step, the step of a
    variable that would leave its upper bound is taken backwards: scipy.optimize._numdiff.approx_derivative with
    ``method='2-point', abs_step=epsilon, bounds=...``) and the signed steps."""
    x = np.asarray(x, dtype=float)
    lo, hi = np.asarray(bounds, dtype=float).T
    h = np.full(x.size, float(epsilon))
    backwards = (x + h > hi) & (x - h >= lo)
    h[backwards] *= -1.0
    probes = np.repeat(x[np.newaxis, :], x.size + 1, axis=0)
    probes[np.arange(1, x.size + 1), np.arange(x.size)] += h
    steps = probes[np.arange(1, x.size + 1), np.arange(x.size)] - x      # the step actually representable
    return probes, steps


def lbfgsb(template, x0, bounds=None, epsilon: float = 1e-4, maxfun: int = 15000, maxiter: int = 15000,
           optimize: bool = True, callback: Optional[Callable] = None, ctx=None):
    """``RefineLBFGSBRun.run`` with batched gradients: returns ``(solution, residual, info)`` like
    ``scipy.optimize.fmin_l_bfgs_b``.  ``info['batches']`` counts the device calls (one per function + gradient
    evaluation, d + 1 trials each)."""
    from scipy.optimize import fmin_l_bfgs_b
    bounds = template.bounds() if bounds is None else np.asarray(bounds, dtype=float)
    x0 = np.clip(np.asarray(x0, dtype=float), bounds[:, 0], bounds[:, 1])
    count = {"batches": 0, "trials": 0}

    def fun_and_grad(x):
        probes, steps = forward_difference_probes(x, epsilon, bounds)
        values = evaluate_generation(template, probes, optimize=optimize, ctx=ctx)
        count["batches"] += 1
        count["trials"] += len(probes)
        return float(values[0]), (values[1:] - values[0]) / steps

    kwargs = dict(bounds=[tuple(b) for b in bounds], maxfun=maxfun, maxiter=maxiter, callback=callback)
    solution, residual, info = fmin_l_bfgs_b(fun_and_grad, x0, **kwargs)
    info.update(count)
    return solution, residual, info


def brute_force_grid(bounds, num_samples: int = 11) -> np.ndarray:
    """the solutions the brute force run generates (custom_brute.py:36-56): a line for one parameter, otherwise a
    num_samples x num_samples grid for every pair of parameters with the others half-way their range"""
    b = np.asarray(bounds, dtype=float)
    mins, ranges = b[:, 0], b[:, 1] - b[:, 0]
    d = len(b)
    out = []
    if d == 1:
        for index in range(num_samples):
            out.append(mins + ranges * np.array([index / float(num_samples - 1)]))
    else:
        for par1, par2 in combinations(range(d), 2):
            indeces = np.ones(d) * 0.5
            for i, j in product(range(num_samples), repeat=2):
                indeces[par1] = i / float(num_samples - 1)
                indeces[par2] = j / float(num_samples - 1)
                out.append(mins + ranges * indeces)
    return np.array(out)


def brute_force(template, num_samples: int = 11, bounds=None, optimize: bool = True, chunk: int = 4096, ctx=None):
    """``RefineBruteForceRun.run``: evaluates the grid in batches of ``chunk`` solutions; returns
    ``(best solution, best residual, solutions, residuals)``"""
    bounds = template.bounds() if bounds is None else np.asarray(bounds, dtype=float)
    grid = brute_force_grid(bounds, num_samples)
    values = np.concatenate([evaluate_generation(template, grid[i:i + chunk], optimize=optimize, ctx=ctx)
                             for i in range(0, len(grid), chunk)])
    best = int(np.argmin(values))
    return grid[best], float(values[best]), grid, values
