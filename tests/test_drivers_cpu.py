"""Population drivers (pyxrd_b200/drivers.py) on the CPU: the batched forward-difference L-BFGS-B takes exactly the
path scipy's own ``approx_grad`` takes (refinement/methods/scipy_runs.py:39-48), the brute-force grid is the
reference's (refinement/methods/custom_brute.py:36-56)."""
from itertools import combinations, product

import numpy as np
from scipy.optimize import fmin_l_bfgs_b

from pyxrd_b200 import drivers


class FakeTemplate(object):
    """stands in for a PopulationTemplate: a smooth objective with a minimum inside / on the bounds"""

    def __init__(self, bounds):
        self._bounds = np.asarray(bounds, dtype=float)
        self.batches = []

    def bounds(self):
        return self._bounds

    @staticmethod
    def f(x):
        x = np.asarray(x, dtype=float)
        return float(np.sum((x - np.array([0.3, 1.7, -0.4])) ** 2 * np.array([1.0, 3.0, 0.5])) + 0.1 * np.sin(3 * x[0]) + 2.0)

    def optimized_residuals(self, x, ctx=None):
        self.batches.append(len(x))
        return np.array([self.f(row) for row in x])

    def residuals(self, x, ctx=None):
        return np.stack([self.optimized_residuals(x)] * 2, axis=1)


def test_batched_lbfgsb_follows_scipy_approx_grad():
    bounds = [(0.0, 1.0), (0.0, 1.5), (-1.0, 1.0)]        # the second variable ends on its upper bound
    x0 = np.array([0.9, 0.2, 0.99995])                     # third start: forward step would leave the bound
    tmpl = FakeTemplate(bounds)
    sol, res, info = drivers.lbfgsb(tmpl, x0, epsilon=1e-4)
    want_sol, want_res, want_info = fmin_l_bfgs_b(FakeTemplate.f, x0, approx_grad=True, bounds=bounds, epsilon=1e-4)
    assert np.allclose(sol, want_sol, rtol=0, atol=1e-12) and abs(res - want_res) < 1e-12
    assert info["nit"] == want_info["nit"]
    assert set(tmpl.batches) == {4} and info["batches"] == len(tmpl.batches) and info["trials"] == 4 * info["batches"]
    assert info["batches"] * 4 == want_info["funcalls"]     # scipy counts every probe as a function call
    sol2, res2, _ = drivers.lbfgsb(tmpl, x0, epsilon=1e-4, optimize=False)
    assert np.allclose(sol2, want_sol, atol=1e-12)


def test_forward_difference_probes_respect_bounds():
    probes, steps = drivers.forward_difference_probes([0.5, 0.99999, 0.0], 1e-4, [(0, 1), (0, 1), (0, 1)])
    assert probes.shape == (4, 3) and np.array_equal(probes[0], [0.5, 0.99999, 0.0])
    assert steps[0] > 0 and steps[1] < 0 and steps[2] > 0
    assert np.all((probes >= 0.0) & (probes <= 1.0))
    for i in range(3):
        others = [j for j in range(3) if j != i]
        assert np.array_equal(probes[i + 1, others], probes[0, others])


def reference_grid(bounds, num_samples):
    """restatement of custom_brute.py:36-56 ``generate()``"""
    npbounds = np.array(bounds, dtype=float)
    npmins, npranges = npbounds[:, 0], npbounds[:, 1] - npbounds[:, 0]
    n = len(bounds)
    if n == 1:
        for index in range(num_samples):
            yield npmins + npranges * np.array([index / float(num_samples - 1)])
    else:
        for par1, par2 in combinations(list(range(n)), 2):
            indeces = np.ones(shape=(n,), dtype=float) * 0.5
            for par_indeces in product(list(range(num_samples)), repeat=2):
                indeces[par1] = par_indeces[0] / float(num_samples - 1)
                indeces[par2] = par_indeces[1] / float(num_samples - 1)
                yield npmins + npranges * indeces


def test_brute_force_grid_is_the_reference_grid():
    for bounds, n in (([(0.0, 2.0)], 5), ([(0, 1), (2, 3), (-1, 1)], 4), ([(0, 1), (5, 9)], 11)):
        assert np.array_equal(drivers.brute_force_grid(bounds, n), np.array(list(reference_grid(bounds, n))))
    tmpl = FakeTemplate([(0.0, 1.0), (0.0, 1.5), (-1.0, 1.0)])
    best, value, grid, values = drivers.brute_force(tmpl, num_samples=7, chunk=50)
    assert len(grid) == 3 * 49 and value == values.min() and np.array_equal(best, grid[np.argmin(values)])
    assert max(tmpl.batches) <= 50 and sum(tmpl.batches) == len(grid)
