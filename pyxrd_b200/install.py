"""Make a running PyXRD use the device path: rebind the hot-path functions of
``pyxrd.calculations.*`` -- and the names that PyXRD's model / refinement layers imported
from them -- to the ``pyxrd_b200.calculations`` implementations.

    import pyxrd_b200.install as b200
    b200.install()            # before or after PyXRD's modules were imported
    b200.uninstall()          # restore the reference functions

The reference has no plugin registry for this path: callers do ``from
pyxrd.calculations.mixture import get_optimized_residual`` etc., so both the defining module
and every importing module's global have to be rebound.  The import sites are the ones listed
in SURVEY.md §8(b).
"""
import importlib
import sys

# module of the reference -> {function name: our module}
_TARGETS = {
    "pyxrd.calculations.atoms": ("atoms", ["get_atomic_scattering_factor", "get_structure_factor"]),
    "pyxrd.calculations.components": ("components", ["get_factors"]),
    "pyxrd.calculations.CSDS": ("CSDS", ["calculate_distribution"]),
    "pyxrd.calculations.goniometer": ("goniometer", ["get_T", "get_lorentz_polarisation_factor",
                                                     "get_fixed_to_ads_correction_range"]),
    "pyxrd.calculations.phases": ("phases", ["get_diffracted_intensity", "get_intensity", "get_structure_factors",
                                             "get_Q_matrices"]),
    "pyxrd.calculations.math_tools": ("math_tools", ["mmult"]),
    "pyxrd.calculations.specimen": ("specimen", ["get_machine_correction_range", "calculate_phase_intensities",
                                                 "calculate_scaled_intensities"]),
    "pyxrd.calculations.statistics": ("statistics", ["R_squared", "Rp", "smooth_pattern", "derive", "Rpder", "Rpw", "Rpe",
                                                    "Rphase"]),
    "pyxrd.calculations.mixture": ("mixture", ["parse_mixture", "calculate_mixture", "optimize_mixture",
                                               "calculate_and_optimize_mixture", "get_residual",
                                               "get_optimized_residual", "_get_residuals"]),
}

# modules that did ``from pyxrd.calculations.X import name`` (reference path:line in SURVEY.md §8(b))
_IMPORT_SITES = [
    "pyxrd.mixture.models.optimizers",        # optimizers.py:10-15
    "pyxrd.refinement.refine_method",         # refine_method.py:14
    "pyxrd.phases.models.abstract_phase",     # abstract_phase.py:19
    "pyxrd.phases.models.component",          # component.py:22
    "pyxrd.atoms.models",                     # models.py:17
    "pyxrd.phases.models.CSDS",               # CSDS.py:10
    "pyxrd.goniometer.models",                # models.py:26-31
    "pyxrd.specimen.models.statistics",       # statistics.py:16 (Rpw, Rp, derive)
    "pyxrd.calculations.specimen",            # specimen.py:12-13 (imports from phases / goniometer)
    "pyxrd.calculations.phases",              # phases.py:15-18
    "pyxrd.calculations.components",          # components.py:13
    "pyxrd.calculations.mixture",             # mixture.py:18
]

_saved = []


def _generation_evaluator(original):
    """``RefineAsyncHelper.do_async_evaluation`` (refinement/refine_async_helper.py:16-23) for the default residual
    path: the generation's solutions are evaluated together from the refiner's binding table
    (``population.template_from_refiner``, discovered once per refiner) instead of one MVC round trip + one task each.
    Custom eval / data functions, refinements a table cannot express and failures of the discovery keep the reference's
    own loop."""
    import numpy as np

    def do_async_evaluation(self, iter_func, eval_func=None, data_func=None, result_func=None):
        refiner = getattr(self, "refiner", None)
        template = None
        if refiner is not None and eval_func is None and data_func is None:
            template = _refiner_template(refiner)
        if not template:
            return original(self, iter_func, eval_func, data_func, result_func)
        solutions = []
        for solution in iter_func():
            solutions.append(solution)
            if self._user_cancelled():
                break
        if solutions:
            values = template.optimized_residuals(np.array([np.asarray(s, dtype=float) for s in solutions]))
            report = result_func if result_func is not None else refiner.update
            for solution, value in zip(solutions, values):
                report(solution, np.array([value]))      # get_optimized_residual returns a 1-element array
        return solutions

    return do_async_evaluation


def _refiner_template(refiner):
    """the binding-table template of a reference Refiner (discovered once, cached on it); False if not expressible"""
    import logging
    template = getattr(refiner, "_b200_template", None)
    if template is None:
        try:
            from .population import mixture_model_of, template_from_refiner
            template = template_from_refiner(refiner, mixture_model_of(refiner))
            reason = template.optimiser_scope()
            if reason:          # fixed fractions, auto_scales off, Rpder ...: the device optimiser would minimise
                raise NotImplementedError(reason)     # another objective than get_optimized_residual
        except Exception as error:          # NotImplementedError: not expressible; anything else: stay safe
            logging.getLogger(__name__).info("pyxrd_b200: this refinement stays on the per-trial path (%s)", error)
            template = False
        refiner._b200_template = template
    return template


def _lbfgsb_run(original):
    """``RefineLBFGSBRun.run`` (refinement/methods/scipy_runs.py:35-52) with the d + 1 probes of every forward-difference
    gradient in one device batch (``drivers.lbfgsb``: the same steps scipy's ``approx_grad`` takes)."""
    def run(self, refiner, maxfun=15000, maxiter=15000, iprint=0, **kwargs):
        template = _refiner_template(refiner)
        if not template:
            return original(self, refiner, maxfun=maxfun, maxiter=maxiter, iprint=iprint, **kwargs)
        from .drivers import lbfgsb
        solution, residual, info = lbfgsb(template, refiner.history.initial_solution, bounds=refiner.ranges, epsilon=1e-4,
                                          maxfun=maxfun, maxiter=maxiter, callback=refiner.update)
        refiner.update(solution, residual=residual)
        return info

    return run


def install(import_reference=True, generations=True, calculations=True):
    """Rebind; returns the list of (module, name) pairs that were replaced.  ``generations``: also evaluate the
    generations of the population methods (CMA-ES, MPSO, PS-CMA-ES, brute force) from a binding table;
    ``calculations=False`` leaves ``pyxrd.calculations.*`` alone (only the generation evaluator is installed)."""
    import pyxrd_b200.calculations as ours
    if _saved:
        return [(m.__name__, n) for m, n, _ in _saved]
    if generations:
        try:
            helper = importlib.import_module("pyxrd.refinement.refine_async_helper") if import_reference \
                else sys.modules.get("pyxrd.refinement.refine_async_helper")
            cls = getattr(helper, "RefineAsyncHelper", None)
            if cls is not None:
                old = cls.__dict__["do_async_evaluation"]
                _saved.append((cls, "do_async_evaluation", old))
                cls.do_async_evaluation = _generation_evaluator(old)
        except Exception:
            pass
        try:
            runs = importlib.import_module("pyxrd.refinement.methods.scipy_runs") if import_reference \
                else sys.modules.get("pyxrd.refinement.methods.scipy_runs")
            cls = getattr(runs, "RefineLBFGSBRun", None)
            if cls is not None:
                old = cls.__dict__["run"]
                _saved.append((cls, "run", old))
                cls.run = _lbfgsb_run(old)
        except Exception:
            pass
    replaced = {}
    for ref_name, (our_name, names) in (_TARGETS.items() if calculations else ()):
        mod = sys.modules.get(ref_name)
        if mod is None and import_reference:
            try:
                mod = importlib.import_module(ref_name)
            except Exception:
                mod = None
        if mod is None:
            continue
        our_mod = getattr(ours, our_name)
        for name in names:
            if hasattr(mod, name) and hasattr(our_mod, name):
                old = getattr(mod, name)
                new = getattr(our_mod, name)
                _saved.append((mod, name, old))
                setattr(mod, name, new)
                replaced[id(old)] = new
    for site in _IMPORT_SITES:
        mod = sys.modules.get(site)
        if mod is None:
            continue
        for name, value in list(vars(mod).items()):
            new = replaced.get(id(value))
            if new is not None and value is not new:
                _saved.append((mod, name, value))
                setattr(mod, name, new)
    return [(m.__name__, n) for m, n, _ in _saved]


def uninstall():
    while _saved:
        mod, name, old = _saved.pop()
        setattr(mod, name, old)
