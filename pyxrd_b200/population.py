"""Populations from a template + parameter binding table.

The reference turns every trial solution into a pattern by writing the d refinable values into
its MVC model (refinement/refiner.py:91-109, refinables/wrapper.py:92-110), rebuilding the whole
DataObject graph (mixture/models/mixture.py:53-85: 4.7-6.4 ms per trial) and handing it to
``get_optimized_residual``.  Here the graph of ONE candidate (the template) is packed once;
:class:`Refinable` rows say which attributes of which DataObjects a solution column drives
(``value = scale * x[param] + offset``), the table is resolved to packed slots once, and a whole
population ``x`` f64[K, d] is expanded natively (``pyxrd_expand_population``: memcpy + K * n multiply
-adds) and evaluated in one C-ABI call.  Everything numeric that depends on the bound values --
probability matrices W / P / valid_probs (``k_probabilities``), CSDS distribution, absolute scale,
z-stretch -- is computed on the device.

Attribute names are the reference's (data_objects.py:22-271).  Inheritance between model objects
(``based_on`` / ``linked_with``, generic/models/properties.py:46-57) is expressed by listing every
DataObject that receives the value as a target of the same Refinable.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import settings
from .packing import (ATOM_D_STRIDE, COMP_D_STRIDE, PHASE_D_STRIDE, PROB_STRIDE, BatchPack, StaticPack,
                      collect_atom_types, prob_model_of)
from .runtime import BINDING_DTYPE, Context, default_context, expand_population

# PYXRD_BIND_* of include/pyxrd_b200.h
BIND_PHASE_D, BIND_COMP_D, BIND_ATOM_D, BIND_MATRIX, BIND_PROB_PARAM, BIND_FRACTION, BIND_SCALE, BIND_BGSHIFT = range(8)

_PHASE_SLOTS = {"sigma_star": 0}
_CSDS_SLOTS = {"average": 1, "alpha_scale": 2, "alpha_offset": 3, "beta_scale": 4, "beta_offset": 5}
_COMP_SLOTS = {"d001": 0, "default_c": 1, "delta_c": 2, "lattice_d": 3, "volume": 4, "weight": 5}
_ATOM_SLOTS = {"pn": 0, "default_z": 1}


class Refinable(object):
    """One column of the solution vector and the DataObject attributes it drives.

    ``targets``: list of ``(object, attribute)`` or ``(object, attribute, scale, offset)`` with
      * PhaseData: ``"sigma_star"``; ``"prob:<n>"`` = n-th parameter of the phase's probability model
        (order of include/pyxrd_b200.h); ``"W:<i>:<j>"`` / ``"P:<i>:<j>"`` = one entry of a fixed matrix;
      * CSDSData: ``"average"`` (+ maximum = int(factor * average) when the template is built with
        ``csds_auto_maximum=True``, phases/models/CSDS.py:127,140), ``"alpha_scale"`` ...;
      * ComponentData: ``"d001" | "default_c" | "delta_c" | "lattice_d" | "volume" | "weight"``;
      * AtomData: ``"pn" | "default_z"``;
      * MixtureData: ``"fractions:<i>" | "scales:<s>" | "bgshifts:<s>"``.
    """

    def __init__(self, name: str, targets: Sequence[tuple], minimum: float = -np.inf, maximum: float = np.inf):
        self.name = name
        self.targets = [tuple(t) + (1.0, 0.0) if len(t) == 2 else tuple(t) for t in targets]
        self.minimum, self.maximum = float(minimum), float(maximum)


class PopulationTemplate(object):
    """Template candidate + resolved binding table."""

    def __init__(self, mixture, refinables: Sequence[Refinable], csds_auto_maximum: bool = True,
                 csds_factor: Optional[float] = None, static: Optional[StaticPack] = None):
        self.mixture = mixture
        self.refinables = list(refinables)
        if static is None:
            types = collect_atom_types(mixture.specimens)
            z = 0
            for spec in mixture.specimens:
                if spec is not None:
                    z = len(spec.z_list)
                    break
            static = StaticPack(mixture.specimens, types, n_patterns=z)
        self.static = static
        self.batch = BatchPack([mixture], static, bgshift_enabled=bool(settings.get("BGSHIFT")))
        self.csds_factor = float(settings.get("LOG_NORMAL_MAX_CSDS_FACTOR") if csds_factor is None else csds_factor)
        self._phases = {}           # id(phase) -> phase, of the template
        csds_phases = {}            # id(CSDS) -> [packed phase index]
        for spec in mixture.specimens:
            if spec is None:
                continue
            for row in spec.phases:
                for phase in row:
                    if phase is None or id(phase) in self._phases:
                        continue
                    self._phases[id(phase)] = phase
                    csds = getattr(phase, "CSDS", None)
                    if csds is not None and id(phase) in self.batch.phase_index:
                        csds_phases.setdefault(id(csds), []).append(self.batch.phase_index[id(phase)])
        rows: List[Tuple[int, int, int, float, float]] = []
        auto = np.zeros(self.batch.phase_i.shape[0], dtype=np.uint8)
        for col, ref in enumerate(self.refinables):
            for obj, attr, scale, offset in ref.targets:
                for target, index in self._resolve(obj, attr, csds_phases):
                    rows.append((col, target, index, float(scale), float(offset)))
                    if csds_auto_maximum and target == BIND_PHASE_D and index % PHASE_D_STRIDE == 1:
                        auto[index // PHASE_D_STRIDE] = 1
        self.bindings = np.zeros(len(rows), dtype=BINDING_DTYPE)
        for i, (col, target, index, scale, offset) in enumerate(rows):
            self.bindings[i] = (col, target, index, 0, scale, offset)
        self.csds_auto = auto if auto.any() else None

    # -- resolution of (object, attribute) to packed slots --------------------------------------
    def _resolve(self, obj, attr, csds_phases):
        b = self.batch
        if obj is self.mixture:
            kind, _, idx = attr.partition(":")
            target = {"fractions": BIND_FRACTION, "scales": BIND_SCALE, "bgshifts": BIND_BGSHIFT}.get(kind)
            if target is None:
                raise KeyError("mixture attribute %r cannot be bound" % (attr,))
            return [(target, int(idx))]
        if id(obj) in b.phase_index:
            pid = b.phase_index[id(obj)]
            if attr in _PHASE_SLOTS:
                return [(BIND_PHASE_D, pid * PHASE_D_STRIDE + _PHASE_SLOTS[attr])]
            kind, _, rest = attr.partition(":")
            if kind == "prob":
                model, params = prob_model_of(obj)
                if not model or not 0 <= int(rest) < len(params):
                    raise KeyError("phase has no probability parameter %r" % (attr,))
                return [(BIND_PROB_PARAM, pid * PROB_STRIDE + int(rest))]
            if kind in ("W", "P"):
                i, j = (int(v) for v in rest.split(":"))
                rank = int(b.phase_i[pid, 1])
                if prob_model_of(obj)[0]:
                    raise KeyError("W / P of a model-driven phase are computed on the device")
                return [(BIND_MATRIX, int(b.phase_i[pid, 6]) + (rank * rank if kind == "P" else 0) + i * rank + j)]
            raise KeyError("phase attribute %r cannot be bound" % (attr,))
        if id(obj) in csds_phases:
            if attr not in _CSDS_SLOTS:
                raise KeyError("CSDS attribute %r cannot be bound" % (attr,))
            return [(BIND_PHASE_D, pid * PHASE_D_STRIDE + _CSDS_SLOTS[attr]) for pid in csds_phases[id(obj)]]
        if id(obj) in b.comp_index:
            if attr not in _COMP_SLOTS:
                raise KeyError("component attribute %r cannot be bound" % (attr,))
            return [(BIND_COMP_D, b.comp_index[id(obj)] * COMP_D_STRIDE + _COMP_SLOTS[attr])]
        if id(obj) in b.atom_index:
            if attr not in _ATOM_SLOTS:
                raise KeyError("atom attribute %r cannot be bound" % (attr,))
            return [(BIND_ATOM_D, b.atom_index[id(obj)] * ATOM_D_STRIDE + _ATOM_SLOTS[attr])]
        raise KeyError("%r is not part of the template mixture" % (obj,))

    # -- population calls -----------------------------------------------------------------------
    @property
    def n_params(self) -> int:
        return len(self.refinables)

    def bounds(self):
        return np.array([[r.minimum, r.maximum] for r in self.refinables])

    def _x(self, x):
        x = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=float)))
        if x.shape[1] != self.n_params:
            raise ValueError("x has %d columns, the template has %d refinables" % (x.shape[1], self.n_params))
        return x

    def expand(self, x):
        """host only: arrays of the expanded K-candidate batch (dict), see runtime.expand_population"""
        return expand_population(self.batch, self._x(x), self.bindings, self.csds_auto, self.csds_factor)

    def residuals(self, x, ctx: Optional[Context] = None):
        """f64[K, 1+S]: ``calculate_mixture`` residuals (mixture.py:221-257) of the K solutions at the
        template's fractions / scales / bgshifts (or the bound ones)"""
        ctx = ctx or default_context()
        with ctx.lock:
            ctx.set_static(self.static)
            ctx.set_residual_method(settings.get("RESIDUAL_METHOD"))
            ctx._resident_epoch = None
            return ctx.evaluate_population(self.batch, self._x(x), self.bindings, self.csds_auto, self.csds_factor)

    def optimized_residuals(self, x, ctx: Optional[Context] = None, outer: int = 0, newton: int = 0,
                            want_solutions: bool = False):
        """f64[K]: ``get_optimized_residual`` (mixture.py:268-276) of the K solutions -- phase patterns,
        then the device mixture optimiser; with ``want_solutions`` also f64[K, M+2S]"""
        ctx = ctx or default_context()
        refine_bg = self.mixture_refines_background()
        with ctx.lock:
            ctx.set_static(self.static)
            ctx._resident_epoch = None
            res, sol = ctx.optimize_population(self.batch, self._x(x), self.bindings, self.csds_auto, self.csds_factor,
                                               refine_bg=refine_bg, outer=outer, newton=newton)
        return (res[:, 0], sol) if want_solutions else res[:, 0]

    def phase_intensities(self, x, ctx: Optional[Context] = None):
        """list over candidates of list over specimens of f64[M, Z, X_s]"""
        ctx = ctx or default_context()
        with ctx.lock:
            ctx.set_static(self.static)
            ctx._resident_epoch = None
            ctx.set_population(self.batch, self._x(x), self.bindings, self.csds_auto, self.csds_factor)
            ctx.compute_intensities()
            return ctx.split_per_specimen(ctx.read_intensities_flat(), (self.batch.M, self.static.Z))

    def mixture_refines_background(self) -> bool:
        mask = getattr(self.mixture, "bgshifts_mask", None)
        return bool(settings.get("BGSHIFT")) and mask is not None and bool(np.any(np.asarray(mask)))

    # -- writing one solution back into the DataObject graph --------------------------------------
    def apply(self, x, probabilities=None, ctx: Optional[Context] = None):
        """Set the template's DataObjects to solution ``x`` (what the reference's ``apply_solution``
        leaves, refinement/refiner.py:91-109).  W / P / valid_probs of model-driven phases come from
        ``probabilities(model, G, params) -> (W, P, valid)``, by default the device leaf
        ``pyxrd_leaf_probabilities``."""
        x = np.asarray(x, dtype=float).ravel()
        if probabilities is None:
            ctx = ctx or default_context()
            probabilities = ctx.leaf_probabilities
        touched = {}
        for col, ref in enumerate(self.refinables):
            for obj, attr, scale, offset in ref.targets:
                value = scale * x[col] + offset
                kind, _, rest = attr.partition(":")
                if kind in ("fractions", "scales", "bgshifts"):
                    arr = np.array(getattr(obj, kind), dtype=float)
                    arr[int(rest)] = value
                    setattr(obj, kind, arr)
                elif kind == "prob":
                    params = list(obj.prob_params)
                    params[int(rest)] = value
                    obj.prob_params = params
                    touched[id(obj)] = obj
                elif kind in ("W", "P"):
                    i, j = (int(v) for v in rest.split(":"))
                    mat = np.array(getattr(obj, kind), dtype=float)
                    mat[i, j] = value
                    setattr(obj, kind, mat)
                else:
                    setattr(obj, attr, value)
                    if attr == "average" and self.csds_auto is not None and hasattr(obj, "maximum"):
                        obj.maximum = int(self.csds_factor * value)
        for phase in touched.values():
            model, params = prob_model_of(phase)
            phase.W, phase.P, phase.valid_probs = probabilities(model, int(phase.G), params)
        self.mixture.parsed = self.mixture.calculated = self.mixture.optimized = False
        return self.mixture
