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

import itertools
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import settings
from .packing import (ATOM_D_STRIDE, COMP_D_STRIDE, PHASE_D_STRIDE, PROB_STRIDE, BatchPack, StaticPack,
                      collect_atom_types, prob_model_of)
from .runtime import BINDING_DTYPE, Context, default_context, expand_population

# PYXRD_BIND_* of include/pyxrd_b200.h
BIND_PHASE_D, BIND_COMP_D, BIND_ATOM_D, BIND_MATRIX, BIND_PROB_PARAM, BIND_FRACTION, BIND_SCALE, BIND_BGSHIFT = range(8)
BIND_MODE_SET, BIND_MODE_MULTIPLY, BIND_MODE_ADD = range(3)      # include/pyxrd_b200.h PYXRD_BIND_MODE_*

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


_serials = itertools.count(1)


class PopulationTemplate(object):
    """Template candidate + resolved binding table."""

    def __init__(self, mixture, refinables: Sequence[Refinable], csds_auto_maximum: bool = True,
                 csds_factor: Optional[float] = None, static: Optional[StaticPack] = None):
        self.mixture = mixture
        self.refinables = list(refinables)
        self._serial = next(_serials)
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

    @classmethod
    def from_table(cls, mixture, bindings, bounds, names=None, csds_auto=None, csds_factor: Optional[float] = None,
                   static: Optional[StaticPack] = None):
        """a template whose binding table is given (``discover_template``, or a table stored with a project): ``bindings``
        structured array / rows ``(param, target, index, scale, offset[, mode])``; ``bounds`` f64[d, 2].  Rows apply in
        order; ``mode`` BIND_MODE_SET (default) ``slot = scale x + offset``, BIND_MODE_MULTIPLY ``slot *= ...``,
        BIND_MODE_ADD ``slot += ...``"""
        bounds = np.asarray(bounds, dtype=float).reshape(-1, 2)
        names = list(names) if names is not None else ["x%d" % i for i in range(len(bounds))]
        self = cls(mixture, [Refinable(n, [], lo, hi) for n, (lo, hi) in zip(names, bounds)], csds_auto_maximum=False,
                   csds_factor=csds_factor, static=static)
        table = bindings if isinstance(bindings, np.ndarray) else None
        if table is None or table.dtype != BINDING_DTYPE:
            if table is not None and table.dtype.names:     # a stored table with other field names: by position
                rows = [(r[0], r[1], r[2], r[4], r[5], r[3]) for r in table]
            else:
                rows = [tuple(r) for r in bindings]
            table = np.zeros(len(rows), dtype=BINDING_DTYPE)
            for i, row in enumerate(rows):
                param, target, index, scale, offset = row[:5]
                mode = int(row[5]) if len(row) > 5 else BIND_MODE_SET
                table[i] = (int(param), int(target), int(index), mode, float(scale), float(offset))
        self.bindings = np.ascontiguousarray(table)
        auto = None if csds_auto is None else np.ascontiguousarray(csds_auto, dtype=np.uint8)
        self.csds_auto = auto if auto is not None and auto.any() else None
        return self

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

    clip_to_bounds = False      # discovered templates: the reference's wrappers clamp a value to its range
                                # (refinables/wrapper.py:101) before the model layer sees it

    def _x(self, x):
        x = np.ascontiguousarray(np.atleast_2d(np.asarray(x, dtype=float)))
        if x.shape[1] != self.n_params:
            raise ValueError("x has %d columns, the template has %d refinables" % (x.shape[1], self.n_params))
        if self.clip_to_bounds:
            b = self.bounds()
            x = np.ascontiguousarray(np.clip(x, b[:, 0], b[:, 1]))
        return x

    def expand(self, x):
        """host only: arrays of the expanded K-candidate batch (dict), see runtime.expand_population"""
        return expand_population(self.batch, self._x(x), self.bindings, self.csds_auto, self.csds_factor)

    def _key(self):
        """identity of what ``pyxrd_set_population`` uploads besides x: a later call with the same key and shape of x only
        uploads the new solution vectors (``pyxrd_update_population``)"""
        return (self._serial, id(self.batch), id(self.bindings), id(self.csds_auto), self.csds_factor)

    def _load(self, ctx: Context, x):
        if ctx.set_static(self.static):
            pass                                  # fresh tables: nothing of an earlier population survives
        ctx.load_population(self.batch, x, self.bindings, self.csds_auto, self.csds_factor, key=self._key())

    def residuals(self, x, ctx: Optional[Context] = None, device: bool = False):
        """f64[K, 1+S]: ``calculate_mixture`` residuals (mixture.py:221-257) of the K solutions at the
        template's fractions / scales / bgshifts (or the bound ones).  ``device=True``: a zero-copy torch view of the
        device buffer instead of a host array (no synchronisation; for the NCCL gather of ``distributed``)"""
        ctx = ctx or default_context()
        with ctx.lock:
            ctx.set_residual_method(settings.get("RESIDUAL_METHOD"))
            self._load(ctx, self._x(x))
            ctx.compute_all(residuals_only=True)
            return ctx.residuals_tensor() if device else ctx.read_residuals()

    def optimized_residuals(self, x, ctx: Optional[Context] = None, outer: int = 0, newton: int = 0,
                            want_solutions: bool = False, device: bool = False):
        """f64[K]: ``get_optimized_residual`` (mixture.py:268-276) of the K solutions -- phase patterns,
        then the device mixture optimiser; with ``want_solutions`` also f64[K, M+2S].  ``device=True``: the zero-copy
        torch view f64[K, 1+S] of the optimiser's residual table on the device"""
        ctx = ctx or default_context()
        reason = self.optimiser_scope()
        if reason:
            raise ValueError("the device optimiser does not minimise what get_optimized_residual minimises for this "
                             "mixture (%s): evaluate its trials one by one through calculations.mixture" % reason)
        refine_bg = self.mixture_refines_background()
        with ctx.lock:
            self._load(ctx, self._x(x))
            ctx.compute_intensities()
            ctx.optimize(refine_bg=refine_bg, outer=outer, newton=newton)
            if device:
                return ctx.residuals_tensor(optimised=True)
            res = ctx.read_opt_residuals()
            sol = ctx.read_solutions() if want_solutions else None
        return (res[:, 0], sol) if want_solutions else res[:, 0]

    def phase_intensities(self, x, ctx: Optional[Context] = None):
        """list over candidates of list over specimens of f64[M, Z, X_s]"""
        ctx = ctx or default_context()
        with ctx.lock:
            self._load(ctx, self._x(x))
            ctx.compute_intensities()
            return ctx.split_per_specimen(ctx.read_intensities_flat(), (self.batch.M, self.static.Z))

    def mixture_refines_background(self) -> bool:
        mask = getattr(self.mixture, "bgshifts_mask", None)
        return bool(settings.get("BGSHIFT")) and mask is not None and bool(np.all(np.asarray(mask)))

    def optimiser_scope(self):
        """None when ``pyxrd_optimize`` minimises exactly what ``get_optimized_residual`` minimises for the template's
        mixture, else the reason as text.  The device optimiser varies ALL fractions and ALL scales (and all background
        shifts or none) and minimises Rp: the reference's default masks (mixture/models/mixture.py:64-73).  A mixture
        with a fixed fraction, ``auto_scales`` off, a partially fixed background, another ``RESIDUAL_METHOD``, more than
        11 phases / 8 specimens or a specimen without observations has to stay on the per-trial path
        (``batched._device_optimisable`` is the same test on a parsed mixture)."""
        mix = self.mixture
        if settings.get("RESIDUAL_METHOD") != "Rp":
            return "RESIDUAL_METHOD is %r, the device optimiser minimises Rp" % (settings.get("RESIDUAL_METHOD"),)
        M, S = self.batch.M, self.batch.S
        if M > 11 or S > 8:
            return "%d phases / %d specimens exceed the optimiser's 11 / 8" % (M, S)
        for name, want in (("fractions_mask", M), ("scales_mask", S)):
            mask = np.asarray(getattr(mix, name, np.ones(want))).ravel()
            if mask.size != want or not np.all(mask):
                return "%s fixes some entries" % name
        bg = np.asarray(getattr(mix, "bgshifts_mask", np.zeros(S))).ravel()
        if np.any(bg) and not np.all(bg):
            return "bgshifts_mask fixes some background shifts and frees others"
        for s, spec in enumerate(mix.specimens):
            if spec is None or np.size(spec.observed_intensity) == 0:
                return "specimen %d has no observations" % s
        return None

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


# ---------------------------------------------------------------------------------------------------------------------
# binding tables discovered from a running model layer
# ---------------------------------------------------------------------------------------------------------------------
# parameter order of the probability models = include/pyxrd_b200.h (the property names of probabilities/models/*.py)
MODEL_PARAMETERS = {
    "R0": ["F1", "F2", "F3", "F4", "F5"],
    "R1G2": ["W1", "P11_or_P22"],
    "R1G3": ["W1", "P11_or_P22", "G1", "G2", "G3", "G4"],
    "R1G4": ["W1", "P11_or_P22", "R1", "R2", "G1", "G2", "G11", "G12", "G21", "G22", "G31", "G32"],
    "R2G2": ["W1", "P112_or_P211", "P21", "P122_or_P221"],
    "R2G3": ["W1", "P111_or_P212", "G1", "G2", "G3", "G4"],
    "R3G2": ["W1", "P1111_or_P2112"],
}


def probability_model_of_object(model):
    """(``prob_model`` name, parameter names, current values) of a reference probability model OBJECT
    (probabilities/models: classes ``R<r>G<g>Model``), or None for anything else"""
    import re
    m = re.match(r"^R(\d)G(\d)Model$", type(model).__name__)
    if not m:
        return None
    R, G = int(m.group(1)), int(m.group(2))
    name = "R0" if R == 0 else "R%dG%d" % (R, G)
    if name not in MODEL_PARAMETERS:
        return None
    names = MODEL_PARAMETERS[name][:G - 1] if R == 0 else MODEL_PARAMETERS[name]
    return name, names, [float(getattr(model, n)) for n in names]


_SLOT_ARRAYS = (("phase_d", BIND_PHASE_D), ("comp_d", BIND_COMP_D), ("atom_d", BIND_ATOM_D), ("matrices", BIND_MATRIX),
                ("prob_params", BIND_PROB_PARAM), ("fractions", BIND_FRACTION), ("scales", BIND_SCALE),
                ("bgshifts", BIND_BGSHIFT))
_STRUCTURE_ARRAYS = ("phase_comp", "comp_i", "atom_type", "job_phase", "prob_model")


def discover_template(get_data_object, x0, bounds, names=None, phase_model=None, csds_factor: Optional[float] = None,
                      verify: int = 3, seed: int = 0):
    """Binding table of a refinement found by PROBING the model layer instead of knowing it.

    ``get_data_object(x) -> MixtureData`` is the reference's own route from a solution to the calculation input
    (``Refiner.get_data_object``, refinement/refiner.py:100-106: it writes x into the MVC model -- inheritance, unit-cell
    and atom-ratio relations included -- and rebuilds the DataObject graph).  The template is the packed graph at ``x0``;
    every refinable is moved twice, the packed arrays are compared, and every slot that moved must have moved affinely
    (``slot = scale * x_i + offset``); a slot that two or more refinables move is probed once more with both moved and
    bound as their product (cell volume a * b * d001) or their sum (component weight over refined atom contents); a
    slot that is quadratic in one refinable with real roots (cell lengths a and b both following an atom ratio) becomes
    two rows of that refinable, confirmed by a third probe;
    the table is then checked against ``verify`` random solutions.  Non-affine slots are
    accepted in two forms only: the CSDS maximum following its average (``int(factor * average)``), and probability
    matrices of phases for which ``phase_model(s, z, m) -> (name, parameter names, values) | None`` names the reference's
    probability model -- those phases become model-driven (W, P computed on the device) and the refinable is bound to the
    model parameter that moved.  Anything else raises ``NotImplementedError`` naming the refinable: the caller then keeps
    evaluating that refinement graph by graph (``batched.get_optimized_residuals``).  (2 d + 2 + verify) graph builds,
    one more per pair of refinables that share a slot."""
    x0 = np.array(x0, dtype=float).ravel()
    bounds = np.asarray(bounds, dtype=float).reshape(-1, 2)
    d = x0.size
    names = list(names) if names is not None else ["x%d" % i for i in range(d)]
    factor = float(settings.get("LOG_NORMAL_MAX_CSDS_FACTOR") if csds_factor is None else csds_factor)
    state = {}

    def annotated(x):
        """the model layer's graph for x, its model-driven phases carrying prob_model / prob_params of THAT solution"""
        data = get_data_object(np.array(x, dtype=float))
        if phase_model is not None:
            for s, spec in enumerate(data.specimens):
                if spec is None:
                    continue
                for z, row in enumerate(spec.phases):
                    for m, phase in enumerate(row):
                        if phase is None or getattr(phase, "type", "Phase") != "Phase":
                            continue
                        info = phase_model(s, z, m)
                        if info is not None:
                            phase.prob_model, phase.prob_params = info[0], list(info[2])
                            phase.prob_param_names = list(info[1])
        return data

    def pack(x):
        data = annotated(x)
        if "static" not in state:
            types = collect_atom_types(data.specimens)
            z = next((len(sp.z_list) for sp in data.specimens if sp is not None), 0)
            state["static"] = StaticPack(data.specimens, types, n_patterns=z)
        batch = BatchPack([data], state["static"], bgshift_enabled=bool(settings.get("BGSHIFT")))
        return data, batch

    def arrays(batch):
        out = {k: np.array(getattr(batch, k), dtype=float).ravel() if getattr(batch, k, None) is not None else np.zeros(0)
               for k, _ in _SLOT_ARRAYS}
        out["maximum"] = np.array(batch.phase_i[:, 3], dtype=np.int64)
        out["structure"] = [np.array(getattr(batch, k)).ravel() if getattr(batch, k, None) is not None else np.zeros(0)
                            for k in _STRUCTURE_ARRAYS] + [np.delete(np.array(batch.phase_i), 3, axis=1).ravel()]
        return out

    def same_structure(a, b):
        return all(u.shape == v.shape and np.array_equal(u, v) for u, v in zip(a["structure"], b["structure"]))

    # a model layer may carry values it only refreshes when something changes (the reference's unit-cell lengths follow
    # atom contents through change signals, unit_cell_prop.py:199-201, and are stale in a freshly loaded project): move
    # every refinable once before the template is taken, so that x0 is seen the way a refinement sees it
    def first_step(i):
        lo, hi = bounds[i]
        return 0.25 * (hi - x0[i]) if (hi - x0[i]) >= (x0[i] - lo) else -0.25 * (x0[i] - lo)

    get_data_object(x0 + np.array([first_step(i) for i in range(d)]))
    data0, batch0 = pack(x0)
    import copy
    template_data = copy.deepcopy(data0)                  # the model layer reuses its DataObjects between calls
    base = arrays(batch0)
    NM = base["matrices"].size
    # matrix-pool ranges of the phases: [offset, offset + 2 r^2) -> packed phase index
    pool_owner = np.full(NM, -1, dtype=np.int64)
    for p, row in enumerate(batch0.phase_i):
        if row[4] & 1 and not row[4] & 8:
            pool_owner[row[6]:row[6] + 2 * row[1] * row[1]] = p
    model_driven = batch0.prob_model if batch0.prob_model is not None else np.zeros(len(batch0.phase_i), dtype=np.int32)
    rows, bound_by = [], {}
    probes, joint_cache = {}, {}          # refinable -> (its first probe value, the arrays there); (j, i) -> arrays

    def joint_probe(j, i):
        if (j, i) not in joint_cache:
            x = x0.copy()
            x[j], x[i] = probes[j][0], probes[i][0]
            joint_cache[(j, i)] = arrays(pack(x)[1])
        return joint_cache[(j, i)]

    third_cache = {}

    def third_probe(i):
        if i not in third_cache:
            x = x0.copy()
            x[i] += 3.0 * first_step(i)
            third_cache[i] = (x[i], arrays(pack(x)[1]))
        return third_cache[i]

    auto = np.zeros(len(batch0.phase_i), dtype=np.uint8)
    for i in range(d):
        step = first_step(i)
        if step == 0.0:
            raise ValueError("refinable %r has an empty range" % (names[i],))
        xa, xb = x0.copy(), x0.copy()
        xa[i] += step
        xb[i] += 2.0 * step
        A, B = arrays(pack(xa)[1]), arrays(pack(xb)[1])
        probes[i] = (xa[i], A)
        if not (same_structure(base, A) and same_structure(base, B)):
            raise NotImplementedError("refinable %r changes the structure of the packed mixture" % (names[i],))
        da, db = xa[i] - x0[i], xb[i] - x0[i]
        for key, target in _SLOT_ARRAYS:
            v0, va, vb = base[key], A[key], B[key]
            for idx in np.nonzero((va != v0) | (vb != v0))[0]:
                if key == "matrices" and model_driven[pool_owner[idx]] != 0:
                    continue                              # W / P of a model-driven phase: computed on the device
                what = "a probability matrix of a phase whose model was not named" if key == "matrices" else key
                scale = (vb[idx] - v0[idx]) / db
                if abs((va[idx] - v0[idx]) - scale * da) <= 1e-9 * max(abs(v0[idx]), abs(vb[idx]), abs(scale * db)):
                    if abs(scale - 1.0) < 1e-12:
                        scale = 1.0
                    offset = v0[idx] - scale * x0[i]
                    if abs(offset) < 1e-12 * max(abs(v0[idx]), 1e-300) or (scale == 1.0 and abs(offset) < 1e-15):
                        offset = 0.0
                    factors = [(float(scale), float(offset))]
                else:
                    # not affine: a product of two affine functions of the refinable?  (cell lengths a and b both follow
                    # an Fe / Al ratio, unit_cell_prop.py:199-201: the volume is quadratic in it.)  v = v0 (1 + alpha t)
                    # (1 + beta t), t = x - x0, from the two probes; a third probe confirms.
                    if key == "matrices" or v0[idx] == 0.0:
                        raise NotImplementedError("refinable %r drives %s[%d] non-affinely" % (names[i], what, idx))
                    c2 = ((vb[idx] - v0[idx]) / db - (va[idx] - v0[idx]) / da) / (db - da)
                    c1 = (va[idx] - v0[idx]) / da - c2 * da
                    p_, q_ = c1 / v0[idx], c2 / v0[idx]
                    disc = p_ * p_ - 4.0 * q_
                    if disc < 0.0 and -disc < 1e-5 * p_ * p_:
                        disc = 0.0                            # a square (volume ~ b^2): the roots coincide up to rounding
                    third = third_probe(i)
                    dc = third[0] - x0[i]
                    vc = third[1][key][idx]
                    if disc < 0.0 or abs(v0[idx] + c1 * dc + c2 * dc * dc - vc) > 1e-9 * max(abs(v0[idx]), abs(vc)):
                        raise NotImplementedError("refinable %r drives %s[%d] non-affinely" % (names[i], what, idx))
                    alpha, beta = 0.5 * (p_ + np.sqrt(disc)), 0.5 * (p_ - np.sqrt(disc))
                    factors = [(float(v0[idx] * alpha), float(v0[idx] * (1.0 - alpha * x0[i]))),
                               (float(beta), float(1.0 - beta * x0[i]))]
                slot = (target, int(idx))
                if slot in bound_by:
                    # a slot two refinables drive: a product (cell volume a * b * d001) or a sum (component weight over
                    # two refined atom contents).  One joint probe tells which; the row multiplies / adds to what
                    # the earlier rows of the slot left.
                    j = bound_by[slot][0]
                    joint = joint_probe(j, i)[key][idx]
                    dj = probes[j][1][key][idx] - v0[idx]
                    di = va[idx] - v0[idx]
                    tol = 1e-9 * max(abs(v0[idx]), abs(joint), abs(di), abs(dj))
                    if key == "phase_d" and idx % PHASE_D_STRIDE == 1:
                        raise NotImplementedError("the CSDS average phase_d[%d] depends on the refinables %r and %r together"
                                                  % (idx, names[j], names[i]))
                    if len(factors) == 1 and abs(joint - (v0[idx] + di + dj)) <= tol:
                        rows.append((i, target, int(idx), factors[0][0], float(-factors[0][0] * x0[i]), BIND_MODE_ADD))
                    elif v0[idx] != 0.0 and abs(joint - (v0[idx] + di) * (v0[idx] + dj) / v0[idx]) <= tol:
                        for n, (f_scale, f_offset) in enumerate(factors):
                            if n == 0:                        # relative to the value the earlier rows leave at x0
                                f_scale, f_offset = f_scale / v0[idx], f_offset / v0[idx]
                                if abs(f_offset) < 1e-12:
                                    f_offset = 0.0
                            rows.append((i, target, int(idx), float(f_scale), float(f_offset), BIND_MODE_MULTIPLY))
                    else:
                        raise NotImplementedError("%s[%d] depends on the refinables %r and %r together, neither as a sum "
                                                  "nor as a product" % (key, idx, names[j], names[i]))
                    bound_by[slot].append(i)
                    continue
                bound_by[slot] = [i]
                for n, (f_scale, f_offset) in enumerate(factors):
                    rows.append((i, target, int(idx), f_scale, f_offset, BIND_MODE_SET if n == 0 else BIND_MODE_MULTIPLY))
        for p in np.nonzero((A["maximum"] != base["maximum"]) | (B["maximum"] != base["maximum"]))[0]:
            avg = [arr["phase_d"][p * PHASE_D_STRIDE + 1] for arr in (base, A, B)]
            if [int(factor * a) for a in avg] != [int(arr["maximum"][p]) for arr in (base, A, B)]:
                raise NotImplementedError("refinable %r moves the CSDS maximum of phase %d in an unknown way" % (names[i], p))
            auto[p] = 1
    tmpl = PopulationTemplate.from_table(template_data, rows, bounds, names=names, csds_auto=auto, csds_factor=factor)
    tmpl.get_data_object = annotated
    tmpl.clip_to_bounds = True
    # the table against the model layer at random solutions
    rng = np.random.default_rng(seed)
    for _ in range(int(verify)):
        x = bounds[:, 0] + rng.uniform(0.0, 1.0, d) * (bounds[:, 1] - bounds[:, 0])
        want = arrays(pack(x)[1])
        got = tmpl.expand(x[np.newaxis, :])
        for key, _ in _SLOT_ARRAYS:
            g = np.array(got[key], dtype=float).ravel() if got.get(key) is not None else np.zeros(0)
            w = want[key]
            if key == "matrices":
                keep = model_driven[pool_owner] == 0 if NM else np.zeros(0, dtype=bool)
                g, w = g[keep], w[keep]
            if g.shape != w.shape or not np.allclose(g, w, rtol=1e-10, atol=1e-300):
                bad = int(np.argmax(np.abs(g - w))) if g.shape == w.shape and g.size else -1
                raise NotImplementedError("binding table does not reproduce %s (entry %d) at a random solution: the "
                                          "refinables interact" % (key, bad))
        if not np.array_equal(got["phase_i"][:, 3], want["maximum"]):
            raise NotImplementedError("binding table does not reproduce the CSDS maxima at a random solution")
    return tmpl


def template_from_refiner(refiner, mixture_model, **kwargs):
    """``discover_template`` for a reference ``Refiner`` (refinement/refiner.py) and the ``Mixture`` MODEL it refines:
    solutions go through ``refiner.get_data_object``; the probability model of the phase in slot (s, m) is read from
    ``mixture_model.phase_matrix`` (mixture/models/mixture.py:87-89)."""
    def phase_model(s, z, m):
        phase = mixture_model.phase_matrix[s, ...].flatten()[m]
        prob = getattr(phase, "probabilities", None) if phase is not None else None
        return probability_model_of_object(prob) if prob is not None else None

    x0 = np.array([r.value for r in refiner.refinables], dtype=float)
    bounds = np.array(refiner.ranges, dtype=float)
    names = [r.text_title for r in refiner.refinables]
    return discover_template(refiner.get_data_object, x0, bounds, names=names, phase_model=phase_model, **kwargs)


def mixture_model_of(refiner):
    """the ``Mixture`` MODEL a reference ``Refiner`` works on: the refiner only holds ``data_callback = lambda:
    self.mixture.data_object`` (refinement/refinement.py:118-127), so the model is taken from that closure, or found
    through the project of the phases listed in ``refiner.metadata``"""
    callback = getattr(refiner, "data_callback", None)
    for cell in getattr(callback, "__closure__", None) or ():
        owner = cell.cell_contents
        mixture = getattr(owner, "mixture", None)
        if mixture is not None and hasattr(mixture, "phase_matrix"):
            return mixture
    phases = (getattr(refiner, "metadata", None) or {}).get("phases") or []
    for phase in phases:
        project = getattr(phase, "parent", None)
        for mixture in getattr(project, "mixtures", None) or []:
            if any(p is phase for p in mixture.phases):
                return mixture
    raise LookupError("cannot find the Mixture model of this refiner")
