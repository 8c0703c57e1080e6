"""Flatten DataObject graphs into the plain arrays of ``include/pyxrd_b200.h``.

This is the host half of the drop-in boundary: the reference
hands ``pyxrd.calculations`` nested Python objects (data_objects.py:22-271); the
C-ABI wants structure-of-arrays buffers.  Only *layout* happens here -- every
per-candidate arithmetic step of the reference runs on the device.

The one numeric thing prepared on the host is the abscissa part of the
wavelength-distribution interpolation (specimen.py:64-69): indices and lerp
weights of ``interp1d(arcsin(stl*lambda_j/2), ., fill_value=0)(theta)``.  They are
candidate-invariant lookup tables; they are built with numpy because whether the
first / last point falls inside the interpolation range is decided by a 1-ulp
comparison of numpy's ``arcsin`` against ``theta``, so only
numpy's own ``arcsin`` reproduces the reference there.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

GONIO_STRIDE = 12
ATYPE_STRIDE = 12
PHASE_I_STRIDE = 8
PHASE_D_STRIDE = 8
COMP_D_STRIDE = 8
COMP_I_STRIDE = 4
ATOM_D_STRIDE = 2

FLAG_VALID_PROBS = 1
FLAG_APPLY_LPF = 2
FLAG_APPLY_CORRECTION = 4
FLAG_RAW_PATTERN = 8

PROB_STRIDE = 12
# PYXRD_PROB_* of include/pyxrd_b200.h: probability models the device evaluates itself
PROB_MODELS = {"fixed": 0, "R0": 1, "R1G2": 2, "R2G2": 3, "R3G2": 4, "R1G3": 5, "R1G4": 6, "R2G3": 7}
_PROB_G = {2: 2, 3: 2, 4: 2, 5: 3, 6: 4, 7: 3}
_PROB_RANK = {1: lambda G: G, 2: lambda G: 2, 3: lambda G: 4, 4: lambda G: 8, 5: lambda G: 3, 6: lambda G: 4, 7: lambda G: 9}
_PROB_NPAR = {1: lambda G: G - 1, 2: lambda G: 2, 3: lambda G: 4, 4: lambda G: 2, 5: lambda G: 6, 6: lambda G: 12, 7: lambda G: 6}

_DIV_MODES = {"FIXED": 0.0, "AUTOMATIC": 1.0}


def prob_model_of(phase):
    """(model code, parameter list) of a phase whose W / P the device derives from its probability
    parameters: ad-hoc attributes ``prob_model`` ("R0" | "R1G2" | "R1G3" | "R1G4" | "R2G2" | "R2G3" | "R3G2" or the code) and
    ``prob_params`` on the PhaseData (a DataObject is an attribute bag, data_objects.py:29-31)."""
    model = getattr(phase, "prob_model", None)
    if model is None:
        return 0, ()
    code = PROB_MODELS[model] if isinstance(model, str) else int(model)
    if code == 0:
        return 0, ()
    G = int(phase.G)
    params = [float(v) for v in getattr(phase, "prob_params", ())]
    if code not in _PROB_RANK or (code != 1 and G != _PROB_G[code]) or (code == 1 and not 1 <= G <= 6):
        raise ValueError("probability model %r is not available for G = %d" % (model, G))
    if len(params) != _PROB_NPAR[code](G):
        raise ValueError("probability model %r with G = %d takes %d parameters, got %d"
                         % (model, G, _PROB_NPAR[code](G), len(params)))
    return code, params


def _atype_key(at, memo: Optional[dict] = None) -> bytes:
    """value key of an atom type; ``memo`` (id -> key, valid for the duration of ONE packing call: the objects are
    alive and unchanged meanwhile) saves rebuilding it for every atom that shares the type object"""
    if memo is not None:
        key = memo.get(id(at))
        if key is None:
            key = memo[id(at)] = _atype_key(at)
        return key
    row = np.empty(ATYPE_STRIDE)
    row[0:5] = np.asarray(at.par_a, dtype=float)
    row[5:10] = np.asarray(at.par_b, dtype=float)
    row[10] = float(at.par_c)
    row[11] = float(at.debye)
    return row.tobytes()


def iter_phases(specimen):
    for row in specimen.phases:
        for phase in row:
            if phase is not None:
                yield phase


def collect_atom_types(mixtures_or_specimens) -> Dict[bytes, int]:
    """value-keyed table of every atom type reachable from the given specimens"""
    table: Dict[bytes, int] = {}
    memo: dict = {}
    for spec in mixtures_or_specimens:
        if spec is None:
            continue
        for phase in iter_phases(spec):
            if getattr(phase, "type", "Phase") != "Phase":
                continue
            if not phase.valid_probs and getattr(phase, "prob_model", None) is None:
                continue
            for comp in phase.components:
                for atom in list(comp.layer_atoms) + list(comp.interlayer_atoms):
                    if atom is not None and atom.atom_type is not None:
                        table.setdefault(_atype_key(atom.atom_type, memo), len(table))
    return table


_WL_CACHE: Dict[tuple, tuple] = {}      # the serial drop-in calls rebuild the same tables for every trial


def wavelength_tables(theta: np.ndarray, wavelength: float, distribution):
    """``_wavelength_tables`` behind a small value-keyed cache (abscissa bytes, wavelength, list): read-only results"""
    key = (np.ascontiguousarray(theta, dtype=float).tobytes(), float(wavelength),
           tuple((float(w), float(f)) for w, f in distribution))
    hit = _WL_CACHE.get(key)
    if hit is None:
        if len(_WL_CACHE) >= 32:
            _WL_CACHE.clear()
        hit = _WL_CACHE[key] = _wavelength_tables(theta, wavelength, distribution)
        for arr in hit:
            arr.setflags(write=False)
    return hit


def _wavelength_tables(theta: np.ndarray, wavelength: float, distribution):
    """index / weights of the shifted-abscissa linear interpolation, per list entry.

    scipy 1.18.1 ``interp1d._call_linear`` + ``_check_bounds``: ``hi =
    searchsorted(x, x_new).clip(1, n-1)``, ``y = w_hi*y[hi] + w_lo*y[hi-1]`` with
    ``w_hi = (x_new-x_lo)/(x_hi-x_lo)``, ``w_lo = (x_hi-x_new)/(x_hi-x_lo)``; 0 outside.
    """
    theta = np.asarray(theta, dtype=float)
    X = theta.size
    n = len(distribution)
    idx = np.full((n, X), -1, dtype=np.int32)
    w_hi = np.zeros((n, X))
    w_lo = np.zeros((n, X))
    frac = np.zeros(n)
    if X == 0:
        return frac, idx, w_hi, w_lo
    stl = 2 * np.sin(theta) / wavelength
    for j, (wl, f) in enumerate(distribution):
        frac[j] = f
        shifted = np.arcsin(stl * wl * 0.5)
        if X < 2 or not np.all(np.isfinite(shifted)) or np.any(np.diff(shifted) <= 0):
            raise ValueError("wavelength %r gives a non-monotonic / invalid shifted 2-theta axis" % (wl,))
        hi = np.searchsorted(shifted, theta).clip(1, X - 1).astype(np.int64)
        x_lo, x_hi = shifted[hi - 1], shifted[hi]
        w_hi[j] = (theta - x_lo) / (x_hi - x_lo)
        w_lo[j] = (x_hi - theta) / (x_hi - x_lo)
        outside = (theta < shifted[0]) | (theta > shifted[-1])
        idx[j] = np.where(outside, -1, hi).astype(np.int32)
    return frac, idx, w_hi, w_lo


def specimen_signature(spec):
    """the scalar part of what a specimen contributes to the candidate-invariant tables, as a cheap comparable tuple
    (goniometer row, wavelength list, sizes); the arrays themselves -- abscissa, observed patterns, selected range --
    are compared value by value in ``StaticPack.same_specimen``: the candidates of one batch must agree on all of it"""
    if spec is None:
        return None
    g = spec.goniometer
    theta = spec.range_theta
    n = int(np.size(theta))
    ends = (float(theta[0]), float(theta[-1])) if n else ()
    return (n, ends, g.wavelength, g.soller1, g.soller2, g.mcr_2theta, g.divergence_mode, g.divergence,
            bool(g.has_absorption_correction), g.absorption, g.sample_surf_density, g.sample_length, g.radius,
            tuple((float(w), float(f)) for w, f in (g.wavelength_distribution or [])), int(np.size(spec.observed_intensity)))


class StaticPack(object):
    """Candidate-invariant tables of a set of specimens (pyxrd_static_desc)."""

    def __init__(self, specimens: Sequence, atom_types: Dict[bytes, int], n_patterns: Optional[int] = None,
                 stl_override: Optional[np.ndarray] = None):
        S = len(specimens)
        self.stl_override = None if stl_override is None else np.ascontiguousarray(stl_override, dtype=float).ravel()
        self.S = S
        self.atom_types = dict(atom_types)
        self.T = len(self.atom_types)
        if n_patterns is None:
            n_patterns = 0
            for spec in specimens:
                if spec is not None:
                    n_patterns = len(spec.z_list)
                    break
        self.Z = int(n_patterns)
        self.specimen_signatures = [specimen_signature(spec) for spec in specimens]
        self._sources = [None] * S         # per specimen: ids of the arrays the tables were built from (identity shortcut)
        self.npts = np.zeros(S, dtype=np.int32)
        self.gonio = np.zeros((S, GONIO_STRIDE))
        self.nwl = np.zeros(S, dtype=np.int32)
        thetas, fracs, idxs, whis, wlos, obss, sels = [], [], [], [], [], [], []
        for s, spec in enumerate(specimens):
            if spec is None:            # None specimen -> residual 0.0 (mixture.py:241-252)
                continue
            theta = np.ascontiguousarray(spec.range_theta, dtype=float)
            X = theta.size
            self.npts[s] = X
            g = spec.goniometer
            mode = _DIV_MODES.get(g.divergence_mode, 2.0)
            self.gonio[s, :11] = [g.wavelength, g.soller1, g.soller2, g.mcr_2theta, mode,
                                  g.divergence if g.divergence is not None else 0.0,
                                  float(g.has_absorption_correction or 0.0), g.absorption,
                                  g.sample_surf_density, g.sample_length if g.sample_length is not None else 0.0,
                                  g.radius if g.radius is not None else 1.0]
            # a specimen without observed data reports residual 0.0 (mixture.py:247-251), whatever its patterns are
            self.gonio[s, 11] = 1.0 if np.size(spec.observed_intensity) > 0 else 0.0
            dist = list(g.wavelength_distribution or [])
            self.nwl[s] = len(dist)
            frac, idx, w_hi, w_lo = wavelength_tables(theta, g.wavelength, dist)
            thetas.append(theta)
            fracs.append(frac)
            idxs.append(idx.ravel())
            whis.append(w_hi.ravel())
            wlos.append(w_lo.ravel())
            obs = np.asarray(spec.observed_intensity, dtype=float)
            if obs.size == 0:
                obs_t = np.zeros((self.Z, X))
                self.no_observations = getattr(self, "no_observations", set()) | {s}
            else:
                obs_t = np.ascontiguousarray(obs.reshape(X, -1).T)
                if obs_t.shape[0] != self.Z:
                    raise ValueError("specimen %d: observed_intensity has %d patterns, z_list has %d"
                                     % (s, obs_t.shape[0], self.Z))
            obss.append(obs_t.ravel())
            sel = spec.selected_range
            sel = np.ones(X, dtype=bool) if sel is None else np.asarray(sel)
            if sel.dtype != np.bool_:     # an index array selects the same points as a mask
                mask = np.zeros(X, dtype=bool)
                mask[sel] = True
                sel = mask
            sels.append(sel.astype(np.uint8))
            self._sources[s] = (id(spec.range_theta), id(spec.observed_intensity), id(spec.selected_range),
                                spec.range_theta, spec.observed_intensity, spec.selected_range)   # keep them alive

        def cat(parts, dtype):
            return np.ascontiguousarray(np.concatenate(parts) if parts else np.zeros(0), dtype=dtype)

        self.theta = cat(thetas, float)
        self.wl_fraction = cat(fracs, float)
        self.wl_index = cat(idxs, np.int32)
        self.wl_w_hi = cat(whis, float)
        self.wl_w_lo = cat(wlos, float)
        self.observed = cat(obss, float)
        self.selected = cat(sels, np.uint8)
        self.offsets = np.concatenate([[0], np.cumsum(self.npts, dtype=np.int64)])
        self.sumX = int(self.offsets[-1])
        self.atype_table = np.zeros((max(self.T, 1), ATYPE_STRIDE))
        for key, i in self.atom_types.items():
            self.atype_table[i] = np.frombuffer(key, dtype=float)
        if not hasattr(self, "no_observations"):
            self.no_observations = set()

    def signature(self):
        """identity of EVERYTHING pyxrd_set_static uploads, to decide whether a resident copy can be reused (the
        gonio row only holds the primary wavelength: the interpolation tables of the wavelength list are part of it).
        Computed once: a pack is not modified after construction."""
        sig = getattr(self, "_signature", None)
        if sig is None:
            sig = self._signature = self._compute_signature()
        return sig

    def _compute_signature(self):
        return (self.S, self.Z, self.T, self.npts.tobytes(), self.gonio.tobytes(), self.theta.tobytes(),
                self.nwl.tobytes(), self.wl_fraction.tobytes(), self.wl_index.tobytes(), self.wl_w_hi.tobytes(),
                self.wl_w_lo.tobytes(), self.observed.tobytes(), self.selected.tobytes(),
                self.atype_table.tobytes(), None if self.stl_override is None else self.stl_override.tobytes())

    def same_specimen(self, s: int, spec) -> bool:
        """does specimen ``spec`` of a candidate describe the same candidate-invariant tables as specimen ``s`` of this
        pack: scalars by ``specimen_signature``, abscissa / observed patterns / selected range value by value (arrays
        that ARE the ones the pack was built from are not compared again)"""
        if specimen_signature(spec) != self.specimen_signatures[s]:
            return False
        if spec is None:
            return True
        src = self._sources[s]
        lo, hi = int(self.offsets[s]), int(self.offsets[s + 1])
        X = hi - lo
        if src is None or id(spec.range_theta) != src[0]:
            if not np.array_equal(np.asarray(spec.range_theta, dtype=float).ravel(), self.theta[lo:hi]):
                return False
        if src is None or id(spec.observed_intensity) != src[1]:
            obs = np.asarray(spec.observed_intensity, dtype=float)
            if obs.size and not np.array_equal(obs.reshape(X, -1).T.ravel(), self.observed[self.Z * lo:self.Z * hi]):
                return False
        if src is None or id(spec.selected_range) != src[2]:
            sel = spec.selected_range
            mask = np.ones(X, dtype=bool) if sel is None else np.asarray(sel)
            if mask.dtype != np.bool_:
                full = np.zeros(X, dtype=bool)
                full[mask] = True
                mask = full
            if not np.array_equal(mask.astype(np.uint8), self.selected[lo:hi]):
                return False
        return True


class BatchPack(object):
    """Per-candidate parameter blocks of K mixtures (pyxrd_batch_desc)."""

    def __init__(self, mixtures: Sequence, static: StaticPack, bgshift_enabled: bool = True):
        K = len(mixtures)
        S, Z = static.S, static.Z
        M = None
        phase_i: List[List[int]] = []
        phase_d: List[List[float]] = []
        phase_comp: List[int] = []
        matrices: List[np.ndarray] = []
        mat_off = 0
        comp_d: List[List[float]] = []
        comp_i: List[List[int]] = []
        atom_d: List[float] = []
        atom_t: List[int] = []
        jobs: List[int] = []
        self.components_of_phase: List[List[object]] = []
        self.prob_model_list: List[int] = []
        self.prob_param_list: List[List[float]] = []
        # id(object) -> packed index, of the FIRST candidate (what a population template binds to)
        self.phase_index: Dict[int, int] = {}
        self.comp_index: Dict[int, int] = {}
        self.atom_index: Dict[int, int] = {}
        self._atype_memo: dict = {}
        fractions = []
        scales = np.zeros((K, S))
        bgshifts = np.zeros((K, S))
        for k, mix in enumerate(mixtures):
            specimens = mix.specimens
            self._atom_index_open = k == 0
            if len(specimens) != S:
                raise ValueError("candidate %d has %d specimens, the static tables have %d" % (k, len(specimens), S))
            fr = np.asarray(mix.fractions, dtype=float).ravel()
            if M is None:
                M = fr.size
            if fr.size != M:
                raise ValueError("candidate %d has %d phases, expected %d" % (k, fr.size, M))
            fractions.append(fr)
            scales[k] = np.asarray(mix.scales, dtype=float).ravel()[:S]
            if bgshift_enabled:
                bgshifts[k] = np.asarray(mix.bgshifts, dtype=float).ravel()[:S]
            phase_ids: Dict[int, int] = {}
            comp_ids: Dict[int, int] = {}
            for s, spec in enumerate(specimens):
                if not static.same_specimen(s, spec):
                    raise ValueError("candidate %d, specimen %d: abscissa / goniometer / observations differ from the "
                                     "resident static tables (specimens are shared by the candidates of a batch)" % (k, s))
                for z in range(Z):
                    row = spec.phases[z] if spec is not None else [None] * M
                    if len(row) != M:
                        raise ValueError("specimen %d pattern %d lists %d phases, expected %d" % (s, z, len(row), M))
                    for phase in row:
                        if phase is None:
                            jobs.append(-1)
                            continue
                        pid = phase_ids.get(id(phase))
                        if pid is None:
                            pid = len(phase_i)
                            phase_ids[id(phase)] = pid
                            mat_off = self._pack_phase(phase, static, phase_i, phase_d, phase_comp, matrices, mat_off,
                                                       comp_d, comp_i, atom_d, atom_t, comp_ids)
                        jobs.append(pid)
            if k == 0:
                self.phase_index, self.comp_index = dict(phase_ids), dict(comp_ids)
        self.K, self.M, self.S, self.Z = K, int(M or 0), S, Z
        self.phase_i = np.ascontiguousarray(np.array(phase_i, dtype=np.int32).reshape(-1, PHASE_I_STRIDE))
        self.phase_d = np.ascontiguousarray(np.array(phase_d, dtype=float).reshape(-1, PHASE_D_STRIDE))
        self.phase_comp = np.ascontiguousarray(np.array(phase_comp, dtype=np.int32))
        self.matrices = np.ascontiguousarray(np.concatenate(matrices) if matrices else np.zeros(0), dtype=float)
        self._atype_memo = {}              # ids are only meaningful while the graphs are being walked
        self.comp_d = np.ascontiguousarray(np.array(comp_d, dtype=float).reshape(-1, COMP_D_STRIDE))
        self.comp_i = np.ascontiguousarray(np.array(comp_i, dtype=np.int32).reshape(-1, COMP_I_STRIDE))
        self.atom_d = np.ascontiguousarray(np.array(atom_d, dtype=float).reshape(-1, ATOM_D_STRIDE))
        self.atom_type = np.ascontiguousarray(np.array(atom_t, dtype=np.int32))
        self.job_phase = np.ascontiguousarray(np.array(jobs, dtype=np.int32))
        self.fractions = np.ascontiguousarray(np.array(fractions, dtype=float).reshape(K, self.M))
        self.scales = np.ascontiguousarray(scales)
        self.bgshifts = np.ascontiguousarray(bgshifts)
        if any(self.prob_model_list):
            self.prob_model = np.ascontiguousarray(np.array(self.prob_model_list, dtype=np.int32))
            self.prob_params = np.zeros((len(self.prob_param_list), PROB_STRIDE))
            for i, row in enumerate(self.prob_param_list):
                self.prob_params[i, :len(row)] = row
        else:
            self.prob_model = self.prob_params = None

    _atom_index_open = True

    def _pack_phase(self, phase, static, phase_i, phase_d, phase_comp, matrices, mat_off,
                    comp_d, comp_i, atom_d, atom_t, comp_ids):
        ptype = getattr(phase, "type", "Phase")
        if ptype == "AbtractPhase":
            raise NotImplementedError          # phases.py:74-75
        if ptype not in ("Phase", "RawPatternPhase"):
            raise NotImplementedError("unknown phase type %r" % (ptype,))
        flags = 0
        if getattr(phase, "apply_lpf", True):
            flags |= FLAG_APPLY_LPF
        if getattr(phase, "apply_correction", True):
            flags |= FLAG_APPLY_CORRECTION
        if ptype == "RawPatternPhase":
            # phases.py:97-104: interp1d sorts its abscissa (stable argsort) before searching; the
            # device kernel expects raw_x ascending followed by raw_y (layout only)
            rx = np.asarray(phase.raw_pattern_x, dtype=float).ravel()
            ry = np.asarray(phase.raw_pattern_y, dtype=float).ravel()
            if rx.size != ry.size:
                raise ValueError("x and y arrays must be equal in length along interpolation axis.")
            order = np.argsort(rx, kind="mergesort")
            phase_i.append([0, 0, int(rx.size), 0, flags | FLAG_VALID_PROBS | FLAG_RAW_PATTERN, 0, mat_off, 0])
            phase_d.append([float(getattr(phase, "sigma_star", 3.0))] + [0.0] * (PHASE_D_STRIDE - 1))
            matrices.append(rx[order])
            matrices.append(ry[order])
            self.components_of_phase.append([])
            self.prob_model_list.append(0)
            self.prob_param_list.append([])
            return mat_off + 2 * rx.size
        model, model_params = prob_model_of(phase)
        if not model and not phase.valid_probs:
            phase_i.append([0, 0, 0, 0, flags, 0, 0, 0])
            phase_d.append([0.0] * PHASE_D_STRIDE)
            self.components_of_phase.append([])
            self.prob_model_list.append(0)
            self.prob_param_list.append([])
            return mat_off
        flags |= FLAG_VALID_PROBS          # model-driven phases: the device decides (k_probabilities)
        G = int(phase.G)
        if model:
            rank = _PROB_RANK[model](G)
            W = P = np.zeros((rank, rank))  # pool entry the device fills
        else:
            W = np.asarray(phase.W, dtype=float)
            P = np.asarray(phase.P, dtype=float)
            rank = P.shape[1]
        self.prob_model_list.append(model)
        self.prob_param_list.append(list(model_params))
        if W.shape != (rank, rank) or P.shape != (rank, rank):
            raise ValueError("W / P must be rank x rank")
        csds = phase.CSDS
        phase_i.append([G, rank, int(csds.minimum), int(csds.maximum), flags, len(phase_comp), mat_off, 0])
        phase_d.append([float(phase.sigma_star), float(csds.average), float(csds.alpha_scale),
                        float(csds.alpha_offset), float(csds.beta_scale), float(csds.beta_offset), 0.0, 0.0])
        matrices.append(W.ravel())
        matrices.append(P.ravel())
        mat_off += 2 * rank * rank
        comps = list(phase.components)
        if len(comps) != G:
            raise ValueError("phase has %d components but G = %d" % (len(comps), G))
        self.components_of_phase.append(comps)
        for comp in comps:
            cid = comp_ids.get(id(comp))
            if cid is None:
                cid = len(comp_d)
                comp_ids[id(comp)] = cid
                # atoms of one group that share a z are made adjacent: the kernels evaluate one
                # complex exponential per run of equal z (layer atoms stay in front, the z-stretch
                # applies to the interlayer group only, components.py:40-45)
                zkey = lambda atom: float(atom.default_z) if atom is not None else 0.0  # noqa: E731
                layer = sorted(comp.layer_atoms, key=zkey)
                inter = sorted(comp.interlayer_atoms, key=zkey)
                comp_d.append([float(comp.d001), float(comp.default_c), float(comp.delta_c), float(comp.lattice_d),
                               float(comp.volume), float(comp.weight), 0.0, 0.0])
                comp_i.append([len(atom_t), len(layer), len(inter), 0])
                for atom in layer + inter:
                    if self._atom_index_open and atom is not None:
                        self.atom_index[id(atom)] = len(atom_t)
                    if atom is None or atom.atom_type is None:
                        atom_d.extend([0.0, 0.0])
                        atom_t.append(-1)
                    else:
                        atom_d.extend([float(atom.pn), float(atom.default_z)])
                        atom_t.append(static.atom_types[_atype_key(atom.atom_type, self._atype_memo)])
            phase_comp.append(cid)
        return mat_off
