"""ctypes binding of ``libpyxrd_b200.so`` (the C-ABI of include/pyxrd_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``pyxrd_b200/csrc/build.py``.  There is no CPU fallback: if the library is
missing or no CUDA device is present, constructing a :class:`Context` raises.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
import sys
import threading
from typing import Optional

import numpy as np

from .packing import BatchPack, StaticPack

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpyxrd_b200.so")

_f64p = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


class StaticDesc(C.Structure):
    _fields_ = [("n_specimens", C.c_int32), ("n_atom_types", C.c_int32), ("n_patterns", C.c_int32),
                ("reserved", C.c_int32), ("spec_npts", _i32p), ("theta", _f64p), ("spec_gonio", _f64p),
                ("spec_nwl", _i32p), ("wl_fraction", _f64p), ("wl_index", _i32p), ("wl_w_hi", _f64p),
                ("wl_w_lo", _f64p), ("observed", _f64p), ("selected", _u8p), ("atom_types", _f64p),
                ("stl_override", _f64p)]


class BatchDesc(C.Structure):
    _fields_ = [("n_candidates", C.c_int32), ("n_phase_slots", C.c_int32), ("n_phases", C.c_int32),
                ("n_components", C.c_int32), ("n_atoms", C.c_int32), ("n_phase_comp", C.c_int32),
                ("n_matrix", C.c_int64), ("phase_i", _i32p), ("phase_d", _f64p), ("phase_comp", _i32p),
                ("matrices", _f64p), ("comp_d", _f64p), ("comp_i", _i32p), ("atom_d", _f64p),
                ("atom_type", _i32p), ("job_phase", _i32p), ("fractions", _f64p), ("scales", _f64p),
                ("bgshifts", _f64p), ("prob_model", _i32p), ("prob_params", _f64p)]


class Binding(C.Structure):
    _fields_ = [("param", C.c_int32), ("target", C.c_int32), ("index", C.c_int32), ("mode", C.c_int32),
                ("scale", C.c_double), ("offset", C.c_double)]


BINDING_DTYPE = np.dtype([("param", "<i4"), ("target", "<i4"), ("index", "<i4"), ("mode", "<i4"),
                          ("scale", "<f8"), ("offset", "<f8")])


class PopulationDesc(C.Structure):
    _fields_ = [("n_candidates", C.c_int32), ("n_params", C.c_int32), ("n_bindings", C.c_int32),
                ("n_specimens", C.c_int32), ("n_patterns", C.c_int32), ("reserved", C.c_int32),
                ("x", _f64p), ("bindings", C.POINTER(Binding)), ("csds_auto_maximum", _u8p),
                ("csds_factor", C.c_double)]


# every symbol include/pyxrd_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "pyxrd_create", "pyxrd_destroy", "pyxrd_last_error", "pyxrd_set_stream", "pyxrd_synchronize",
    "pyxrd_launch_count", "pyxrd_set_option", "pyxrd_set_static", "pyxrd_set_batch", "pyxrd_set_solution", "pyxrd_set_intensities",
    "pyxrd_compute_intensities", "pyxrd_compute_residuals", "pyxrd_compute_all", "pyxrd_set_residual_method", "pyxrd_optimize", "pyxrd_read_solutions",
    "pyxrd_read_opt_residuals", "pyxrd_intensity_count", "pyxrd_total_count",
    "pyxrd_read_intensities", "pyxrd_read_scaled", "pyxrd_read_totals", "pyxrd_read_residuals",
    "pyxrd_read_correction", "pyxrd_device_buffers", "pyxrd_evaluate_host", "pyxrd_optimize_host", "pyxrd_leaf_asf",
    "pyxrd_leaf_lpf", "pyxrd_leaf_csds", "pyxrd_leaf_component", "pyxrd_stage_times", "pyxrd_dfma_peak",
    "pyxrd_expand_population", "pyxrd_expanded_desc", "pyxrd_expanded_free", "pyxrd_expand_error",
    "pyxrd_set_population", "pyxrd_update_population", "pyxrd_device_results", "pyxrd_put_rows_to_peers", "pyxrd_evaluate_population", "pyxrd_optimize_population", "pyxrd_leaf_probabilities",
    "pyxrd_leaf_statistic", "pyxrd_leaf_derive", "pyxrd_leaf_lorentz", "pyxrd_leaf_mmult", "pyxrd_leaf_matrix_powers",
]

_lib = None
_lib_lock = threading.Lock()
_token_counter = itertools.count(1)
# set in a child that was forked after its parent had initialised CUDA (os.register_at_fork below)
_forked_after_cuda = False


class NativeLibraryMissing(RuntimeError):
    pass


class DeviceError(RuntimeError):
    pass


def load_library():
    """dlopen the in-tree shared library (no compute, no CUDA initialisation)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("PYXRD_B200_LIB", LIB_PATH)      # A/B builds of the same source (tools/)
        if not os.path.exists(path):
            raise NativeLibraryMissing(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(pyxrd_b200 has no CPU fallback)" % path)
        lib = C.CDLL(path)
        vp = C.c_void_p
        lib.pyxrd_create.argtypes = [C.c_int, C.POINTER(vp)]
        lib.pyxrd_destroy.argtypes = [vp]
        lib.pyxrd_destroy.restype = None
        lib.pyxrd_last_error.argtypes = [vp]
        lib.pyxrd_last_error.restype = C.c_char_p
        lib.pyxrd_set_stream.argtypes = [vp, vp]
        lib.pyxrd_synchronize.argtypes = [vp]
        lib.pyxrd_launch_count.argtypes = [vp]
        lib.pyxrd_launch_count.restype = C.c_int64
        lib.pyxrd_set_option.argtypes = [vp, C.c_char_p, C.c_double]
        lib.pyxrd_set_static.argtypes = [vp, C.POINTER(StaticDesc)]
        lib.pyxrd_set_batch.argtypes = [vp, C.POINTER(BatchDesc)]
        lib.pyxrd_set_solution.argtypes = [vp, _f64p, _f64p, _f64p]
        lib.pyxrd_set_intensities.argtypes = [vp, _f64p, C.c_int64]
        lib.pyxrd_compute_intensities.argtypes = [vp]
        lib.pyxrd_compute_residuals.argtypes = [vp, C.c_int]
        lib.pyxrd_compute_all.argtypes = [vp, C.c_int]
        lib.pyxrd_set_residual_method.argtypes = [vp, C.c_int]
        lib.pyxrd_optimize.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        lib.pyxrd_read_solutions.argtypes = [vp, _f64p, C.c_int64]
        lib.pyxrd_read_opt_residuals.argtypes = [vp, _f64p, C.c_int64]
        lib.pyxrd_intensity_count.argtypes = [vp]
        lib.pyxrd_intensity_count.restype = C.c_int64
        lib.pyxrd_total_count.argtypes = [vp]
        lib.pyxrd_total_count.restype = C.c_int64
        for name in ("pyxrd_read_intensities", "pyxrd_read_scaled", "pyxrd_read_totals", "pyxrd_read_residuals",
                     "pyxrd_read_correction"):
            getattr(lib, name).argtypes = [vp, _f64p, C.c_int64]
        lib.pyxrd_device_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        lib.pyxrd_evaluate_host.argtypes = [vp, C.POINTER(BatchDesc), _f64p, _f64p]
        lib.pyxrd_optimize_host.argtypes = [vp, C.POINTER(BatchDesc), C.c_int, C.c_int, C.c_int, _f64p, _f64p]
        lib.pyxrd_leaf_asf.argtypes = [vp, _f64p, C.c_int64, _f64p, _f64p]
        lib.pyxrd_leaf_statistic.argtypes = [vp, C.c_int, _f64p, _f64p, _f64p, C.c_int64, C.c_double, _f64p]
        lib.pyxrd_leaf_derive.argtypes = [vp, _f64p, C.c_int64, C.c_int, _f64p]
        lib.pyxrd_leaf_mmult.argtypes = [vp, _f64p, _f64p, C.c_int64, C.c_int, _f64p]
        lib.pyxrd_leaf_matrix_powers.argtypes = [vp, _f64p, C.c_int64, C.c_int, C.c_int, _f64p]
        lib.pyxrd_leaf_lorentz.argtypes = [vp, _f64p, C.c_int64, C.c_double, C.c_double, C.c_double, _f64p]
        lib.pyxrd_leaf_lpf.argtypes = [vp, _f64p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, _f64p]
        lib.pyxrd_leaf_csds.argtypes = [vp, _f64p, C.c_int, C.c_int, _f64p, _f64p, _f64p]
        lib.pyxrd_leaf_component.argtypes = [vp, _f64p, C.c_int64, _f64p, C.c_int, C.c_int, _f64p, _f64p, _u8p,
                                             _f64p, _f64p, _f64p]
        lib.pyxrd_stage_times.argtypes = [vp, _f64p]
        lib.pyxrd_expand_population.argtypes = [C.POINTER(BatchDesc), C.POINTER(PopulationDesc), C.POINTER(vp)]
        lib.pyxrd_expanded_desc.argtypes = [vp, C.POINTER(BatchDesc)]
        lib.pyxrd_expanded_desc.restype = None
        lib.pyxrd_expanded_free.argtypes = [vp]
        lib.pyxrd_expanded_free.restype = None
        lib.pyxrd_expand_error.argtypes = []
        lib.pyxrd_expand_error.restype = C.c_char_p
        lib.pyxrd_set_population.argtypes = [vp, C.POINTER(BatchDesc), C.POINTER(PopulationDesc)]
        lib.pyxrd_update_population.argtypes = [vp, _f64p]
        lib.pyxrd_device_results.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
        lib.pyxrd_put_rows_to_peers.argtypes = [vp, vp, C.c_int64, C.c_int, C.POINTER(vp), C.c_int, C.c_int64]
        lib.pyxrd_evaluate_population.argtypes = [vp, C.POINTER(BatchDesc), C.POINTER(PopulationDesc), _f64p]
        lib.pyxrd_optimize_population.argtypes = [vp, C.POINTER(BatchDesc), C.POINTER(PopulationDesc), C.c_int, C.c_int,
                                                  C.c_int, _f64p, _f64p]
        lib.pyxrd_leaf_probabilities.argtypes = [vp, C.c_int, C.c_int, _f64p, _f64p, _f64p, C.POINTER(C.c_int),
                                                 C.POINTER(C.c_int)]
        lib.pyxrd_dfma_peak.argtypes = [vp, C.c_int, _f64p, _f64p]
        _lib = lib
        return lib


def _p(arr, typ):
    return arr.ctypes.data_as(typ)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context(object):
    """One device context = one CUDA stream + resident tables / batch / results."""

    def __init__(self, device: int = 0):
        if _forked_after_cuda:
            raise DeviceError(
                "this process was forked from a process that had already initialised CUDA; CUDA cannot be used in such "
                "a child.  Evaluate trials in the parent through pyxrd_b200.provider.B200AsyncServerProvider (one "
                "batched device call per generation) or start the workers with the 'spawn' start method")
        self._lib = load_library()
        self._h = C.c_void_p()
        rc = self._lib.pyxrd_create(int(device), C.byref(self._h))
        if rc != 0 or not self._h:
            self._h = C.c_void_p()
            raise DeviceError("pyxrd_create(device=%d) failed with status %d: no usable CUDA device "
                              "(pyxrd_b200 has no CPU fallback)" % (device, rc))
        self.device = int(device)
        self.lock = threading.RLock()
        self._static: Optional[StaticPack] = None
        self._static_sig = None
        self._batch: Optional[BatchPack] = None
        self._keep = []
        # which parsed mixture the resident batch + intensities belong to (calculations/mixture.py); every call that
        # replaces the static tables, the batch or its intensities clears it
        self._resident_token = None
        # which population (template, table, K) is resident for pyxrd_update_population; cleared together with the token
        self._population_key = None

    def new_resident_token(self):
        """mark the resident batch as belonging to one caller-side object; unique across processes and contexts (the
        token is stored on a picklable DataObject and may travel to another process)"""
        self._resident_token = (os.getpid(), id(self), next(_token_counter))
        return self._resident_token

    def invalidate_residency(self):
        self._resident_token = None
        self._population_key = None

    # -- plumbing ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.pyxrd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.pyxrd_last_error(self._h)
            raise DeviceError("%s failed: %s" % (what, msg.decode("utf-8", "replace") if msg else rc))

    def set_stream(self, cuda_stream_handle: Optional[int]):
        self._check(self._lib.pyxrd_set_stream(self._h, C.c_void_p(cuda_stream_handle or 0)), "set_stream")
        self._stream_handle = int(cuda_stream_handle or 0)

    def synchronize(self):
        self._check(self._lib.pyxrd_synchronize(self._h), "synchronize")

    @property
    def launch_count(self) -> int:
        return int(self._lib.pyxrd_launch_count(self._h))

    def set_option(self, name: str, value: float):
        """measurement / experiment switches (include/pyxrd_b200.h: pyxrd_set_option); nothing is read from the environment"""
        self._check(self._lib.pyxrd_set_option(self._h, name.encode("ascii"), float(value)), "set_option")

    # -- upload -----------------------------------------------------------------------
    def set_static(self, static: StaticPack, force: bool = False):
        if not force and static is self._static:          # the very pack that is resident
            return False
        sig = static.signature()
        if not force and self._static_sig == sig:
            self._static = static
            return False
        d = StaticDesc(static.S, static.T, static.Z, 0, _p(static.npts, _i32p), _p(static.theta, _f64p),
                       _p(static.gonio, _f64p), _p(static.nwl, _i32p), _p(static.wl_fraction, _f64p),
                       _p(static.wl_index, _i32p), _p(static.wl_w_hi, _f64p), _p(static.wl_w_lo, _f64p),
                       _p(static.observed, _f64p), _p(static.selected, _u8p), _p(static.atype_table, _f64p),
                       _p(static.stl_override, _f64p) if static.stl_override is not None else None)
        self.invalidate_residency()
        self._check(self._lib.pyxrd_set_static(self._h, C.byref(d)), "set_static")
        self._static, self._static_sig, self._batch = static, sig, None
        return True

    @staticmethod
    def batch_desc(b: BatchPack) -> BatchDesc:
        return BatchDesc(b.K, b.M, b.phase_i.shape[0], b.comp_d.shape[0], b.atom_type.shape[0],
                         b.phase_comp.shape[0], b.matrices.shape[0], _p(b.phase_i, _i32p), _p(b.phase_d, _f64p),
                         _p(b.phase_comp, _i32p), _p(b.matrices, _f64p), _p(b.comp_d, _f64p), _p(b.comp_i, _i32p),
                         _p(b.atom_d, _f64p), _p(b.atom_type, _i32p), _p(b.job_phase, _i32p),
                         _p(b.fractions, _f64p), _p(b.scales, _f64p), _p(b.bgshifts, _f64p),
                         _p(b.prob_model, _i32p) if getattr(b, "prob_model", None) is not None else None,
                         _p(b.prob_params, _f64p) if getattr(b, "prob_params", None) is not None else None)

    def set_batch(self, batch: BatchPack):
        self.invalidate_residency()
        d = self.batch_desc(batch)
        self._check(self._lib.pyxrd_set_batch(self._h, C.byref(d)), "set_batch")
        self._batch = batch

    def set_solution(self, fractions, scales, bgshifts):
        b = self._batch
        f, s, g = _f64(fractions).reshape(b.K, b.M), _f64(scales).reshape(b.K, b.S), _f64(bgshifts).reshape(b.K, b.S)
        self._check(self._lib.pyxrd_set_solution(self._h, _p(f, _f64p), _p(s, _f64p), _p(g, _f64p)), "set_solution")

    def set_intensities(self, flat):
        self.invalidate_residency()
        flat = _f64(flat).ravel()
        self._check(self._lib.pyxrd_set_intensities(self._h, _p(flat, _f64p), flat.size), "set_intensities")

    # -- compute ----------------------------------------------------------------------
    def compute_intensities(self):
        self._check(self._lib.pyxrd_compute_intensities(self._h), "compute_intensities")

    def set_residual_method(self, name: str = "Rp"):
        """settings.RESIDUAL_METHOD (mixture.py:22-27,89).  "Rp" and "Rpder" run on the device; "Rpw" and
        "Rphase" raise exactly what the reference raises for them (its own implementations cannot be
        called with the 2-D clipped patterns mixture.py:88-89 passes: statistics.py:58-61,81)."""
        if name == "Rpw":
            raise ValueError("The truth value of an array with more than one element is ambiguous. "
                             "Use a.any() or a.all()")
        if name == "Rphase":
            raise TypeError("Rphase() missing 1 required positional argument: 'phase'")
        code = {"Rp": 0, "Rpder": 1}[name]          # KeyError for an unknown name, like the reference's dict
        if getattr(self, "_residual_method", 0) != code:
            self._check(self._lib.pyxrd_set_residual_method(self._h, code), "set_residual_method")
            self._residual_method = code

    def compute_residuals(self, want_scaled: bool = False, residuals_only: bool = False):
        """``residuals_only``: the total patterns are not stored (PYXRD_RESIDUALS_ONLY)"""
        flags = int(bool(want_scaled)) | (2 if residuals_only else 0)
        self._check(self._lib.pyxrd_compute_residuals(self._h, flags), "compute_residuals")

    def compute_all(self, want_scaled: bool = False, residuals_only: bool = False):
        """phase patterns + wavelength distribution + totals / residuals of the resident batch (the last wavelength
        pass runs inside the residual kernel)"""
        flags = int(bool(want_scaled)) | (2 if residuals_only else 0)
        self._check(self._lib.pyxrd_compute_all(self._h, flags), "compute_all")

    def optimize(self, refine_bg: bool = True, outer: int = 0, newton: int = 0):
        """device optimiser over the resident batch (needs computed intensities)"""
        self._check(self._lib.pyxrd_optimize(self._h, int(bool(refine_bg)), int(outer), int(newton)), "optimize")

    def read_solutions(self):
        b = self._batch
        return self._read(self._lib.pyxrd_read_solutions, b.K * (b.M + 2 * b.S)).reshape(b.K, b.M + 2 * b.S)

    def read_opt_residuals(self):
        b = self._batch
        return self._read(self._lib.pyxrd_read_opt_residuals, b.K * (b.S + 1)).reshape(b.K, b.S + 1)

    def evaluate_host(self, batch: BatchPack, want_intensities: bool = False, out_residuals=None,
                      out_intensities=None):
        """one C-ABI call: upload K candidates, compute, return residuals [K, 1+S] (+ intensities).
        ``out_*`` may be preallocated (e.g. pinned) float64 arrays."""
        self.invalidate_residency()
        d = self.batch_desc(batch)
        res = np.empty(batch.K * (batch.S + 1)) if out_residuals is None else out_residuals
        inten = out_intensities
        if inten is None and want_intensities:
            inten = np.empty(batch.K * batch.M * batch.Z * self._static.sumX)
        self._check(self._lib.pyxrd_evaluate_host(self._h, C.byref(d), _p(res, _f64p),
                                                  _p(inten, _f64p) if inten is not None else None), "evaluate_host")
        self._batch = batch
        return res.reshape(batch.K, batch.S + 1), inten

    def optimize_host(self, batch: BatchPack, refine_bg: bool = True, outer: int = 0, newton: int = 0,
                      out_residuals=None, out_solutions=None):
        """one C-ABI call per generation: K candidates in (host), optimised residuals [K, 1+S] and
        solutions [K, M+2S] out (host)"""
        self.invalidate_residency()
        d = self.batch_desc(batch)
        res = np.empty(batch.K * (batch.S + 1)) if out_residuals is None else out_residuals
        sol = np.empty(batch.K * (batch.M + 2 * batch.S)) if out_solutions is None else out_solutions
        self._check(self._lib.pyxrd_optimize_host(self._h, C.byref(d), int(bool(refine_bg)), int(outer), int(newton),
                                                  _p(res, _f64p), _p(sol, _f64p)), "optimize_host")
        self._batch = batch
        return res.reshape(batch.K, batch.S + 1), sol.reshape(batch.K, batch.M + 2 * batch.S)

    # -- results ----------------------------------------------------------------------
    def _read(self, fn, count):
        out = np.empty(int(count))
        self._check(fn(self._h, _p(out, _f64p), int(count)), fn.__name__)
        return out

    def read_intensities_flat(self):
        return self._read(self._lib.pyxrd_read_intensities, self._lib.pyxrd_intensity_count(self._h))

    def read_scaled_flat(self):
        return self._read(self._lib.pyxrd_read_scaled, self._lib.pyxrd_intensity_count(self._h))

    def read_totals_flat(self):
        return self._read(self._lib.pyxrd_read_totals, self._lib.pyxrd_total_count(self._h))

    def read_residuals(self):
        b = self._batch
        return self._read(self._lib.pyxrd_read_residuals, b.K * (b.S + 1)).reshape(b.K, b.S + 1)

    def read_correction_flat(self):
        return self._read(self._lib.pyxrd_read_correction, self._static.sumX)

    def split_per_specimen(self, flat, lead):
        """flat [K][S][lead...][X_s] -> list over k of list over s of arrays [lead..., X_s]"""
        st, b = self._static, self._batch
        n_lead = int(np.prod(lead))
        out, pos = [], 0
        for _ in range(b.K):
            per = []
            for s in range(st.S):
                X = int(st.npts[s])
                per.append(flat[pos:pos + n_lead * X].reshape(tuple(lead) + (X,)))
                pos += n_lead * X
            out.append(per)
        return out

    def device_buffers(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._lib.pyxrd_device_buffers(self._h, C.byref(a), C.byref(b), C.byref(c)), "device_buffers")
        return a.value, b.value, c.value

    # -- leaf helpers -----------------------------------------------------------------
    def leaf_asf(self, s, atom_type_row):
        s = _f64(s).ravel()
        row = _f64(atom_type_row).ravel()
        out = np.empty_like(s)
        self._check(self._lib.pyxrd_leaf_asf(self._h, _p(s, _f64p), s.size, _p(row, _f64p), _p(out, _f64p)), "leaf_asf")
        return out

    def leaf_statistic(self, kind: int, exp, calc=None, phase=None, extra: float = 0.0) -> float:
        """statistics.py on plain arrays: kind 0 Rp, 1 Rpw, 2 R_squared, 3 Rpe (extra = parameters), 4 Rphase"""
        e = _f64(exp).ravel()
        c = _f64(calc).ravel() if calc is not None else None
        p = _f64(phase).ravel() if phase is not None else None
        for other in (c, p):
            if other is not None and other.size != e.size:
                raise ValueError("operands could not be broadcast together with shapes (%d,) (%d,)" % (e.size, other.size))
        out = np.zeros(1)
        self._check(self._lib.pyxrd_leaf_statistic(self._h, int(kind), _p(e, _f64p), _p(c, _f64p) if c is not None else None,
                                                   _p(p, _f64p) if p is not None else None, e.size, float(extra),
                                                   _p(out, _f64p)), "leaf_statistic")
        return float(out[0])

    def leaf_derive(self, pattern, smooth_only: bool = False):
        v = _f64(pattern).ravel()
        if v.size < 31:
            raise ValueError("Input vector needs to be bigger than window size.")        # math_tools.py:87-88
        out = np.empty_like(v)
        self._check(self._lib.pyxrd_leaf_derive(self._h, _p(v, _f64p), v.size, int(bool(smooth_only)), _p(out, _f64p)),
                    "leaf_derive")
        return out

    def leaf_mmult(self, A, B):
        """math_tools.mmult: batched complex matrix product [X, r, r] x [X, r, r] -> complex128[X, r, r]"""
        A = np.ascontiguousarray(A, dtype=np.complex128)
        B = np.ascontiguousarray(B, dtype=np.complex128)
        if A.ndim != 3 or A.shape != B.shape or A.shape[1] != A.shape[2]:
            raise ValueError("operands could not be broadcast together with shapes %r %r" % (A.shape, B.shape))
        out = np.empty_like(A)
        self._check(self._lib.pyxrd_leaf_mmult(self._h, A.ctypes.data_as(_f64p), B.ctypes.data_as(_f64p), A.shape[0], A.shape[1],
                                               out.ctypes.data_as(_f64p)), "leaf_mmult")
        return out

    def leaf_matrix_powers(self, Q, n_max: int):
        """phases.get_Q_matrices: complex128[n_max + 1, X, r, r], entry n = Q^(n + 1)"""
        Q = np.ascontiguousarray(Q, dtype=np.complex128)
        if Q.ndim != 3 or Q.shape[1] != Q.shape[2]:
            raise ValueError("Q must be [X, r, r], got %r" % (Q.shape,))
        out = np.empty((int(n_max) + 1,) + Q.shape, dtype=np.complex128)
        self._check(self._lib.pyxrd_leaf_matrix_powers(self._h, Q.ctypes.data_as(_f64p), Q.shape[0], Q.shape[1], int(n_max),
                                                       out.ctypes.data_as(_f64p)), "leaf_matrix_powers")
        return out

    def leaf_lpf(self, theta, sigma_star, soller1, soller2, mcr_2theta):
        t = _f64(theta).ravel()
        out = np.empty_like(t)
        self._check(self._lib.pyxrd_leaf_lpf(self._h, _p(t, _f64p), t.size, float(sigma_star), float(soller1),
                                             float(soller2), float(mcr_2theta), _p(out, _f64p)), "leaf_lpf")
        return out

    def leaf_T(self, theta, sigma_star, soller1, soller2):
        t = _f64(theta).ravel()
        out = np.empty_like(t)
        self._check(self._lib.pyxrd_leaf_lorentz(self._h, _p(t, _f64p), t.size, float(sigma_star), float(soller1), float(soller2),
                                           _p(out, _f64p)), "leaf_T")
        return out

    def leaf_csds(self, average, alpha_scale, alpha_offset, beta_scale, beta_offset, minimum, maximum):
        p = _f64([average, alpha_scale, alpha_offset, beta_scale, beta_offset])
        arr = np.empty(int(maximum) + 1)
        mean = np.empty(1)
        coef = np.empty(int(maximum) + 2)
        self._check(self._lib.pyxrd_leaf_csds(self._h, _p(p, _f64p), int(minimum), int(maximum), _p(arr, _f64p),
                                              _p(mean, _f64p), _p(coef, _f64p)), "leaf_csds")
        return arr, float(mean[0]), coef

    def leaf_component(self, stl, comp_row, n_layer, n_inter, atom_d, atom_rows, has_type):
        """(SF complex[n], phi complex[n], z[NA]) of one component on an arbitrary stl axis"""
        stl = _f64(stl).ravel()
        comp_row = _f64(comp_row).ravel()
        na = int(n_layer) + int(n_inter)
        atom_d = _f64(atom_d).reshape(max(na, 0), 2) if na else np.zeros((1, 2))
        atom_rows = _f64(atom_rows).reshape(na, 12) if na else np.zeros((1, 12))
        has = np.ascontiguousarray(has_type, dtype=np.uint8) if na else np.zeros(1, dtype=np.uint8)
        sf = np.zeros(stl.size, dtype=np.complex128)
        phi = np.zeros(stl.size, dtype=np.complex128)
        z = np.zeros(max(na, 1))
        self._check(self._lib.pyxrd_leaf_component(self._h, _p(stl, _f64p), stl.size, _p(comp_row, _f64p), int(n_layer),
                                                   int(n_inter), _p(atom_d, _f64p), _p(atom_rows, _f64p), _p(has, _u8p),
                                                   sf.ctypes.data_as(_f64p), phi.ctypes.data_as(_f64p), _p(z, _f64p)),
                    "leaf_component")
        return sf, phi, z[:na]

    def leaf_probabilities(self, model, G, params):
        """(W [r, r], P [r, r], valid_probs) of one probability model (PYXRD_PROB_*), evaluated on the
        device as probabilities/models/*.py update() + solve() + validate() do"""
        from .packing import PROB_MODELS, PROB_STRIDE
        code = PROB_MODELS[model] if isinstance(model, str) else int(model)
        p = np.zeros(PROB_STRIDE)
        params = _f64(params).ravel()
        p[:params.size] = params
        W, P = np.zeros(81), np.zeros(81)
        valid, rank = C.c_int(), C.c_int()
        self._check(self._lib.pyxrd_leaf_probabilities(self._h, code, int(G), _p(p, _f64p), _p(W, _f64p), _p(P, _f64p),
                                                       C.byref(valid), C.byref(rank)), "leaf_probabilities")
        r = rank.value
        return W[:r * r].reshape(r, r).copy(), P[:r * r].reshape(r, r).copy(), bool(valid.value)

    # -- populations from a template + binding table ----------------------------------------
    def _population_desc(self, template: BatchPack, x, bindings, csds_auto=None, csds_factor=2.5):
        self.invalidate_residency()               # every population call replaces the resident batch
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError("x must be [K, d]")
        bindings = np.ascontiguousarray(bindings, dtype=BINDING_DTYPE)
        auto = None if csds_auto is None else np.ascontiguousarray(csds_auto, dtype=np.uint8)
        pop = PopulationDesc(x.shape[0], x.shape[1], bindings.size, template.S, template.Z, 0, _p(x, _f64p),
                             bindings.ctypes.data_as(C.POINTER(Binding)), _p(auto, _u8p) if auto is not None else None,
                             float(csds_factor))
        return pop, (x, bindings, auto)

    def evaluate_population(self, template: BatchPack, x, bindings, csds_auto=None, csds_factor=2.5):
        """residuals f64[K, 1+S] of the K candidates ``x`` [K, d] describes (one C-ABI call)"""
        d = self.batch_desc(template)
        pop, keep = self._population_desc(template, x, bindings, csds_auto, csds_factor)
        K = keep[0].shape[0]
        res = np.empty(K * (template.S + 1))
        self._check(self._lib.pyxrd_evaluate_population(self._h, C.byref(d), C.byref(pop), _p(res, _f64p)),
                    "evaluate_population")
        self._batch = _Shape(K, template.M, template.S, template.Z)
        return res.reshape(K, template.S + 1)

    def set_population(self, template: BatchPack, x, bindings, csds_auto=None, csds_factor=2.5, key=None):
        """upload template + solution vectors + table, expand on the device, run the batch prologue.  ``key``: any
        hashable identity of (template, table); a later ``update_population`` of the same key and K re-uses the
        resident template"""
        d = self.batch_desc(template)
        pop, keep = self._population_desc(template, x, bindings, csds_auto, csds_factor)
        self._check(self._lib.pyxrd_set_population(self._h, C.byref(d), C.byref(pop)), "set_population")
        self._batch = _Shape(keep[0].shape[0], template.M, template.S, template.Z)
        self._population_key = None if key is None else (key, keep[0].shape)

    def population_is_resident(self, key, shape) -> bool:
        return key is not None and self._population_key == (key, tuple(shape))

    def update_population(self, x=None):
        """the next generation of the resident population: only x f64[K, d] is uploaded (None: the resident vectors
        are expanded again)"""
        if self._population_key is None and x is not None:
            raise DeviceError("update_population: no resident population")
        if x is not None:
            x = np.ascontiguousarray(x, dtype=np.float64)
            if x.shape != self._population_key[1]:
                raise ValueError("update_population: x is %r, the resident population is %r" % (x.shape, self._population_key[1]))
        keep_key, keep_batch = self._population_key, self._batch
        self._resident_token = None
        try:
            self._check(self._lib.pyxrd_update_population(self._h, _p(x, _f64p) if x is not None else None), "update_population")
        except DeviceError:
            self._population_key = None
            raise
        self._population_key, self._batch = keep_key, keep_batch

    def load_population(self, template: BatchPack, x, bindings, csds_auto=None, csds_factor=2.5, key=None):
        """``update_population`` when (key, x.shape) is resident, else ``set_population``"""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if self.population_is_resident(key, x.shape):
            try:
                return self.update_population(x)
            except DeviceError:
                pass                # e.g. dropped on the library side: upload everything again
        self.set_population(template, x, bindings, csds_auto, csds_factor, key=key)

    def optimize_population(self, template: BatchPack, x, bindings, csds_auto=None, csds_factor=2.5,
                            refine_bg: bool = True, outer: int = 0, newton: int = 0):
        """(optimised residuals [K, 1+S], solutions [K, M+2S]) of the K candidates (one C-ABI call)"""
        d = self.batch_desc(template)
        pop, keep = self._population_desc(template, x, bindings, csds_auto, csds_factor)
        K = keep[0].shape[0]
        res = np.empty(K * (template.S + 1))
        sol = np.empty(K * (template.M + 2 * template.S))
        self._check(self._lib.pyxrd_optimize_population(self._h, C.byref(d), C.byref(pop), int(bool(refine_bg)),
                                                        int(outer), int(newton), _p(res, _f64p), _p(sol, _f64p)),
                    "optimize_population")
        self._batch = _Shape(K, template.M, template.S, template.Z)
        return res.reshape(K, template.S + 1), sol.reshape(K, template.M + 2 * template.S)

    def stage_times(self):
        """dict of CUDA-event milliseconds of the latest phase / wavelength / residual / prologue stages"""
        out = np.zeros(4)
        self._check(self._lib.pyxrd_stage_times(self._h, _p(out, _f64p)), "stage_times")
        return dict(phase=out[0], wavelength=out[1], residual=out[2], prologue=out[3])

    def device_results(self):
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self._lib.pyxrd_device_results(self._h, C.byref(a), C.byref(b)), "device_results")
        return a.value, b.value

    def put_rows_to_peers(self, src_ptr: int, n_rows: int, cols: int, peer_ptrs, dst_row0: int):
        """this rank's rows (device pointer) into the gathered table of every rank (device pointers of the peers' tables,
        mapped by the caller): stores over NVLink on the context's stream"""
        arr = (C.c_void_p * len(peer_ptrs))(*[int(p) for p in peer_ptrs])
        self._check(self._lib.pyxrd_put_rows_to_peers(self._h, C.c_void_p(int(src_ptr)), int(n_rows), int(cols), arr,
                                                      len(peer_ptrs), int(dst_row0)), "put_rows_to_peers")

    def stream_is_torch_current(self) -> bool:
        import torch
        return getattr(self, "_stream_handle", 0) != 0 and \
            torch.cuda.current_stream(self.device).cuda_stream == self._stream_handle

    def residuals_tensor(self, optimised: bool = False):
        """zero-copy torch view f64[K, 1+S] of the device residual buffer (``optimised``: of the optimiser's residuals),
        for the NCCL gather; valid until the next call that replaces the batch"""
        import torch
        b = self._batch
        res = self.device_results()[0] if optimised else self.device_buffers()[2]
        # the library's kernels run on the context's stream; unless that IS torch's current stream, whatever torch does
        # with the view (a collective, a copy) must not start before they finish
        if getattr(self, "_stream_handle", 0) == 0 or \
                torch.cuda.current_stream(self.device).cuda_stream != self._stream_handle:
            self.synchronize()

        class _View(object):
            __cuda_array_interface__ = dict(shape=(b.K, b.S + 1), typestr="<f8", data=(int(res), False), version=3)

        return torch.as_tensor(_View(), device="cuda:%d" % self.device)

    def dfma_peak(self, iterations: int = 4096):
        t, ms = C.c_double(), C.c_double()
        self._check(self._lib.pyxrd_dfma_peak(self._h, int(iterations), C.byref(t), C.byref(ms)), "dfma_peak")
        return t.value, ms.value


class _Shape(object):
    """K / M / S / Z of a batch that was expanded natively (nothing else of it exists on the Python side)"""

    def __init__(self, K, M, S, Z):
        self.K, self.M, self.S, self.Z = K, M, S, Z


def expand_population(template: BatchPack, x, bindings, csds_auto=None, csds_factor=2.5):
    """Host-only (no device): the arrays of the K-candidate batch ``pyxrd_expand_population`` builds from the
    template and the binding table, as a dict of numpy copies (tests, debugging)."""
    lib = load_library()
    d = Context.batch_desc(template)
    x = np.ascontiguousarray(x, dtype=np.float64)
    bindings = np.ascontiguousarray(bindings, dtype=BINDING_DTYPE)
    auto = None if csds_auto is None else np.ascontiguousarray(csds_auto, dtype=np.uint8)
    pop = PopulationDesc(x.shape[0], x.shape[1], bindings.size, template.S, template.Z, 0, _p(x, _f64p),
                         bindings.ctypes.data_as(C.POINTER(Binding)), _p(auto, _u8p) if auto is not None else None,
                         float(csds_factor))
    handle = C.c_void_p()
    if lib.pyxrd_expand_population(C.byref(d), C.byref(pop), C.byref(handle)) != 0:
        raise ValueError(lib.pyxrd_expand_error().decode("utf-8", "replace"))
    try:
        e = BatchDesc()
        lib.pyxrd_expanded_desc(handle, C.byref(e))
        K, S, Z, M = e.n_candidates, template.S, template.Z, e.n_phase_slots

        def arr(ptr, n, dtype):
            if not ptr or n == 0:
                return np.zeros(0, dtype=dtype)
            return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)

        out = dict(
            K=K, M=M,
            phase_i=arr(e.phase_i, e.n_phases * 8, np.int32).reshape(-1, 8),
            phase_d=arr(e.phase_d, e.n_phases * 8, float).reshape(-1, 8),
            phase_comp=arr(e.phase_comp, e.n_phase_comp, np.int32),
            matrices=arr(e.matrices, e.n_matrix, float),
            comp_d=arr(e.comp_d, e.n_components * 8, float).reshape(-1, 8),
            comp_i=arr(e.comp_i, e.n_components * 4, np.int32).reshape(-1, 4),
            atom_d=arr(e.atom_d, e.n_atoms * 2, float).reshape(-1, 2),
            atom_type=arr(e.atom_type, e.n_atoms, np.int32),
            job_phase=arr(e.job_phase, K * S * Z * M, np.int32),
            fractions=arr(e.fractions, K * M, float).reshape(K, M),
            scales=arr(e.scales, K * S, float).reshape(K, S),
            bgshifts=arr(e.bgshifts, K * S, float).reshape(K, S),
            prob_model=arr(e.prob_model, e.n_phases, np.int32) if e.prob_model else None,
            prob_params=arr(e.prob_params, e.n_phases * 12, float).reshape(-1, 12) if e.prob_params else None)
        return out
    finally:
        lib.pyxrd_expanded_free(handle)


_contexts = {}
_ctx_lock = threading.Lock()


def _after_fork_in_child():
    """CUDA state does not survive fork(): a child of a process that had initialised CUDA (through this library or
    through torch) can neither use the parent's contexts nor create its own.  The inherited handles are dropped WITHOUT
    calling into the CUDA runtime and the child is marked, so that the first attempt to compute raises a DeviceError
    that says what to do instead (the reference evaluates trials in forked Pool workers, pyxrd_server.py:57-64)."""
    global _forked_after_cuda, _ctx_lock
    _ctx_lock = threading.Lock()
    had_cuda = bool(_contexts)
    torch = sys.modules.get("torch")
    if torch is not None:
        try:
            had_cuda = had_cuda or bool(torch.cuda.is_initialized())
        except Exception:
            pass
    for ctx in _contexts.values():
        ctx._h = C.c_void_p()                     # never pyxrd_destroy() a handle of the parent in the child
        ctx.lock = threading.RLock()
    _contexts.clear()
    if had_cuda:
        _forked_after_cuda = True


if hasattr(os, "register_at_fork"):
    os.register_at_fork(after_in_child=_after_fork_in_child)


def default_context(device: Optional[int] = None) -> Context:
    """Process-wide context, created lazily (after any fork: SURVEY.md §8(b) threading model)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("PYXRD_B200_DEVICE") is None \
            else int(os.environ["PYXRD_B200_DEVICE"])
    key = (os.getpid(), device)
    with _ctx_lock:
        ctx = _contexts.get(key)
        if ctx is None:
            ctx = Context(device)
            _contexts[key] = ctx
        return ctx
