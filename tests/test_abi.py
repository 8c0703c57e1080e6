"""CPU checks of the boundary: the shared library loads and exports every symbol the
header declares; the host-side packing produces the documented layouts."""
import os
import re

import numpy as np
import pytest

from helpers import ROOT
from pyxrd_b200 import packing, runtime
from pyxrd_b200 import synthetic as syn


def _ensure_built():
    if not os.path.exists(runtime.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()


def test_library_exports_every_declared_symbol():
    _ensure_built()
    header = open(os.path.join(ROOT, "include", "pyxrd_b200.h")).read()
    declared = set(re.findall(r"\b(pyxrd_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(runtime.EXPORTS)
    lib = runtime.load_library()
    for name in sorted(declared):
        assert hasattr(lib, name), name


def test_struct_layout_matches_header():
    header = open(os.path.join(ROOT, "include", "pyxrd_b200.h")).read()
    for struct, cls in (("pyxrd_static_desc", runtime.StaticDesc), ("pyxrd_batch_desc", runtime.BatchDesc)):
        end = header.index("} %s;" % struct)
        body = header[header.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = re.findall(r"\*?\s*\*?([a-z_0-9]+)\s*;", body)
        assert names == [f[0] for f in cls._fields_], struct


def test_no_cuda_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _ensure_built()
    with pytest.raises(runtime.DeviceError):
        runtime.Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "pyxrd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py") or f.endswith(".cu") or f.endswith(".cuh"):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyxrd_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f


def test_packing_layout_c4():
    mixes = syn.CONFIGS["c4"].population(7, 3)
    types = packing.collect_atom_types([s for m in mixes for s in m.specimens])
    static = packing.StaticPack(mixes[0].specimens, types)
    batch = packing.BatchPack(mixes, static)
    assert static.S == 2 and static.Z == 1 and static.sumX == 4200
    assert batch.K == 3 and batch.M == 6
    assert batch.job_phase.shape == (3 * 2 * 1 * 6,)
    assert batch.phase_i.shape == (36, packing.PHASE_I_STRIDE)
    # rank / G of the six phases of one specimen
    assert [(int(r[0]), int(r[1])) for r in batch.phase_i[:6]] == [(1, 1), (1, 1), (1, 1), (2, 2), (2, 2), (2, 4)]
    assert static.wl_index.shape == (2 * 2 * 2100,)
    # the excluded window of specimen 2 is masked
    assert static.selected[:2100].all() and not static.selected[2100:].all()
    # matrices pool: W then P per phase
    off = batch.phase_i[3, 6]
    W = batch.matrices[off:off + 4].reshape(2, 2)
    P = batch.matrices[off + 4:off + 8].reshape(2, 2)
    assert np.allclose(P.sum(axis=1), 1.0) and abs(np.trace(W) - 1.0) < 1e-12


def test_wavelength_tables_match_oracle_interpolation():
    from oracle import pyxrd_oracle as orc
    theta = syn.theta_grid(3.0, 45.0, 0.02)
    rng = np.random.default_rng(5)
    y = rng.uniform(0.5, 2.0, theta.size)
    for wl0, dist in ((syn.CU_KA1, [(syn.CU_KA1, 1.0)]), (syn.CU_PAIR[0][0], syn.CU_PAIR)):
        frac, idx, w_hi, w_lo = packing.wavelength_tables(theta, wl0, dist)
        stl = 2 * np.sin(theta) / wl0
        for j, (wl, f) in enumerate(dist):
            want = orc._interp_linear_fill0(np.arcsin(stl * wl * 0.5), y * f, theta)
            hi = idx[j]
            got = np.where(hi >= 0, w_hi[j] * (y * f)[np.maximum(hi, 1)] + w_lo[j] * (y * f)[np.maximum(hi, 1) - 1], 0.0)
            assert np.array_equal(got, want)


def test_candidates_of_a_batch_must_share_their_specimens():
    mixes = syn.CONFIGS["c4"].population(7, 3)
    types = packing.collect_atom_types([s for m in mixes for s in m.specimens])
    static = packing.StaticPack(mixes[0].specimens, types)
    packing.BatchPack(mixes, static)
    mixes[2].specimens[1].goniometer.wavelength_distribution = [(syn.CU_KA1, 1.0)]
    with pytest.raises(ValueError, match="shared by the candidates"):
        packing.BatchPack(mixes, static)
    mixes = syn.CONFIGS["c4"].population(7, 2)
    mixes[1].specimens[0].range_theta = mixes[1].specimens[0].range_theta[:-5]
    with pytest.raises(ValueError, match="shared by the candidates"):
        packing.BatchPack(mixes, static)
