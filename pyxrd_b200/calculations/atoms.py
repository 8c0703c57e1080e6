"""Device versions of ``pyxrd/calculations/atoms.py``."""
import logging

import numpy as np

from ..runtime import default_context

logger = logging.getLogger(__name__)


def _type_row(atom_type):
    row = np.empty(12)
    row[0:5] = np.asarray(atom_type.par_a, dtype=float)
    row[5:10] = np.asarray(atom_type.par_b, dtype=float)
    row[10] = float(atom_type.par_c)
    row[11] = float(atom_type.debye)
    return row


def get_atomic_scattering_factor(angstrom_range, atom_type):
    """atoms.py:13-38: ``angstrom_range`` = (stl*0.05)**2 values; None type -> zeros + warning."""
    s = np.asarray(angstrom_range, dtype=float)
    if atom_type is None:
        logger.warning("get_atomic_scattering_factor reports: 'None found!'")
        return np.zeros_like(s)
    ctx = default_context()
    with ctx.lock:
        return ctx.leaf_asf(s, _type_row(atom_type)).reshape(s.shape)


def get_structure_factor(range_stl, atom):
    """atoms.py:40-67: asf * pn * exp(2 pi i z stl) using the atom's current ``z``."""
    stl = np.asarray(range_stl, dtype=float)
    if atom is None or atom.atom_type is None:
        logger.warning("get_structure_factor reports: 'None found!'")
        return np.zeros_like(stl)
    ctx = default_context()
    comp_row = np.zeros(8)           # a bare layer atom: no stretch, no phase factor needed
    comp_row[1] = 1.0
    with ctx.lock:
        sf, _, _ = ctx.leaf_component(stl, comp_row, 1, 0, [[atom.pn, atom.z]], [_type_row(atom.atom_type)], [1])
    return sf.reshape(stl.shape)
