"""Device versions of ``pyxrd/calculations/components.py``."""
import numpy as np

from ..runtime import default_context
from .atoms import _type_row


def calculate_z(default_z, lattice_d, z_factor):
    """components.py:15-16 (scalar bookkeeping; the kernels evaluate the same expression)."""
    return lattice_d + z_factor * (default_z - lattice_d)


def get_factors(range_stl, component):
    """components.py:18-54 -> (SF, phi) complex arrays; writes the stretched ``atom.z`` back."""
    stl = np.asarray(range_stl, dtype=float)
    layer = list(component.layer_atoms)
    inter = list(component.interlayer_atoms)
    atoms = layer + inter
    atom_d = np.zeros((max(len(atoms), 1), 2))
    rows = np.zeros((max(len(atoms), 1), 12))
    has = np.zeros(max(len(atoms), 1), dtype=np.uint8)
    for i, atom in enumerate(atoms):
        if atom is None:
            continue
        atom_d[i] = [atom.pn if atom.pn is not None else 0.0, atom.default_z]
        if atom.atom_type is not None:
            rows[i] = _type_row(atom.atom_type)
            has[i] = 1
    comp_row = np.zeros(8)
    comp_row[:4] = [component.d001, component.default_c, component.delta_c, component.lattice_d]
    ctx = default_context()
    with ctx.lock:
        sf, phi, z = ctx.leaf_component(stl, comp_row, len(layer), len(inter), atom_d[:len(atoms)], rows[:len(atoms)],
                                        has[:len(atoms)])
    for atom, zi in zip(atoms, z):
        if atom is not None:
            atom.z = float(zi)
    return sf.reshape(stl.shape), phi.reshape(stl.shape)
