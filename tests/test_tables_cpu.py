"""Generated tables of the phase kernels: the committed sincos_table.inc holds the correctly rounded values and the
constants the device code assumes, and the table-assisted evaluation (emulated with numpy, as
tools/make_sincos_table.py does) stays within 2 ulp of 1 of libm on the argument range of the kernels."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "pyxrd_b200", "csrc", "sincos_table.inc")


def _arrays():
    text = open(INC).read()
    consts = re.search(r"PX_ST\[10\] = \{(.*?)\};", text, re.S).group(1)
    table = re.search(r"PX_SCT\[256\] = \{(.*?)\};", text, re.S).group(1)
    num = lambda block: np.array([float(v) for v in re.findall(r"-?\d[\d.e+-]*", block)])     # noqa: E731
    return num(consts), num(table).reshape(128, 2)


def test_sincos_table_values():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60
    st, tab = _arrays()
    assert st.size == 10 and st[1] == 6755399441055744.0
    assert st[0] == float(64 / mp.pi) and st[2] == float(-mp.pi / 64)
    assert st[3] == float(-mp.pi / 64 - mp.mpf(st[2]))
    for k in range(128):
        assert tab[k, 0] == float(mp.sin(k * mp.pi / 64)) and tab[k, 1] == float(mp.cos(k * mp.pi / 64))


def test_table_assisted_sincos_accuracy():
    st, tab = _arrays()
    rng = np.random.default_rng(5)
    a = np.concatenate([rng.uniform(-200.0, 200.0, 200000), rng.uniform(-1e-3, 1e-3, 1000), [0.0, np.pi / 128, -np.pi / 128]])
    t = a * st[0] + st[1]
    k = t - st[1]
    ld = np.longdouble                       # (the device uses fused multiply-adds: exact products)
    r = np.float64(ld(a) + ld(k) * ld(st[2]))
    r = np.float64(ld(r) + ld(k) * ld(st[3]))
    assert np.abs(r).max() <= np.pi / 128 * (1 + 1e-9)
    r2 = r * r
    ps = (st[4] * r2 + st[5]) * r2 + st[6]
    pc = (st[7] * r2 + st[8]) * r2 + st[9]
    sr, cm1 = (r * r2) * ps + r, r2 * pc
    idx = k.astype(np.int64) & 127
    S, C = tab[idx, 0], tab[idx, 1]
    s, c = C * sr + (S * cm1 + S), -S * sr + (C * cm1 + C)
    want_s, want_c = np.float64(np.sin(ld(a))), np.float64(np.cos(ld(a)))
    assert np.abs(s - want_s).max() < 4.5e-16 and np.abs(c - want_c).max() < 4.5e-16
