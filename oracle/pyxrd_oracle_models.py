"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the probability models of SURVEY.md §8(f) rank 2:
numpy restatement of every probability model ``pyxrd/probabilities/models`` (v0.8.4) ships
(probabilities/models/__init__.py:9-12) -- R0 G1..G6, R1 G2 / G3 / G4, R2 G2 / G3, R3 G2.  Only ``tests/`` may import this module.

Pinned against the live reference model classes (``tests/golden/make_golden.py models`` ->
``tests/golden/models.npz``; the reference runs headless behind ``tests/golden/_live_models.py``).

What the calculation layer consumes (phases/models/phase.py:43-52): ``W`` = the top-level diagonal
distribution matrix, ``P`` = the top-level probability matrix, ``valid_probs`` = all(P_valid) and
all(W_valid).  State convention: the reference keeps its level matrices between ``update()`` calls
and ``solve()`` (base_models.py:142-162) fills level ``num`` from level ``num + 1`` *before* that
level has been refreshed, so for R = 3 the level-1 matrices lag one ``update()`` behind; the oracle
(and the device kernel) return the converged state = what a second ``update()`` with the same
parameters leaves.  W, P themselves never depend on that history.
"""
from itertools import product

import numpy as np

R0, R1G2, R2G2, R3G2, R1G3, R1G4, R2G3 = 1, 2, 3, 4, 5, 6, 7          # = PYXRD_PROB_* of include/pyxrd_b200.h
PROB_CODES = {"R0": R0, "R1G2": R1G2, "R2G2": R2G2, "R3G2": R3G2, "R1G3": R1G3, "R1G4": R1G4, "R2G3": R2G3}
N_PARAMS = {R1G2: 2, R2G2: 4, R3G2: 2, R1G3: 6, R1G4: 12, R2G3: 6}
SHAPE = {R1G2: (1, 2), R2G2: (2, 2), R3G2: (3, 2), R1G3: (1, 3), R1G4: (1, 4), R2G3: (2, 3)}


class Levels(object):
    """level matrices of base_models.py:100-128 with the multi-index access of :243-285"""

    def __init__(self, R, G):
        self.R, self.G = R, G
        Rp = max(R, 1)
        self.lW = [np.zeros((G ** (r + 1),) * 2) for r in range(Rp)] + [np.zeros((G ** Rp,) * 2)]
        self.lP = [np.zeros((G ** (r + 1),) * 2) for r in range(Rp)]

    def _xy(self, ind, Rl):
        x = y = 0
        for i in range(1, Rl + 1):
            f = self.G ** (Rl - i)
            x += ind[i - 1] * f
            y += ind[i] * f
        return x, y

    def wpos(self, ind):
        ind = tuple(np.atleast_1d(ind))
        l = len(ind)
        if l == max(self.R, 1) + 1:
            return l - 1, self._xy(ind, max(l - 1, 1))
        x = 0
        for i in range(l):
            x += ind[i] * self.G ** (l - (i + 1))
        return l - 1, (x, x)

    def ppos(self, ind):
        ind = tuple(ind)
        l = len(ind)
        return l - 2, self._xy(ind, max(l - 1, 1))

    def W(self, *ind):
        r, xy = self.wpos(ind)
        return self.lW[r][xy]

    def setW(self, ind, value):
        r, xy = self.wpos(ind)
        self.lW[r][xy] = value

    def P(self, *ind):
        r, xy = self.ppos(ind)
        return self.lP[r][xy]

    def setP(self, ind, value):
        r, xy = self.ppos(ind)
        self.lP[r][xy] = value

    def solve(self):
        """base_models.py:142-162"""
        G = self.G
        for num in range(1, self.R):
            for base in product(range(G), repeat=num):
                acc = 0
                for i in range(G):
                    acc += self.W(*((i,) + base))
                self.setW(base, acc)
            for base in product(range(G), repeat=num + 1):
                w = self.W(*base[:-1])
                with np.errstate(all="ignore"):
                    self.setP(base, self.W(*base) / w if w > 0 else 0.0)
        self.lW[-1][:] = np.dot(self.lW[-2], self.lP[-1])

    def valid(self):
        """base_models.py:164-241: the sum rules compare against 1e4 / 1e6 and never fire; what is
        left is every entry of every level matrix (and of W.P) inside [0, 1] (NaN passes)."""
        ok = True
        for M in self.lW + self.lP:
            ok = ok and not bool(np.any((M < 0.0) | (M > 1.0)))
        return ok


# lower bound of the first parameter (W1): R2models.py:115 (0.5), R3models.py:119 (2/3); everything else [0, 1]
W1_MINIMUM = {R0: 0.0, R1G2: 0.0, R2G2: 0.5, R3G2: 2.0 / 3.0, R1G3: 0.0, R1G4: 0.0, R2G3: 0.5}   # R2models.py:370


def _clamp(v, lo=0.0):
    """property setters clamp to [minimum, maximum] (mvc/models/properties/cast_property.py:44-47)"""
    return min(max(float(v), lo), 1.0)


def _update(model, G, params, lv):
    p = [_clamp(v, W1_MINIMUM[model] if i == 0 else 0.0) for i, v in enumerate(params)]
    with np.errstate(all="ignore"):
        if model == R0:                                   # R0models.py:68-83
            if G > 1:
                for i in range(G - 1):
                    if i > 0:
                        lv.setW((i,), p[i] * (1.0 - np.sum(np.diag(lv.lW[0])[0:i])))
                    else:
                        lv.setW((i,), p[i])
                lv.setW((G - 1,), 1.0 - np.sum(np.diag(lv.lW[0])[:-1]))
            else:
                lv.setW((0,), 1.0)
            lv.lP[0][:] = np.repeat(np.diag(lv.lW[0])[np.newaxis, :], G, 0)
        elif model == R1G2:                               # R1models.py:125-139
            W1, P11_or_P22 = p
            lv.setW((0,), W1)
            lv.setW((1,), 1.0 - lv.W(0))
            if lv.W(0) <= 0.5:
                lv.setP((0, 0), P11_or_P22)
                lv.setP((0, 1), 1.0 - lv.P(0, 0))
                lv.setP((1, 0), lv.W(0) * lv.P(0, 1) / lv.W(1))
                lv.setP((1, 1), 1.0 - lv.P(1, 0))
            else:
                lv.setP((1, 1), P11_or_P22)
                lv.setP((1, 0), 1.0 - lv.P(1, 1))
                lv.setP((0, 1), lv.W(1) * lv.P(1, 0) / lv.W(0))
                lv.setP((0, 0), 1.0 - lv.P(0, 1))
        elif model == R2G2:                               # R2models.py:195-233
            W1, P112_or_P211, P21, P122_or_P221 = p
            lv.setW((0,), W1)
            lv.setW((1,), 1.0 - lv.W(0))
            lv.setP((1, 0), P21)
            lv.setP((1, 1), 1.0 - lv.P(1, 0))
            lv.setW((1, 0), lv.W(1) * lv.P(1, 0))
            lv.setW((1, 1), lv.W(1) * lv.P(1, 1))
            lv.setW((0, 1), lv.W(1, 0))
            lv.setW((0, 0), lv.W(0) - lv.W(1, 0))
            if lv.W(0) <= 2.0 / 3.0:
                lv.setP((0, 0, 1), P112_or_P211)
                lv.setP((1, 0, 0), 0.0 if lv.W(1, 0) == 0.0 else lv.P(0, 0, 1) * lv.W(0, 0) / lv.W(1, 0))
            else:
                lv.setP((1, 0, 0), P112_or_P211)
                lv.setP((0, 0, 1), 0.0 if lv.W(0, 0) == 0.0 else lv.P(1, 0, 0) * lv.W(1, 0) / lv.W(0, 0))
            lv.setP((1, 0, 1), 1.0 - lv.P(1, 0, 0))
            lv.setP((0, 0, 0), 1.0 - lv.P(0, 0, 1))
            if lv.P(1, 0) <= 0.5:
                lv.setP((0, 1, 1), P122_or_P221)
                lv.setP((1, 1, 0), lv.P(0, 1, 1) * lv.W(0, 1) / lv.W(1, 1))
            else:
                lv.setP((1, 1, 0), P122_or_P221)
                lv.setP((0, 1, 1), lv.P(1, 1, 0) * lv.W(1, 1) / lv.W(0, 1))
            lv.setP((0, 1, 0), 1.0 - lv.P(0, 1, 1))
            lv.setP((1, 1, 1), 1.0 - lv.P(1, 1, 0))
        elif model == R3G2:                               # R3models.py:158-198
            W1, P1111_or_P2112 = p
            c01 = lambda v: max(min(v, 1.0), 0.0)          # noqa: E731
            lv.setW((0,), W1)
            lv.setW((1,), 1.0 - W1)
            if lv.W(0) <= 0.75:
                lv.setP((0, 0, 0, 0), P1111_or_P2112)
                lv.setP((0, 0, 0, 1), c01(1.0 - lv.P(0, 0, 0, 0)))
                lv.setP((1, 0, 0, 0), c01(lv.P(0, 0, 0, 1) * (lv.W(0) - 2 * lv.W(1)) / lv.W(1)))
                lv.setP((1, 0, 0, 1), c01(1.0 - lv.P(1, 0, 0, 0)))
            else:
                lv.setP((1, 0, 0, 1), P1111_or_P2112)
                lv.setP((1, 0, 0, 0), c01(1.0 - lv.P(1, 0, 0, 1)))
                lv.setP((0, 0, 0, 0), c01(1.0 - lv.P(1, 0, 0, 0) * lv.W(1) / (lv.W(0) - 2 * lv.W(1))))
                lv.setP((0, 0, 0, 1), c01(1.0 - lv.P(0, 0, 0, 0)))
            for a, b, c in ((0, 0, 1), (0, 1, 0), (0, 1, 1), (1, 0, 1), (1, 1, 0), (1, 1, 1)):
                lv.setP((a, b, c, 0), 1.0)
                lv.setP((a, b, c, 1), 0.0)
            lv.setW((0, 0, 0), c01(3 * lv.W(0) - 2))
            for t in ((1, 0, 1), (1, 1, 0), (1, 1, 1), (0, 1, 1)):
                lv.setW(t, 0.0)
            for t in ((1, 0, 0), (0, 1, 0), (0, 0, 1)):
                lv.setW(t, c01(1 - lv.W(0)))
        elif model == R1G3:                               # R1models.py:367-410
            W1, P11_or_P22, G1, G2, G3, G4 = p
            lv.setW((0,), W1)
            lv.setW((1,), (1 - lv.W(0)) * G1)
            lv.setW((2,), 1.0 - lv.W(0) - lv.W(1))
            W0inv = 1.0 / lv.W(0) if lv.W(0) > 0.0 else 0.0
            if lv.W(0) <= 0.5:
                lv.setP((0, 0), P11_or_P22)
                Wxx = lv.W(0) * (lv.P(0, 0) - 1) + lv.W(1) + lv.W(2)
            else:
                Wxx = (1.0 - lv.W(0)) * P11_or_P22
            lv.setW((1, 1), Wxx * G2 * G3)
            lv.setW((1, 2), Wxx * G2 * (1 - G3))
            lv.setP((1, 1), lv.W(1, 1) / lv.W(1) if lv.W(1) > 0.0 else 0.0)
            lv.setW((2, 1), Wxx * (1 - G2) * G4)
            lv.setW((2, 2), Wxx * (1 - G2) * (1 - G4))
            lv.setP((1, 2), (lv.W(1, 2) / lv.W(1)) if lv.W(1) > 0.0 else 0.0)
            lv.setP((1, 0), 1 - lv.P(1, 1) - lv.P(1, 2))
            lv.setP((2, 1), (lv.W(2, 1) / lv.W(2)) if lv.W(2) > 0.0 else 0.0)
            lv.setP((2, 2), (lv.W(2, 2) / lv.W(2)) if lv.W(2) > 0.0 else 0.0)
            lv.setP((2, 0), 1 - lv.P(2, 1) - lv.P(2, 2))
            lv.setP((0, 1), (lv.W(1) - lv.W(1, 1) - lv.W(2, 1)) * W0inv)
            lv.setP((0, 2), (lv.W(2) - lv.W(1, 2) - lv.W(2, 2)) * W0inv)
            if lv.W(0) > 0.5:
                lv.setP((0, 0), 1 - lv.P(0, 1) - lv.P(0, 2))
            # :405-407 rewrite the two-index weights, which solve() replaces by W . P anyway
        elif model == R1G4:                               # R1models.py:747-797
            W1, P11_or_P22, R1, R2, G1, G2, G11, G12, G21, G22, G31, G32 = p
            lv.setW((0,), W1)
            lv.setW((1,), (1.0 - lv.W(0)) * R1)
            lv.setW((2,), (1.0 - lv.W(0) - lv.W(1)) * R2)
            lv.setW((3,), 1.0 - lv.W(0) - lv.W(1) - lv.W(2))
            W0inv = 1.0 / lv.W(0) if lv.W(0) > 0.0 else 0.0
            if lv.W(0) < 0.5:
                lv.setP((0, 0), P11_or_P22)
                Wxx = lv.W(0) * (lv.P(0, 0) - 1) + lv.W(1) + lv.W(2) + lv.W(3)
            else:
                Wxx = P11_or_P22 * (lv.W(1) + lv.W(2) + lv.W(3))
            W1x = Wxx * G1
            Wyx = (Wxx - W1x)
            W2x = Wyx * G2
            W3x = Wyx - W2x
            for i, (Ga, Gb, Wix) in enumerate(((G11, G12, W1x), (G21, G22, W2x), (G31, G32, W3x)), start=1):
                lv.setW((i, 1), Ga * Wix)
                lv.setW((i, 2), Gb * (Wix - lv.W(i, 1)))
                lv.setW((i, 3), Wix - lv.W(i, 1) - lv.W(i, 2))
            for i in range(1, 4):
                lv.setP((i, 0), 1)
                for j in range(1, 4):
                    lv.setP((i, j), lv.W(i, j) / lv.W(i) if lv.W(i) > 0 else 0)
                    lv.setP((i, 0), lv.P(i, 0) - lv.P(i, j))
                lv.setW((i, 0), lv.W(i) * lv.P(i, 0))
            lv.setP((0, 1), (lv.W(1) - lv.W(1, 1) - lv.W(2, 1) - lv.W(3, 1)) * W0inv)
            lv.setP((0, 2), (lv.W(2) - lv.W(1, 2) - lv.W(2, 2) - lv.W(3, 2)) * W0inv)
            lv.setP((0, 3), (lv.W(3) - lv.W(1, 3) - lv.W(2, 3) - lv.W(3, 3)) * W0inv)
            if lv.W(0) >= 0.5:
                lv.setP((0, 0), 1 - lv.P(0, 1) - lv.P(0, 2) - lv.P(0, 3))
        elif model == R2G3:                               # R2models.py:485-535
            W1, P111_or_P212, G1, G2, G3, G4 = p
            lv.setW((0,), W1)
            lv.setW((1,), (1.0 - lv.W(0)) * G1)
            lv.setW((2,), 1.0 - lv.W(0) - lv.W(1))
            for t in ((1, 1), (1, 2), (2, 1), (2, 2)):
                lv.setW(t, 0)
            for t in ((1, 0), (0, 1), (0, 1, 0)):
                lv.setW(t, lv.W(1))
            for t in ((2, 0), (0, 2), (0, 2, 0)):
                lv.setW(t, lv.W(2))
            lv.setW((0, 0), lv.W(0) - lv.W(0, 1) - lv.W(0, 2))
            Wx = lv.W(1) + lv.W(2)
            if lv.W(0) < 2.0 / 3.0:
                lv.setP((0, 0, 0), P111_or_P212)
                Px0x = 1 - (lv.W(0) - Wx) / Wx * (1 - lv.P(0, 0, 0)) if Wx != 0 else 0.0
            else:
                Px0x = P111_or_P212
                lv.setP((0, 0, 0), 1 - Wx / (lv.W(0) - Wx) * (1 - Px0x) if (lv.W(0) - Wx) != 0 else 0.0)
            Wx0x = Wx * Px0x
            W10x = G2 * Wx0x
            W20x = Wx0x - W10x
            lv.setW((1, 0, 1), G3 * W10x)
            lv.setW((1, 0, 2), (1 - G3) * W10x)
            lv.setW((1, 0, 0), lv.W(1, 0) - lv.W(1, 0, 1) - lv.W(1, 0, 2))
            lv.setW((2, 0, 1), G4 * W20x)
            lv.setW((2, 0, 2), (1 - G4) * W20x)
            lv.setW((2, 0, 0), lv.W(2, 0) - lv.W(2, 0, 1) - lv.W(2, 0, 2))
            lv.setW((0, 0, 0), lv.W(0, 0) * lv.P(0, 0, 0))
            lv.setW((0, 0, 1), lv.W(0, 1) - lv.W(1, 0, 1) - lv.W(2, 0, 1))
            lv.setW((0, 0, 2), lv.W(0, 2) - lv.W(1, 0, 2) - lv.W(2, 0, 2))
            for i in range(3):
                for j in range(3):
                    for k in range(3):
                        lv.setP((i, j, k), lv.W(i, j, k) / lv.W(i, j) if lv.W(i, j) > 0 else 0.0)
        else:
            raise ValueError("unknown probability model %r" % (model,))
        lv.solve()


def probabilities(model, G, params):
    """-> (W [r, r] diagonal, P [r, r], valid) of the converged model state (module docstring)."""
    R = 0 if model == R0 else SHAPE[model][0]
    if model != R0 and G != SHAPE[model][1]:
        raise ValueError("model %d is a %d-component model" % (model, SHAPE[model][1]))
    lv = Levels(R, G)
    _update(model, G, params, lv)
    _update(model, G, params, lv)          # second update(): level matrices no longer lag (R = 3)
    return lv.lW[-2].copy(), lv.lP[-1].copy(), lv.valid()
