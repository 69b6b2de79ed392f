"""Synthetic OT problem generators and the bundled-input loader (host side, NumPy only).

The reference ships no generator (``prob_set`` appears only in a comment,
Class1/APD_SsN_Class1.m:25), so the synthetic configurations of BASELINE.json are defined here
exactly as SURVEY.md section 8d states them.
"""
import os

import numpy as np


def grid_points(g):
    """``((a+1/2)/g, (b+1/2)/g)`` for a,b in 0..g-1, index ``i = a*g + b``."""
    a, b = np.meshgrid(np.arange(g), np.arange(g), indexing="ij")
    return np.stack([(a.reshape(-1) + 0.5) / g, (b.reshape(-1) + 0.5) / g], axis=1)


def grid_cost(g, dtype=np.float64):
    """Normalised squared-distance cost ``C_ij = |x_i - y_j|^2 / max C`` (m = n = g*g), as a
    column-major flattened vector ``c = C(:)``."""
    pts = grid_points(g)
    sq = (pts ** 2).sum(1)
    C = sq[:, None] + sq[None, :] - 2.0 * pts @ pts.T
    np.maximum(C, 0.0, out=C)
    C /= C.max()
    return np.asfortranarray(C, dtype=dtype).reshape(-1, order="F")


def grid_marginals(g, seed=0, balanced=True):
    """``l = U(0,1)+0.1``, ``r = U(0,1)+0.1`` from ``RandomState(seed)`` (l first); balanced
    OT rescales ``r *= sum(l)/sum(r)``."""
    m = n = g * g
    rs = np.random.RandomState(seed)
    l = rs.random_sample(m) + 0.1
    r = rs.random_sample(n) + 0.1
    if balanced:
        r = r * (l.sum() / r.sum())
    return r, l


def grid_problem(g, seed=0):
    """Config 2/4/5 of BASELINE.json: balanced OT between two g x g grids."""
    m = n = g * g
    r, l = grid_marginals(g, seed, balanced=True)
    return {"c": grid_cost(g), "r": r, "l": l, "p": np.ones(m), "q": np.ones(n),
            "gama": np.inf, "m": m, "n": n}


def grid_problem_pot(g, seed=0, mass_fraction=0.65):
    """Config 3: partial OT on g x g grids, ``mu = 0.65*min(sum r, sum l)``, ``phi = 1``."""
    m = n = g * g
    r, l = grid_marginals(g, seed, balanced=False)
    return {"c": grid_cost(g), "r": r, "l": l, "p": np.ones(m), "q": np.ones(n),
            "phi": np.ones(m * n), "mu": mass_fraction * min(r.sum(), l.sum()), "m": m, "n": n}


def random_problem(m, n, seed=0):
    """Random-cost OT of the bundled example's kind (c ~ U(0,1), p = q = 1, gama = Inf)."""
    rs = np.random.RandomState(seed)
    c = rs.random_sample(m * n)
    l = rs.random_sample(m) + 0.1
    r = rs.random_sample(n) + 0.1
    r = r * (l.sum() / r.sum())
    return {"c": c, "r": r, "l": l, "p": np.ones(m), "q": np.ones(n), "gama": np.inf,
            "m": m, "n": n}


def load_bundled_class1(path):
    """Reads the reference's Class1/InputData/data1-500.mat (MAT v5 with integer-compressed doubles:
    p, q stored as uint8 and m, n as uint16) through the C-side reader (``matio.py``, libssnmat.so)."""
    from . import matio
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    P = matio.load_problem(path)
    return {k: P[k] for k in ("c", "r", "l", "p", "q", "gama", "m", "n")}


def load_bundled_class2(path):
    """Class2/InputData/data4-500.mat: as above plus ``phi`` and ``mu`` (no ``gama``; the m x n matrix
    ``C`` of the file is ``c`` reshaped and is not returned)."""
    from . import matio
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    P = matio.load_problem(path)
    return {k: P[k] for k in ("c", "r", "l", "p", "q", "phi", "mu", "m", "n")}
