"""Seeded synthetic inputs for the optimizer-step hot path.

Shared by the tests, ``bench.py`` and the golden-vector generator so that the
CUDA path, the NumPy oracle and the reference see bit-identical inputs.
Pure NumPy, no reference or oracle imports.
"""
from __future__ import annotations

import numpy as np

ELEMENT_CYCLE = ("C", "H", "O", "N")


def grid_geometry(natoms: int, rng: np.random.Generator, spacing: float = 2.8,
                  jitter: float = 0.2) -> np.ndarray:
    """N atoms on a ceil(N^(1/3))^3 cubic grid (Bohr) plus Gaussian jitter."""
    m = int(np.ceil(natoms ** (1.0 / 3.0) - 1e-12))
    pts = np.array([(i, j, k) for i in range(m) for j in range(m) for k in range(m)],
                   dtype=np.float64)[:natoms]
    return pts * spacing + rng.normal(0.0, jitter, size=(natoms, 3))


def elements(natoms: int, all_sulfur: bool = False) -> list[str]:
    if all_sulfur:
        return ["S"] * natoms
    return [ELEMENT_CYCLE[i % 4] for i in range(natoms)]


def spd_hessian(n: int, rng: np.random.Generator, neg_lowest: bool = False) -> np.ndarray:
    """A A^T / n + 0.1 I; optionally with its lowest eigenvalue set to -0.05."""
    A = rng.standard_normal((n, n))
    H = A @ A.T / n + 0.1 * np.eye(n)
    if neg_lowest:
        w, V = np.linalg.eigh(H)
        w[0] = -0.05
        H = (V * w) @ V.T
        H = 0.5 * (H + H.T)
    return H


def structure(config: int, b: int, natoms: int, saddle: bool = False):
    """One synthetic structure: (x0 (n,), H (n,n), g0 (n,), rng) with
    seed = 1000 * config + b."""
    rng = np.random.default_rng(1000 * config + b)
    x0 = grid_geometry(natoms, rng).reshape(-1)
    n = 3 * natoms
    H = spd_hessian(n, rng, neg_lowest=saddle)
    g0 = rng.normal(0.0, 1e-2, size=n)
    return x0, H, g0, rng


def second_point(x0, H, g0, move0, rng):
    """x1 = x0 - move0 (caller convention, optimizer.py:798),
    g1 = g0 + H (x1 - x0) + N(0, 1e-4^2); E1 - E0 = -1e-3."""
    x1 = x0 - move0
    g1 = g0 + H @ (x1 - x0) + rng.normal(0.0, 1e-4, size=x0.size)
    return x1, g1


def batch(config: int, B: int, natoms: int, saddle: bool = False):
    """Stacked batch: x0 (B,n), H (B,n,n), g0 (B,n), list of rngs."""
    xs, Hs, gs, rngs = [], [], [], []
    for b in range(B):
        x0, H, g0, rng = structure(config, b, natoms, saddle)
        xs.append(x0); Hs.append(H); gs.append(g0); rngs.append(rng)
    return np.stack(xs), np.stack(Hs), np.stack(gs), rngs
