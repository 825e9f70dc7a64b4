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


# ---- config 3 (NEB 64 x 30) and config 4 (conformer / AFIR batch 8192 x N) shapes, SURVEY §8d ----------------------
def neb_chain(nimg: int, natoms: int, seed: int = 3000):
    """A NEB chain: linear interpolation between two jittered end points + N(0, 0.05^2), energies from a 1-D
    double well along the path (uphill, downhill and extremum branches of the tangent rule all occur), seeded
    gradients and per-image Hessians.  -> X (nimg, n), E (nimg,), G (nimg, n), H (nimg, n, n)."""
    rng = np.random.default_rng(seed)
    n = 3 * natoms
    xa = grid_geometry(natoms, rng).reshape(-1)
    xb = xa + rng.normal(0.0, 0.3, n)
    t = np.linspace(0.0, 1.0, nimg)
    X = xa[None, :] + (xb - xa)[None, :] * t[:, None] + rng.normal(0.0, 0.05, (nimg, n))
    E = 0.05 * (16.0 * t ** 2 * (1.0 - t) ** 2 - 0.3 * t) - 0.02 * np.sin(6.0 * np.pi * t)   # two wells, ripples
    G = rng.normal(0.0, 1e-2, (nimg, n))
    H = np.stack([spd_hessian(n, np.random.default_rng(seed + 40 + i)) for i in range(nimg)])
    return X, E, G, H


def neb_next(X, G, H, delta, rng):
    """The chain after one NEB move (x - delta, as the caller applies it) with gradients of the local quadratic
    models plus noise: gives the second iteration its quasi-Newton history."""
    X1 = X - delta
    G1 = G + np.einsum("bij,bj->bi", H, X1 - X) + rng.normal(0.0, 1e-4, X.shape)
    return X1, G1


def s8_crown(rng: np.random.Generator, jitter: float) -> np.ndarray:
    """One S8 crown ring in Bohr (S-S 2.06 A, S-S-S 108 deg: ring radius 2.357 A, puckering +-0.497 A), randomly
    rotated about its axis, plus Gaussian jitter (A)."""
    R, h = 2.357, 0.9946
    phi = rng.uniform(0.0, 2.0 * np.pi) + np.arange(8) * (np.pi / 4.0)
    p = np.stack([R * np.cos(phi), R * np.sin(phi), 0.5 * h * (-1.0) ** np.arange(8)], axis=1)
    return (p + rng.normal(0.0, jitter, size=(8, 3))) / 0.52917721067


def conformer_batch(B: int, natoms: int, seed: int = 4000, jitter: float = 0.05):
    """Config 4: B S8-like conformers - natoms / 8 crown rings stacked 4 A apart along z with a random lateral offset,
    every atom jittered (the model Hessian of the reference then has a handful of mildly negative modes, as for a real
    distorted ring; a cubic grid of sulfur atoms puts the LJ terms of lindh.py:118-130 at -8 Hartree / Bohr^2) - with raw
    gradients.  -> xyz (B, natoms, 3) in Bohr, g (B, 3 natoms)."""
    assert natoms % 8 == 0
    xyz = np.empty((B, natoms, 3)); g = np.empty((B, 3 * natoms))
    for b in range(B):
        rng = np.random.default_rng(seed + b)
        rings = []
        for r in range(natoms // 8):
            ring = s8_crown(rng, jitter)
            ring += np.array([rng.normal(0.0, 0.3), rng.normal(0.0, 0.3), 4.0 * r]) / 0.52917721067
            rings.append(ring)
        xyz[b] = np.concatenate(rings)
        g[b] = rng.normal(0.0, 1e-2, 3 * natoms)
    return xyz, g


# ---- a minimal constraint object for CRSIRFO (the reference's comes from its projection-constraint subsystem) -------
class DistanceConstraints:
    """Interface CRSIRFO calls on ``constraints_obj`` (Optimizer/crsirfo.py:21,68): bond-length constraints between
    atom pairs (0-based).  ``_get_all_constraint_vectors`` returns the UNNORMALISED gradients of the distances scaled
    by ``weights`` (CRSIRFO normalises them itself); ``adjust_init_coord`` is one symmetric SHAKE pass towards
    ``targets`` (or the identity when ``targets`` is None)."""

    def __init__(self, pairs, targets=None, weights=None):
        self.pairs = [tuple(p) for p in pairs]
        self.targets = targets
        self.weights = [1.0] * len(self.pairs) if weights is None else list(weights)

    def _get_all_constraint_vectors(self, geom):
        geom = np.asarray(geom, dtype=np.float64)
        rows = np.zeros((len(self.pairs), geom.size))
        for r, ((i, j), w) in enumerate(zip(self.pairs, self.weights)):
            d = geom[i] - geom[j]
            u = w * d / np.linalg.norm(d)
            rows[r, 3 * i:3 * i + 3] = u
            rows[r, 3 * j:3 * j + 3] = -u
        return rows

    def adjust_init_coord(self, geom):
        out = np.array(geom, dtype=np.float64, copy=True)
        if self.targets is None:
            return out
        for (i, j), t in zip(self.pairs, self.targets):
            d = out[i] - out[j]
            r = np.linalg.norm(d)
            shift = 0.5 * (t - r) * d / r
            out[i] += shift
            out[j] -= shift
        return out
