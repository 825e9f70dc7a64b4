"""CPU: oracle restatements of connectivity / Fischer / AFIR vs golden vectors from the reference."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O
from multioptpy_b200.Parameters.tables import covalent_radius

RTOL = 1e-10


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def load(golden_dir):
    z = np.load(os.path.join(golden_dir, "producers.npz"))
    return z, [str(s) for s in z["names"]]


def unpad(tab, cnt):
    return [list(map(int, r)) for r in tab[:cnt]]


@pytest.mark.parametrize("idx", range(8))
def test_connectivity_tables_bit_exact(golden_dir, idx):
    z, names = load(golden_dir)
    name = names[idx]
    elems = [str(e) for e in z[f"{name}/elements"]]
    radii = np.array([covalent_radius(e) for e in elems])
    b, a, d = O.connectivity_tables(z[f"{name}/xyz"], radii)
    c = z[f"{name}/counts"]
    assert b == unpad(z[f"{name}/bonds"], c[0])
    assert a == unpad(z[f"{name}/angles"], c[1])
    assert d == unpad(z[f"{name}/dihedrals"], c[2])


def test_survey_fixture_counts(golden_dir):
    """SURVEY §8c known answers: aldol_rxn 9/10/6, s8 7/6/5 and the aldol bond list."""
    z, _ = load(golden_dir)
    assert list(z["aldol_rxn/counts"]) == [9, 10, 6]
    assert list(z["s8/counts"]) == [7, 6, 5]
    assert unpad(z["aldol_rxn/bonds"], 9) == [[0, 1], [0, 2], [0, 3], [4, 5], [4, 6], [4, 9], [5, 7], [5, 8], [7, 10]]


@pytest.mark.parametrize("idx", [0, 1, 3, 6])
def test_fischer_matches_reference(golden_dir, idx):
    z, names = load(golden_dir)
    name = names[idx]
    elems = [str(e) for e in z[f"{name}/elements"]]
    radii = np.array([covalent_radius(e) for e in elems])
    H = O.fischer_hessian(z[f"{name}/xyz"], radii)
    assert rel(H, z[f"{name}/fischer"]) < RTOL
    if name == "aldol_rxn":   # SURVEY §8c known answer
        assert abs(np.linalg.norm(H) - 3.218500608080904) < 1e-9
        assert abs(np.trace(H) - 1.010961596554204e+01) < 1e-9


@pytest.mark.parametrize("idx", [0, 1, 6])
def test_afir_matches_reference(golden_dir, idx):
    z, names = load(golden_dir)
    name = names[idx]
    elems = [str(e) for e in z[f"{name}/elements"]]
    radii = [covalent_radius(e) for e in elems]
    for c in range(3):
        f1 = [int(v) - 1 for v in z[f"{name}/afir_f1"][c] if v > 0]
        f2 = [int(v) - 1 for v in z[f"{name}/afir_f2"][c] if v > 0]
        E, g, H = O.afir_egh(z[f"{name}/xyz"], f1, f2, radii, float(z[f"{name}/afir_gamma"][c]))
        assert abs(E - z[f"{name}/afir_E"][c]) <= RTOL * abs(z[f"{name}/afir_E"][c])
        assert rel(g, z[f"{name}/afir_g"][c]) < RTOL
        assert rel(H, z[f"{name}/afir_H"][c]) < RTOL
