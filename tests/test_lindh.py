"""Lindh model Hessian, decomposed parity: the diagonal redundant-internal force
constants and project(B^T diag(k) B) vs the reference; the reference's K term is ill-posed and
vanishes for a zero gradient, which is what main() is compared at."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O
from multioptpy_b200.ModelHessian.lindh import lindh_atom_params

RTOL = 1e-10


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


@pytest.mark.parametrize("idx", range(4))
def test_oracle_lindh_vs_reference(golden_dir, idx):
    z = np.load(os.path.join(golden_dir, "lindh.npz"))
    name = str(z["names"][idx])
    elems = [str(e) for e in z[f"{name}/elements"]]
    prm = lindh_atom_params(elems)
    assert rel(O.lindh_kdiag(z[f"{name}/xyz"], prm), z[f"{name}/kdiag"]) < RTOL
    assert rel(O.lindh_hessian_bkb(z[f"{name}/xyz"], prm), z[f"{name}/H_bkb"]) < RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(4))
def test_gpu_lindh_vs_reference(golden_dir, idx):
    from multioptpy_b200.ModelHessian.lindh import LindhApproxHessian
    from multioptpy_b200.ModelHessian.approx_hessian import ApproxHessian
    z = np.load(os.path.join(golden_dir, "lindh.npz"))
    name = str(z["names"][idx])
    elems = [str(e) for e in z[f"{name}/elements"]]
    xyz = z[f"{name}/xyz"]
    L = LindhApproxHessian(device="cuda:0")
    assert rel(L.guess_lindh_diagonal(xyz, elems), z[f"{name}/kdiag"]) < RTOL
    H = ApproxHessian(device="cuda:0").main(xyz, elems, np.zeros_like(xyz), "lindh")
    assert rel(H, z[f"{name}/H_bkb"]) < RTOL
    assert rel(H, z[f"{name}/H_main0"]) < RTOL      # reference main() with a zero gradient
