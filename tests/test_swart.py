"""Swart model Hessian: oracle vs goldens generated from the reference
(oracle/gen_golden.py swart), CUDA kernel vs both."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O

RTOL = 1e-10


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "swart.npz"))
    return z, [str(s) for s in z["names"]]


def test_oracle_swart_vs_reference(golden_dir):
    z, names = _load(golden_dir)
    for name in names:
        if name == "grid50":
            continue   # python loops: keep the CPU suite short
        x, r = z[f"{name}/xyz"], z[f"{name}/radii"]
        assert rel(O.swart_raw_hessian(x, r), z[f"{name}/H_raw"]) < 1e-13, name
        assert rel(O.swart_hessian(x, r), z[f"{name}/H"]) < 1e-12, name


def test_swart_radii_table(golden_dir):
    from multioptpy_b200.ModelHessian.swart import swart_radii
    z, names = _load(golden_dir)
    for name in names:
        assert np.array_equal(swart_radii([str(e) for e in z[f"{name}/elements"]]), z[f"{name}/radii"]), name


@pytest.mark.gpu
def test_gpu_swart_vs_golden(golden_dir):
    from multioptpy_b200.ModelHessian.swart import SwartApproxHessian
    z, names = _load(golden_dir)
    for name in names:
        S = SwartApproxHessian(device="cuda:0")
        H = S.main(z[f"{name}/xyz"], [str(e) for e in z[f"{name}/elements"]], None)
        assert rel(S.cart_hess, z[f"{name}/H_raw"]) < RTOL, name
        assert rel(H, z[f"{name}/H"]) < RTOL, name
        assert np.array_equal(H, H.T)


@pytest.mark.gpu
def test_gpu_swart_batched_and_large(golden_dir):
    """Tensor mode over a batch of distinct geometries (gather kernel, natoms <= 100) and the
    scatter kernel with global-memory accumulation (natoms > 100) against the oracle."""
    import torch
    from multioptpy_b200 import synthetic
    from multioptpy_b200.ModelHessian.approx_hessian import ApproxHessian
    from multioptpy_b200.ModelHessian.swart import swart_radii
    for natoms, B in ((24, 5), (60, 2), (104, 1)):
        elems = synthetic.elements(natoms)
        xs = np.stack([synthetic.grid_geometry(natoms, np.random.default_rng(90 + b), spacing=2.6, jitter=0.25) for b in range(B)])
        H = ApproxHessian(device="cuda:0").main(torch.from_numpy(xs).cuda(), elems, None, "swart").cpu().numpy()
        for b in range(B):
            assert rel(H[b], O.swart_hessian(xs[b], swart_radii(elems))) < RTOL, (natoms, b)


@pytest.mark.gpu
def test_gpu_swart_nonfinite_fallback():
    """A NaN coordinate poisons every term: status reports the stretch-only fallback was taken."""
    import torch
    from multioptpy_b200 import ops
    x = np.array([[[0.0, 0, 0], [2.0, 0, 0], [np.nan, 1.0, 0]]])
    H, _, st = ops.swart_hessian(torch.from_numpy(x).cuda(), np.array([1.4, 1.4, 0.59]), want_raw=True)
    assert int(st[0]) == 1
