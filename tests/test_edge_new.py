"""Edge cases of the round-2 entry points (empty batches, smallest sizes, degenerate inputs) through the C ABI."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import np_oracle as O  # noqa: E402


def rel(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


@pytest.mark.gpu
def test_empty_batches_and_ranges():
    import torch
    from multioptpy_b200 import ops
    dev = "cuda:0"
    f64 = torch.float64
    # no structures
    z = lambda *s: torch.zeros(*s, dtype=f64, device=dev)
    Hp, gp, rank = ops.constraint_project(z(0, 2, 9), z(0, 9, 9), z(0, 9))
    assert Hp.shape == (0, 9, 9) and gp.shape == (0, 9) and rank.numel() == 0
    H, Hraw, st = ops.swart_hessian(z(0, 4, 3), np.ones(4))
    assert H.shape == (0, 12, 12)
    assert ops.hessian_sr_correction(z(0, 12, 12), z(0, 4, 3), np.ones(4), np.zeros(4)).shape == (0, 12, 12)
    # a rank that owns no image of the chain
    chain = torch.arange(5 * 3 * 3, dtype=f64, device=dev).reshape(5, 3, 3)
    assert ops.neb_redistribute(chain, 2, 0).shape == (0, 3, 3)
    torch.cuda.synchronize()


@pytest.mark.gpu
def test_smallest_sizes():
    import torch
    from multioptpy_b200 import ops, synthetic
    from multioptpy_b200.ModelHessian.swart import swart_radii
    dev = "cuda:0"
    # Swart on two and three atoms (no bends / one bend per centre) against the oracle
    for elems, xyz in ((["O", "H"], np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 1.8]])),
                       (["O", "H", "H"], np.array([[0.0, 0.0, 0.0], [0.0, 1.43, 1.1], [0.0, -1.43, 1.1]]))):
        r = np.array(swart_radii(elems))
        H, _, st = ops.swart_hessian(torch.from_numpy(xyz[None]).to(dev), r)
        assert rel(H[0].cpu().numpy(), O.swart_hessian(xyz, r)) < 1e-10, elems
    # a chain of two images: both end points stay
    X = np.random.default_rng(0).normal(size=(2, 4, 3))
    out = ops.neb_redistribute(torch.from_numpy(X).to(dev)).cpu().numpy()
    assert np.array_equal(out, X)
    # one zero constraint row = no constraint: P = I, Hp = sym(H), gp = g
    rng = np.random.default_rng(1)
    H = rng.normal(size=(3, 9, 9)); g = rng.normal(size=(3, 9))
    Hp, gp, rank = ops.constraint_project(torch.zeros(3, 1, 9, dtype=torch.float64, device=dev), torch.from_numpy(H).to(dev),
                                          torch.from_numpy(g).to(dev))
    assert int(rank.abs().sum()) == 0
    assert np.array_equal(gp.cpu().numpy(), g)
    assert rel(Hp.cpu().numpy(), 0.5 * (H + H.transpose(0, 2, 1))) < 1e-15


@pytest.mark.gpu
def test_constraint_projection_properties():
    """Random constraint rows incl. a dependent one: rank = number of independent rows, gp orthogonal to every row, Hp
    annihilates nothing but maps the rows to sigma times themselves, the subspace spectrum equals the oracle's."""
    import torch
    from multioptpy_b200 import ops
    dev = "cuda:0"
    rng = np.random.default_rng(7)
    B, n, k = 4, 30, 5
    C = rng.normal(size=(B, k, n)); C[:, 4] = 2.0 * C[:, 1] - 0.5 * C[:, 2]      # row 4 depends on rows 1, 2
    A = rng.normal(size=(B, n, n)); H = A + A.transpose(0, 2, 1)
    g = rng.normal(size=(B, n))
    Hp, gp, rank = ops.constraint_project(torch.from_numpy(C).to(dev), torch.from_numpy(H).to(dev), torch.from_numpy(g).to(dev))
    Hp, gp = Hp.cpu().numpy(), gp.cpu().numpy()
    assert (rank.cpu().numpy() == 4).all()
    for b in range(B):
        assert np.abs(C[b] @ gp[b]).max() < 1e-12 * np.abs(C[b]).max() * np.linalg.norm(g[b]) * 10
        U = O.constraint_null_space(C[b])
        assert U.shape[1] == n - 4
        lam_sub = np.linalg.eigvalsh(U.T @ H[b] @ U)
        lam_full = np.linalg.eigvalsh(Hp[b])
        sigma = np.linalg.norm(H[b]) + 1.0
        assert rel(lam_full[:n - 4], lam_sub) < 1e-12
        assert np.allclose(lam_full[n - 4:], sigma, rtol=1e-12)
        assert np.array_equal(Hp[b], Hp[b].T)
