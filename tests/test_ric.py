"""Redundant internal coordinates: oracle vs goldens generated from the reference
(oracle/gen_golden.py ric), CUDA kernels vs both."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O

RTOL = 1e-10


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def _load(golden_dir):
    z = np.load(os.path.join(golden_dir, "ric.npz"))
    return z, [str(s) for s in z["names"]]


def _tabs(z, name, wc=False):
    """wc: dihedrals within 0.05 rad of planar left out.  Their second derivatives (acos'' near +-1)
    are roundoff-dominated in the reference itself: two evaluation orders of the same formula differ
    by 1e-6 .. 1e-9 relative there, so the 1e-10 bar is applied to the well-conditioned table and a
    1e-5 bar to the full one."""
    c = z[f"{name}/counts"]
    if wc:
        return [z[f"{name}/bonds"][: c[0]], z[f"{name}/angles"][: c[1]], z[f"{name}/dihedrals_wc"][: int(z[f"{name}/n_dih_wc"])]]
    return [z[f"{name}/bonds"][: c[0]], z[f"{name}/angles"][: c[1]], z[f"{name}/dihedrals"][: c[2]]]


def test_oracle_ric_vs_reference(golden_dir):
    z, names = _load(golden_dir)
    for name in names:
        x = z[f"{name}/xyz"]
        Bm = O.ric_bmatrix(x)
        assert np.array_equal(Bm, z[f"{name}/Bmat"]), name
        K = O.ric_kmatrix(x, _tabs(z, name), z[f"{name}/q"])
        assert rel(K, z[f"{name}/K"]) < 1e-12, name
        assert rel(O.ric_kmatrix(x, _tabs(z, name, wc=True), z[f"{name}/q"]), z[f"{name}/K_wc"]) < 1e-12, name
        assert rel(Bm.T @ z[f"{name}/Hric"] @ Bm + K, z[f"{name}/Hc_full"]) < 1e-13, name
        rows = np.array([O.ric_partial_row(x, [int(a) for a in lab if a > 0]) for lab in z[f"{name}/labels"]])
        assert rel(rows, z[f"{name}/rows"]) < 1e-14, name
        ig = O.ric_int_grad(z[f"{name}/pB"], z[f"{name}/g"])
        assert rel(ig, z[f"{name}/int_grad"]) < 1e-12, name
        assert rel(O.ric_cart_grad(z[f"{name}/pB"], ig), z[f"{name}/cart_grad"]) < 1e-12, name
    for name in [str(s) for s in z["special_names"]]:
        lab = [int(a) for a in z[f"special/{name}/labels"] if a > 0]
        assert rel(O.ric_partial_row(z[f"special/{name}/xyz"], lab), z[f"special/{name}/row"]) < 1e-14, name


@pytest.mark.gpu
def test_gpu_ric_vs_golden(golden_dir):
    import torch
    from multioptpy_b200 import ops
    from multioptpy_b200.Coordinate import redundant_coordinate as rc
    z, names = _load(golden_dir)
    R = rc.RedundantInternalCoordinates(device="cuda:0")
    for name in names:
        x = z[f"{name}/xyz"]
        tabs = _tabs(z, name)
        assert rel(R.B_matrix(x), z[f"{name}/Bmat"]) < 1e-15, name
        assert rel(R.RICgrad2cartgrad(z[f"{name}/q"], coord=x), z[f"{name}/gq"]) < RTOL, name
        assert rel(R.K_matrix(x, _tabs(z, name, wc=True), z[f"{name}/q"]), z[f"{name}/K_wc"]) < RTOL, name
        assert rel(R.K_matrix(x, tabs, z[f"{name}/q"]), z[f"{name}/K"]) < 1e-5, name
        Kref = torch.from_numpy(z[f"{name}/K"][None]).cuda()
        Hfull = ops.ric_hess_to_cart(torch.from_numpy(x[None]).cuda(), torch.from_numpy(z[f"{name}/Hric"][None]).cuda(), Kref)
        assert rel(Hfull[0].cpu().numpy(), z[f"{name}/Hc_full"]) < RTOL, name
        assert rel(R.RIChess2carthess(x, tabs, z[f"{name}/Hric"], None, z[f"{name}/q"]), z[f"{name}/Hc_full"]) < 1e-5, name
        xd = torch.from_numpy(x[None]).cuda()
        Hd = ops.ric_hess_to_cart(xd, torch.from_numpy(z[f"{name}/hdiag"][None]).cuda(),
                                  torch.from_numpy(z[f"{name}/K"][None]).cuda())[0].cpu().numpy()
        assert rel(Hd, z[f"{name}/Hc_diag"]) < RTOL, name
        for lab, row in zip(z[f"{name}/labels"], z[f"{name}/rows"]):
            lab = [int(a) for a in lab if a > 0]
            f = {2: rc.partial_stretch_B_matirx, 3: rc.partial_bend_B_matrix, 4: rc.partial_torsion_B_matrix}[len(lab)]
            tol = RTOL
            if len(lab) == 4:   # phi = acos(c) loses digits like eps / phi^2 near planarity (same in the reference)
                phi = float(O._ric_coordinate_torch(torch.tensor(x[[a - 1 for a in lab]], dtype=torch.float64)))
                if not 0.05 < phi < np.pi - 0.05:
                    tol = 1e-6
            assert rel(f(x, *lab).ravel(), row) < tol, (name, lab)
        ig = rc.calc_int_grad_from_pBmat(z[f"{name}/g"].reshape(-1, 1), z[f"{name}/pB"])
        assert ig.shape == (len(z[f"{name}/pB"]), 1)
        assert rel(ig.ravel(), z[f"{name}/int_grad"]) < RTOL, name
        cg = rc.calc_cart_grad_from_pBmat(ig, z[f"{name}/pB"])
        assert rel(cg.ravel(), z[f"{name}/cart_grad"]) < RTOL, name


@pytest.mark.gpu
def test_gpu_ric_special_branches(golden_dir):
    from multioptpy_b200.Coordinate import redundant_coordinate as rc
    z, _ = _load(golden_dir)
    for name in [str(s) for s in z["special_names"]]:
        lab = [int(a) for a in z[f"special/{name}/labels"] if a > 0]
        f = {3: rc.partial_bend_B_matrix, 4: rc.partial_torsion_B_matrix}[len(lab)]
        got = f(z[f"special/{name}/xyz"], *lab).ravel()
        ref = z[f"special/{name}/row"]
        assert np.abs(got - ref).max() <= 1e-10 * max(np.abs(ref).max(), 1e-6), (name, got, ref)


@pytest.mark.gpu
def test_gpu_ric_batched_vs_oracle():
    """Tensor mode: batch of distinct geometries, per-structure connectivity from the device tables."""
    import torch
    from multioptpy_b200 import ops, synthetic
    from multioptpy_b200.Utils.bond_connectivity import radii_array
    B, N = 4, 20
    elems = synthetic.elements(N)
    xs = np.stack([synthetic.grid_geometry(N, np.random.default_rng(60 + b), spacing=2.6, jitter=0.25) for b in range(B)])
    xd = torch.from_numpy(xs).cuda()
    bonds, angles, dihs, counts, st = ops.connectivity(xd, radii_array(elems))
    M = N * (N - 1) // 2
    rng = np.random.default_rng(3)
    q = rng.normal(0, 1e-2, size=(B, M)); hd = np.abs(rng.normal(0.3, 0.1, size=(B, M)))
    K = ops.ric_kmatrix(xd, bonds, angles, dihs, counts, torch.from_numpy(q).cuda())
    H = ops.ric_hess_to_cart(xd, torch.from_numpy(hd).cuda(), K).cpu().numpy()
    g = ops.ric_grad_to_cart(xd, torch.from_numpy(q).cuda()).cpu().numpy()
    cn = counts.cpu().numpy(); bo, an, di = bonds.cpu().numpy(), angles.cpu().numpy(), dihs.cpu().numpy()
    for b in range(B):
        Bm = O.ric_bmatrix(xs[b])
        tabs = [bo[b, : cn[b, 0]], an[b, : cn[b, 1]], di[b, : cn[b, 2]]]
        Kref = O.ric_kmatrix(xs[b], tabs, q[b])
        assert rel(H[b], Bm.T @ np.diag(hd[b]) @ Bm + Kref) < 1e-7, b   # random cloud: some near-planar torsions
        assert rel(g[b], Bm.T @ q[b]) < RTOL, b
