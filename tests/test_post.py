"""Kabsch alignment and the convergence test: oracle vs goldens generated from the
reference (oracle/gen_golden.py post), CUDA kernels vs both."""
import os
import types

import numpy as np
import pytest

from oracle import np_oracle as O


def _z(golden_dir):
    return np.load(os.path.join(golden_dir, "post.npz"))


def test_oracle_post_vs_reference(golden_dir):
    z = _z(golden_dir)
    for c, N in enumerate(z["kabsch/natoms"]):
        Pa, Qc = O.kabsch(z["kabsch/P"][c, :N], z["kabsch/Q"][c, :N])
        assert np.abs(Pa - z["kabsch/P_aligned"][c, :N]).max() < 1e-13
        assert np.abs(Qc - z["kabsch/Q_centred"][c, :N]).max() < 1e-13
    t = z["conv/thresholds"]
    for c in range(len(z["conv/ok"])):
        ok, mdt, rdt, _ = O.check_convergence(z["conv/grad"][c], z["conv/disp"][c], *t)
        assert int(ok) == z["conv/ok"][c] and mdt == z["conv/max_disp_thr"][c] and rdt == z["conv/rms_disp_thr"][c]


@pytest.mark.gpu
def test_gpu_kabsch_vs_golden(golden_dir):
    from multioptpy_b200.Utils.calc_tools import Calculationtools
    z = _z(golden_dir)
    ct = Calculationtools(device="cuda:0")
    for c, N in enumerate(z["kabsch/natoms"]):
        P, Q = z["kabsch/P"][c, :N].copy(), z["kabsch/Q"][c, :N].copy()
        Pa, Qc = ct.kabsch_algorithm(P, Q)
        ref = z["kabsch/P_aligned"][c, :N]
        assert np.abs(Pa - ref).max() <= 1e-10 * np.abs(ref).max(), c
        assert np.abs(Qc - z["kabsch/Q_centred"][c, :N]).max() < 1e-13
        assert Qc is Q and abs(P.mean()) < 1e-13       # both arguments centred in place, as the reference


@pytest.mark.gpu
def test_gpu_kabsch_batched_and_collinear():
    import torch
    from multioptpy_b200 import ops, synthetic
    rng = np.random.default_rng(4)
    B, N = 64, 17
    Q = np.stack([synthetic.grid_geometry(N, np.random.default_rng(b)) for b in range(B)])
    P = Q + rng.normal(0, 0.1, Q.shape)
    P[3] = np.outer(np.arange(N), [1.0, 2.0, 3.0]); Q[3] = np.outer(np.arange(N), [2.0, -1.0, 0.5])   # collinear
    Pa, Qc, st = ops.kabsch(torch.from_numpy(P).cuda(), torch.from_numpy(Q).cuda())
    Pa, st = Pa.cpu().numpy(), st.cpu().numpy()
    assert st[3] == 1 and st.sum() == 1
    for b in range(B):
        if b == 3:
            continue
        ref, _ = O.kabsch(P[b], Q[b])
        assert np.abs(Pa[b] - ref).max() <= 1e-10 * np.abs(ref).max(), b


@pytest.mark.gpu
def test_gpu_convergence_vs_golden(golden_dir):
    import torch
    from multioptpy_b200 import ops
    from multioptpy_b200.Utils.calc_tools import ConvergenceChecker
    z = _z(golden_dir)
    t = z["conv/thresholds"]
    conv, out = ops.check_convergence(torch.from_numpy(z["conv/grad"]).cuda(), torch.from_numpy(z["conv/disp"]).cuda(), *t)
    assert np.array_equal(conv.cpu().numpy(), z["conv/ok"])
    out = out.cpu().numpy()
    assert np.array_equal(out[:, 1], z["conv/max_disp_thr"]) and np.allclose(out[:, 2], z["conv/rms_disp_thr"], rtol=1e-14, atol=0)
    cfg = types.SimpleNamespace(MAX_FORCE_THRESHOLD=t[0], RMS_FORCE_THRESHOLD=t[1], MAX_DISPLACEMENT_THRESHOLD=t[2],
                                RMS_DISPLACEMENT_THRESHOLD=t[3])
    chk = ConvergenceChecker(cfg, device="cuda:0")
    for c in (0, 5, 11):
        st = types.SimpleNamespace(effective_gradient=z["conv/grad"][c].reshape(-1, 3))
        ok, mdt, rdt = chk.check_convergence(st, z["conv/disp"][c].reshape(-1, 3), [])
        assert int(ok) == z["conv/ok"][c] and mdt == z["conv/max_disp_thr"][c]


def test_fix_atoms_oracle_matches_reference_expression():
    """The oracle restates optimization.py:1325-1343 verbatim; its defining properties: rows / columns of the fixed
    coordinates vanish (to the 1e-10 regularisation) and the free block is the Schur complement."""
    from oracle import np_oracle as O
    rng = np.random.default_rng(5)
    A = rng.standard_normal((18, 18)); H = A @ A.T + np.eye(18)
    He = O.fix_atoms_effective_hessian(H.copy(), [2, 5])
    fix = [3, 4, 5, 12, 13, 14]
    free = [i for i in range(18) if i not in fix]
    assert np.abs(He[fix]).max() < 1e-8 and np.abs(He[:, fix]).max() < 1e-8
    S = H[np.ix_(free, free)] - H[np.ix_(free, fix)] @ np.linalg.solve(H[np.ix_(fix, fix)], H[np.ix_(fix, free)])
    assert np.abs(He[np.ix_(free, free)] - S).max() < 1e-8


@pytest.mark.gpu
def test_fix_atoms_effective_hessian_vs_oracle():
    import torch
    from multioptpy_b200 import ops
    from oracle import np_oracle as O
    rng = np.random.default_rng(6)
    B, n = 5, 33
    Hs = []
    for b in range(B):
        A = rng.standard_normal((n, n)); Hs.append(A @ A.T / n + 0.1 * np.eye(n))
    Hs[3][np.ix_(range(6, 9), range(n))] = 0.0; Hs[3][np.ix_(range(n), range(6, 9))] = 0.0    # a singular fixed block: pinv cut-off
    H = np.stack(Hs)
    Hd = torch.from_numpy(H.copy()).to("cuda:0")
    ops.fix_atoms_effective_hessian(Hd, [3, 7, 10])
    for b in range(B):
        ref = O.fix_atoms_effective_hessian(H[b].copy(), [3, 7, 10])
        assert np.linalg.norm(Hd[b].cpu().numpy() - ref) <= 1e-10 * np.linalg.norm(ref)
