"""Large-n eigensolver (MOP_EIGH_LARGE: cluster tridiagonalisation streamed from L2, bisection,
twisted factorisation, register back-transform) against LAPACK, and the P-RFO step at
BASELINE config 5's size (n = 600) against the oracle."""
import numpy as np
import pytest

from oracle import np_oracle as O

pytestmark = pytest.mark.gpu


def _check(A, evals, evecs, tol_scale=1.0):
    n = A.shape[-1]
    for b in range(A.shape[0]):
        ref = np.linalg.eigvalsh(A[b])
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(evals[b] - ref).max() <= 1e-13 * scale * max(1, n / 10) * tol_scale, b
        Vb = evecs[b].T
        assert np.abs(Vb.T @ Vb - np.eye(n)).max() < 5e-12 * tol_scale, b
        assert np.abs(A[b] @ Vb - Vb * evals[b]).max() < 1e-12 * scale * n * tol_scale, b


@pytest.mark.parametrize("n,cl", [(162, 0), (200, 1), (200, 2), (201, 4), (384, 8), (600, 0), (600, 4), (1024, 8)])
def test_eigh_large_vs_lapack(n, cl):
    import torch
    from multioptpy_b200 import ops, synthetic, _lib
    rng = np.random.default_rng(n + cl)
    B = 4
    A = rng.standard_normal((B, n, n))
    A = 0.5 * (A + A.transpose(0, 2, 1))
    A[1] = synthetic.spd_hessian(n, rng, neg_lowest=True)
    if n % 3 == 0:       # projected Hessian: exact 6-dimensional null space (cluster at zero)
        x = synthetic.grid_geometry(n // 3, rng).reshape(-1)
        A[2] = O.project_hessian_trrot(synthetic.spd_hessian(n, rng), x)
    A[3] = np.diag(np.linspace(-1.0, 2.0, n)) + 1e-3 * A[3]
    lib = _lib.load()
    lib.mop_priv_large_cluster(cl)
    try:
        evals, evecs, st = ops.eigh(torch.from_numpy(A).cuda(), "large")
        torch.cuda.synchronize()
    finally:
        lib.mop_priv_large_cluster(0)
    st = st.cpu().numpy()
    assert not (st & ops.ST_EIG_NOCONV).any()
    _check(A, evals.cpu().numpy(), evecs.cpu().numpy())


def test_eigh_large_nonfinite_and_zero():
    import torch
    from multioptpy_b200 import ops
    n = 192
    A = np.zeros((3, n, n))
    A[1] = np.eye(n) * 2.0
    A[2, 5, 7] = A[2, 7, 5] = np.nan
    evals, evecs, st = ops.eigh(torch.from_numpy(A).cuda(), "large")
    evals = evals.cpu().numpy()
    assert np.array_equal(evals[0], np.zeros(n)) and np.allclose(evals[1], 2.0)
    assert not np.isfinite(evals[2]).all()


def test_rsprfo_n600_vs_oracle():
    """BASELINE config 5 shape: P-RFO + Bofill, N = 200 atoms, two consecutive steps."""
    import torch
    from multioptpy_b200 import synthetic
    from multioptpy_b200.Optimizer.rsprfo import EnhancedRSPRFO
    B, natoms = 3, 200
    x0, H0, g0, rngs = synthetic.batch(7, B, natoms, saddle=True)
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    opt = EnhancedRSPRFO(method="rsprfo_bofill", saddle_order=1, device=dev, display_flag=False)
    opt.set_hessian(T(H0)); opt.set_bias_hessian(None)
    oracles = []
    for b in range(B):
        o = O.RSPRFOOracle(method="rsprfo_bofill", saddle_order=1)
        o.set_hessian(H0[b]); oracles.append(o)
    x, g = x0.copy(), g0.copy()
    xp = gp = mp = None
    for it in range(2):
        Be = torch.full((B,), -1e-3 * it, dtype=torch.float64, device=dev)
        if it == 0:
            mv = opt.run(T(x), T(g), B_e=Be).cpu().numpy().copy()
        else:
            mv = opt.run(T(x), T(g), pre_B_g=T(gp), pre_geom=T(xp), B_e=Be, pre_move_vector=T(mp)).cpu().numpy().copy()
        for b, o in enumerate(oracles):
            m = o.run(x[b], g[b], xp[b] if it else None, gp[b] if it else None, -1e-3 * it, mp[b] if it else None)
            err = np.linalg.norm(mv[b] - m) / np.linalg.norm(m)
            assert err < 1e-10, (it, b, err)
        xp, gp, mp = x.copy(), g.copy(), mv.copy()
        x = x - mv
        g = np.stack([g0[b] + H0[b] @ (x[b] - x0[b]) for b in range(B)])


def test_rsirfo_n300_vs_oracle():
    """RS-I-RFO beyond the shared-memory eigensolver (n = 300): update + projection + factored
    eigendecomposition, two consecutive steps, minimum and first-order saddle search."""
    import torch
    from multioptpy_b200 import synthetic
    from multioptpy_b200.Optimizer.rsirfo import RSIRFO
    B, natoms = 3, 100
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for so in (0, 1):
        x0, H0, g0, rngs = synthetic.batch(8 + so, B, natoms, saddle=so > 0)
        opt = RSIRFO(method="rsirfo_bofill", saddle_order=so, device=dev)
        opt.set_hessian(T(H0)); opt.set_bias_hessian(T(np.zeros_like(H0)))
        oracles = []
        for b in range(B):
            o = O.RSIRFOOracle(method="rsirfo_bofill", saddle_order=so)
            o.set_hessian(H0[b].copy()); oracles.append(o)
        x, g = x0.copy(), g0.copy()
        xp = gp = None
        for it in range(2):
            Be = torch.full((B,), -1e-3 * it, dtype=torch.float64, device=dev)
            if it == 0:
                mv = opt.run(T(x), T(g), B_e=Be, g=T(g)).cpu().numpy().copy()
            else:
                mv = opt.run(T(x), T(g), pre_B_g=T(gp), pre_geom=T(xp), B_e=Be, g=T(g), pre_g=T(gp)).cpu().numpy().copy()
            for b, o in enumerate(oracles):
                m = o.run(x[b], g[b], g[b], xp[b] if it else None, gp[b] if it else None, -1e-3 * it)
                err = np.linalg.norm(mv[b] - m) / np.linalg.norm(m)
                assert err < 1e-10, (so, it, b, err)
            xp, gp = x.copy(), g.copy()
            x = x - mv
            g = np.stack([g0[b] + H0[b] @ (x[b] - x0[b]) for b in range(B)])
