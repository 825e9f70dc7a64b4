"""GPU parity: CUDA kernels (through the C ABI) vs the NumPy oracle and the golden
vectors produced by the unmodified reference.  Tolerance: 1e-10 relative on
Hessians, step vectors, energies; max|dlambda| <= 1e-10 max|lambda| on spectra
(north_star / SURVEY §8c)."""
import os

import numpy as np
import pytest
import torch

from multioptpy_b200 import ops, synthetic
from multioptpy_b200.Optimizer.rsirfo import RSIRFO
from multioptpy_b200.Optimizer.hessian_update import ModelHessianUpdate, BlockHessianUpdate
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-10
DEV = "cuda:0"


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


# --------------------------------------------------------------------------- update
def test_update_deltas_vs_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "update_deltas.npz"))
    for mid in np.unique(z["method"]):
        sel = np.where(z["method"] == mid)[0]
        H, s, y, ref = z["H"][sel], z["s"][sel], z["y"][sel], z["delta"][sel]
        d, st = ops.hessian_update(T(H), T(s), T(y), int(mid))
        d1, st1 = ops.hessian_update(T(H), T(s), T(y), int(mid), multi_cta=False)   # single-kernel path
        assert torch.equal(st, st1) and float((d - d1).abs().max()) <= 1e-12 * max(float(d1.abs().max()), 1e-300)
        d = d.cpu().numpy()
        for i in range(len(sel)):
            scale = max(np.linalg.norm(ref[i]), 1e-3 * np.linalg.norm(H[i]))
            err = np.linalg.norm(d[i] - ref[i]) / scale
            assert err < RTOL, (int(mid), O.UPDATE_NAMES[int(mid)], int(z["kind"][sel[i]]), err)


@pytest.mark.parametrize("n", [9, 33, 150, 200])
@pytest.mark.parametrize("method", [15, 23, 11, 13, 22, 25, 1])
def test_update_inplace_vs_oracle(n, method):
    rng = np.random.default_rng(n * 100 + method)
    B = 5
    H = np.stack([synthetic.spd_hessian(n, rng) for _ in range(B)])
    s = rng.normal(0, 0.05, (B, n))
    y = np.einsum("bij,bj->bi", H, s) + rng.normal(0, 5e-3, (B, n))
    y[1] = -y[1]            # negative curvature -> skipped
    s[2] *= 1e-12           # tiny step -> skipped
    Hd = T(H)
    Hd1 = T(H)
    _, st1 = ops.hessian_update(Hd1, T(s), T(y), method, inplace=True, rsirfo_guards=True, multi_cta=False)
    _, st = ops.hessian_update(Hd, T(s), T(y), method, inplace=True, rsirfo_guards=True)
    assert torch.equal(st, st1)
    assert float((Hd - Hd1).abs().max()) <= 1e-12 * float(Hd1.abs().max())
    assert torch.equal(Hd, Hd.transpose(1, 2))
    got, st = Hd.cpu().numpy(), st.cpu().numpy()
    for b in range(B):
        exp, upd = O.rsirfo_update_hessian(H[b], s[b], y[b], np.zeros(n), np.zeros(n), method)
        assert bool(st[b] & ops.ST_UPDATED) == upd
        assert rel(got[b], exp) < RTOL
    assert st[1] & ops.ST_UPD_SKIP_CURV and st[2] & ops.ST_UPD_SKIP_SMALL


def test_hessian_update_operator_classes(golden_dir):
    z = np.load(os.path.join(golden_dir, "update_deltas.npz"))
    i = int(np.where((z["method"] == 23) & (z["kind"] == 0))[0][0])
    d = ModelHessianUpdate(device=DEV).Bofill_hessian_update(z["H"][i], z["s"][i].reshape(-1, 1), z["y"][i].reshape(-1, 1))
    assert rel(d, z["delta"][i]) < RTOL
    i = int(np.where((z["method"] == 11) & (z["kind"] == 0))[0][0])
    d = BlockHessianUpdate(device=DEV).block_FSB_hessian_update(z["H"][i], z["s"][i].reshape(-1, 1), z["y"][i].reshape(-1, 1))
    assert np.linalg.norm(d - z["delta"][i]) / np.linalg.norm(z["H"][i]) < RTOL


# ----------------------------------------------------------------------- projection
def test_projection_vs_reference_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "projection.npz"))
    for i, na in enumerate(z["natoms"]):
        n = 3 * int(na)
        x, H, g = z["x"][i, :n], z["H"][i, :n, :n], z["g"][i, :n]
        Hp, gp, st = ops.project_trrot(T(H[None]), T(x[None]), g=T(g[None]))
        assert rel(Hp[0].cpu().numpy(), z["Hp"][i, :n, :n]) < 1e-12
        assert rel(gp[0].cpu().numpy(), z["gp"][i, :n]) < 1e-12
        Hp_np = Hp[0].cpu().numpy()
        assert np.array_equal(Hp_np, Hp_np.T)


def test_projection_with_bias_and_asymmetric_input():
    rng = np.random.default_rng(5)
    n = 90
    x = synthetic.grid_geometry(30, rng).reshape(-1)
    H = synthetic.spd_hessian(n, rng) + 1e-3 * rng.standard_normal((n, n))   # not symmetric
    Hb = 0.05 * synthetic.spd_hessian(n, rng)
    Hp, _, _ = ops.project_trrot(T(H[None]), T(x[None]), Hbias=T(Hb[None]))
    assert rel(Hp[0].cpu().numpy(), O.project_hessian_trrot(H + Hb, x)) < 1e-12


# --------------------------------------------------------------------------- eigh
@pytest.mark.parametrize("algo", ["jacobi", "tridiag", "auto"])
@pytest.mark.parametrize("n", [3, 24, 33, 90, 149, 150, 200])
def test_eigh_vs_lapack(n, algo):
    rng = np.random.default_rng(n)
    B = 6
    A = rng.standard_normal((B, n, n))
    A = 0.5 * (A + A.transpose(0, 2, 1))
    if n % 3 == 0:       # projected Hessian: exact 6-dimensional null space
        x = synthetic.grid_geometry(n // 3, rng).reshape(-1)
        A[0] = O.project_hessian_trrot(synthetic.spd_hessian(n, rng), x)
    A[1] = np.diag(np.arange(n, dtype=float))            # already diagonal
    w, V = np.linalg.eigh(A[2]); w[: n // 2] = 0.5       # heavy degeneracy
    A[2] = (V * w) @ V.T; A[2] = 0.5 * (A[2] + A[2].T)
    if algo == "tridiag" and n > 158:
        pytest.skip("shared-memory tridiagonal path supports n <= 158")
    evals, evecs, st = ops.eigh(T(A), algo)
    evals, evecs = evals.cpu().numpy(), evecs.cpu().numpy()
    assert not (st.cpu().numpy() & ops.ST_EIG_NOCONV).any()
    for b in range(B):
        ref = np.linalg.eigvalsh(A[b])
        scale = max(np.abs(ref).max(), 1e-300)
        assert np.abs(evals[b] - ref).max() <= 1e-13 * scale * max(1, n / 10)
        Vb = evecs[b].T                                   # columns = eigenvectors
        assert np.abs(Vb.T @ Vb - np.eye(n)).max() < 1e-12
        assert np.abs(A[b] @ Vb - Vb * evals[b]).max() < 1e-12 * scale * n


# ------------------------------------------------------------------------ rsirfo
def _trace_names(golden_dir):
    z = np.load(os.path.join(golden_dir, "rsirfo_traces.npz"))
    return z, [str(s) for s in z["names"]]


@pytest.mark.parametrize("idx", range(20))
def test_rsirfo_trace_vs_reference_golden(golden_dir, idx):
    """Drop-in class, reference calling convention (NumPy (n,1) arrays)."""
    z, names = _trace_names(golden_dir)
    name = names[idx]
    so, natoms, nsteps, bias, neb = [int(v) for v in z[f"{name}/meta"]]
    opt = RSIRFO(method=str(z[f"{name}/method"]), saddle_order=so, element_list=["C"] * natoms,
                 trust_radius_max=(0.1 if so > 0 else 0.5), trust_radius_min=0.01, device=DEV)
    if neb:
        opt.switch_NEB_mode()
    Hstate = z[f"{name}/H0"].copy()
    opt.set_hessian(Hstate)
    opt.set_bias_hessian(z[f"{name}/Hb"].copy())
    X, BG, G, BE = z[f"{name}/x"], z[f"{name}/Bg"], z[f"{name}/g"], z[f"{name}/Be"]
    col = lambda a: a.reshape(-1, 1).copy()
    for k in range(nsteps):
        if k == 0:
            mv = opt.run(col(X[k]), col(BG[k]), [], [], float(BE[k]), 0.0, [], col(X[0]), col(G[k]), [])
        else:
            mv = opt.run(col(X[k]), col(BG[k]), [], col(X[k - 1]), float(BE[k]), 0.0, [], col(X[0]),
                         col(G[k]), col(G[k - 1]))
        assert mv.shape == (3 * natoms, 1)
        assert rel(mv.ravel(), z[f"{name}/move"][k]) < RTOL, (name, k)
        assert rel(opt.hessian, z[f"{name}/H_after"][k]) < RTOL, (name, k)
        assert opt.hessian is Hstate                      # aliasing kept
        assert abs(opt.trust_radius - z[f"{name}/trust"][k]) < 1e-14, (name, k)
        p = z[f"{name}/pred"][k]
        assert abs(opt.predicted_energy_changes[-1] - p) <= RTOL * abs(p) + 1e-16, (name, k)


@pytest.mark.parametrize("natoms,saddle,method", [(50, 0, "rsirfo_bfgs"), (30, 1, "rsirfo_block_bofill"),
                                                    (24, 0, "rsirfo_block_fsb"), (8, 0, "rsirfo_bofill")])
def test_rsirfo_batched_vs_oracle(natoms, saddle, method):
    """Tensor mode: B structures per launch, two consecutive steps, vs the oracle."""
    B = 24
    n = 3 * natoms
    x0, H0, g0, rngs = synthetic.batch(2, B, natoms, saddle=saddle > 0)
    opt = RSIRFO(method=method, saddle_order=saddle, device=DEV)
    Hd = T(H0)
    opt.set_hessian(Hd)
    opt.set_bias_hessian(None)
    mv0 = opt.run(T(x0), T(g0), B_e=torch.zeros(B, dtype=torch.float64, device=DEV), g=T(g0)).cpu().numpy().copy()
    x1 = np.empty_like(x0); g1 = np.empty_like(g0)
    oracles = []
    for b in range(B):
        o = O.RSIRFOOracle(method=method, saddle_order=saddle)
        o.set_hessian(H0[b].copy()); o.set_bias_hessian(None)
        m = o.run(x0[b], g0[b], g0[b], None, None, 0.0)
        assert rel(mv0[b], m) < RTOL, ("step0", b)
        x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], m, rngs[b])
        oracles.append(o)
    mv1 = opt.run(T(x1), T(g1), pre_geom=T(x0), B_e=torch.full((B,), -1e-3, dtype=torch.float64, device=DEV),
                  g=T(g1), pre_g=T(g0)).cpu().numpy()
    Hn = Hd.cpu().numpy()
    lam = opt._out["eigvals"].cpu().numpy()
    for b, o in enumerate(oracles):
        m = o.run(x1[b], g1[b], g1[b], x0[b], g0[b], -1e-3)
        assert rel(mv1[b], m) < RTOL, ("step1", b)
        assert rel(Hn[b], o.hessian) < RTOL
        ref = o.last["eigvals"]
        assert np.abs(lam[b] - ref).max() <= RTOL * np.abs(ref).max()


def test_fast_reciprocal_accuracy():
    """The MUFU-seeded Newton reciprocal of the twisted factorisation is ~1 ulp."""
    from multioptpy_b200 import _lib
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.standard_normal(100000), 10.0 ** rng.uniform(-18, 18, 100000) * rng.choice([-1, 1], 100000)])
    xd = T(x); out = torch.empty_like(xd)
    _lib.check(_lib.load().mop_priv_fast_rcp(xd.data_ptr(), out.data_ptr(), x.size, None))
    torch.cuda.synchronize()
    err = np.abs(out.cpu().numpy() * x - 1.0)
    assert err.max() < 4 * 2.220446049250313e-16, err.max()


# ------------------------------------------------------------------------ producers
def _producers(golden_dir):
    z = np.load(os.path.join(golden_dir, "producers.npz"))
    return z, [str(s) for s in z["names"]]


@pytest.mark.parametrize("idx", range(8))
def test_connectivity_bit_exact_vs_reference(golden_dir, idx):
    from multioptpy_b200.Utils.bond_connectivity import BondConnectivity
    z, names = _producers(golden_dir)
    name = names[idx]
    elems = [str(e) for e in z[f"{name}/elements"]]
    tabs = BondConnectivity(device=DEV).connectivity_table(z[f"{name}/xyz"], elems)
    c = z[f"{name}/counts"]
    assert [len(t) for t in tabs] == list(c)
    assert tabs[0] == z[f"{name}/bonds"][:c[0]].tolist()
    assert tabs[1] == z[f"{name}/angles"][:c[1]].tolist()
    assert tabs[2] == z[f"{name}/dihedrals"][:c[2]].tolist()


@pytest.mark.parametrize("idx", range(8))
def test_fischer_vs_reference(golden_dir, idx):
    from multioptpy_b200.ModelHessian.approx_hessian import ApproxHessian
    z, names = _producers(golden_dir)
    name = names[idx]
    elems = [str(e) for e in z[f"{name}/elements"]]
    xyz = z[f"{name}/xyz"]
    H = ApproxHessian(device=DEV).main(xyz, elems, np.zeros_like(xyz), "fischer")
    assert rel(H, z[f"{name}/fischer"]) < RTOL
    assert np.array_equal(H, H.T)


def test_fischer_batched_jittered():
    """A batch of distinct conformers in one launch vs the oracle."""
    from multioptpy_b200.ModelHessian.fischer import FischerApproxHessian
    from multioptpy_b200.Parameters.tables import covalent_radius
    rng = np.random.default_rng(12)
    N, B = 24, 16
    elems = synthetic.elements(N)
    radii = np.array([covalent_radius(e) for e in elems])
    xyz = np.stack([synthetic.grid_geometry(N, rng, spacing=2.6, jitter=0.25) for _ in range(B)])
    H = FischerApproxHessian(device=DEV).main(T(xyz), elems).cpu().numpy()
    for b in range(0, B, 5):
        assert rel(H[b], O.fischer_hessian(xyz[b], radii)) < RTOL


@pytest.mark.parametrize("idx", range(8))
def test_afir_vs_reference(golden_dir, idx):
    from multioptpy_b200.Potential.AFIR_potential import AFIRPotential
    z, names = _producers(golden_dir)
    name = names[idx]
    elems = [str(e) for e in z[f"{name}/elements"]]
    xyz = z[f"{name}/xyz"]
    for c in range(3):
        f1 = [int(v) for v in z[f"{name}/afir_f1"][c] if v > 0]
        f2 = [int(v) for v in z[f"{name}/afir_f2"][c] if v > 0]
        pot = AFIRPotential(AFIR_Fragm_1=f1, AFIR_Fragm_2=f2, element_list=elems, device=DEV)
        gam = torch.tensor([float(z[f"{name}/afir_gamma"][c])], dtype=torch.float64)
        E, g, H = pot.calc_energy_grad_hess(xyz, gam)
        Eref = z[f"{name}/afir_E"][c]
        assert abs(float(E) - Eref) <= RTOL * abs(Eref)
        assert abs(float(pot.calc_energy(torch.tensor(xyz), gam)) - Eref) <= RTOL * abs(Eref)
        assert rel(g.cpu().numpy(), z[f"{name}/afir_g"][c]) < RTOL
        assert rel(H.cpu().numpy(), z[f"{name}/afir_H"][c]) < RTOL


def test_bias_potential_aggregator_aldol(golden_dir):
    """BiasPotentialCalculation.main with `-ma 95 1 5 50 3 11` on aldol_rxn.xyz."""
    from multioptpy_b200.Potential.potential import BiasPotentialCalculation
    z, _ = _producers(golden_dir)
    elems = [str(e) for e in z["aldol_rxn/elements"]]
    xyz = z["aldol_rxn/xyz"]
    fd = {"AFIR_gamma": [[95.0], [50.0]], "AFIR_Fragm_1": [[1], [3]], "AFIR_Fragm_2": [[5], [11]]}
    bg, Be, Bg, Hb = BiasPotentialCalculation(device=DEV).main(0.0, np.zeros_like(xyz), xyz, elems, fd)
    assert abs(Be - 3.794394592266868e-01) < 1e-10
    assert abs(np.linalg.norm(Bg) - 3.827240360969133e-02) < 1e-11
    assert abs(np.linalg.norm(Hb) - 7.463054721281174e-03) < 1e-12


# ------------------------------------------------------------------ packed lower-triangular storage
@pytest.mark.gpu
@pytest.mark.parametrize("natoms,method,bias", [(11, "rsirfo_bofill", False), (30, "rsirfo_block_fsb", True),
                                                (50, "rsirfo_bfgs", False), (8, "rsirfo_flowchart", True)])
def test_rsirfo_packed_storage_vs_oracle(natoms, method, bias):
    """mop_rsirfo_step_packed (Hessians as packed lower triangles, n (n + 1) / 2 doubles per structure): two steps
    vs the oracle (1e-10), the updated triangle vs the oracle's Hessian, pack / unpack round trip bit-exact."""
    import torch
    from multioptpy_b200 import ops, synthetic
    from multioptpy_b200.Optimizer.rsirfo import RSIRFO
    B = 6
    n = 3 * natoms
    x0, H0, g0, rngs = synthetic.batch(77, B, natoms)
    rng = np.random.default_rng(5)
    Hb = np.zeros_like(H0)
    if bias:
        for b in range(B):
            M = rng.standard_normal((n, 3)); Hb[b] = 0.02 * (M @ M.T)
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    P = ops.pack_lower(T(H0))
    assert torch.equal(ops.unpack_lower(P, n), T(H0))
    opt = RSIRFO(method=method, saddle_order=0, device=dev)
    opt.set_hessian_packed(P); opt.set_bias_hessian(ops.pack_lower(T(Hb)) if bias else None)
    z = torch.zeros(B, dtype=torch.float64, device=dev)
    mv0 = opt.run(T(x0), T(g0), B_e=z, g=T(g0)).cpu().numpy().copy()
    x1 = np.empty_like(x0); g1 = np.empty_like(g0); oracles = []
    for b in range(B):
        o = O.RSIRFOOracle(method=method, saddle_order=0)
        o.set_hessian(H0[b].copy()); o.set_bias_hessian(Hb[b].copy() if bias else None)
        m = o.run(x0[b], g0[b], g0[b], None, None, 0.0)
        assert rel(mv0[b], m) < RTOL, b
        x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], m, rngs[b])
        oracles.append(o)
    mv1 = opt.run(T(x1), T(g1), pre_geom=T(x0), B_e=z - 1e-3, g=T(g1), pre_g=T(g0)).cpu().numpy()
    Hfull = opt.get_hessian().cpu().numpy()
    for b, o in enumerate(oracles):
        m = o.run(x1[b], g1[b], g1[b], x0[b], g0[b], -1e-3)
        assert rel(mv1[b], m) < RTOL, b
        assert rel(Hfull[b], o.hessian) < RTOL, b
        assert np.array_equal(Hfull[b], Hfull[b].T)


@pytest.mark.gpu
@pytest.mark.parametrize("natoms", [12, 11])   # 11: an odd triangle (561 doubles) - chunks start on 8-byte boundaries (TMA phase match)
def test_host_pipeline_two_phase_vs_oracle_and_single_call(natoms):
    """HostStepPipeline (pinned host buffers, chunked copies, mop_rsirfo_step_packed_begin per chunk +
    mop_rsirfo_step_packed_finish once): two steps vs the oracle (1e-10) and equal to the one-call
    mop_rsirfo_step_packed on the same inputs to rounding (the one-call path reduces in stages whose warps sum their
    partial products in a different order; status words and states must agree exactly); uneven chunks incl. a
    single-structure chunk."""
    import torch
    from multioptpy_b200 import ops, synthetic
    from multioptpy_b200.host_pipeline import HostStepPipeline, pack_lower_host
    B, method = 7, "rsirfo_bfgs"
    n = 3 * natoms
    mid = ops.resolve_update_method(method)
    x0, H0, g0, rngs = synthetic.batch(91, B, natoms)
    dev = "cuda:0"
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pipe = HostStepPipeline(B, n, mid, device=dev, chunks=[1, 4, 2], nstream=2)
    hst = ops.new_rsirfo_state(B, 0.5, "cpu").pin_memory()
    h_move = torch.empty(B, n, dtype=torch.float64).pin_memory()
    h_stat = torch.empty(B, dtype=torch.int32).pin_memory()
    hx0, hg0 = pin(x0), pin(g0)
    pipe.step(hx0, hg0, hg0, hst, h_move, h_stat, hH=pin(pack_lower_host(H0)), hBe=pin(np.zeros(B)), state_back=True)
    mv0 = h_move.numpy().copy()
    # the same step through the single-call packed entry
    Pd = ops.pack_lower(T(H0)); std = ops.new_rsirfo_state(B, 0.5, dev)
    zero = torch.zeros(B, dtype=torch.float64, device=dev)
    o0 = ops.rsirfo_step(Pd, T(x0), T(g0), T(g0), std, method=mid, Be=zero, packed=True)
    assert rel(o0["move"].cpu().numpy(), mv0) < 1e-12
    x1 = np.empty_like(x0); g1 = np.empty_like(g0); oracles = []
    for b in range(B):
        o = O.RSIRFOOracle(method=method, saddle_order=0)
        o.set_hessian(H0[b].copy())
        m = o.run(x0[b], g0[b], g0[b], None, None, 0.0)
        assert rel(mv0[b], m) < RTOL, b
        x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], m, rngs[b])
        oracles.append(o)
    hx1, hg1 = pin(x1), pin(g1)
    # second step on the RESIDENT Hessians (hH = None), update active
    pipe.step(hx1, hg1, hg1, hst, h_move, h_stat, hx_prev=hx0, hg_prev=hg0, hBe=pin(np.full(B, -1e-3)), state_back=True)
    o1 = ops.rsirfo_step(Pd, T(x1), T(g1), T(g1), std, method=mid, x_prev=T(x0), g_prev=T(g0), Be=zero - 1e-3, packed=True)
    assert rel(o1["move"].cpu().numpy(), h_move.numpy()) < 1e-12
    assert np.array_equal(o1["status"].cpu().numpy(), h_stat.numpy())
    assert rel(std.cpu().numpy(), hst.numpy()) < 1e-12
    Hfull = pipe.hessians().cpu().numpy()
    for b, o in enumerate(oracles):
        m = o.run(x1[b], g1[b], g1[b], x0[b], g0[b], -1e-3)
        assert rel(h_move.numpy()[b], m) < RTOL, b
        assert rel(Hfull[b], o.hessian) < RTOL, b
    assert (h_stat.numpy() & ops.ST_UPDATED).all()
