"""Parity at the BASELINE config shapes: configs[2] NEB 64 x 30 against NEBRFOOracle over two
iterations, configs[3] the AFIR + Lindh + rsirfo_block_fsb chain at 8192 x N=24 and x N=8 against the oracle chain on a
strided sample, configs[1] full size with the update ACTIVE (step 1).  The chains are the ones bench.py times
(bench_configs.py)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

RTOL = 1e-10


def rel(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


def test_config3_chain_inputs_exercise_every_tangent_branch():
    """The synthetic NEB chain must hit the uphill, downhill and extremum branches of the BNEB tangent rule
    (pathopt_bneb_force.py:70-98) - otherwise the config-shape parity below would not cover them."""
    from multioptpy_b200 import synthetic
    X, E, G, H = synthetic.neb_chain(64, 30)
    up = sum(E[i - 1] < E[i] < E[i + 1] for i in range(1, 63))
    down = sum(E[i - 1] > E[i] > E[i + 1] for i in range(1, 63))
    assert up > 3 and down > 3 and 62 - up - down > 1
    assert X.shape == (64, 90) and H.shape == (64, 90, 90)


@pytest.mark.gpu
def test_config3_neb_64x30_two_iterations_vs_oracle():
    import torch
    import bench_configs as bc
    from multioptpy_b200.Optimizer.rfo_neb import RFOOptimizer
    ref = bc.neb_reference_two_iterations(64, 30)
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    opt = RFOOptimizer(64, 30, device=dev)
    opt.set_hessians(T(ref["H"]))
    d0 = opt.rfo_move_vectors(T(ref["X"]), T(ref["E"]), T(ref["G"])).cpu().numpy()
    assert rel(d0, ref["delta0"]) < RTOL
    d1 = opt.rfo_move_vectors(T(ref["X1"]), T(ref["E1"]), T(ref["G1"])).cpu().numpy()
    assert rel(d1, ref["delta1"]) < RTOL
    assert rel(opt.last["force"].cpu().numpy(), ref["force1"]) < RTOL
    assert rel(opt.hessian.cpu().numpy(), ref["H_after"]) < RTOL
    from multioptpy_b200 import ops
    st = opt.last["status"].cpu().numpy()
    assert np.all(st & ops.ST_UPDATED)          # iteration 2 runs with the quasi-Newton update active


@pytest.mark.gpu
@pytest.mark.parametrize("natoms", [24, 8])
def test_config4_afir_lindh_block_fsb_chain_vs_oracle(natoms):
    import torch
    import bench_configs as bc
    from multioptpy_b200 import ops, synthetic
    B = 8192
    xyz, g = synthetic.conformer_batch(B, natoms, seed=4000 + 7 * natoms)
    ch = bc.C4Chain(xyz, g, "cuda:0")
    r = ch.two_iterations()
    st1 = r["status1"].cpu().numpy()
    assert np.mean((st1 & ops.ST_UPDATED) != 0) > 0.5     # the update is active for most of the batch (s.y > 0)
    H1 = r["H_after"]
    assert torch.equal(H1, H1.transpose(1, 2))
    # MOP_ST_ALPHA_UNSTABLE: the secular root within ~1e-6 |pole| of its pole; the reference's own step then depends on
    # the summation order of its sums (rfo_secular.cuh) - rare, reported, excluded from the 1e-10 comparison
    unst = ((r["status0"] | r["status1"]).cpu().numpy() & ops.ST_ALPHA_UNSTABLE) != 0
    assert unst.mean() < 0.1
    assert torch.isfinite(r["move1"]).all()
    picks = [b for b in range(0, B, B // 12) if not unst[b]][:8]
    assert len(picks) == 8
    for b in picks:
        ref = bc.c4_chain_reference(xyz[b], g[b], ch.elems, ch.frag1, ch.frag2, 100.0)
        assert abs(float(r["E_afir"][b]) - ref["E_afir"]) <= RTOL * abs(ref["E_afir"]), b
        assert rel(r["g_afir"][b].cpu().numpy(), ref["g_afir"]) < RTOL, b
        assert rel(r["H_afir"][b].cpu().numpy(), ref["H_afir"]) < RTOL, b
        assert rel(r["H_model"][b].cpu().numpy(), ref["H_model"]) < RTOL, b
        assert rel(r["move0"][b].cpu().numpy(), ref["move0"]) < RTOL, b
        assert rel(r["move1"][b].cpu().numpy(), ref["move1"]) < RTOL, b
        assert rel(H1[b].cpu().numpy(), ref["H_after"]) < RTOL, b


@pytest.mark.gpu
def test_alpha_unstable_flag_on_pole_hugging_roots():
    """Sulfur atoms on a cubic grid: the LJ terms of the reference's Lindh Hessian (lindh.py:118-130) give modes at
    -8 Hartree / Bohr^2, the RFO root sits ~1e-6 below its pole and the alpha loop of rsirfo.py:986-1248 wanders for up
    to 40 micro-cycles between step norms that differ by 0.1 ... 1 (of ~2000).  The device replays the loop and reports
    MOP_ST_ALPHA_UNSTABLE; structures without the flag must match the oracle to 1e-10, the flagged ones are finite and
    either match or are another member of the same family (the trust-radius steepest-descent fallback, or the RFO step
    along the lowest mode)."""
    import torch
    import bench_configs as bc
    from multioptpy_b200 import ops, synthetic
    B, natoms = 64, 8
    xyz = np.empty((B, natoms, 3)); g = np.empty((B, 3 * natoms))
    for b in range(B):
        rng = np.random.default_rng(4056 + b)
        xyz[b] = synthetic.grid_geometry(natoms, rng, spacing=3.9, jitter=0.25)
        g[b] = rng.normal(0.0, 1e-2, 3 * natoms)
    ch = bc.C4Chain(xyz, g, "cuda:0")
    x0 = ch.xyz.reshape(B, -1).contiguous()
    Eb, gb, Hb = ch.afir(ch.xyz)
    H = ch.lindh(ch.xyz).clone()
    st = ops.new_rsirfo_state(B, 0.5, "cuda:0")
    o0 = ch.step(H, x0, ch.g, Eb, gb, Hb, st)
    mv = o0["move"].cpu().numpy(); stat = o0["status"].cpu().numpy()
    assert np.isfinite(mv).all()
    unst = (stat & ops.ST_ALPHA_UNSTABLE) != 0
    assert unst.mean() > 0.25 and np.all(stat[unst] & ops.ST_ALPHA_SEARCH)
    nmatch = 0
    for b in range(B):
        ref = bc.c4_chain_reference(xyz[b], g[b], ch.elems, ch.frag1, ch.frag2, 100.0, first_only=True)
        r = rel(mv[b], ref["move0_raw"])
        if not unst[b]:
            assert r < RTOL, b
        else:
            nmatch += r < 1e-6
            nd, no = np.linalg.norm(mv[b]), np.linalg.norm(ref["move0_raw"])
            cos = abs(mv[b] @ ref["move0_raw"]) / (nd * no)
            # same family: the RFO step along the pole's mode (norm within 1e-3 of the oracle's) or the fallback at |step| = trust
            assert abs(nd - 0.5) < 1e-12 or abs(no - 0.5) < 1e-12 or (cos > 1 - 1e-6 and abs(nd / no - 1) < 1e-2), b
    assert nmatch >= 0.5 * unst.sum()      # the replay usually lands on the reference's exit all the same


@pytest.mark.gpu
def test_config2_full_size_step1_update_active_vs_oracle():
    """1024 x N=50, rsirfo_bfgs: step 0, then step 1 with the Hessian update active; 12 structures vs the oracle
    (move, updated Hessian), batch-wide properties on all 1024 (status, symmetry, secant equation H' s = y for BFGS)."""
    import torch
    import bench
    from multioptpy_b200 import ops, synthetic
    from oracle import np_oracle as O
    B = 1024
    x0, H0, g0, rngs = bench.make_inputs(B, 0)
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    m = ops.resolve_update_method("rsirfo_bfgs")
    H = T(H0); st = ops.new_rsirfo_state(B, 0.5, torch.device(dev))
    z = torch.zeros(B, dtype=torch.float64, device=dev)
    mv0 = ops.rsirfo_step(H, T(x0), T(g0), T(g0), st, method=m, Be=z)["move"].cpu().numpy().copy()
    x1 = np.empty_like(x0); g1 = np.empty_like(g0)
    for b in range(B):
        x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
    out = ops.rsirfo_step(H, T(x1), T(g1), T(g1), st, method=m, x_prev=T(x0), g_prev=T(g0), Be=z - 1e-3)
    stat = out["status"].cpu().numpy()
    assert np.all(stat & ops.ST_UPDATED) and not np.any(stat & (ops.ST_EIG_FALLBACK | ops.ST_EIG_NOCONV))
    Hn = H.cpu().numpy()
    assert np.array_equal(Hn, Hn.transpose(0, 2, 1))
    s = x1 - x0; y = g1 - g0
    sec = np.linalg.norm(np.einsum("bij,bj->bi", Hn, s) - y, axis=1) / np.linalg.norm(y, axis=1)
    assert sec.max() < 1e-12                      # BFGS satisfies the secant equation
    mv1 = out["move"].cpu().numpy()
    assert np.isfinite(mv1).all()
    for b in range(0, B, 86):
        o = O.RSIRFOOracle(method="rsirfo_bfgs", saddle_order=0)
        o.set_hessian(H0[b].copy()); o.set_bias_hessian(None)
        o.run(x0[b], g0[b], g0[b], None, None, 0.0)
        mo = o.run(x1[b], g1[b], g1[b], x0[b], g0[b], -1e-3)
        assert rel(mv1[b], mo) < RTOL, b
        assert rel(Hn[b], o.hessian) < RTOL, b
