"""Constrained RS-I-RFO (Optimizer/crsirfo.py): the oracle restatement and the CUDA path
(mop_constraint_project + mop_rsirfo_spectral_step + mop_crsirfo_finalize behind the CRSIRFO drop-in) against
multi-step traces of the reference's CRSIRFO.run recorded with synthetic.DistanceConstraints as the constraint object
(oracle/gen_golden.py crsirfo): two / three bond constraints, a duplicated row (rank truncation), saddle order 1, an
active SHAKE correction with its gradient transport term, a bias Hessian (added into self.hessian as the reference
does), the explicit subspace convergence exit, no constraints."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import np_oracle as O  # noqa: E402

RTOL = 1e-10


def rel(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


def cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "crsirfo_traces.npz"))
    return z, [str(n) for n in z["names"]]


def test_oracle_replays_reference_crsirfo(golden_dir):
    z, names = cases(golden_dir)
    for name in names:
        so, natoms, nsteps, bias = [int(v) for v in z[f"{name}/meta"]]
        o = O.CRSIRFOOracle(method=str(z[f"{name}/method"]), saddle_order=so, trust_radius_max=(0.1 if so > 0 else 0.5))
        o.set_hessian(z[f"{name}/H0"].copy())
        if bias:
            o.set_bias_hessian(z[f"{name}/Hb"].copy())
        unconstrained = name == "c_unconstrained"
        for k in range(nsteps):
            xp = z[f"{name}/x_in"][k - 1] if k else None
            gp = z[f"{name}/g"][k - 1] if k else None
            mv = o.run(z[f"{name}/x"][k], z[f"{name}/Bg"][k], z[f"{name}/g"][k], xp, gp, float(z[f"{name}/Be"][k]),
                       rows=None if unconstrained else z[f"{name}/rows"][k], shake=z[f"{name}/shake"][k])
            assert bool(z[f"{name}/converged"][k]) == o.converged_sub, (name, k)
            if o.converged_sub:
                assert np.all(mv == 0.0) and np.all(z[f"{name}/move"][k] == 0.0), (name, k)
            else:
                assert rel(mv, z[f"{name}/move"][k]) < 1e-9, (name, k)
                assert abs(o.last["pred"] - float(z[f"{name}/pred"][k])) <= 1e-9 * abs(float(z[f"{name}/pred"][k])), (name, k)
            assert rel(o.hessian, z[f"{name}/H_after"][k]) < 1e-12, (name, k)
            assert abs(o.trust_radius - float(z[f"{name}/trust"][k])) <= 1e-12, (name, k)


def test_traces_cover_the_branches(golden_dir):
    z, names = cases(golden_dir)
    assert int(z["c_min_converged/converged"].sum()) == 1
    assert np.linalg.norm(z["c_min_shake/shake"][0]) > 1e-3           # the SHAKE transport term is active
    assert all(int(z[f"{n}/converged"].sum()) == 0 for n in names if n != "c_min_converged")


@pytest.mark.gpu
def test_gpu_crsirfo_vs_reference_traces(golden_dir):
    import torch
    from multioptpy_b200 import ops, synthetic
    from multioptpy_b200.Optimizer.crsirfo import CRSIRFO
    z, names = cases(golden_dir)
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for name in names:
        so, natoms, nsteps, bias = [int(v) for v in z[f"{name}/meta"]]
        n = 3 * natoms
        opt = CRSIRFO(constraints=None, method=str(z[f"{name}/method"]), saddle_order=so, device=dev,
                      trust_radius_max=(0.1 if so > 0 else 0.5), trust_radius_min=0.01)
        H = T(z[f"{name}/H0"].reshape(1, n, n).copy())
        opt.set_hessian(H)
        if bias:
            opt.set_bias_hessian(T(z[f"{name}/Hb"].reshape(1, n, n)))
        for k in range(nsteps):
            one = lambda key, kk=k: T(z[f"{name}/{key}"][kk].reshape(1, -1))
            pre_x = one("x_in", k - 1) if k else []
            pre_g = one("g", k - 1) if k else []
            Be = torch.tensor([float(z[f"{name}/Be"][k])], dtype=torch.float64, device=dev)
            mv = opt.run(one("x"), one("Bg"), pre_geom=pre_x, B_e=Be, g=one("g"), pre_g=pre_g,
                         constraint_vectors=T(z[f"{name}/rows"][k][None]), shake_displacement=one("shake"))
            st = int(opt.last_status[0].item())
            conv = bool(z[f"{name}/converged"][k])
            assert conv == bool(st & ops.ST_CONSTR_CONVERGED), (name, k)
            ref_mv = z[f"{name}/move"][k]
            if conv:
                assert np.all(mv.cpu().numpy() == 0.0), (name, k)
            else:
                assert rel(mv[0].cpu().numpy(), ref_mv) < 1e-9, (name, k, rel(mv[0].cpu().numpy(), ref_mv))
            assert rel(opt.hessian[0].cpu().numpy(), z[f"{name}/H_after"][k]) < RTOL, (name, k)
            trust = float(opt.state_tensor[0, ops.RS_TRUST].item())
            assert abs(trust - float(z[f"{name}/trust"][k])) <= 1e-12, (name, k, trust)


@pytest.mark.gpu
def test_gpu_crsirfo_numpy_mode_with_constraint_object(golden_dir):
    """Reference calling convention: (n, 1) arrays, the constraint object called on the host, Hessian aliasing."""
    from multioptpy_b200 import synthetic
    from multioptpy_b200.Optimizer.crsirfo import CRSIRFO
    z, _ = cases(golden_dir)
    name = "c_min_shake"
    so, natoms, nsteps, bias = [int(v) for v in z[f"{name}/meta"]]
    x0 = z[f"{name}/x_in"][0].reshape(-1, 3)
    pairs = [(0, 1), (4, 6)]
    tg = [float(np.linalg.norm(x0[i] - x0[j]) + 0.03 * (1 + q)) for q, (i, j) in enumerate(pairs)]
    cons = synthetic.DistanceConstraints(pairs, targets=tg)
    opt = CRSIRFO(constraints=cons, method=str(z[f"{name}/method"]), saddle_order=so, device="cuda:0")
    H = z[f"{name}/H0"].copy()
    opt.set_hessian(H)
    col = lambda a: np.asarray(a).reshape(-1, 1).copy()
    for k in range(nsteps):
        pre_x = col(z[f"{name}/x_in"][k - 1]) if k else []
        pre_g = col(z[f"{name}/g"][k - 1]) if k else []
        mv = opt.run(col(z[f"{name}/x_in"][k]), col(z[f"{name}/Bg"][k]), [], pre_x, float(z[f"{name}/Be"][k]), 0.0, [],
                     col(x0), col(z[f"{name}/g"][k]), pre_g)
        assert mv.shape == (3 * natoms, 1)
        assert rel(mv.ravel(), z[f"{name}/move"][k]) < 1e-9, k
        assert rel(H, z[f"{name}/H_after"][k]) < RTOL, k        # written back into the caller's array
