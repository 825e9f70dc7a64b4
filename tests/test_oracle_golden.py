"""Pin the NumPy oracle (oracle/np_oracle.py) against golden vectors produced by
the unmodified reference (oracle/gen_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O

RTOL = 1e-10  # north_star tolerance (FP64 step vectors / Hessians / energies)


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def test_update_deltas_match_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "update_deltas.npz"))
    worst = 0.0
    for i in range(z["method"].size):
        mid = int(z["method"][i])
        H, s, y, ref = z["H"][i], z["s"][i], z["y"][i], z["delta"][i]
        got = O.hessian_update_delta(mid, H, s, y)
        # deltas of the block family are (B + d) - B: absolute rounding ~ eps * |H|
        scale = max(np.linalg.norm(ref), 1e-3 * np.linalg.norm(H))
        err = np.linalg.norm(got - ref) / scale
        worst = max(worst, err)
        assert err < RTOL, (mid, O.UPDATE_NAMES[mid], int(z["kind"][i]), err)
    assert worst < RTOL


def test_projection_matches_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "projection.npz"))
    for i, na in enumerate(z["natoms"]):
        n = 3 * int(na)
        x, H, g = z["x"][i, :n], z["H"][i, :n, :n], z["g"][i, :n]
        assert rel(O.project_hessian_trrot(H, x), z["Hp"][i, :n, :n]) < 1e-12
        assert rel(O.project_grad_trrot(g, x), z["gp"][i, :n]) < 1e-12


def _names(z):
    return [str(s) for s in z["names"]]


@pytest.mark.parametrize("idx", range(20))
def test_rsirfo_trace_matches_reference(golden_dir, idx):
    z = np.load(os.path.join(golden_dir, "rsirfo_traces.npz"))
    name = _names(z)[idx]
    so, natoms, nsteps, bias, neb = [int(v) for v in z[f"{name}/meta"]]
    opt = O.RSIRFOOracle(method=str(z[f"{name}/method"]), saddle_order=so,
                         trust_radius_max=(0.1 if so > 0 else 0.5), trust_radius_min=0.01)
    opt.NEB_mode = bool(neb)
    opt.set_hessian(z[f"{name}/H0"].copy())
    opt.set_bias_hessian(z[f"{name}/Hb"].copy())
    X, BG, G, BE = z[f"{name}/x"], z[f"{name}/Bg"], z[f"{name}/g"], z[f"{name}/Be"]
    for k in range(nsteps):
        xp = X[k - 1] if k > 0 else None
        gp = G[k - 1] if k > 0 else None
        mv = opt.run(X[k], BG[k], G[k], xp, gp, float(BE[k]))
        assert rel(mv, z[f"{name}/move"][k]) < RTOL, (name, k)
        assert rel(opt.hessian, z[f"{name}/H_after"][k]) < RTOL, (name, k)
        assert abs(opt.trust_radius - z[f"{name}/trust"][k]) < 1e-14, (name, k)
        p = z[f"{name}/pred"][k]
        assert abs(opt.pred[-1] - p) <= RTOL * abs(p) + 1e-16, (name, k)
