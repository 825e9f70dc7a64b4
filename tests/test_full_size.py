"""Full-size (BASELINE.json configs[1]: B = 1024, 3N = 150) checks of the fused RS-I-RFO path through
size-independent properties, plus an oracle spot check on a sample of the batch.

The oracle cannot replay 1024 structures in seconds, so the whole batch is held to invariants of the
algorithm (rsirfo.py:360-490): the spectrum of the TR/ROT-projected Hessian has the trace and the Frobenius
norm of the matrix, six (near-)zero modes, and the returned step is the RFO step in the space of the kept
modes: (Hp - mu I) s = -gp restricted to that space for one scalar mu per structure, s orthogonal to the
TR/ROT null space.  Tolerances are written beside each assertion."""
import numpy as np
import pytest
import torch

from multioptpy_b200 import ops, synthetic
from oracle import np_oracle as O

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0") if torch.cuda.is_available() else None
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_c2_batch_properties_and_sample_parity():
    B, natoms = 1024, 50
    n = 3 * natoms
    x0, H0, g0, rngs = synthetic.batch(1001, B, natoms)
    Hd = T(H0)
    st = ops.new_rsirfo_state(B, 0.5, DEV)
    zero = torch.zeros(B, dtype=torch.float64, device=DEV)
    method = ops.resolve_update_method("rsirfo_bfgs")
    out = ops.rsirfo_step(Hd, T(x0), T(g0), T(g0), st, method=method, Be=zero)
    mv = out["move"].cpu().numpy()
    lam = out["eigvals"].cpu().numpy()
    status = out["status"].cpu().numpy()
    assert not (status & (ops.ST_EIG_NOCONV | ops.ST_EIG_NONFINITE)).any()
    Hp, gp, _ = ops.project_trrot(T(H0), T(x0), g=T(g0))
    Hp, gp = Hp.cpu().numpy(), gp.cpu().numpy()

    # spectrum: ascending, trace and Frobenius norm of Hp (1e-12 relative), six null modes (1e-10 of the scale)
    assert (np.diff(lam, axis=1) >= 0).all()
    tr = np.trace(Hp, axis1=1, axis2=2)
    fro2 = (Hp * Hp).sum(axis=(1, 2))
    scale = np.abs(lam).max(axis=1)
    assert np.abs(lam.sum(axis=1) - tr).max() <= 1e-12 * n * scale.max()
    assert np.abs((lam * lam).sum(axis=1) - fro2).max() <= 1e-12 * fro2.max()
    assert ((np.abs(lam) < 1e-10 * scale[:, None]).sum(axis=1) == 6).all()

    # step: s = -move solves the level-shifted system on the kept modes; mu from the Rayleigh-type quotient
    s = -mv
    Hs = np.einsum("bij,bj->bi", Hp, s)
    mu = ((Hs + gp) * s).sum(axis=1) / (s * s).sum(axis=1)
    resid = Hs - mu[:, None] * s + gp
    rel = np.linalg.norm(resid, axis=1) / np.linalg.norm(gp, axis=1)
    assert rel.max() < 1e-9, rel.max()          # gp has no component outside the kept modes for these inputs
    assert (mu < lam[:, 6] + 1e-12).all()        # the RFO shift lies below the lowest kept eigenvalue

    # oracle spot check (1e-10 relative on the step, the contract of the path) on a spread sample
    for b in list(range(0, B, 97)) + [B - 1]:
        o = O.RSIRFOOracle(method="rsirfo_bfgs", saddle_order=0)
        o.set_hessian(H0[b].copy()); o.set_bias_hessian(None)
        m = o.run(x0[b], g0[b], g0[b], None, None, 0.0)
        assert np.linalg.norm(mv[b] - m) / np.linalg.norm(m) < 1e-10, b
        ref = o.last["eigvals"]
        assert np.abs(lam[b] - ref).max() <= 1e-10 * np.abs(ref).max(), b
