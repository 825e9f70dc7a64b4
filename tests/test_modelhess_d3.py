"""fischerd3old / fischerd3 model Hessians and the ts / clip modifiers: the oracle against goldens
produced by the reference's own dispatcher (oracle/gen_golden.py modelhess_d3), the CUDA path (k_model_hessian kinds 2
and 3, k_ts_modify, k_clip_recompose, through the ApproxHessian drop-in) against the same goldens."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import np_oracle as O  # noqa: E402
from multioptpy_b200.ModelHessian.fischerd3old import d3_atom_params, d3_dynamic_atom_params  # noqa: E402
from multioptpy_b200.Parameters.tables import covalent_radius  # noqa: E402

RTOL = 1e-10


def rel(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


def load(golden_dir):
    z = np.load(os.path.join(golden_dir, "modelhess_d3.npz"))
    return z, [str(n) for n in z["names"]]


def oracle_hessian(kind, xyz, elems):
    if kind.startswith("fischerd3old"):
        H = O.fischerd3old_hessian(xyz, d3_atom_params(elems))
    elif kind.startswith("fischerd3"):
        H = O.fischerd3_hessian(xyz, d3_dynamic_atom_params(elems))
    else:
        H = O.fischer_hessian(xyz, np.array([covalent_radius(e) for e in elems]))
    if "ts" in kind:
        H = O.ts_hessian(H)
    if "sr" in kind:
        H = O.sr_hessian(H, xyz, elems, np.array([covalent_radius(e) for e in elems]))
    if "clip" in kind:
        H = O.clip_hessian(H)
    return H


def test_oracle_matches_reference(golden_dir):
    z, names = load(golden_dir)
    for name in names:
        if name == "grid50":
            continue            # (covered on the GPU; the pure-Python oracle needs seconds per call at 50 atoms)
        elems = [str(e) for e in z[f"{name}/elements"]]
        for kind in [str(t) for t in z["types"]]:
            assert rel(oracle_hessian(kind, z[f"{name}/xyz"], elems), z[f"{name}/{kind}"]) < RTOL, (name, kind)


def test_linear_cases_hit_the_skip_branches(golden_dir):
    """CO2 / HCCH are exactly linear: every bend is skipped (|cos| > 0.9999) in the D3 variants, so their Hessians have
    no bending stiffness - unlike plain Fischer, which keeps the (ill-defined) bend rows."""
    z, _ = load(golden_dir)
    H = z["co2/fischerd3old"]
    assert np.abs(H[1::3, 1::3]).max() < 1e-3 and np.abs(H[0::3, 0::3]).max() > 0.1


def test_dispatch_raises_for_unsupported_types():
    from multioptpy_b200.ModelHessian.approx_hessian import ApproxHessian
    from multioptpy_b200._lib import MopError
    x = np.zeros((3, 3))
    for t in ("fischerd4", "schlegel", "lindh2007d3", "gfnff"):
        with pytest.raises(MopError):
            ApproxHessian(device="cuda:0").main(x, ["O", "C", "O"], np.zeros((3, 3)), t)


@pytest.mark.gpu
def test_cuda_matches_reference(golden_dir):
    from multioptpy_b200.ModelHessian.approx_hessian import ApproxHessian
    z, names = load(golden_dir)
    for name in names:
        elems = [str(e) for e in z[f"{name}/elements"]]
        xyz = z[f"{name}/xyz"]
        for kind in [str(t) for t in z["types"]]:
            H = ApproxHessian(device="cuda:0").main(xyz, elems, np.zeros_like(xyz), kind)
            tol = RTOL
            if "ts" in kind:
                # the reflection goes through ONE eigenvector: it is defined to eps ||H|| / gap only, in LAPACK as here
                base = ApproxHessian(device="cuda:0").main(xyz, elems, np.zeros_like(xyz), kind.replace("ts", ""))
                lam = np.linalg.eigvalsh(base)
                t = int(np.argmax(np.abs(lam) >= 1e-8))
                gap = min(abs(lam[t] - lam[q]) for q in range(len(lam)) if q != t)
                tol = max(RTOL, 100 * 2.2e-16 * np.abs(lam).max() / max(gap, 1e-300))
            assert rel(H, z[f"{name}/{kind}"]) < tol, (name, kind, tol)


@pytest.mark.gpu
def test_cuda_batched_jittered_vs_oracle():
    import torch
    from multioptpy_b200 import ops, synthetic
    B, N = 16, 24
    elems = synthetic.elements(N)
    xyz = np.stack([synthetic.grid_geometry(N, np.random.default_rng(900 + b), spacing=2.6, jitter=0.3) for b in range(B)])
    xd = torch.from_numpy(xyz).to("cuda:0")
    H2, _, st2 = ops.fischer_d3old_hessian(xd, d3_atom_params(elems))
    H3, _, st3 = ops.fischer_d3old_hessian(xd, d3_dynamic_atom_params(elems), dynamic=True)
    assert int(st2.max()) == 0 and int(st3.max()) == 0
    for b in (0, 5, 15):
        assert rel(H2[b].cpu().numpy(), O.fischerd3old_hessian(xyz[b], d3_atom_params(elems))) < RTOL
        assert rel(H3[b].cpu().numpy(), O.fischerd3_hessian(xyz[b], d3_dynamic_atom_params(elems))) < RTOL
    Hts, mod = ops.hessian_ts_modify(H2)
    assert int(mod.sum()) > 0
    assert rel(Hts[0].cpu().numpy(), O.ts_hessian(H2[0].cpu().numpy())) < RTOL
