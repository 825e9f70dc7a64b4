"""Restraint bias potentials: oracle vs goldens generated from the reference
(oracle/gen_golden.py keep), CUDA kernel vs both, aggregator sum."""
import os

import numpy as np
import pytest

from oracle import np_oracle as O

RTOL = 1e-10


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def _z(golden_dir):
    z = np.load(os.path.join(golden_dir, "keep.npz"))
    return z, [str(s) for s in z["names"]]


def test_oracle_keep_vs_reference(golden_dir):
    z, names = _z(golden_dir)
    for name in names:
        k, p = z[f"{name}/kp"]
        if int(z[f"{name}/kind"]) == 4:
            p = float(z[f"{name}/phi0"])      # radians, float32- or float64-derived as the reference call did
        E, g, H = O.keep_egh(z[f"{name}/xyz"], int(z[f"{name}/kind"]), list(z[f"{name}/f1"]), list(z[f"{name}/f2"]), k, p)
        assert abs(E - float(z[f"{name}/E"])) <= 1e-14 * max(abs(E), 1e-300), name
        assert rel(g, z[f"{name}/g"]) < 1e-13 and rel(H, z[f"{name}/H"]) < 1e-13, name


@pytest.mark.gpu
def test_gpu_keep_vs_golden(golden_dir):
    from multioptpy_b200.Potential.keep_potential import (StructKeepAnglePotential, StructKeepDihedralAnglePotential,
                                                           StructKeepPotential, StructKeepPotentialv2)
    import torch
    z, names = _z(golden_dir)
    for name in names:
        kind = int(z[f"{name}/kind"]); k, p = z[f"{name}/kp"]
        f1 = [int(a) + 1 for a in z[f"{name}/f1"]]; f2 = [int(a) + 1 for a in z[f"{name}/f2"]]
        if kind == 1:
            pot = StructKeepPotential(device="cuda:0", keep_pot_spring_const=k, keep_pot_distance=p, keep_pot_atom_pairs=[f1[0], f2[0]])
        elif kind == 2:
            pot = StructKeepPotentialv2(device="cuda:0", keep_pot_v2_spring_const=k, keep_pot_v2_distance=p,
                                        keep_pot_v2_fragm1=f1, keep_pot_v2_fragm2=f2)
        elif kind == 3:
            pot = StructKeepAnglePotential(device="cuda:0", keep_angle_atom_pairs=f1, keep_angle_spring_const=k, keep_angle_angle=p)
        else:
            pot = StructKeepDihedralAnglePotential(device="cuda:0", keep_dihedral_angle_atom_pairs=f1,
                                                   keep_dihedral_angle_spring_const=float(k), keep_dihedral_angle_angle=float(p))   # Python floats, as the
            # reference is configured: torch.tensor(float) is float32, torch.tensor(np.float64) would not be
        if kind == 4:   # parameters as the reference call that made the golden: none (configured, float32) or float64 tensor
            params = [] if name.endswith("_cfg") else torch.tensor([k, p], dtype=torch.float64)
            E, g, H = pot.calc_energy_grad_hess(z[f"{name}/xyz"], params)
            Eref = float(z[f"{name}/E"])
            assert abs(float(E[0]) - Eref) <= RTOL * max(abs(Eref), 1e-12), name
            assert rel(g[0].cpu().numpy(), z[f"{name}/g"].ravel()) < RTOL, name
            assert rel(H[0].cpu().numpy(), z[f"{name}/H"]) < RTOL, name
            continue
        E, g, H = pot.calc_energy_grad_hess(z[f"{name}/xyz"], [k, p])
        Eref = float(z[f"{name}/E"])
        assert abs(float(E[0]) - Eref) <= RTOL * max(abs(Eref), 1e-12), name
        assert rel(g[0].cpu().numpy(), z[f"{name}/g"].ravel()) < RTOL, name
        assert rel(H[0].cpu().numpy(), z[f"{name}/H"]) < RTOL, name
        assert abs(float(pot.calc_energy(z[f"{name}/xyz"])) - Eref) <= RTOL * max(abs(Eref), 1e-12), name


@pytest.mark.gpu
def test_gpu_bias_aggregator_afir_plus_restraints(golden_dir):
    """BiasPotentialCalculation.main with an AFIR term and three restraints = sum of the parts."""
    from multioptpy_b200.Potential.potential import BiasPotentialCalculation
    z, _ = _z(golden_dir)
    xyz = z["keep_1_5/xyz"]; N = len(xyz)
    elems = ["C", "H", "H", "H", "C", "C", "H", "O", "H", "H", "H"][:N]
    fd = {"AFIR_gamma": [[95.0]], "AFIR_Fragm_1": [[1]], "AFIR_Fragm_2": [[5]],
          "keep_pot_spring_const": [0.4, 0.0], "keep_pot_distance": [1.6, 2.0], "keep_pot_atom_pairs": [[1, 5], [2, 3]],
          "keep_pot_v2_spring_const": [[0.7]], "keep_pot_v2_distance": [[2.5]], "keep_pot_v2_fragm1": [[1, 2, 3, 4]],
          "keep_pot_v2_fragm2": [[5, 6, 7, 8, 9]],
          "keep_angle_spring_const": [0.3], "keep_angle_angle": [109.5], "keep_angle_atom_pairs": [[2, 1, 3]]}
    bg, Be, Bg, Hb = BiasPotentialCalculation(device="cuda:0").main(0.0, np.zeros((N, 3)), xyz, elems, fd)
    fd_afir = {k: v for k, v in fd.items() if k.startswith("AFIR")}
    bg0, Be0, _, Hb0 = BiasPotentialCalculation(device="cuda:0").main(0.0, np.zeros((N, 3)), xyz, elems, fd_afir)
    Es = Be0; gs = bg0.copy(); Hs = Hb0.copy()
    for name in ("keep_1_5", "keepv2", "angle_gen"):
        Es += float(z[f"{name}/E"]); gs += z[f"{name}/g"]; Hs += z[f"{name}/H"]
    assert abs(Be - Es) <= RTOL * abs(Es) and rel(bg, gs) < RTOL and rel(Hb, Hs) < RTOL


@pytest.mark.gpu
def test_gpu_keep_batched_vs_oracle():
    import torch
    from multioptpy_b200 import ops, synthetic
    B, N = 6, 12
    xs = np.stack([synthetic.grid_geometry(N, np.random.default_rng(70 + b)) for b in range(B)])
    terms = [(ops.BIAS_KEEP, [0], [5], 0.3, 1.8), (ops.BIAS_KEEP_V2, [1, 2, 3], [7, 8], 0.9, 2.2), (ops.BIAS_KEEP_ANGLE, [4, 6, 9], [], 0.25, 100.0),
             (ops.BIAS_KEEP_DIHEDRAL, [2, 5, 8, 11], [], 0.35, 1.1), (ops.BIAS_KEEP_DIHEDRAL, [0, 3, 7, 10], [], 0.2, -2.9)]
    E, g, H = ops.bias_terms(torch.from_numpy(xs).cuda(), ops.pack_bias_terms(terms, torch.device("cuda:0")), len(terms))
    for b in range(B):
        Er, gr, Hr = 0.0, np.zeros((N, 3)), np.zeros((3 * N, 3 * N))
        for kind, f1, f2, k, p in terms:
            e1, g1, h1 = O.keep_egh(xs[b], kind, f1, f2, k, p)
            Er += e1; gr += g1; Hr += h1
        assert abs(float(E[b]) - Er) <= RTOL * abs(Er) and rel(g[b].cpu().numpy(), gr.ravel()) < RTOL and rel(H[b].cpu().numpy(), Hr) < RTOL, b


def test_every_reference_activation_key_is_handled_or_raises(golden_dir):
    """a force_data key the reference would turn into a potential must never be dropped silently.
    tests/golden/potential_keys.json lists every activation test of the reference aggregator
    (Potential/potential.py, generated by oracle/gen_golden.py potkeys)."""
    import json
    from multioptpy_b200.Potential import potential as P
    keys = json.load(open(os.path.join(golden_dir, "potential_keys.json")))
    assert set(keys) <= set(P._ACTIVATION)
    for key, (rule, _line) in keys.items():
        assert P._ACTIVATION[key] == rule, key
        fd = {key: [[1.0, 2.0]] if rule == "all" else [1.0]}
        assert P.active_keys(fd) == [key]
        off = {key: [[0.0, 2.0]] if rule == "all" else ([0.0] if rule == "nz" else [])}
        assert P.active_keys(off) == []
        if key not in P._HANDLED:
            with pytest.raises(P.MopError):
                P.BiasPotentialCalculation("").main(0.0, np.zeros((3, 3)), np.zeros((3, 3)), ["H"] * 3, fd)
