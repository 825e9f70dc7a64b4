"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.gen_golden            # all sets
    python -m oracle.gen_golden update rsirfo

The produced files are committed; the GPU box and the CPU test-suite only read
the .npz files.  Inputs come from ``multioptpy_b200.synthetic`` (seeded).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.np_oracle import UPDATE_DISPATCH  # noqa: E402  (ids/names only)
from multioptpy_b200 import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield


# ---------------------------------------------------------------- update ----
REF_UPDATE_FN = {
    1: ("M", "flowchart_hessian_update"),
    2: ("B", "block_CFD_FSB_hessian_update_dd"),
    3: ("B", "block_CFD_FSB_hessian_update_weighted"),
    4: ("B", "block_CFD_FSB_hessian_update"),
    5: ("B", "block_CFD_Bofill_hessian_update_weighted"),
    6: ("B", "block_CFD_Bofill_hessian_update"),
    7: ("B", "block_BFGS_hessian_update_dd"),
    8: ("B", "block_BFGS_hessian_update"),
    9: ("B", "block_FSB_hessian_update_dd"),
    10: ("B", "block_FSB_hessian_update_weighted"),
    11: ("B", "block_FSB_hessian_update"),
    12: ("B", "block_Bofill_hessian_update_weighted"),
    13: ("B", "block_Bofill_hessian_update"),
    14: ("M", "BFGS_hessian_update_dd"),
    15: ("M", "BFGS_hessian_update"),
    16: ("M", "SR1_hessian_update"),
    18: ("M", "CFD_FSB_hessian_update_dd"),
    19: ("M", "CFD_FSB_hessian_update"),
    20: ("M", "CFD_Bofill_hessian_update"),
    21: ("M", "FSB_hessian_update_dd"),
    22: ("M", "FSB_hessian_update"),
    23: ("M", "Bofill_hessian_update"),
    24: ("M", "PSB_hessian_update"),
    25: ("M", "MSP_hessian_update"),
}


def update_inputs(n, rng, kind):
    H = synthetic.spd_hessian(n, rng)
    s = rng.normal(0.0, 0.05, size=n)
    if kind == "normal":
        y = H @ s + rng.normal(0.0, 5e-3, size=n)
    elif kind == "near_secant":      # y ~ H s : tiny SR1 / phi denominators
        y = H @ s + rng.normal(0.0, 1e-9, size=n)
    elif kind == "low_curvature":    # 0 < s.y < 0.2 s.s : double damping active
        y = 0.05 * s + rng.normal(0.0, 1e-4, size=n)
    elif kind == "negative":         # s.y < 0
        y = -(H @ s)
    elif kind == "tiny_step":        # ||s|| below the block rank guards
        s = s * 1e-8
        y = H @ s
    elif kind == "flow_sr1":         # flowchart: z.s strongly negative
        H = H * 8.0
        y = H @ s * 0.3
    else:
        raise ValueError(kind)
    return H, s, y


def gen_update():
    hu = ref_shim.ref("Optimizer.hessian_update")
    bu = ref_shim.ref("Optimizer.block_hessian_update")
    kinds = ["normal", "near_secant", "low_curvature", "negative", "tiny_step", "flow_sr1"]
    mids, Hs, ss, ys, ds, ks = [], [], [], [], [], []
    n = 18
    for mid, (cls, fname) in sorted(REF_UPDATE_FN.items()):
        for ki, kind in enumerate(kinds):
            for rep in range(2):
                rng = np.random.default_rng(77000 + 100 * mid + 10 * ki + rep)
                H, s, y = update_inputs(n, rng, kind)
                obj = hu.ModelHessianUpdate() if cls == "M" else bu.BlockHessianUpdate()
                args = (H.copy(), s.reshape(-1, 1).copy(), y.reshape(-1, 1).copy())
                if mid == 1:
                    args = args + ("auto",)
                with quiet():
                    d = getattr(obj, fname)(*args)
                mids.append(mid); Hs.append(H); ss.append(s); ys.append(y)
                ds.append(np.asarray(d, dtype=np.float64)); ks.append(ki)
    np.savez_compressed(os.path.join(GOLD, "update_deltas.npz"), method=np.array(mids, np.int32),
                        kind=np.array(ks, np.int32), H=np.stack(Hs), s=np.stack(ss), y=np.stack(ys),
                        delta=np.stack(ds), kinds=np.array(kinds))
    print("update_deltas:", len(mids), "cases")


# ---------------------------------------------------------------- rsirfo ----
RSIRFO_CASES = [
    # (name, method, saddle_order, natoms, nsteps, bias, neb_mode, seed)
    ("bfgs_min_n24", "rsirfo_bfgs", 0, 8, 5, False, False, 1),
    ("bofill_min_n33", "rsirfo_bofill", 0, 11, 5, True, False, 2),
    ("fsb_min_n30", "rsirfo_fsb", 0, 10, 4, False, False, 3),
    ("msp_min_n24", "rsirfo_msp", 0, 8, 4, False, False, 4),
    ("psb_min_n24", "rsirfo_psb", 0, 8, 4, True, False, 5),
    ("sr1_min_n24", "rsirfo_sr1", 0, 8, 4, False, False, 6),
    ("cfd_bofill_min_n24", "rsirfo_cfd_bofill", 0, 8, 4, False, False, 7),
    ("cfd_fsb_min_n24", "rsirfo_cfd_fsb", 0, 8, 4, False, False, 8),
    ("blockfsb_min_n36", "rsirfo_block_fsb", 0, 12, 5, True, False, 9),
    ("blockbofill_ts_n36", "rsirfo_block_bofill", 1, 12, 5, False, False, 10),
    ("bofill_ts_n33", "rsirfo_bofill", 1, 11, 5, True, False, 11),
    ("blockbofill_neb_n30", "rsirfo_block_bofill", 1, 10, 4, False, True, 12),
    ("blockfsb_neb0_n30", "rsirfo_block_fsb", 0, 10, 4, False, True, 13),
    ("bofill_ts2_n24", "rsirfo_bofill", 2, 8, 4, False, False, 14),
    ("auto_min_n24", "rsirfo", 0, 8, 4, False, False, 15),
    ("blockbfgs_min_n24", "rsirfo_block_bfgs", 0, 8, 4, False, False, 16),
    ("bfgs_min_n72", "rsirfo_bfgs", 0, 24, 4, False, False, 17),
    ("bfgs_min_n150", "rsirfo_bfgs", 0, 50, 3, False, False, 18),
    ("fsb_dd_min_n24", "rsirfo_fsb_dd", 0, 8, 4, False, False, 19),
    ("blockcfdbofill_min_n24", "rsirfo_block_cfd_bofill", 0, 8, 4, False, False, 20),
]


class QuadraticPES:
    """E(x) = E0 + g0.(x-x0) + 1/2 (x-x0)^T Ht (x-x0) + c3 * sum((x-x0)^3) and a
    constant-Hessian harmonic bias; only used to feed consistent (E, g)."""

    def __init__(self, x0, g0, Ht, Hb, rng):
        self.x0, self.g0, self.Ht, self.Hb = x0, g0, Ht, Hb
        self.xc = x0 + rng.normal(0.0, 0.05, size=x0.size)
        self.c3 = 0.02

    def raw(self, x):
        d = x - self.x0
        e = self.g0 @ d + 0.5 * d @ self.Ht @ d + self.c3 * np.sum(d ** 3)
        g = self.g0 + self.Ht @ d + 3 * self.c3 * d ** 2
        return e, g

    def bias(self, x):
        d = x - self.xc
        return 0.5 * d @ self.Hb @ d, self.Hb @ d


def run_rsirfo_case(case):
    name, method, so, natoms, nsteps, bias, neb, seed = case
    rs = ref_shim.ref("Optimizer.rsirfo")
    rng = np.random.default_rng(424200 + seed)
    n = 3 * natoms
    x0 = synthetic.grid_geometry(natoms, rng).reshape(-1)
    H0 = synthetic.spd_hessian(n, rng, neg_lowest=so > 0)
    if so > 1:
        w, V = np.linalg.eigh(H0)
        w[1] = -0.02
        H0 = (V * w) @ V.T
        H0 = 0.5 * (H0 + H0.T)
    E = rng.standard_normal((n, n))
    Ht = H0 + 0.05 * (E + E.T) / np.sqrt(n)
    g0 = rng.normal(0.0, 2e-2, size=n)
    if bias:
        Bm = rng.standard_normal((n, 4))
        Hb = 0.02 * (Bm @ Bm.T)
    else:
        Hb = np.zeros((n, n))
    pes = QuadraticPES(x0, g0, Ht, Hb, rng)
    opt = rs.RSIRFO(method=method, saddle_order=so, element_list=["C"] * natoms,
                    trust_radius_max=(0.1 if so > 0 else 0.5), trust_radius_min=0.01)
    if neb:
        opt.switch_NEB_mode()
    opt.set_hessian(H0.copy())
    opt.set_bias_hessian(Hb.copy())
    rec = {k: [] for k in ("x", "Bg", "g", "Be", "move", "H_after", "trust", "pred")}
    x = x0.copy()
    x_prev = g_prev = None
    for k in range(nsteps):
        e, g = pes.raw(x)
        eb, gb = pes.bias(x)
        Be, Bg = e + eb, g + gb
        col = lambda a: a.reshape(-1, 1).copy()
        with quiet():
            if x_prev is None:
                mv = opt.run(col(x), col(Bg), [], [], Be, 0.0, [], col(x0), col(g), [])
            else:
                mv = opt.run(col(x), col(Bg), [], col(x_prev), Be, 0.0, [], col(x0), col(g), col(g_prev))
        mv = np.asarray(mv, float).ravel()
        rec["x"].append(x.copy()); rec["Bg"].append(Bg); rec["g"].append(g); rec["Be"].append(Be)
        rec["move"].append(mv); rec["H_after"].append(np.array(opt.hessian, float))
        rec["trust"].append(float(opt.trust_radius)); rec["pred"].append(float(opt.predicted_energy_changes[-1]))
        x_prev, g_prev = x.copy(), g.copy()
        step_cap = 0.1 if so > 0 else 0.5           # caller clamp, optimizer.py:792
        nrm = np.linalg.norm(mv)
        x = x - (mv * (step_cap / nrm) if nrm > step_cap else mv)
    out = {f"{name}/{k}": np.array(v) for k, v in rec.items()}
    out[f"{name}/H0"] = H0
    out[f"{name}/Hb"] = Hb
    out[f"{name}/meta"] = np.array([so, natoms, nsteps, int(bias), int(neb)], np.int64)
    out[f"{name}/method"] = np.array(method)
    return out


def gen_rsirfo():
    blob = {}
    for case in RSIRFO_CASES:
        blob.update(run_rsirfo_case(case))
        print("rsirfo case", case[0])
    blob["names"] = np.array([c[0] for c in RSIRFO_CASES])
    np.savez_compressed(os.path.join(GOLD, "rsirfo_traces.npz"), **blob)


def gen_projection():
    """TR/ROT projection of gradient and Hessian (calc_tools.py:249, rsirfo.py:128)."""
    rs = ref_shim.ref("Optimizer.rsirfo")
    ct = ref_shim.ref("Utils.calc_tools")
    xs, Hs, gs, Hps, gps = [], [], [], [], []
    for seed, natoms in enumerate([3, 5, 8, 11]):
        rng = np.random.default_rng(9100 + seed)
        n = 3 * natoms
        x = synthetic.grid_geometry(natoms, rng).reshape(-1)
        H = synthetic.spd_hessian(n, rng)
        g = rng.standard_normal(n)
        opt = rs.RSIRFO(method="rsirfo_bfgs", saddle_order=0)
        gp = opt._project_grad_tr_rot(g.copy(), x.reshape(-1, 1).copy())
        Hp = ct.Calculationtools().project_out_hess_tr_and_rot_for_coord(
            H.copy(), x.reshape(-1, 3).copy(), x.reshape(-1, 3).copy(), False)
        pad = 33 - n
        xs.append(np.pad(x, (0, pad))); gs.append(np.pad(g, (0, pad))); gps.append(np.pad(gp, (0, pad)))
        Hs.append(np.pad(H, ((0, pad), (0, pad)))); Hps.append(np.pad(Hp, ((0, pad), (0, pad))))
    np.savez_compressed(os.path.join(GOLD, "projection.npz"), natoms=np.array([3, 5, 8, 11]),
                        x=np.stack(xs), H=np.stack(Hs), g=np.stack(gs), Hp=np.stack(Hps), gp=np.stack(gps))
    print("projection: 4 cases")


RANKDEF_GEOMS = {
    # rank-deficient TR/ROT sets.  "exact": the dependent raw vector is exactly zero (axis-aligned), so the reference's
    # Householder Q is reproducible; "noise": the dependent column is rounding noise in the reference itself.
    "diatomic_z": (np.array([[0.0, 0.0, -1.1], [0.0, 0.0, 1.1]]), "exact"),
    "diatomic_x": (np.array([[-0.9, 0.0, 0.0], [1.3, 0.0, 0.0]]), "exact"),
    "diatomic_general": (np.array([[0.3, -0.2, 0.5], [1.1, 0.9, -0.7]]), "noise"),
    "linear3_z": (np.array([[0.0, 0.0, -2.2], [0.0, 0.0, 0.1], [0.0, 0.0, 2.3]]), "exact"),
    "linear3_x": (np.array([[-2.2, 0.0, 0.0], [0.1, 0.0, 0.0], [2.3, 0.0, 0.0]]), "exact"),
    "linear4_y": (np.array([[0.0, -3.1, 0.0], [0.0, -1.0, 0.0], [0.0, 1.2, 0.0], [0.0, 3.4, 0.0]]), "exact"),
    "linear3_general": (np.array([[-1.0, -1.0, -1.0], [0.1, 0.1, 0.1], [1.3, 1.3, 1.3]]), "noise"),
}


def gen_rankdef():
    """Gradient / Hessian projection and full RSIRFO / EnhancedRSPRFO steps on two-atom and linear geometries
    (rsirfo.py:128-190 reduced QR keeps six columns; rsprfo.py:244-285 drops |R_jj| <= 1e-10 columns and skips
    fewer than three atoms; calc_tools.py:249-316 Gram-Schmidt with drop)."""
    rs = ref_shim.ref("Optimizer.rsirfo")
    rp = ref_shim.ref("Optimizer.rsprfo")
    ct = ref_shim.ref("Utils.calc_tools")
    blob = {"names": np.array(list(RANKDEF_GEOMS)), "kind": np.array([v[1] for v in RANKDEF_GEOMS.values()])}
    for si, (name, (xyz, kind)) in enumerate(RANKDEF_GEOMS.items()):
        rng = np.random.default_rng(7700 + si)
        natoms = xyz.shape[0]; n = 3 * natoms
        x = xyz.reshape(-1)
        H = synthetic.spd_hessian(n, rng)
        g = rng.normal(0.0, 1e-2, size=n)
        with quiet():
            o1 = rs.RSIRFO(method="rsirfo_bfgs", saddle_order=0)
            o2 = rp.EnhancedRSPRFO(method="rsprfo_bofill", saddle_order=1, element_list=["C"] * natoms)
            gp1 = o1._project_grad_tr_rot(g.copy(), x.reshape(-1, 1).copy())
            gp2 = o2._project_grad_tr_rot(g.copy(), x.reshape(-1, 1).copy())
            Hp = ct.Calculationtools().project_out_hess_tr_and_rot_for_coord(
                H.copy(), x.reshape(-1, 3).copy(), x.reshape(-1, 3).copy(), False)
            # two full steps of each optimizer (the second with an update) on a quadratic surface
            col = lambda a: np.asarray(a, float).reshape(-1, 1).copy()
            o1.set_hessian(H.copy()); o1.set_bias_hessian(np.zeros((n, n)))
            mv1a = np.asarray(o1.run(col(x), col(g), [], [], 0.0, 0.0, [], col(x), col(g), []), float).ravel()
            x1 = x - mv1a; g1 = g + H @ (x1 - x) * 1.05
            mv1b = np.asarray(o1.run(col(x1), col(g1), col(g), col(x), -1e-3, 0.0, col(mv1a), col(x), col(g1), col(g)), float).ravel()
            Hn = synthetic.spd_hessian(n, rng, neg_lowest=True)
            o2.set_hessian(Hn.copy()); o2.set_bias_hessian(np.zeros((n, n)))
            mv2a = np.asarray(o2.run(col(x), col(g), [], [], 0.0, 0.0, [], col(x), col(g), []), float).ravel()
        blob.update({f"{name}/x": x, f"{name}/H": H, f"{name}/g": g, f"{name}/gp_rsirfo": gp1, f"{name}/gp_rsprfo": gp2,
                     f"{name}/Hp": Hp, f"{name}/move_rsirfo0": mv1a, f"{name}/x1": x1, f"{name}/g1": g1,
                     f"{name}/move_rsirfo1": mv1b, f"{name}/H_rsirfo1": np.array(o1.hessian, float),
                     f"{name}/Hn": Hn, f"{name}/move_rsprfo0": mv2a})
        print("rankdef", name, kind, "|gp1|", np.linalg.norm(gp1), "|gp2|", np.linalg.norm(gp2))
    np.savez_compressed(os.path.join(GOLD, "rankdef.npz"), **blob)


def gen_potkeys():
    """Activation keys of BiasPotentialCalculation (Potential/potential.py): every force_data key the reference tests
    before it builds a potential term, with the form of the test."""
    import json, re
    src = open(os.path.join(ref_shim.REF_ROOT, "multioptpy", "Potential", "potential.py")).read().splitlines()
    keys = {}
    for ln, line in enumerate(src, 1):
        t = line.strip()
        if not t.startswith("if "):
            continue
        for m in re.finditer(r'force_data\["([A-Za-z0-9_]+)"\]', t):
            k = m.group(1)
            if re.search(r'not 0\.0 in force_data\["%s"\]' % k, t):
                keys.setdefault(k, ["all", ln])
            elif re.search(r'force_data\["%s"\]\[i\] != 0\.0' % k, t):
                keys.setdefault(k, ["nz", ln])
            elif re.search(r'len\(force_data\["%s"\]\) > 0' % k, t):
                keys.setdefault(k, ["len", ln])
    with open(os.path.join(GOLD, "potential_keys.json"), "w") as f:
        json.dump(keys, f, indent=1, sort_keys=True)
    print("potkeys:", len(keys), "activation keys")


def gen_neb_full():
    """The UNMODIFIED reference RFOOptimizer.optimize (Optimizer/rfo_neb.py:104-208) driven for three NEB iterations:
    BNEB force, per-image RS-I-RFO with Ayala update, TR_calc, FIRE move, RFO / FIRE combine -> new geometry."""
    import tempfile, types
    bn = ref_shim.ref("MEP.pathopt_bneb_force")
    rn = ref_shim.ref("Optimizer.rfo_neb")
    nimg, natoms = 9, 10
    n = 3 * natoms
    tmp = tempfile.mkdtemp() + "/"
    cfg = types.SimpleNamespace(NEB_FOLDER_DIRECTORY=tmp, fix_init_edge=False, fix_end_edge=False,
                                apply_convergence_criteria=False, element_list=synthetic.elements(natoms),
                                bohr2angstroms=0.52917721067, dt=0.5, a=0.10, n_reset=0, FIRE_N_accelerate=5,
                                FIRE_f_inc=1.10, FIRE_f_accelerate=0.99, FIRE_f_decelerate=0.5, FIRE_a_start=0.1,
                                FIRE_dt_max=3.0)
    rngH = np.random.default_rng(78)
    H_init = np.stack([synthetic.spd_hessian(n, rngH) for _ in range(nimg)])
    for i in range(nimg):
        np.save(os.path.join(tmp, f"tmp_hessian_{i}.npy"), H_init[i])
    opt = rn.RFOOptimizer(cfg)
    calc = bn.CaluculationBNEB()
    X, E, G = neb_chain(nimg, natoms, 41)
    rec = {k: [] for k in ("X", "E", "G", "V", "Vprev", "new_geom_ang")}
    V = np.zeros((nimg, natoms, 3)); Vprev = np.zeros((nimg, natoms, 3))
    prevX = prevG = None
    rngv = np.random.default_rng(5)
    for it in range(3):
        geoms = X.reshape(nimg, natoms, 3).copy(); grads = G.reshape(nimg, natoms, 3).copy()
        with quiet():
            new_ang = opt.optimize(geoms, grads, None if prevX is None else prevX.reshape(nimg, natoms, 3),
                                   None if prevG is None else prevG.reshape(nimg, natoms, 3), it, E.copy(), E.copy(),
                                   Vprev.copy(), V.copy(), None, None, calc)
        rec["X"].append(X.copy()); rec["E"].append(E.copy()); rec["G"].append(G.copy())
        rec["V"].append(V.copy()); rec["Vprev"].append(Vprev.copy()); rec["new_geom_ang"].append(np.asarray(new_ang, float))
        prevX, prevG = X.copy(), G.copy()
        Xn, En, Gn = neb_chain(nimg, natoms, 42 + it)
        X = (np.asarray(new_ang, float) / cfg.bohr2angstroms).reshape(nimg, n)
        E = En; G = G + 0.3 * (Gn - G)
        Vprev = V.copy(); V = rngv.normal(0.0, 0.02, (nimg, natoms, 3))
    blob = {k: np.array(v) for k, v in rec.items()}
    blob["H_init"] = H_init
    blob["H_final"] = np.stack([np.load(os.path.join(tmp, f"tmp_hessian_{i}.npy")) for i in range(nimg)])
    blob["meta"] = np.array([nimg, natoms], np.int64)
    np.savez_compressed(os.path.join(GOLD, "neb_full.npz"), **blob)
    print("neb_full: |dx| per iteration", [float(np.abs(rec["new_geom_ang"][i] / cfg.bohr2angstroms - rec["X"][i].reshape(nimg, natoms, 3)).max()) for i in range(3)])


def read_xyz(path):
    """Minimal xyz reader (Angstrom) -> (elements, coords in Bohr)."""
    lines = [l.split() for l in open(path).read().strip().splitlines()]
    rows = [l for l in lines if len(l) >= 4 and l[0][0].isalpha()]
    elems = [r[0] for r in rows]
    xyz = np.array([[float(v) for v in r[1:4]] for r in rows]) / 0.52917721067
    return elems, xyz


PRODUCER_CASES = [
    ("aldol_rxn", "test/aldol_rxn.xyz"), ("s8", "test/s8_for_confomation_search_test.xyz"),
    ("aldol_rxn_PT", "test/aldol_rxn_PT.xyz"), ("reductive_elimination", "test/reductive_elimination.xyz"),
    ("claisen", "test/claisen_rearrengment.xyz"), ("SN2", "test/SN2.xyz"),
    ("grid24", None), ("grid50", None),
]


def producer_geometry(name, path):
    if path is not None:
        return read_xyz(os.path.join(ref_shim.REF_ROOT, path))
    natoms = int(name[4:])
    rng = np.random.default_rng(5150 + natoms)
    # a compact, chemically plausible cloud: grid spacing 2.6 Bohr so that bonds/angles/dihedrals exist
    xyz = synthetic.grid_geometry(natoms, rng, spacing=2.6, jitter=0.25)
    return synthetic.elements(natoms), xyz


def pad_table(tab, width):
    a = np.full((max(len(tab), 1), width), -1, np.int32)
    for i, row in enumerate(tab):
        a[i] = row
    return a


def gen_producers():
    import torch
    bc = ref_shim.ref("Utils.bond_connectivity")
    fi = ref_shim.ref("ModelHessian.fischer")
    af = ref_shim.ref("Potential.AFIR_potential")
    blob = {"names": np.array([c[0] for c in PRODUCER_CASES])}
    for name, path in PRODUCER_CASES:
        elems, xyz = producer_geometry(name, path)
        N = len(elems)
        with quiet():
            tabs = bc.BondConnectivity().connectivity_table(xyz.copy(), elems)
            Hf = fi.FischerApproxHessian().main(xyz.copy(), elems, np.zeros((N, 3)))
        blob[f"{name}/elements"] = np.array(elems)
        blob[f"{name}/xyz"] = xyz
        blob[f"{name}/counts"] = np.array([len(t) for t in tabs], np.int32)
        blob[f"{name}/bonds"] = pad_table(tabs[0], 2)
        blob[f"{name}/angles"] = pad_table(tabs[1], 3)
        blob[f"{name}/dihedrals"] = pad_table(tabs[2], 4)
        blob[f"{name}/fischer"] = np.asarray(Hf, float)
        # AFIR: a single atom pair and a half/half fragment split (1-based indices as on the CLI)
        half = N // 2
        afir_cases = [([1], [min(5, N)], 95.0), (list(range(1, half + 1)), list(range(half + 1, N + 1)), 100.0),
                      ([2, 3], [N], -40.0)]
        Es, Gs, Hs, F1, F2, GAM = [], [], [], [], [], []
        for f1, f2, gamma in afir_cases:
            pot = af.AFIRPotential(AFIR_Fragm_1=f1, AFIR_Fragm_2=f2, element_list=elems)
            geom = torch.tensor(xyz, dtype=torch.float64, requires_grad=True)
            par = torch.tensor([gamma], dtype=torch.float64, requires_grad=True)
            E = pot.calc_energy(geom, par)
            g = torch.func.jacrev(pot.calc_energy, argnums=0)(geom, par)
            H = torch.func.hessian(pot.calc_energy, argnums=0)(geom, par).reshape(3 * N, 3 * N)
            Es.append(float(E)); Gs.append(g.detach().numpy()); Hs.append(H.detach().numpy())
            F1.append(np.pad(np.array(f1, np.int32), (0, N - len(f1)), constant_values=-1))
            F2.append(np.pad(np.array(f2, np.int32), (0, N - len(f2)), constant_values=-1))
            GAM.append(gamma)
        blob[f"{name}/afir_E"] = np.array(Es); blob[f"{name}/afir_g"] = np.stack(Gs); blob[f"{name}/afir_H"] = np.stack(Hs)
        blob[f"{name}/afir_f1"] = np.stack(F1); blob[f"{name}/afir_f2"] = np.stack(F2); blob[f"{name}/afir_gamma"] = np.array(GAM)
        print("producers case", name, N, "atoms, tables", [len(t) for t in tabs])
    np.savez_compressed(os.path.join(GOLD, "producers.npz"), **blob)


def gen_c1_trace():
    """Config 1 (optmain test/aldol_rxn.xyz -ma 95 1 5 50 3 11 -opt rsirfo_bofill -modelhess fischer):
    the reference CalculateMoveVector.calc_move_vector boundary replayed on recorded gradients.
    GFN2-xTB is not installable offline, so the raw (E, g) come from a seeded analytic PES; the
    bias, the model Hessian, the optimizer and the caller are the unmodified reference."""
    import torch
    opt_mod = ref_shim.ref("optimizer")
    fi = ref_shim.ref("ModelHessian.fischer")
    af = ref_shim.ref("Potential.AFIR_potential")
    elems, xyz0 = read_xyz(os.path.join(ref_shim.REF_ROOT, "test/aldol_rxn.xyz"))
    N = len(elems); n = 3 * N
    rng = np.random.default_rng(20260101)
    with quiet():
        H0 = np.asarray(fi.FischerApproxHessian().main(xyz0.copy(), elems, np.zeros((N, 3))), float)
    E = rng.standard_normal((n, n))
    Ht = H0 + 0.02 * (E + E.T) / np.sqrt(n) + 0.05 * np.eye(n)
    g0 = rng.normal(0.0, 5e-3, size=n)
    pes = QuadraticPES(xyz0.reshape(-1).copy(), g0, Ht, np.zeros((n, n)), rng)
    terms = [([1], [5], 95.0), ([3], [11], 50.0)]
    pots = [af.AFIRPotential(AFIR_Fragm_1=f1, AFIR_Fragm_2=f2, element_list=elems) for f1, f2, _ in terms]

    def bias(x):
        Eb, gb, Hb = 0.0, np.zeros((N, 3)), np.zeros((n, n))
        for pot, (_, _, gam) in zip(pots, terms):
            geom = torch.tensor(x.reshape(N, 3), dtype=torch.float64, requires_grad=True)
            par = torch.tensor([gam], dtype=torch.float64, requires_grad=True)
            Eb += float(pot.calc_energy(geom, par))
            gb = gb + torch.func.jacrev(pot.calc_energy, argnums=0)(geom, par).detach().numpy()
            Hb = Hb + torch.func.hessian(pot.calc_energy, argnums=0)(geom, par).reshape(n, n).detach().numpy()
        return Eb, gb, Hb

    with quiet():
        CMV = opt_mod.CalculateMoveVector(0.5, elems, saddle_order=0, FC_COUNT=-1, temperature=0.0,
                                          model_hess_flag="fischer", max_trust_radius=None, min_trust_radius=None)
        insts = CMV.initialization(["rsirfo_bofill"])
    Model_hess = H0.copy()
    geom = xyz0.copy()
    pre = dict(B_g=np.zeros((N, 3)), geom=np.zeros((N, 3)), B_e=0.0, move=np.zeros((N, 3)), g=np.zeros((N, 3)))
    rec = {k: [] for k in ("geom", "B_g", "g", "B_e", "Hb", "new_geom", "move", "trust", "H_after")}
    for it in range(6):
        e, g = pes.raw(geom.reshape(-1))
        g = g.reshape(N, 3)
        Eb, gb, Hb = bias(geom.reshape(-1))
        B_e, B_g = e + Eb, g + gb
        with quiet():
            insts[0].set_hessian(Model_hess)          # every iteration, by reference
            insts[0].set_bias_hessian(Hb)
            new_geom, move, insts = CMV.calc_move_vector(it, geom.copy(), B_g.copy(), pre["B_g"].copy(), pre["geom"].copy(),
                                                         B_e, pre["B_e"], pre["move"].copy(), xyz0.copy(), g.copy(),
                                                         pre["g"].copy(), insts, print_flag=False)
        rec["geom"].append(geom.copy()); rec["B_g"].append(B_g.copy()); rec["g"].append(g.copy()); rec["B_e"].append(B_e)
        rec["Hb"].append(Hb.copy()); rec["new_geom"].append(np.asarray(new_geom, float).copy())
        rec["move"].append(np.asarray(move, float).copy()); rec["trust"].append(float(CMV.trust_radii))
        rec["H_after"].append(np.asarray(insts[0].hessian, float).copy())
        pre = dict(B_g=B_g, geom=geom.copy(), B_e=B_e, move=np.asarray(move, float).copy(), g=g)
        geom = np.asarray(new_geom, float) / 0.52917721067
    blob = {k: np.array(v) for k, v in rec.items()}
    blob["elements"] = np.array(elems); blob["H0"] = H0; blob["xyz0"] = xyz0
    np.savez_compressed(os.path.join(GOLD, "c1_calc_move_vector.npz"), **blob)
    print("c1 trace: move norms", [float(np.linalg.norm(m)) for m in rec["move"]], "trust", rec["trust"])
    print("first move row", rec["move"][0][0])


def neb_chain(nimg, natoms, seed):
    """Synthetic NEB chain (SURVEY §8d, config 3 shape): linear interpolation between two
    jittered endpoints + noise; double-well energies so that up-hill, down-hill and extremum
    tangent branches all occur; gradients consistent with a quadratic model + noise."""
    rng = np.random.default_rng(seed)
    a = synthetic.grid_geometry(natoms, rng).reshape(-1)
    b = a + rng.normal(0.0, 0.6, size=a.size)
    t = np.linspace(0.0, 1.0, nimg)
    X = np.stack([(1 - s) * a + s * b for s in t]) + rng.normal(0.0, 0.05, size=(nimg, a.size))
    E = 0.05 * np.sin(2.5 * np.pi * t) + 0.02 * t + rng.normal(0.0, 1e-4, size=nimg)
    G = rng.normal(0.0, 2e-2, size=(nimg, a.size))
    return X, E, G


def gen_neb():
    import tempfile, types
    bn = ref_shim.ref("MEP.pathopt_bneb_force")
    rn = ref_shim.ref("Optimizer.rfo_neb")
    rs = ref_shim.ref("Optimizer.rsirfo")
    trn = ref_shim.ref("Optimizer.trust_radius_neb")
    nimg, natoms = 9, 10
    n = 3 * natoms
    elems = synthetic.elements(natoms)
    tmp = tempfile.mkdtemp() + "/"

    class Helper(rn.OptimizationAlgorithm):
        def optimize(self, *a, **k):
            pass
    helper = Helper()
    calc = bn.CaluculationBNEB()
    tr = trn.TR_NEB(NEB_FOLDER_DIRECTORY=tmp, fix_init_edge=False, fix_end_edge=False, apply_convergence_criteria=False)
    opts = []
    for num in range(nimg):
        with quiet():
            if num == 0 or num == nimg - 1:
                o = rs.RSIRFO(method="rsirfo_block_fsb", saddle_order=0, trust_radius=0.5)
            else:
                o = rs.RSIRFO(method="rsirfo_block_bofill", saddle_order=0, trust_radius=0.2)
                o.switch_NEB_mode()
        opts.append(o)
    rngH = np.random.default_rng(77)
    H = [synthetic.spd_hessian(n, rngH) for _ in range(nimg)]
    rec = {k: [] for k in ("X", "E", "G", "force", "tau", "gamma", "H_after", "rfo_delta", "rfo_move")}
    X, E, G = neb_chain(nimg, natoms, 31)
    prevX = prevG = None
    for it in range(3):
        geoms = X.reshape(nimg, natoms, 3)
        grads = G.reshape(nimg, natoms, 3)
        with quiet():
            force = calc.calc_force(geoms, E, grads, it, elems)
        taus = np.array([calc.get_tau(i).reshape(-1) for i in range(nimg)])
        gammas, deltas = [], []
        for num in range(nimg):
            h0 = H[num].copy()
            with quiet():
                hess = helper._apply_ayala_hessian_update(H[num], num, nimg, geoms, E, grads, calc)
            gammas.append(0.0 if num in (0, nimg - 1) else float(np.sum((hess - h0) * np.outer(taus[num], taus[num])) / max(np.sum(np.outer(taus[num], taus[num]) ** 2), 1e-300)))
            o = opts[num]
            o.set_hessian(hess)
            o.set_bias_hessian(np.zeros((n, n)))
            col = lambda v: v.reshape(-1, 1).copy()
            with quiet():
                if it == 0:
                    mv = o.run(col(X[num]), col(G[num]), None, None, 0.0, 0.0, [], [], col(G[num]), None)
                else:
                    mv = o.run(col(X[num]), col(G[num]), col(prevG[num]), col(prevX[num]), 0.0, 0.0, [], [], col(G[num]), col(prevG[num]))
                mv = helper._limit_step_size(mv, num == 0 or num == nimg - 1)
            deltas.append(np.asarray(mv, float).reshape(natoms, 3))
            H[num] = np.asarray(o.get_hessian(), float).copy()
            o.set_hessian(None); o.set_bias_hessian(None)
        with quiet():
            mvs = tr.TR_calc(geoms, grads, [d.copy() for d in deltas], E, E, None)
        rec["X"].append(X.copy()); rec["E"].append(E.copy()); rec["G"].append(G.copy())
        rec["force"].append(np.asarray(force, float).reshape(nimg, n)); rec["tau"].append(taus)
        rec["gamma"].append(np.array(gammas)); rec["H_after"].append(np.stack(H))
        rec["rfo_delta"].append(np.stack(deltas).reshape(nimg, n)); rec["rfo_move"].append(np.stack([np.asarray(m, float) for m in mvs]).reshape(nimg, n))
        prevX, prevG = X.copy(), G.copy()
        # next point of the chain: move against the RFO vectors, new seeded energies / gradients
        Xn, En, Gn = neb_chain(nimg, natoms, 32 + it)
        X = X - np.stack([np.asarray(m, float) for m in mvs]).reshape(nimg, n)
        E = En; G = G + 0.3 * (Gn - G)
    blob = {k: np.array(v) for k, v in rec.items()}
    blob["H0"] = np.stack([synthetic.spd_hessian(n, np.random.default_rng(77)) for _ in range(1)])
    rngH = np.random.default_rng(77)
    blob["H_init"] = np.stack([synthetic.spd_hessian(n, rngH) for _ in range(nimg)])
    blob["meta"] = np.array([nimg, natoms], np.int64)
    np.savez_compressed(os.path.join(GOLD, "neb_rfo.npz"), **blob)
    print("neb: gammas", np.round(rec["gamma"][0], 4), "move norms", np.round(np.linalg.norm(rec["rfo_move"][1], axis=1), 4))


def gen_lindh():
    """Lindh model Hessian, decomposed parity: diagonal RIC force constants,
    B^T diag(k) B projected, and main() with a zero gradient (its K term vanishes there)."""
    li = ref_shim.ref("ModelHessian.lindh")
    ric = ref_shim.ref("Coordinate.redundant_coordinate")
    ct = ref_shim.ref("Utils.calc_tools")
    blob = {"names": np.array(["aldol_rxn", "s8", "claisen", "grid24"])}
    for name, path in [c for c in PRODUCER_CASES if c[0] in ("aldol_rxn", "s8", "claisen", "grid24")]:
        elems, xyz = producer_geometry(name, path)
        N = len(elems)
        L = li.LindhApproxHessian()
        with quiet():
            B = ric.RedundantInternalCoordinates().B_matrix(xyz)
            L.RIC_variable_num = len(B)
            kd = np.diag(L.guess_lindh_hessian(xyz.copy(), elems)).copy()
            Hbkb = ct.Calculationtools().project_out_hess_tr_and_rot_for_coord(B.T @ np.diag(kd) @ B, elems, xyz.copy(), False)
            try:
                Hmain = np.asarray(li.LindhApproxHessian().main(xyz.copy(), elems, np.zeros((N, 3))), float)
            except Exception as exc:
                print("  main(zero gradient) failed:", exc)
                Hmain = np.full((3 * N, 3 * N), np.nan)
        # main() with a gradient: the internal gradient of the singular solve is recorded as an input
        gq = np.random.default_rng(7 + N).normal(0, 1e-2, size=(N, 3))
        with quiet():
            ig = ric.RedundantInternalCoordinates().cartgrad2RICgrad(gq.reshape(3 * N, 1), B)
            try:
                Hmain_g = np.asarray(li.LindhApproxHessian().main(xyz.copy(), elems, gq.copy()), float)
            except Exception as exc:
                print("  main(gradient) failed:", exc); Hmain_g = np.full((3 * N, 3 * N), np.nan)
        blob[f"{name}/grad"] = gq; blob[f"{name}/int_grad"] = np.asarray(ig, float).ravel(); blob[f"{name}/H_main_g"] = Hmain_g
        print("   main(g): |int_grad| max", np.abs(ig).max(), "|H|", np.linalg.norm(Hmain_g))
        blob[f"{name}/elements"] = np.array(elems); blob[f"{name}/xyz"] = xyz
        blob[f"{name}/kdiag"] = kd; blob[f"{name}/H_bkb"] = np.asarray(Hbkb, float); blob[f"{name}/H_main0"] = Hmain
        print("lindh case", name, "kdiag range", kd.min(), kd.max(), "main0 vs BkB",
              np.linalg.norm(Hmain - Hbkb) / np.linalg.norm(Hbkb))
    np.savez_compressed(os.path.join(GOLD, "lindh.npz"), **blob)


def swart_extra_geometries():
    """Geometries that exercise the near-linear branches of the Swart angle term."""
    out = {}
    # exactly linear along x (cross product zero -> second reference axis) and along a general direction
    out["lin_x"] = (["O", "C", "O", "H"], np.array([[-2.2, 0, 0], [0, 0, 0], [2.2, 0, 0], [4.1, 0, 0.0]]))
    u = np.array([1.0, 2.0, -0.5]); u /= np.linalg.norm(u)
    out["lin_gen"] = (["H", "C", "N"], np.array([-2.0 * u, 0 * u, 2.2 * u]))
    # nearly linear (th1 < tolth, cos < 0) plus a sharp angle (cos > 0.8) at a heavy centre
    out["near_lin"] = (["C", "C", "C", "H", "H"], np.array([[-2.5, 0.1, 0], [0, 0, 0], [2.5, 0.25, 0.1], [0.3, 2.0, 0.2], [0.9, 1.9, 0.3]]))
    out["sharp"] = (["Pd", "H", "H", "P", "Cl"], np.array([[0, 0, 0], [3.0, 0.7, 0], [3.0, -0.7, 0.1], [-4.0, 0.5, 0.3], [0.2, 4.2, -0.3]]))
    rng = np.random.default_rng(5157)
    out["cloud30"] = (synthetic.elements(30), synthetic.grid_geometry(30, rng, spacing=2.4, jitter=0.3))
    return out


def gen_swart():
    """Swart model Hessian (ModelHessian/swart.py) on the producer molecules and on geometries
    that hit the near-linear / sharp-angle branches; raw (unprojected) Hessian recorded as well."""
    sw = ref_shim.ref("ModelHessian.swart")
    cases = [(name,) + producer_geometry(name, path) for name, path in PRODUCER_CASES]
    cases += [(name, e, x) for name, (e, x) in swart_extra_geometries().items()]
    blob = {"names": np.array([c[0] for c in cases])}
    for name, elems, xyz in cases:
        xyz = np.asarray(xyz, float)
        S = sw.SwartApproxHessian()
        with quiet():
            Hp = np.asarray(S.main(xyz.copy(), list(elems), np.zeros((len(elems), 3))), float)
        blob[f"{name}/elements"] = np.array(elems); blob[f"{name}/xyz"] = xyz
        blob[f"{name}/radii"] = S._get_radii_array(list(elems))
        blob[f"{name}/H_raw"] = S.cart_hess.copy(); blob[f"{name}/H"] = Hp
        print("swart case", name, len(elems), "|H|", np.linalg.norm(Hp))
    np.savez_compressed(os.path.join(GOLD, "swart.npz"), **blob)


def gen_ric():
    """Redundant-internal-coordinate helpers (Coordinate/redundant_coordinate.py): all-pairs B matrix,
    partial stretch / bend / torsion rows (incl. the linear and planar special branches), Wilson
    back-transformation B^T H B + K with a supplied RIC gradient, pseudo-inverse gradient transforms."""
    rc = ref_shim.ref("Coordinate.redundant_coordinate")
    bc = ref_shim.ref("Utils.bond_connectivity")
    blob = {"names": np.array(["aldol_rxn", "s8", "grid24"])}
    rng = np.random.default_rng(181818)
    for name, path in [c for c in PRODUCER_CASES if c[0] in ("aldol_rxn", "s8", "grid24")]:
        elems, xyz = producer_geometry(name, path)
        N = len(elems)
        R = rc.RedundantInternalCoordinates()
        with quiet():
            Bm = R.B_matrix(xyz)
            tabs = bc.BondConnectivity().connectivity_table(xyz.copy(), elems)
        M = len(Bm)
        nt = sum(len(t) for t in tabs)
        q = rng.normal(0, 1e-2, size=M)                  # an internal gradient (input)
        hd = np.abs(rng.normal(0.3, 0.1, size=M))
        A = rng.standard_normal((M, 6)); Hric = np.diag(hd) + 0.01 * (A @ A.T)
        with quiet():
            Hc_diag = R.RIChess2carthess(xyz, tabs, np.diag(hd), Bm, q)
            Hc_full = R.RIChess2carthess(xyz, tabs, Hric, Bm, q)
            K = Hc_diag - Bm.T @ np.diag(hd) @ Bm
            gq = R.RICgrad2cartgrad(q, Bm)
        # near-planar dihedrals make K roundoff-sensitive in the reference itself (acos'' near +-1):
        # a second K with those terms left out carries the tight parity check
        def well_conditioned(row):
            import torch
            v = float(rc.TorchDerivatives().dihedral_angle(torch.tensor(xyz[list(row)], dtype=torch.float64)))
            return 0.05 < v < np.pi - 0.05
        dih_wc = [list(r) for r in tabs[2] if well_conditioned(r)]
        tabs_wc = [tabs[0], tabs[1], dih_wc]
        with quiet():
            K_wc = R.RIChess2carthess(xyz, tabs_wc, np.diag(hd), Bm, q) - Bm.T @ np.diag(hd) @ Bm
        blob[f"{name}/dihedrals_wc"] = pad_table(dih_wc, 4); blob[f"{name}/n_dih_wc"] = np.int32(len(dih_wc))
        blob[f"{name}/K_wc"] = K_wc
        labels = []
        for t in tabs:
            for row in t[: 6]:
                labels.append([int(a) + 1 for a in row])
        rows = []
        for lab in labels:
            f = {2: rc.partial_stretch_B_matirx, 3: rc.partial_bend_B_matrix, 4: rc.partial_torsion_B_matrix}[len(lab)]
            rows.append(f(xyz, *lab)[0])
        pB = np.array(rows[: min(len(rows), 5)])
        g = rng.normal(0, 1e-2, size=3 * N)
        ig = rc.calc_int_grad_from_pBmat(g.reshape(-1, 1), pB).ravel()
        cg = rc.calc_cart_grad_from_pBmat(ig.reshape(-1, 1), pB).ravel()
        blob[f"{name}/xyz"] = xyz; blob[f"{name}/elements"] = np.array(elems)
        blob[f"{name}/Bmat"] = Bm; blob[f"{name}/q"] = q; blob[f"{name}/hdiag"] = hd; blob[f"{name}/Hric"] = Hric
        blob[f"{name}/K"] = K; blob[f"{name}/Hc_diag"] = Hc_diag; blob[f"{name}/Hc_full"] = Hc_full; blob[f"{name}/gq"] = gq
        blob[f"{name}/bonds"] = pad_table(tabs[0], 2); blob[f"{name}/angles"] = pad_table(tabs[1], 3)
        blob[f"{name}/dihedrals"] = pad_table(tabs[2], 4)
        blob[f"{name}/counts"] = np.array([len(t) for t in tabs], np.int32)
        lab_arr = np.zeros((len(labels), 4), np.int32)
        for r_, lab in enumerate(labels):
            lab_arr[r_, : len(lab)] = lab
        blob[f"{name}/labels"] = lab_arr; blob[f"{name}/rows"] = np.array(rows)
        blob[f"{name}/pB"] = pB; blob[f"{name}/g"] = g; blob[f"{name}/int_grad"] = ig; blob[f"{name}/cart_grad"] = cg
        print("ric case", name, "M", M, "terms", nt, "|K|", np.linalg.norm(K), "rows", len(rows))
    # special branches of the partial rows: linear bend, planar (phi = 0, pi) torsions
    sp = {
        "bend_linear": (np.array([[-2.0, 0, 0], [0, 0, 0], [2.2, 0, 0.0]]), [1, 2, 3]),
        "tors_trans": (np.array([[1.0, 1.0, 0], [0, 0, 0], [2.5, 0, 0], [3.4, -1.1, 0.0]]), [1, 2, 3, 4]),
        "tors_cis": (np.array([[1.0, 1.0, 0], [0, 0, 0], [2.5, 0, 0], [3.4, 1.2, 0.0]]), [1, 2, 3, 4]),
        "tors_gen": (np.array([[1.0, 1.0, 0.2], [0, 0, 0], [2.5, 0, 0], [3.4, 0.3, -1.2]]), [1, 2, 3, 4]),
        "tors_neg": (np.array([[1.0, 1.0, 0.2], [0, 0, 0], [2.5, 0, 0], [3.4, 0.3, 1.2]]), [4, 3, 2, 1]),
    }
    blob["special_names"] = np.array(list(sp))
    for name, (x, lab) in sp.items():
        f = {3: rc.partial_bend_B_matrix, 4: rc.partial_torsion_B_matrix}[len(lab)]
        blob[f"special/{name}/xyz"] = x
        blob[f"special/{name}/labels"] = np.array(lab + [0] * (4 - len(lab)), np.int32)
        blob[f"special/{name}/row"] = f(x, *lab)[0]
    np.savez_compressed(os.path.join(GOLD, "ric.npz"), **blob)


def gen_keep():
    """Restraint bias potentials (Potential/keep_potential.py, keep_angle_potential.py): E, gradient and
    Hessian by torch.func as the aggregator computes them (Potential/potential.py:127-137)."""
    import torch
    kp = ref_shim.ref("Potential.keep_potential")
    ka = ref_shim.ref("Potential.keep_angle_potential")
    kd = ref_shim.ref("Potential.keep_dihedral_angle_potential")
    elems, xyz = read_xyz(os.path.join(ref_shim.REF_ROOT, "test/aldol_rxn.xyz"))
    N = len(elems)
    lin = xyz.copy(); lin[1] = lin[0] + np.array([2.0, 0, 0]); lin[2] = lin[0] + np.array([4.3, 1e-5, 0])     # i=0? see cases
    cases = [
        ("keep_1_5", 1, [0], [4], 0.4, 1.6, xyz),
        ("keep_3_11", 1, [2], [10], 1.2, 3.1, xyz),
        ("keepv2", 2, [0, 1, 2, 3], [4, 5, 6, 7, 8], 0.7, 2.5, xyz),
        ("angle_gen", 3, [1, 0, 2], [], 0.3, 109.5, xyz),
        ("angle_to_180", 3, [5, 4, 9], [], 0.2, 180.0, xyz),
        ("angle_to_0", 3, [5, 4, 9], [], 0.2, 0.0, xyz),
        ("angle_near_pi", 3, [1, 0, 2], [], 0.5, 120.0, lin),      # atoms 1-0-2 almost linear: Taylor branch at pi
        ("angle_near_pi_180", 3, [1, 0, 2], [], 0.5, 180.0, lin),
        # dihedrals (kind 4): p in degrees here; the blob stores the radians the reference derived from it.
        # "_cfg" cases call calc_energy without parameters (configured angle -> float32 deg2rad)
        ("dihedral_gen", 4, [1, 0, 4, 5], [], 0.3, 60.0, xyz),
        ("dihedral_wrap", 4, [6, 4, 5, 7], [], 0.25, 175.0, xyz),
        ("dihedral_neg", 4, [2, 0, 4, 9], [], 0.4, -123.4, xyz),
        ("dihedral_cfg", 4, [1, 0, 4, 5], [], 0.3, 60.1, xyz),
    ]
    lin[1] = lin[0] + np.array([-2.0, 0, 0]); lin[2] = lin[0] + np.array([2.3, 4e-4, 0])
    blob = {"names": np.array([c[0] for c in cases])}
    for name, kind, f1, f2, k, p, geom in cases:
        g = torch.tensor(geom, dtype=torch.float64)
        par = torch.tensor([k, p], dtype=torch.float64)
        if kind == 1:
            pot = kp.StructKeepPotential(keep_pot_spring_const=k, keep_pot_distance=p, keep_pot_atom_pairs=[f1[0] + 1, f2[0] + 1])
        elif kind == 2:
            pot = kp.StructKeepPotentialv2(keep_pot_v2_spring_const=k, keep_pot_v2_distance=p,
                                           keep_pot_v2_fragm1=[a + 1 for a in f1], keep_pot_v2_fragm2=[a + 1 for a in f2])
        elif kind == 3:
            pot = ka.StructKeepAnglePotential(keep_angle_atom_pairs=[a + 1 for a in f1], keep_angle_spring_const=k, keep_angle_angle=p)
        else:
            pot = kd.StructKeepDihedralAnglePotential(keep_dihedral_angle_atom_pairs=[a + 1 for a in f1],
                                                      keep_dihedral_angle_spring_const=k, keep_dihedral_angle_angle=p)
            if name.endswith("_cfg"):
                par = []                                                      # configured angle: float32 deg2rad
                blob[f"{name}/phi0"] = float(torch.deg2rad(torch.tensor(p)))
            else:
                blob[f"{name}/phi0"] = float(torch.deg2rad(par[1]))           # float64, as the aggregator passes it
        E = float(pot.calc_energy(g, par))
        gr = torch.func.jacrev(pot.calc_energy, argnums=0)(g, par).numpy()
        H = torch.func.hessian(pot.calc_energy, argnums=0)(g, par).reshape(3 * N, 3 * N).numpy()
        blob[f"{name}/xyz"] = geom; blob[f"{name}/kind"] = np.int32(kind)
        blob[f"{name}/f1"] = np.array(f1, np.int32); blob[f"{name}/f2"] = np.array(f2, np.int32)
        blob[f"{name}/kp"] = np.array([k, p]); blob[f"{name}/E"] = E; blob[f"{name}/g"] = gr; blob[f"{name}/H"] = H
        print("keep case", name, "E", E, "|g|", np.linalg.norm(gr), "|H|", np.linalg.norm(H))
    np.savez_compressed(os.path.join(GOLD, "keep.npz"), **blob)


def gen_bias2():
    """LJ repulsive (scale / value), anharmonic keep, fragment well and out-of-plane-angle potentials
    (Potential/LJ_repulsive_potential.py, anharmonic_keep_potential.py, switching_potential.py,
    keep_outofplain_angle_potential.py): E, gradient, Hessian by torch.func as the aggregator computes them."""
    import torch
    lj = ref_shim.ref("Potential.LJ_repulsive_potential")
    an = ref_shim.ref("Potential.anharmonic_keep_potential")
    sw = ref_shim.ref("Potential.switching_potential")
    oop = ref_shim.ref("Potential.keep_outofplain_angle_potential")
    ang = ref_shim.ref("Potential.keep_angle_potential")
    dih = ref_shim.ref("Potential.keep_dihedral_angle_potential")
    elems, xyz = read_xyz(os.path.join(ref_shim.REF_ROOT, "test/aldol_rxn.xyz"))
    N = len(elems)
    cen = lambda f: xyz[[a - 1 for a in f]].mean(axis=0)
    f1w, f2w = [1, 2, 3, 4], [5, 6, 7, 8, 9]
    dw = np.linalg.norm(cen(f1w) - cen(f2w)) * 0.52917721067      # centroid distance in Angstrom
    flat = xyz.copy(); flat[2] = flat[0] + 1.7 * (flat[1] - flat[0])          # atoms 1, 2, 3 collinear: undefined plane
    def lims_for(dists_bohr):     # [a, b, c, d] in Angstrom: the middle target inside the flat part, the others on the two switching walls
        ds = np.sort(np.asarray(dists_bohr)) * 0.52917721067
        return [float(ds[0] - 0.2), float(ds[0] + 0.15), float(ds[2] - 0.15), float(ds[2] + 0.2)]
    flat2 = xyz.copy(); flat2[2] = flat2[1] + 0.8 * (flat2[1] - flat2[0])     # 1 - 2 - 3 exactly linear, vertex 2
    perp = np.cross(xyz[1] - xyz[0], [0.3, -0.2, 0.9]); perp /= np.linalg.norm(perp)
    near_pi = flat2.copy(); near_pi[2] = near_pi[2] + 4e-4 * np.linalg.norm(flat2[2] - flat2[1]) * perp      # 4e-4 rad off linear
    near_pi2 = flat2.copy(); near_pi2[2] = near_pi2[2] + 2e-4 * np.linalg.norm(flat2[2] - flat2[1]) * perp    # (exactly linear: u clamps, all derivatives 0)
    near_0 = xyz.copy(); near_0[2] = near_0[1] + 1.3 * (near_0[0] - near_0[1]) + 5e-4 * 1.3 * np.linalg.norm(near_0[0] - near_0[1]) * perp
    cases = [
        ("lj_scale", dict(cls="lj_scale", well=1.3, dist=0.9, f1=[1, 2, 3], f2=[5, 6, 10, 11]), xyz),
        ("lj_value", dict(cls="lj_value", well=2.5, dist=3.2, f1=[1, 4], f2=[6, 7, 8]), xyz),
        ("anharmonic", dict(cls="anh", k=0.35, depth=0.12, pair=[1, 5], dist=1.7), xyz),
        ("anharmonic_far", dict(cls="anh", k=0.9, depth=0.05, pair=[3, 11], dist=2.4), xyz),
        ("well_inside", dict(cls="well", wall=40.0, f1=f1w, f2=f2w, lim=[dw - 1.5, dw - 0.8, dw + 0.9, dw + 1.6]), xyz),
        ("well_short_switch", dict(cls="well", wall=40.0, f1=f1w, f2=f2w, lim=[dw - 0.3, dw + 0.4, dw + 1.9, dw + 2.6]), xyz),
        ("well_short_linear", dict(cls="well", wall=25.0, f1=f1w, f2=f2w, lim=[dw + 0.2, dw + 0.9, dw + 1.9, dw + 2.6]), xyz),
        ("well_long_switch", dict(cls="well", wall=40.0, f1=f1w, f2=f2w, lim=[dw - 2.6, dw - 1.9, dw - 0.4, dw + 0.3]), xyz),
        ("well_long_linear", dict(cls="well", wall=25.0, f1=f1w, f2=f2w, lim=[dw - 2.6, dw - 1.9, dw - 0.9, dw - 0.2]), xyz),
        ("oop_gen", dict(cls="oop", k=0.3, atoms=[1, 2, 3, 5], angle=12.0), xyz),
        ("oop_neg", dict(cls="oop", k=0.2, atoms=[5, 6, 7, 10], angle=-25.0), xyz),
        ("oop_undefined", dict(cls="oop", k=0.3, atoms=[1, 5, 2, 3], angle=10.0), flat),
        # fragment-centroid (v2) restraints: general angle, exactly linear theta0 = 180 in its expansion region and in its
        # regular region, theta0 = 0, the quadratic continuations near pi and near 0, dihedral and out-of-plane angle
        ("angle_v2_gen", dict(cls="ang2", k=0.3, f=[[1, 2], [5], [6, 7, 8]], angle=100.0), xyz),
        ("angle_v2_lin180", dict(cls="ang2", k=0.25, f=[[1], [2], [3]], angle=180.0), near_pi2),
        ("angle_v2_lin180_bent", dict(cls="ang2", k=0.25, f=[[1, 4], [2], [3, 9]], angle=180.0), xyz),
        ("angle_v2_zero", dict(cls="ang2", k=0.2, f=[[1], [5, 6], [3]], angle=0.0), xyz),
        ("angle_v2_near_pi", dict(cls="ang2", k=0.3, f=[[1], [2], [3]], angle=120.0), near_pi),
        ("angle_v2_near_0", dict(cls="ang2", k=0.3, f=[[1], [2], [3]], angle=30.0), near_0),
        ("dihedral_v2", dict(cls="dih2", k=0.2, f=[[1, 2], [3], [5, 6], [7]], angle=35.0), xyz),
        ("oop_v2", dict(cls="oop2", k=0.3, f=[[1], [2, 3], [5], [6, 7]], angle=12.0), xyz),
        # the other wells of switching_potential.py: |y| of three atoms against a wall pair, three atoms around a fixed
        # point (stored as float32 by the reference), three atoms around the centroid of a centre fragment; the limits
        # are placed around the middle distance so that the targets fall into different regions of the well
        ("wall_well", dict(cls="wallw", wall=30.0, direction="y", targets=[1, 4, 7], lim=lims_for(np.abs(xyz[[0, 3, 6], 1]))), xyz),
        ("vp_well", dict(cls="vpw", wall=35.0, point=[0.3, -1.2, 0.7123456789], targets=[2, 5, 9],
                         lim=lims_for(np.linalg.norm(xyz[[1, 4, 8]] - np.float32([0.3, -1.2, 0.7123456789]).astype(float), axis=1))), xyz),
        ("around_well", dict(cls="arw", wall=25.0, center=[1, 2, 3], targets=[6, 8, 10],
                             lim=lims_for(np.linalg.norm(xyz[[5, 7, 9]] - xyz[[0, 1, 2]].mean(axis=0), axis=1))), xyz),
    ]
    blob = {"names": np.array([c[0] for c in cases]), "elements": np.array(elems)}
    for name, cfg, geom in cases:
        g = torch.tensor(geom, dtype=torch.float64)
        par = []
        if cfg["cls"] == "lj_scale":
            pot = lj.LJRepulsivePotentialScale(repulsive_potential_well_scale=cfg["well"], repulsive_potential_dist_scale=cfg["dist"],
                                               repulsive_potential_Fragm_1=cfg["f1"], repulsive_potential_Fragm_2=cfg["f2"], element_list=elems)
        elif cfg["cls"] == "lj_value":
            pot = lj.LJRepulsivePotentialValue(repulsive_potential_well_value=cfg["well"], repulsive_potential_dist_value=cfg["dist"],
                                               repulsive_potential_Fragm_1=cfg["f1"], repulsive_potential_Fragm_2=cfg["f2"], element_list=elems)
        elif cfg["cls"] == "anh":
            pot = an.StructAnharmonicKeepPotential(anharmonic_keep_pot_spring_const=cfg["k"], anharmonic_keep_pot_potential_well_depth=cfg["depth"],
                                                   anharmonic_keep_pot_atom_pairs=cfg["pair"], anharmonic_keep_pot_distance=cfg["dist"])
        elif cfg["cls"] == "well":
            pot = sw.WellPotential(well_pot_wall_energy=cfg["wall"], well_pot_fragm_1=cfg["f1"], well_pot_fragm_2=cfg["f2"], well_pot_limit_dist=cfg["lim"])
        elif cfg["cls"] == "wallw":
            pot = sw.WellPotentialWall(wall_well_pot_wall_energy=cfg["wall"], wall_well_pot_direction=cfg["direction"],
                                       wall_well_pot_limit_dist=cfg["lim"], wall_well_pot_target=cfg["targets"])
        elif cfg["cls"] == "vpw":
            pot = sw.WellPotentialVP(void_point_well_pot_wall_energy=cfg["wall"], void_point_well_pot_coordinate=list(cfg["point"]),
                                     void_point_well_pot_limit_dist=cfg["lim"], void_point_well_pot_target=cfg["targets"])
        elif cfg["cls"] == "arw":
            pot = sw.WellPotentialAround(around_well_pot_wall_energy=cfg["wall"], around_well_pot_center=cfg["center"],
                                         around_well_pot_limit_dist=cfg["lim"], around_well_pot_target=cfg["targets"])
        elif cfg["cls"] == "ang2":
            f = cfg["f"]
            pot = ang.StructKeepAnglePotentialv2(keep_angle_v2_spring_const=cfg["k"], keep_angle_v2_angle=cfg["angle"],
                                                 keep_angle_v2_fragm1=f[0], keep_angle_v2_fragm2=f[1], keep_angle_v2_fragm3=f[2])
            par = torch.tensor([cfg["k"], cfg["angle"]], dtype=torch.float64)     # process_keep_angle_v2, potential.py:328-344
        elif cfg["cls"] == "dih2":
            f = cfg["f"]
            pot = dih.StructKeepDihedralAnglePotentialv2(keep_dihedral_angle_v2_spring_const=cfg["k"], keep_dihedral_angle_v2_angle=cfg["angle"],
                                                         keep_dihedral_angle_v2_fragm1=f[0], keep_dihedral_angle_v2_fragm2=f[1],
                                                         keep_dihedral_angle_v2_fragm3=f[2], keep_dihedral_angle_v2_fragm4=f[3])
            par = torch.tensor([cfg["k"], cfg["angle"]], dtype=torch.float64)
        elif cfg["cls"] == "oop2":
            f = cfg["f"]
            pot = oop.StructKeepOutofPlainAnglePotentialv2(keep_out_of_plain_angle_v2_spring_const=cfg["k"], keep_out_of_plain_angle_v2_angle=cfg["angle"],
                                                           keep_out_of_plain_angle_v2_fragm1=f[0], keep_out_of_plain_angle_v2_fragm2=f[1],
                                                           keep_out_of_plain_angle_v2_fragm3=f[2], keep_out_of_plain_angle_v2_fragm4=f[3])
            par = torch.tensor([cfg["k"], cfg["angle"]], dtype=torch.float64)
        else:
            pot = oop.StructKeepOutofPlainAnglePotential(keep_out_of_plain_angle_spring_const=cfg["k"],
                                                         keep_out_of_plain_angle_atom_pairs=cfg["atoms"], keep_out_of_plain_angle_angle=cfg["angle"])
            par = torch.tensor([cfg["k"], cfg["angle"]], dtype=torch.float64)     # as the aggregator passes it (potential.py:802)
        E = float(pot.calc_energy(g, par))
        gr = torch.func.jacrev(pot.calc_energy, argnums=0)(g, par).numpy()
        H = torch.func.hessian(pot.calc_energy, argnums=0)(g, par).reshape(3 * N, 3 * N).numpy()
        blob[f"{name}/xyz"] = geom; blob[f"{name}/E"] = E; blob[f"{name}/g"] = gr; blob[f"{name}/H"] = H
        blob[f"{name}/cfg"] = np.array(json.dumps(cfg))
        print("bias2 case", name, "E", E, "|g|", np.linalg.norm(gr), "|H|", np.linalg.norm(H))
    np.savez_compressed(os.path.join(GOLD, "bias2.npz"), **blob)


def gen_fire():
    """FIRE optimizer of the NEB driver (Optimizer/fire_neb.py): 8-iteration trace on a synthetic chain with
    a quadratic force field, (dt, a, n_reset) schedule included."""
    import tempfile, types
    fn = ref_shim.ref("Optimizer.fire_neb")
    nimg, natoms = 9, 10
    X, E, G = neb_chain(nimg, natoms, 41)
    X = X.reshape(nimg, natoms, 3)
    cfg = types.SimpleNamespace(dt=0.5, a=0.10, n_reset=0, FIRE_N_accelerate=2, FIRE_f_inc=1.10, FIRE_f_accelerate=0.99,
                                FIRE_f_decelerate=0.5, FIRE_a_start=0.1, FIRE_dt_max=3.0,
                                NEB_FOLDER_DIRECTORY=tempfile.mkdtemp() + "/", fix_init_edge=False, fix_end_edge=False,
                                apply_convergence_criteria=False, bohr2angstroms=0.52917721067)
    with quiet():
        opt = fn.FIREOptimizer(cfg)
    X0 = X.copy()
    K = 0.05 + 0.02 * np.random.default_rng(3).random((nimg, natoms, 1))
    force = lambda x: -K * (x - X0 * 0.97) * 0.2
    V = np.zeros_like(X); Vp = np.zeros_like(X)
    rec = {k: [] for k in ("X", "F", "V", "Vprev", "move", "state", "have_prev")}
    geom = X.copy()
    for it in range(8):
        F = force(geom) * (-1.0 if it == 5 else 1.0)      # iteration 5: negative power -> reset branch
        pre = Vp if it > 0 else []
        rec["X"].append(geom.copy()); rec["F"].append(F.copy()); rec["V"].append(V.copy()); rec["Vprev"].append(Vp.copy())
        rec["have_prev"].append(int(it > 0))
        # the reference computes total_velocity internally and returns only the new geometry: re-derive what
        # the driver carries over (NEB.execute keeps total_velocity from the optimizer's attribute-free maths)
        dt_before, a_before = opt.dt, opt.a
        with quiet():
            new_geom_ang = opt.optimize(geom, F, pre, it, V, [], E, E, geom)
        move = np.asarray(new_geom_ang) / cfg.bohr2angstroms - geom
        rec["move"].append(move.copy()); rec["state"].append([opt.dt, opt.a, opt.n_reset])
        # velocity bookkeeping as in the reference driver: restate the velocity update with the recorded state
        fnm = np.linalg.norm(F, axis=2, keepdims=True); vnm = np.linalg.norm(V, axis=2, keepdims=True)
        with np.errstate(all="ignore"):
            blend = (1.0 - a_before) * V + a_before * (vnm / fnm) * F
        vneb = np.where(fnm > 1e-10, blend, V)
        if opt.n_reset == 0:
            vneb = vneb * 0
        Vnew = vneb + opt.dt * F
        Vp = V.copy() if False else Vnew.copy() * 0 + (Vnew if it == 0 else Vnew)
        Vp = Vnew.copy(); V = Vnew.copy()
        geom = geom + move
    blob = {k: np.array(v) for k, v in rec.items()}
    blob["cfg"] = np.array([cfg.dt, 0.10, 0, cfg.FIRE_N_accelerate, cfg.FIRE_f_inc, cfg.FIRE_f_decelerate, cfg.FIRE_a_start, cfg.FIRE_dt_max])
    print("fire: states", [tuple(np.round(s_, 4)) for s_ in rec["state"]])
    np.savez_compressed(os.path.join(GOLD, "fire_neb.npz"), **blob)


def gen_post():
    """Kabsch alignment and the convergence test (Utils/calc_tools.py:412-425, optimization.py:1244-1289)."""
    import types
    ct = ref_shim.ref("Utils.calc_tools")
    # optimization.py pulls plotting / engine packages that are absent offline: stub whatever is missing
    from unittest.mock import MagicMock
    for _ in range(60):
        try:
            opt = ref_shim.ref("optimization")
            break
        except ModuleNotFoundError as exc:
            parts = exc.name.split(".")
            for k in range(1, len(parts) + 1):
                sys.modules.setdefault(".".join(parts[:k]), MagicMock())
    rng = np.random.default_rng(9090)
    blob = {}
    Ps, Qs, Pa, Qa = [], [], [], []
    for case in range(12):
        N = [5, 11, 24, 30][case % 4]
        Q = synthetic.grid_geometry(N, rng)
        th = rng.normal(size=3); th /= np.linalg.norm(th); ang = rng.uniform(0.2, 3.0)
        K = np.array([[0, -th[2], th[1]], [th[2], 0, -th[0]], [-th[1], th[0], 0]])
        R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
        P = (R @ Q.T).T + rng.normal(0, 0.05, Q.shape) + rng.normal(0, 1.0, 3)
        if case % 5 == 4:
            P = P * np.array([1.0, 1.0, -1.0])        # mirror image: exercises the det < 0 branch
        P0, Q0 = P.copy(), Q.copy()
        with quiet():
            Pr, Qr = ct.Calculationtools().kabsch_algorithm(P, Q)
        pad = lambda a: np.pad(a, ((0, 30 - len(a)), (0, 0)))
        Ps.append(pad(P0)); Qs.append(pad(Q0)); Pa.append(pad(np.asarray(Pr))); Qa.append(pad(np.asarray(Qr)))
    blob["kabsch/natoms"] = np.array([[5, 11, 24, 30][c % 4] for c in range(12)], np.int32)
    blob["kabsch/P"] = np.array(Ps); blob["kabsch/Q"] = np.array(Qs)
    blob["kabsch/P_aligned"] = np.array(Pa); blob["kabsch/Q_centred"] = np.array(Qa)
    # convergence test: gradients / displacements around the thresholds
    chk = opt.ConvergenceChecker.__new__(opt.ConvergenceChecker)
    rows = []
    for case in range(40):
        thr = dict(MAX_FORCE_THRESHOLD=3e-4, RMS_FORCE_THRESHOLD=2e-4, MAX_DISPLACEMENT_THRESHOLD=1.5e-3,
                   RMS_DISPLACEMENT_THRESHOLD=1.0e-3)
        chk.config = types.SimpleNamespace(**thr)
        n = 36
        g = rng.normal(0, 10 ** rng.uniform(-5.5, -3.0), n); d = rng.normal(0, 10 ** rng.uniform(-4.5, -2.5), n)
        g[rng.integers(0, n, 5)] = 0.0; d[rng.integers(0, n, 5)] = 1e-11        # entries the rms filter drops
        state = types.SimpleNamespace(effective_gradient=g.reshape(-1, 3))
        ok, mdt, rdt = chk.check_convergence(state, d.reshape(-1, 3), [])
        rows.append((g, d, ok, mdt, rdt))
    blob["conv/grad"] = np.array([r[0] for r in rows]); blob["conv/disp"] = np.array([r[1] for r in rows])
    blob["conv/ok"] = np.array([r[2] for r in rows], np.int32)
    blob["conv/max_disp_thr"] = np.array([r[3] for r in rows]); blob["conv/rms_disp_thr"] = np.array([r[4] for r in rows])
    blob["conv/thresholds"] = np.array([3e-4, 2e-4, 1.5e-3, 1.0e-3])
    print("post: kabsch cases", len(Ps), "convergence cases", len(rows), "converged", int(blob["conv/ok"].sum()))
    np.savez_compressed(os.path.join(GOLD, "post.npz"), **blob)


RSPRFO_CASES = [
    # (name, method, saddle_order, natoms, nsteps, bias, seed)
    ("prfo_bofill_ts_n36", "rsprfo_bofill", 1, 12, 6, False, 1),
    ("prfo_blockbofill_ts_n33", "rsprfo_block_bofill", 1, 11, 6, True, 2),
    ("prfo_fsb_min_n30", "rsprfo_fsb", 0, 10, 5, False, 3),
    ("prfo_bofill_ts2_n30", "rsprfo_bofill", 2, 10, 5, False, 4),
    ("prfo_bofill_ts_n150", "rsprfo_bofill", 1, 50, 4, False, 5),
    ("prfo_msp_ts_n24", "rsprfo_msp", 1, 8, 6, True, 6),
]


def run_rsprfo_case(case):
    name, method, so, natoms, nsteps, bias, seed = case
    rp = ref_shim.ref("Optimizer.rsprfo")
    rng = np.random.default_rng(515100 + seed)
    n = 3 * natoms
    x0 = synthetic.grid_geometry(natoms, rng).reshape(-1)
    H0 = synthetic.spd_hessian(n, rng, neg_lowest=so > 0)
    if so > 1:
        w, V = np.linalg.eigh(H0); w[1] = -0.02
        H0 = (V * w) @ V.T; H0 = 0.5 * (H0 + H0.T)
    E = rng.standard_normal((n, n))
    Ht = H0 + 0.05 * (E + E.T) / np.sqrt(n)
    g0 = rng.normal(0.0, 1e-2, size=n)
    Hb = np.zeros((n, n))
    if bias:
        Bm = rng.standard_normal((n, 4)); Hb = 0.02 * (Bm @ Bm.T)
    pes = QuadraticPES(x0, g0, Ht, Hb, rng)
    with quiet():
        opt = rp.EnhancedRSPRFO(method=method, saddle_order=so, element_list=["C"] * natoms,
                                trust_radius_max=(0.3 if so > 0 else 0.5), trust_radius_min=0.01)
        opt.set_hessian(H0.copy()); opt.set_bias_hessian(Hb.copy())
    rec = {k: [] for k in ("x", "Bg", "Be", "move", "H_after", "trust", "pred")}
    x = x0.copy(); x_prev = Bg_prev = mv_prev = None
    for k in range(nsteps):
        e, g = pes.raw(x); eb, gb = pes.bias(x)
        Be, Bg = e + eb, g + gb
        col = lambda a: a.reshape(-1, 1).copy()
        with quiet():
            if x_prev is None:
                mv = opt.run(col(x), col(Bg), [], [], Be, 0.0, [], col(x0), col(g), [])
            else:
                mv = opt.run(col(x), col(Bg), col(Bg_prev), col(x_prev), Be, 0.0, col(mv_prev), col(x0), col(g), [])
        mv = np.asarray(mv, float).ravel()
        rec["x"].append(x.copy()); rec["Bg"].append(Bg.copy()); rec["Be"].append(Be); rec["move"].append(mv)
        rec["H_after"].append(np.array(opt.hessian, float)); rec["trust"].append(float(opt.trust_radius))
        rec["pred"].append(float(opt.predicted_energy_changes[-1]))
        x_prev, Bg_prev, mv_prev = x.copy(), Bg.copy(), mv.copy()
        x = x - mv
    out = {f"{name}/{k}": np.array(v) for k, v in rec.items()}
    out[f"{name}/H0"] = H0; out[f"{name}/Hb"] = Hb
    out[f"{name}/meta"] = np.array([so, natoms, nsteps, int(bias)], np.int64)
    out[f"{name}/method"] = np.array(method)
    return out


def gen_rsprfo():
    blob = {}
    for case in RSPRFO_CASES:
        blob.update(run_rsprfo_case(case))
        print("rsprfo case", case[0])
    blob["names"] = np.array([c[0] for c in RSPRFO_CASES])
    np.savez_compressed(os.path.join(GOLD, "rsprfo_traces.npz"), **blob)


RSPRFO_REJECT_CASES = [
    # (name, method, saddle_order, natoms, nsteps, spike_step, big_modes, seed)
    # spike_step: the gradient fed at that step makes s.y tiny and ||y|| ~ 1, so the BFGS term y y^T / (s.y) pushes the
    # updated spectrum beyond 1e6 -> EnhancedRSPRFO.update_hessian reverts (rsprfo.py:1242-1250)
    ("prfo_reject_bfgs_n30", "rsprfo_bfgs", 1, 10, 5, 2, 0, 11),
    ("prfo_reject_sr1_n24", "rsprfo_sr1", 1, 8, 5, 3, 0, 12),
    # big_modes: six modes at 6e5 -> ||H||_F = 1.47e6 > 1e6 but max |lambda| = 6e5: accepted (only the exact spectrum decides)
    ("prfo_bigmodes_accept_n30", "rsprfo_bofill", 1, 10, 4, -1, 6, 13),
]


def run_rsprfo_reject_case(case):
    name, method, so, natoms, nsteps, spike, big, seed = case
    rp = ref_shim.ref("Optimizer.rsprfo")
    rng = np.random.default_rng(616100 + seed)
    n = 3 * natoms
    x0 = synthetic.grid_geometry(natoms, rng).reshape(-1)
    H0 = synthetic.spd_hessian(n, rng, neg_lowest=so > 0)
    if big:
        Q, _ = np.linalg.qr(rng.standard_normal((n, big)))
        H0 = H0 + 6e5 * (Q @ Q.T); H0 = 0.5 * (H0 + H0.T)
    E = rng.standard_normal((n, n))
    Ht = H0 + 0.05 * (E + E.T) / np.sqrt(n)
    g0 = rng.normal(0.0, 1e-2, size=n)
    Hb = np.zeros((n, n))
    pes = QuadraticPES(x0, g0, Ht, Hb, rng)
    with quiet():
        opt = rp.EnhancedRSPRFO(method=method, saddle_order=so, element_list=["C"] * natoms,
                                trust_radius_max=0.3, trust_radius_min=0.01)
        opt.set_hessian(H0.copy()); opt.set_bias_hessian(Hb.copy())
    rec = {k: [] for k in ("x", "Bg", "Be", "move", "H_after", "trust", "pred")}
    x = x0.copy(); x_prev = Bg_prev = mv_prev = None
    col = lambda a: a.reshape(-1, 1).copy()
    for k in range(nsteps):
        e, g = pes.raw(x)
        if k == spike:
            s = x - x_prev
            u = rng.standard_normal(n); u -= (u @ s) / (s @ s) * s; u /= np.linalg.norm(u)
            base = np.asarray(opt.hessian, float) @ s if "sr1" in method else 0.0    # SR1: (y - H s).s = 1e-8 instead
            g = Bg_prev + base + u + 1e-8 * s / (s @ s)   # s.y = 1e-8, ||y|| ~ 1
        Be, Bg = e, g
        with quiet():
            if x_prev is None:
                mv = opt.run(col(x), col(Bg), [], [], Be, 0.0, [], col(x0), col(g), [])
            else:
                mv = opt.run(col(x), col(Bg), col(Bg_prev), col(x_prev), Be, 0.0, col(mv_prev), col(x0), col(g), [])
        mv = np.asarray(mv, float).ravel()
        rec["x"].append(x.copy()); rec["Bg"].append(Bg.copy()); rec["Be"].append(Be); rec["move"].append(mv)
        rec["H_after"].append(np.array(opt.hessian, float)); rec["trust"].append(float(opt.trust_radius))
        rec["pred"].append(float(opt.predicted_energy_changes[-1]))
        x_prev, Bg_prev, mv_prev = x.copy(), Bg.copy(), mv.copy()
        x = x - mv
    if spike >= 0:   # the spike step must have left the Hessian untouched
        assert np.array_equal(rec["H_after"][spike], rec["H_after"][spike - 1]), name
    out = {f"{name}/{k}": np.array(v) for k, v in rec.items()}
    out[f"{name}/H0"] = H0; out[f"{name}/Hb"] = Hb
    out[f"{name}/meta"] = np.array([so, natoms, nsteps, spike], np.int64)
    out[f"{name}/method"] = np.array(method)
    return out


def gen_rsprfo_reject():
    blob = {}
    for case in RSPRFO_REJECT_CASES:
        blob.update(run_rsprfo_reject_case(case))
        print("rsprfo reject case", case[0])
    blob["names"] = np.array([c[0] for c in RSPRFO_REJECT_CASES])
    np.savez_compressed(os.path.join(GOLD, "rsprfo_reject.npz"), **blob)


MODELHESS_EXTRA = {   # exactly linear and near-linear molecules: the skip / damping branches of the D3 variants
    "co2": (["O", "C", "O"], np.array([[-2.2, 0.0, 0.0], [0.0, 0.0, 0.0], [2.2, 0.0, 0.0]])),
    "hcch": (["H", "C", "C", "H"], np.array([[-3.15, 0.0, 0.0], [-1.14, 0.0, 0.0], [1.14, 0.0, 0.0], [3.15, 0.0, 0.0]])),
    "hccf_bent": (["H", "C", "C", "F"], np.array([[-3.15, 0.03, 0.0], [-1.14, 0.0, 0.0], [1.14, 0.0, 0.01], [3.5, -0.02, 0.0]])),
    "h2o2": (["O", "O", "H", "H"], np.array([[1.6, 0.0, -4.0], [1.6, 0.46, -2.64], [2.43, 0.05, -2.32], [0.79, -0.52, -4.02]]) / 0.52917721067),
}
MODELHESS_TYPES = ["fischerd3old", "fischerd3", "fischerts", "fischerclip", "fischerd3oldtsclip", "fischerd3clip", "fischersr",
                   "fischerd3oldtssrclip"]


def gen_modelhess_d3():
    """fischerd3old / fischerd3 and the ts / clip modifiers through the reference's own dispatcher
    (ApproxHessian.main, ModelHessian/approx_hessian.py:30-112)."""
    ah = ref_shim.ref("ModelHessian.approx_hessian")
    cases = [(n, *producer_geometry(n, p)) for n, p in PRODUCER_CASES] + [(n, e, x) for n, (e, x) in MODELHESS_EXTRA.items()]
    blob = {"names": np.array([c[0] for c in cases]), "types": np.array(MODELHESS_TYPES)}
    for name, elems, xyz in cases:
        N = len(elems)
        blob[f"{name}/elements"] = np.array(elems)
        blob[f"{name}/xyz"] = np.asarray(xyz, float)
        for t in MODELHESS_TYPES:
            with quiet():
                H = ah.ApproxHessian().main(np.asarray(xyz, float).copy(), list(elems), np.zeros((N, 3)), t)
            blob[f"{name}/{t}"] = np.asarray(H, float)
        print("modelhess_d3 case", name, N, "atoms")
    np.savez_compressed(os.path.join(GOLD, "modelhess_d3.npz"), **blob)



def gen_redistribute():
    """distribute_geometry (Interpolation/linear_interpolation.py:308) on NEB chains: the config-3 shape, an unevenly
    spaced chain, a chain with a repeated image (zero-length segment), three images, a collapsed chain."""
    try:
        LI = ref_shim.ref("Interpolation.linear_interpolation")
    except Exception:   # the module imports scipy-only helpers; fall back to the two functions' own module deps
        raise
    CT = ref_shim.ref("Utils.calc_tools")
    cases = {}
    X, _, _ = neb_chain(64, 30, 11)
    cases["c3_64x30"] = X.reshape(64, 30, 3)
    X, _, _ = neb_chain(17, 9, 12)
    X = X.reshape(17, 9, 3)
    t = np.linspace(0.0, 1.0, 17) ** 2.5          # crowded near the first image
    cases["uneven_17x9"] = np.stack([(1 - u) * X[0] + u * X[-1] for u in t]) + 0.02 * np.random.default_rng(5).normal(size=X.shape)
    Y = X.copy(); Y[6] = Y[5]
    cases["repeat_17x9"] = Y
    cases["three_3x5"] = neb_chain(3, 5, 13)[0].reshape(3, 5, 3)
    cases["collapsed_6x4"] = np.repeat(neb_chain(2, 4, 14)[0].reshape(2, 4, 3)[:1], 6, axis=0)
    blob = {"names": np.array(list(cases))}
    for name, X in cases.items():
        with quiet():
            out = LI.distribute_geometry([x.copy() for x in X])
            pl = CT.calc_path_length_list([x.copy() for x in X])
        blob[f"{name}/X"] = X
        blob[f"{name}/out"] = np.array(out)
        blob[f"{name}/path_length"] = np.array(pl)
    np.savez_compressed(os.path.join(GOLD, "neb_redistribute.npz"), **blob)
    print("neb_redistribute.npz", len(cases), "cases")



# name, method, saddle order, natoms, steps, constraint pairs, row weights, SHAKE targets (or None), bias, seed, converge-at-step
CRSIRFO_CASES = [
    ("c_min_2bonds", "rsirfo_bofill", 0, 8, 4, [(0, 1), (2, 5)], [1.0, 3.5], None, False, 1, None),
    ("c_min_dup_row", "rsirfo_bfgs", 0, 9, 3, [(0, 1), (3, 4), (0, 1)], [1.0, 0.2, 2.0], None, False, 2, None),
    ("c_ts_3bonds", "rsirfo_bofill", 1, 10, 4, [(0, 3), (1, 2), (6, 7)], [1.0, 1.0, 1.0], None, False, 3, None),
    ("c_min_shake", "rsirfo_fsb", 0, 8, 4, [(0, 1), (4, 6)], [1.0, 1.0], "perturbed", False, 4, None),
    ("c_min_bias", "rsirfo_bofill", 0, 8, 3, [(1, 2)], [1.0], None, True, 5, None),
    ("c_min_converged", "rsirfo_bofill", 0, 8, 3, [(0, 1), (2, 3)], [1.0, 1.0], None, False, 6, 1),
    ("c_unconstrained", "rsirfo_bofill", 0, 7, 3, [], [], None, False, 7, None),
]


def run_crsirfo_case(case):
    """CRSIRFO.run (Optimizer/crsirfo.py) with synthetic.DistanceConstraints as the constraint object."""
    name, method, so, natoms, nsteps, pairs, weights, targets, bias, seed, conv_at = case
    cr = ref_shim.ref("Optimizer.crsirfo")
    rng = np.random.default_rng(515100 + seed)
    n = 3 * natoms
    x0 = synthetic.grid_geometry(natoms, rng).reshape(-1)
    H0 = synthetic.spd_hessian(n, rng, neg_lowest=so > 0)
    E = rng.standard_normal((n, n))
    Ht = H0 + 0.05 * (E + E.T) / np.sqrt(n)
    g0 = rng.normal(0.0, 2e-2, size=n)
    Hb = None
    if bias:
        Bm = rng.standard_normal((n, 3))
        Hb = 0.02 * (Bm @ Bm.T)
    pes = QuadraticPES(x0, g0, Ht, np.zeros((n, n)) if Hb is None else Hb, rng)
    tg = None
    if targets == "perturbed":      # targets a little off the current distances: the SHAKE pass moves the atoms by ~1e-2
        tg = [float(np.linalg.norm(x0.reshape(-1, 3)[i] - x0.reshape(-1, 3)[j]) + 0.03 * (1 + q)) for q, (i, j) in enumerate(pairs)]
    cons = synthetic.DistanceConstraints(pairs, targets=tg, weights=weights) if pairs else None
    opt = cr.CRSIRFO(constraints=cons, method=method, saddle_order=so, element_list=["C"] * natoms,
                     trust_radius_max=(0.1 if so > 0 else 0.5), trust_radius_min=0.01)
    opt.set_hessian(H0.copy())
    if Hb is not None:
        opt.set_bias_hessian(Hb.copy())
    keys = ("x_in", "x", "shake", "rows", "Bg", "g", "Be", "move", "H_after", "trust", "pred", "converged")
    rec = {k: [] for k in keys}
    x = x0.copy()
    x_prev = g_prev = None
    for k in range(nsteps):
        e, g = pes.raw(x)
        eb, gb = pes.bias(x) if Hb is not None else (0.0, np.zeros(n))
        Be, Bg = e + eb, g + gb
        if conv_at is not None and k == conv_at:      # gradient inside the span of the constraint rows: subspace gradient 0
            rows_now = cons._get_all_constraint_vectors(x.reshape(-1, 3))
            Bg = 0.3 * rows_now[0] - 0.1 * rows_now[1]
            g = Bg.copy()
        xc = cons.adjust_init_coord(x.reshape(-1, 3)).ravel() if cons is not None else x.copy()
        rows = cons._get_all_constraint_vectors(xc.reshape(-1, 3)) if cons is not None else np.zeros((1, n))
        col = lambda a: a.reshape(-1, 1).copy()
        npred = len(opt.predicted_energy_changes)
        with quiet():
            if x_prev is None:
                mv = opt.run(col(x), col(Bg), [], [], Be, 0.0, [], col(x0), col(g), [])
            else:
                mv = opt.run(col(x), col(Bg), [], col(x_prev), Be, 0.0, [], col(x0), col(g), col(g_prev))
        mv = np.asarray(mv, float).ravel()
        rec["x_in"].append(x.copy()); rec["x"].append(xc); rec["shake"].append(xc - x); rec["rows"].append(rows)
        rec["Bg"].append(Bg); rec["g"].append(g); rec["Be"].append(Be); rec["move"].append(mv)
        rec["H_after"].append(np.array(opt.hessian, float)); rec["trust"].append(float(opt.trust_radius))
        conv = bool(opt.proj_grad_converged) and len(opt.predicted_energy_changes) == npred
        rec["converged"].append(int(conv))
        rec["pred"].append(float(opt.predicted_energy_changes[-1]) if opt.predicted_energy_changes else 0.0)
        opt.proj_grad_converged = False
        x_prev, g_prev = xc.copy(), g.copy()      # the caller sees the corrected geometry only through the move; it passes its own x
        x_prev = x.copy()
        cap = 0.1 if so > 0 else 0.5
        nrm = np.linalg.norm(mv)
        x = xc - (mv * (cap / nrm) if nrm > cap else mv)
    out = {f"{name}/{k}": np.array(v) for k, v in rec.items()}
    out[f"{name}/H0"] = H0
    out[f"{name}/Hb"] = np.zeros((n, n)) if Hb is None else Hb
    out[f"{name}/meta"] = np.array([so, natoms, nsteps, int(bias)], np.int64)
    out[f"{name}/method"] = np.array(method)
    return out


def gen_crsirfo():
    blob = {}
    for case in CRSIRFO_CASES:
        blob.update(run_crsirfo_case(case))
        print("crsirfo case", case[0])
    blob["names"] = np.array([c[0] for c in CRSIRFO_CASES])
    np.savez_compressed(os.path.join(GOLD, "crsirfo_traces.npz"), **blob)


SETS = {"crsirfo": gen_crsirfo, "redistribute": gen_redistribute, "bias2": gen_bias2, "modelhess_d3": gen_modelhess_d3, "keep": gen_keep, "fire": gen_fire, "post": gen_post, "ric": gen_ric, "swart": gen_swart, "update": gen_update, "rsirfo": gen_rsirfo, "projection": gen_projection, "producers": gen_producers,
        "c1": gen_c1_trace, "neb": gen_neb, "lindh": gen_lindh, "rsprfo": gen_rsprfo, "rsprfo_reject": gen_rsprfo_reject, "rankdef": gen_rankdef, "potkeys": gen_potkeys, "neb_full": gen_neb_full}

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or list(SETS)
    for w in which:
        SETS[w]()
