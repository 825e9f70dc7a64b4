"""Per-row measurement of the hot-path kernels (SURVEY §8 a-rows) on one B200: device time per call
(CUDA events, median of repeats, inputs resident and larger than L2 where the config is), algorithmic
bytes / flops, fraction of the measured roofline, and the CPU oracle timed on one host core beside it.
Writes gpurun_out/r2_row_measurements.json (copied to profiles/) and prints a markdown table.  Not the bench line: bench.py is."""
import json, os, sys, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic
from multioptpy_b200.Utils.bond_connectivity import radii_array
from multioptpy_b200.ModelHessian.lindh import lindh_atom_params
from multioptpy_b200.ModelHessian.swart import swart_radii
from oracle import np_oracle as O

dev = torch.device("cuda:0")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = float(peaks.get("hbm_gbs", 6650.0))
FP64 = 36.7  # TFLOP/s, mop_priv_bench_dfma (bench.py measures it in-run)
T = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev) if dt is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
rows = []


def gpu_ms(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def cpu_s(fn, reps=2):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def add(row, what, B, ms, bytes_per=None, flops_per=None, cpu_per=None, note=""):
    r = {"row": row, "kernel": what, "structures": B, "ms": ms, "structures_per_s": B / (ms * 1e-3)}
    if bytes_per is not None:
        r["GBs"] = B * bytes_per / (ms * 1e-3) / 1e9; r["hbm_frac"] = r["GBs"] / HBM
    if flops_per is not None:
        r["TFLOPs"] = B * flops_per / (ms * 1e-3) / 1e12; r["fp64_frac"] = r["TFLOPs"] / FP64
    if cpu_per is not None:
        r["cpu_structures_per_s_per_core"] = 1.0 / cpu_per
        r["gpu_over_one_core"] = r["structures_per_s"] * cpu_per
    r["note"] = note
    rows.append(r)
    print(r, flush=True)


def geoms(B, N, seed=0, spacing=2.6, jitter=0.25):
    return np.stack([synthetic.grid_geometry(N, np.random.default_rng(seed + b), spacing=spacing, jitter=jitter) for b in range(B)])


# ---- a1-a3 update, a4-a5 projection, a6 eigh, a7-a10 step: config 2 sizes ------------------------
B, N = 1024, 50; n = 3 * N
rng = np.random.default_rng(1)
H0 = np.stack([synthetic.spd_hessian(n, np.random.default_rng(10 + b)) for b in range(8)])
Hd = T(np.tile(H0, (B // 8, 1, 1)))
s = T(rng.normal(0, 0.05, (B, n))); y = torch.einsum("bij,bj->bi", Hd, s) + T(rng.normal(0, 1e-3, (B, n)))
x = T(geoms(B, N).reshape(B, n)); g = T(rng.normal(0, 1e-2, (B, n)))
Hs = [Hd.clone() for _ in range(3)]
it = [0]
def upd():
    it[0] += 1
    ops.hessian_update(Hs[it[0] % 3], s, y, 23, inplace=True, rsirfo_guards=True)
cpu = cpu_s(lambda: O.hessian_update_delta(23, H0[0], s[0].cpu().numpy(), y[0].cpu().numpy()))
add("a1-a3", "mop_hessian_update (Bofill, in place, single-kernel ABI path)", B, gpu_ms(upd), 16 * n * n + 32 * n, None, cpu, "C2 size")
cpu = cpu_s(lambda: O.project_hessian_trrot(H0[0], x[0].cpu().numpy()))
add("a4-a5", "mop_project_trrot (Hessian + gradient)", B, gpu_ms(lambda: ops.project_trrot(Hd, x, g=g)), 16 * n * n + 32 * n, None, cpu, "C2 size")
Hp, gp, _ = ops.project_trrot(Hd, x, g=g)
cpu = cpu_s(lambda: np.linalg.eigh(Hp[0].cpu().numpy()))
add("a6", "mop_eigh auto (n=150: shared-memory tridiagonal path, V formed)", B, gpu_ms(lambda: ops.eigh(Hp)), None, 9.0 * n ** 3, cpu, "C2 size; cpu = numpy.linalg.eigh")

# large n
Bl, nl = 64, 600
Al = np.stack([synthetic.spd_hessian(nl, np.random.default_rng(70 + b), neg_lowest=True) for b in range(4)])
Ald = T(np.tile(Al, (Bl // 4, 1, 1)))
cpu = cpu_s(lambda: np.linalg.eigh(Al[0]), reps=1)
add("a6", "mop_eigh large (n=600: cluster tridiagonalisation, V formed)", Bl, gpu_ms(lambda: ops.eigh(Ald), reps=3, warm=1), None, 9.0 * nl ** 3, cpu, "C5 size; cpu = numpy.linalg.eigh")

# ---- a11 caller ----------------------------------------------------------------------------------
mv = T(rng.normal(0, 0.1, (B, n))); tr = torch.full((B,), 0.3, dtype=torch.float64, device=dev)
add("a11", "mop_clamp_and_move", B, gpu_ms(lambda: ops.clamp_and_move(x, mv, tr)), 32 * n, None, None, "C2 size")

# ---- a12 P-RFO small n -----------------------------------------------------------------------------
z = lambda *sh: torch.zeros(*sh, dtype=torch.float64, device=dev)
Hsad = np.stack([synthetic.spd_hessian(n, np.random.default_rng(30 + b), neg_lowest=True) for b in range(8)])
Hsd = T(np.tile(Hsad, (B // 8, 1, 1)))
st = dict(state=z(B, ops.PRFO_STATE), prev_grad=z(B, n), prev_move=z(B, n), ts_vec=z(B, n)); st["state"][:, 0] = 0.1
o0 = ops.rsprfo_step(Hsd.clone(), x, g, st, method=23, saddle_order=1, Be=z(B))
x1 = x - o0["move"]; g1 = g + torch.einsum("bij,bj->bi", Hsd, x1 - x)
st_ref = {k: v.clone() for k, v in st.items()}; Hc = Hsd.clone(); mv0 = o0["move"].clone(); Be1 = z(B) - 1e-3
def prfo():
    Hc.copy_(Hsd)
    for k in st: st[k].copy_(st_ref[k])
    ops.rsprfo_step(Hc, x1, g1, st, method=23, saddle_order=1, x_prev=x, Bg_prev=g, pre_move=mv0, Be=Be1)
orc = O.RSPRFOOracle(method="rsprfo_bofill", saddle_order=1); orc.set_hessian(Hsad[0])
xx, gg = x[0].cpu().numpy(), g[0].cpu().numpy()
m0 = orc.run(xx, gg, None, None, 0.0, None)
import copy
def prfo_cpu():
    o2 = copy.deepcopy(orc)
    o2.run(x1[0].cpu().numpy(), g1[0].cpu().numpy(), xx, gg, -1e-3, m0)
add("a12", "mop_rsprfo_step (P-RFO + Bofill, n=150, incl. state reset copies)", B, gpu_ms(prfo), None, (9 + 4 / 3) * n ** 3 + 40 * n * n, cpu_s(prfo_cpu), "C2 size, saddle order 1")

# ---- producers: a13 Fischer, a14 Swart, a16 connectivity at N=50; a15 Lindh, a17 AFIR at config 4 ---
elems = synthetic.elements(N)
xyz = T(geoms(B, N, spacing=2.8, jitter=0.2)); rad = radii_array(elems)   # SURVEY §8d geometry
cpu = cpu_s(lambda: O.connectivity_tables(xyz[0].cpu().numpy(), rad))
add("a16", "mop_connectivity", B, gpu_ms(lambda: ops.connectivity(xyz, rad)), 24 * N, None, cpu, "N=50")
cpu = cpu_s(lambda: O.fischer_hessian(xyz[0].cpu().numpy(), rad), reps=1)
add("a13", "mop_fischer_hessian (incl. projection)", B, gpu_ms(lambda: ops.fischer_hessian(xyz, rad)), 24 * n * n, None, cpu, "N=50")
srad = swart_radii(elems)
cpu = cpu_s(lambda: O.swart_hessian(xyz[0].cpu().numpy(), srad), reps=1)
add("a14", "mop_swart_hessian (incl. projection)", B, gpu_ms(lambda: ops.swart_hessian(xyz, srad)), 24 * n * n, None, cpu, "N=50")

B4, N4 = 8192, 24; n4 = 3 * N4
el4 = synthetic.elements(N4, all_sulfur=True)
# config 4's own batch (S8-like crown-ring conformers).  Until round 2 this row used the dense all-sulfur GRID of the other
# producer rows (every atom bonded to every neighbour: 99 bonds, 382 angles and a dihedral table that saturates its 1536
# slots per structure - not a molecule): k_lindh takes 8.0 ms there against 1.8 ms on the conformers (tools/prof_lindh_rows.py).
xyz4 = T(synthetic.conformer_batch(B4, N4, seed=4168)[0])
prm = lindh_atom_params(el4)
cpu = cpu_s(lambda: O.lindh_hessian_bkb(xyz4[0].cpu().numpy(), prm), reps=1)
add("a15", "mop_lindh_hessian (force constants + B^T k B + projection)", B4, gpu_ms(lambda: ops.lindh_hessian(xyz4, prm)), 24 * n4 * n4, None, cpu, "config 4: 8192 x N=24 S8-like conformers (dense all-sulfur grid: 9.1 ms, dihedral table saturated)")
f1 = torch.arange(0, N4 // 2, dtype=torch.int32, device=dev); f2 = torch.arange(N4 // 2, N4, dtype=torch.int32, device=dev)
r32 = torch.tensor(radii_array(el4), dtype=torch.float32, device=dev); gam = torch.full((B4,), 100.0, dtype=torch.float64, device=dev)
cpu = cpu_s(lambda: O.afir_egh(xyz4[0].cpu().numpy(), list(range(N4 // 2)), list(range(N4 // 2, N4)), radii_array(el4), 100.0), reps=1)
add("a17", "mop_afir (energy + gradient + Hessian)", B4, gpu_ms(lambda: ops.afir(xyz4, f1, f2, r32, gam)), 8 * n4 * n4 + 16 * n4, None, cpu, "config 4: 8192 x N=24, 12+12 fragments; cpu = torch.func jacrev/hessian")

# ---- a18 RIC ---------------------------------------------------------------------------------------
Br = 1024
xr = xyz4[:Br].contiguous()
bo, an, di, cn, _ = ops.connectivity(xr, radii_array(el4))
M4 = N4 * (N4 - 1) // 2
nterm_max = int(cn.sum(dim=1).max().item())
q = T(rng.normal(0, 1e-2, (Br, max(M4, nterm_max)))); hd = T(np.abs(rng.normal(0.3, 0.1, (Br, M4))))
def ric():
    K = ops.ric_kmatrix(xr, bo, an, di, cn, q)
    ops.ric_hess_to_cart(xr, hd, K)
c0 = cn[0].cpu().numpy(); tabs0 = [bo[0, :c0[0]].cpu().numpy(), an[0, :c0[1]].cpu().numpy(), di[0, :c0[2]].cpu().numpy()]
def ric_cpu():
    Bm = O.ric_bmatrix(xr[0].cpu().numpy())
    Bm.T @ np.diag(hd[0].cpu().numpy()) @ Bm + O.ric_kmatrix(xr[0].cpu().numpy(), tabs0, q[0].cpu().numpy())
add("a18", "mop_ric_kmatrix + mop_ric_hess_to_cart (B^T k B + K)", Br, gpu_ms(ric), 16 * n4 * n4, None, cpu_s(ric_cpu, reps=1), "N=24; cpu = torch.func.hessian per internal coordinate")

# ---- a19-a20 NEB: config 3, 64 images x 30 atoms on one GPU -------------------------------------------
nimg, Nn = 64, 30; nn_ = 3 * Nn
xa = synthetic.grid_geometry(Nn, np.random.default_rng(3)).reshape(-1); xb = xa + np.random.default_rng(4).normal(0, 0.3, nn_)
X = np.stack([xa + (xb - xa) * t for t in np.linspace(0, 1, nimg)]) + np.random.default_rng(5).normal(0, 0.02, (nimg, nn_))
E = -np.sin(np.linspace(0, np.pi, nimg)) * 0.05; G = np.random.default_rng(6).normal(0, 1e-2, (nimg, nn_))
pad = lambda a: np.concatenate([a[:1], a, a[-1:]])
xh, Eh, gh = T(pad(X)), T(pad(E)), T(pad(G)); gd = T(G)
Hn = T(np.stack([synthetic.spd_hessian(nn_, np.random.default_rng(40 + i)) for i in range(nimg)]))
def neb():
    f, tau = ops.bneb_force(nimg, 0, xh, Eh, gd)
    ops.neb_ayala(nimg, 0, xh, Eh, gh, tau, Hn)
    d = f.clone()
    ops.neb_limit_tr(nimg, 0, xh, gd, d)
cpu = cpu_s(lambda: O.bneb_force(X, E, G), reps=1)
add("a19-a20", "mop_bneb_force + mop_neb_ayala + mop_neb_limit_tr (one NEB iteration's path kernels, 64 images)", nimg,
    gpu_ms(neb), None, None, cpu / nimg, "config 3 on one GPU; unit = image; cpu = oracle bneb_force only")

json.dump({"hbm_peak_gbs": HBM, "fp64_peak_tflops": FP64, "rows": rows}, open(os.path.join(ROOT, "gpurun_out" if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else "profiles", "r2_row_measurements.json"), "w"), indent=1)
print("\n| row | kernel | units | ms | units/s | GB/s (frac of HBM) | TFLOP/s alg. (frac) | CPU units/s/core | note |")
print("|---|---|---|---|---|---|---|---|---|")
for r in rows:
    gb = f"{r['GBs']:.0f} ({r['hbm_frac']:.2f})" if "GBs" in r else "-"
    tf = f"{r['TFLOPs']:.2f} ({r['fp64_frac']:.3f})" if "TFLOPs" in r else "-"
    cp = f"{r['cpu_structures_per_s_per_core']:.3g}" if "cpu_structures_per_s_per_core" in r else "-"
    print(f"| {r['row']} | {r['kernel']} | {r['structures']} | {r['ms']:.3f} | {r['structures_per_s']:.3g} | {gb} | {tf} | {cp} | {r['note']} |")
