"""Diagnostics on the GPU box: status-bit histogram and per-phase timings of one
C2 bench step (not part of the product or the tests)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic
import bench

B = int(os.environ.get("DIAG_B", "1024")); n = 150
dev = torch.device("cuda:0")
x0, H0, g0, rngs = bench.make_inputs(B, 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
H = T(H0); st = ops.new_rsirfo_state(B, 0.5, dev)
zero = torch.zeros(B, dtype=torch.float64, device=dev)
m = ops.resolve_update_method("rsirfo_bfgs")
out = ops.rsirfo_step(H, T(x0), T(g0), T(g0), st, method=m, Be=zero)
mv0 = out["move"].cpu().numpy()
s0 = out["status"].cpu().numpy()
x1 = np.empty_like(x0); g1 = np.empty_like(g0)
for b in range(B):
    x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
Hs = [H.clone() for _ in range(4)]; sts = [st.clone() for _ in range(4)]
x1d, g1d, x0d, g0d = T(x1), T(g1), T(x0), T(g0)
def bits(s):
    names = ["UPDATED","SKIP_SMALL","SKIP_CURV","TERM_ZEROED","LEVEL_SHIFT","EIG_NONFINITE","ALPHA_SEARCH","STEP_NAN_SD","HARD_CASE","TRROT_RANKDEF","BRENT","EIG_NOCONV","EIG_FALLBACK","NO_HISTORY"]
    return {nm: int(((s >> i) & 1).sum()) for i, nm in enumerate(names) if ((s >> i) & 1).any()}
print("step0 status:", bits(s0))
for i in range(3):
    torch.cuda.synchronize(); t = time.perf_counter()
    out = ops.rsirfo_step(Hs[i], x1d, g1d, g1d, sts[i], method=m, x_prev=x0d, g_prev=g0d, Be=zero - 1e-3)
    torch.cuda.synchronize(); print("step1 ms", (time.perf_counter() - t) * 1e3)
print("step1 status:", bits(out["status"].cpu().numpy()))

from multioptpy_b200 import _lib
lib = _lib.load()
if os.environ.get('TRI_THREADS'):
    lib.mop_debug_tri_threads(int(os.environ['TRI_THREADS']))
if os.environ.get('TRI_ABLATE'):
    lib.mop_debug_tri_ablate(int(os.environ['TRI_ABLATE']))
dbg = torch.zeros(B, 16, dtype=torch.int64, device=dev)
lib.mop_debug_tri_timing(dbg.data_ptr())
out = ops.rsirfo_step(Hs[3], x1d, g1d, g1d, sts[3], method=m, x_prev=x0d, g_prev=g0d, Be=zero - 1e-3)
torch.cuda.synchronize()
lib.mop_debug_tri_timing(None)
d = dbg.cpu().numpy().astype(float)
names = ["load", "tridiag(+Qtg)", "spill/scale/split", "multisection", "twisted", "cluster MGS", "rfo core", "Zc + Qy"]
print("phase cycles (median over CTAs):")
for i, nm in enumerate(names):
    print(f"  {nm:20s} {np.median(d[:, i]):12.0f}  max {d[:, i].max():12.0f}")
print("  total median", np.median(d[:, :8].sum(1)))
print("phase-1 segments warp0   [top..A, A..B(symv), B..C(reduce), C..E-arrive(update), E wait]:", [int(np.median(d[:, 8 + i])) for i in range(5)])
print("phase-1 segments warp15  [top..A, A..B(symv), B..C(reduce), C..E-arrive(update), E wait]:", [int(np.median(d[:, 13 + i])) for i in range(3)])

# k_spectrum_step phases (default path)
dbg2 = torch.zeros(B, 16, dtype=torch.int64, device=dev)
lib.mop_debug_spectrum_timing(dbg2.data_ptr())
Hx = H.clone(); stx = st.clone()
out = ops.rsirfo_step(Hx, x1d, g1d, g1d, stx, method=m, x_prev=x0d, g_prev=g0d, Be=zero - 1e-3)
torch.cuda.synchronize()
lib.mop_debug_spectrum_timing(None)
d2 = dbg2.cpu().numpy().astype(float)
names2 = ["load/scale/split", "eigenvalues", "twisted", "cluster CGS2", "gamma + rfo core", "Z c", "Q y"]
print("k_spectrum_step phase cycles (median / max over CTAs):")
for i, nm in enumerate(names2):
    print(f"  {nm:20s} {np.median(d2[:, i]):12.0f}  max {d2[:, i].max():12.0f}")
print("  total median", np.median(d2[:, :7].sum(1)))
print("status:", bits(out["status"].cpu().numpy()))
