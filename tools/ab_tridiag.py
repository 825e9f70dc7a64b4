"""A/B of the tridiagonalisation kernels inside the C2 step (diagnostics): blocked DMMA kernel vs k_tridiag_rwf.
Prints ms per 1024-structure step, the max relative difference of the moves and the per-phase clocks of the
blocked kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic, _lib
import bench
B = int(os.environ.get("DIAG_B", "1024")); dev = torch.device("cuda:0")
x0, H0, g0, rngs = bench.make_inputs(B, 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
lib = _lib.load()
import ctypes
lib.mop_priv_tridiag_blk_timing.argtypes = [ctypes.c_void_p]
H = T(H0); st = ops.new_rsirfo_state(B, 0.5, dev)
zero = torch.zeros(B, dtype=torch.float64, device=dev)
m = ops.resolve_update_method("rsirfo_bfgs")
out = ops.rsirfo_step(H.clone(), T(x0), T(g0), T(g0), st, method=m, Be=zero)
mv0 = out["move"].cpu().numpy()
x1 = np.empty_like(x0); g1 = np.empty_like(g0)
for b in range(B):
    x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
x0d, g0d, x1d, g1d = T(x0), T(g0), T(x1), T(g1)
ref = None
for blocked, fused in ((0, 0), (1, 0), (1, 1)):
    lib.mop_debug_packed_blocked(blocked); lib.mop_debug_front_fused(fused)
    Hs = [H.clone() for _ in range(4)]; sts = [st.clone() for _ in range(4)]
    o = None
    for i in range(2):
        o = ops.rsirfo_step(Hs[i], x1d, g1d, g1d, sts[i], method=m, x_prev=x0d, g_prev=g0d, Be=zero - 1e-3, out=o)
    torch.cuda.synchronize()
    mv = o["move"].clone()
    if ref is None:
        ref = mv
    err = float(((mv - ref).norm(dim=1) / ref.norm(dim=1)).max())
    Hs = [H.clone() for _ in range(8)]; sts = [st.clone() for _ in range(8)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        o = ops.rsirfo_step(Hs[i], x1d, g1d, g1d, sts[i], method=m, x_prev=x0d, g_prev=g0d, Be=zero - 1e-3, out=o)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 8
    fb = int((o["status"] & ops.ST_EIG_FALLBACK).ne(0).sum())
    print(f"blocked={blocked} fused={fused}: {ms:.3f} ms/step = {B / ms * 1e3:.0f} steps/s; max rel diff vs rwf {err:.2e}; fallbacks {fb}", flush=True)
# phase clocks of the blocked kernel
lib.mop_debug_packed_blocked(1); lib.mop_debug_front_fused(0)
dbg = torch.zeros(B, 16, dtype=torch.int64, device=dev)
lib.mop_priv_tridiag_blk_timing(dbg.data_ptr())
Hs = H.clone(); s2 = st.clone()
ops.rsirfo_step(Hs, x1d, g1d, g1d, s2, method=m, x_prev=x0d, g_prev=g0d, Be=zero - 1e-3)
torch.cuda.synchronize()
lib.mop_priv_tridiag_blk_timing(None); lib.mop_debug_front_fused(1)
d = dbg.cpu().numpy().astype(float)
names = ["(a) panel rows + c", "(b) symv", "(c) reduction", "(d) scalars/w/v", "trailing DMMA"]
tot = d[:, :5].sum(1).mean()
print(f"blocked kernel, thread 0, mean SM cycles per structure {tot:.0f} ({tot / 148:.0f} per column)")
for q, nm in enumerate(names):
    print(f"  {nm:22s} {d[:, q].mean():10.0f}  {100 * d[:, q].mean() / tot:5.1f} %   thread 96: {d[:, 8 + q].mean():10.0f}")
