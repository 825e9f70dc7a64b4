"""Minimal single-step driver for ncu captures (B structures of the C2 workload)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic
import bench
B = int(os.environ.get("DIAG_B", "148")); dev = torch.device("cuda:0")
x0, H0, g0, rngs = bench.make_inputs(B, 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
H = T(H0); st = ops.new_rsirfo_state(B, 0.5, dev)
zero = torch.zeros(B, dtype=torch.float64, device=dev)
m = ops.resolve_update_method("rsirfo_bfgs")
out = ops.rsirfo_step(H, T(x0), T(g0), T(g0), st, method=m, Be=zero)
mv0 = out["move"].cpu().numpy()
x1 = np.empty_like(x0); g1 = np.empty_like(g0)
for b in range(B):
    x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
dbg = None
if os.environ.get("SP_PHASES"):   # phase clocks of k_spectrum_step (cycles per structure)
    from multioptpy_b200 import _lib
    dbg = torch.zeros(B, 16, dtype=torch.int64, device=dev)
    _lib.load().mop_priv_spectrum_timing(dbg.data_ptr())
out = ops.rsirfo_step(H, T(x1), T(g1), T(g1), st, method=m, x_prev=T(x0), g_prev=T(g0), Be=zero - 1e-3)
torch.cuda.synchronize()
if dbg is not None:
    _lib.load().mop_priv_spectrum_timing(0)
    d = dbg.cpu().numpy().astype(float)
    names = ["load/scale/split", "eigenvalues", "twisted vectors", "cluster CGS2", "gamma + rfo_core", "y = Z c", "Q y"]
    print("mean cycles per structure", d[:, :7].sum(1).mean())
    for q, nm in enumerate(names):
        print(f"  {nm:20s} mean {d[:, q].mean():10.0f}  p50 {np.median(d[:, q]):10.0f}  max {d[:, q].max():10.0f}")
if os.environ.get("PROF"):   # per-kernel device times of one more step (torch.profiler)
    from torch.profiler import profile, ProfilerActivity
    H2 = T(H0); st2 = ops.new_rsirfo_state(B, 0.5, dev)
    ops.rsirfo_step(H2, T(x0), T(g0), T(g0), st2, method=m, Be=zero)
    xx1, gg1, xx0, gg0 = T(x1), T(g1), T(x0), T(g0)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ops.rsirfo_step(H2, xx1, gg1, gg1, st2, method=m, x_prev=xx0, g_prev=gg0, Be=zero - 1e-3)
        torch.cuda.synchronize()
    for e in sorted(prof.events(), key=lambda e: e.time_range.start):
        if e.device_time_total > 0:
            print(f"{e.device_time_total/1e3:9.3f} ms  {e.name[:100]}")
print("ok", int(out["status"].sum().item()))
