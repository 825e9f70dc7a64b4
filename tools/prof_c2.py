"""Per-kernel durations of one C2 step (torch.profiler), diagnostics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from multioptpy_b200 import ops, synthetic, _lib
import bench
B = int(os.environ.get("DIAG_B", "1024")); dev = torch.device("cuda:0")
x0, H0, g0, rngs = bench.make_inputs(B, 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
lib = _lib.load()
lib.mop_debug_stream_chunk(int(os.environ.get("CHUNK", "256")))
lib.mop_debug_tri_packed(int(os.environ.get("PACKED", "1"))); lib.mop_debug_packed_threads(int(os.environ.get("PKT", "256")))
H = T(H0); st = ops.new_rsirfo_state(B, 0.5, dev)
zero = torch.zeros(B, dtype=torch.float64, device=dev)
m = ops.resolve_update_method("rsirfo_bfgs")
out = ops.rsirfo_step(H.clone(), T(x0), T(g0), T(g0), st, method=m, Be=zero)
mv0 = out["move"].cpu().numpy()
x1 = np.empty_like(x0); g1 = np.empty_like(g0)
for b in range(B):
    x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
x0d, g0d, x1d, g1d = T(x0), T(g0), T(x1), T(g1)
Hs = [H.clone() for _ in range(3)]; sts = [st.clone() for _ in range(3)]
o = ops.rsirfo_step(Hs[0], x1d, g1d, g1d, sts[0], method=m, x_prev=x0d, g_prev=g0d, Be=zero - 1e-3)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    o = ops.rsirfo_step(Hs[1], x1d, g1d, g1d, sts[1], method=m, x_prev=x0d, g_prev=g0d, Be=zero - 1e-3, out=o)
    torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total > 0:
        print(f"{e.device_time_total/1e3:10.3f} ms  x{e.count:<3d} {e.key[:100]}")
