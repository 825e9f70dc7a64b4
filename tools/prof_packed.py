"""One C2 step on PACKED Hessians (mop_rsirfo_step_packed, the e2e path's kernels) - ncu target for the TMA bulk copies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic
import bench
B = int(os.environ.get("DIAG_B", "1024")); dev = torch.device("cuda:0")
x0, H0, g0, rngs = bench.make_inputs(B, 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
P = ops.pack_lower(T(H0)); st = ops.new_rsirfo_state(B, 0.5, dev)
zero = torch.zeros(B, dtype=torch.float64, device=dev)
m = ops.resolve_update_method("rsirfo_bfgs")
out = ops.rsirfo_step(P, T(x0), T(g0), T(g0), st, method=m, Be=zero, packed=True)
mv0 = out["move"].cpu().numpy()
x1 = np.empty_like(x0); g1 = np.empty_like(g0)
for b in range(B):
    x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
out = ops.rsirfo_step(P, T(x1), T(g1), T(g1), st, method=m, x_prev=T(x0), g_prev=T(g0), Be=zero - 1e-3, packed=True)
torch.cuda.synchronize()
print("ok", int((out["status"] & ops.ST_UPDATED).ne(0).sum().item()), "updated")
