"""Per-kernel durations of one P-RFO step (torch.profiler / CUPTI), diagnostics."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from multioptpy_b200 import ops, synthetic
natoms = int(sys.argv[1]); B = int(sys.argv[2]); n = 3 * natoms
x0, H0, g0, _ = synthetic.structure(5, 0, natoms, saddle=True)
H = torch.from_numpy(H0).repeat(B, 1, 1).contiguous().cuda()
x = torch.from_numpy(x0).repeat(B, 1).contiguous().cuda(); g = torch.from_numpy(g0).repeat(B, 1).contiguous().cuda()
z = lambda *s: torch.zeros(*s, dtype=torch.float64, device="cuda")
st = dict(state=z(B, ops.PRFO_STATE), prev_grad=z(B, n), prev_move=z(B, n), ts_vec=z(B, n)); st["state"][:, 0] = 0.1
out = ops.rsprfo_step(H.clone(), x, g, st, method=23, saddle_order=1, Be=z(B))
x1 = x - out["move"]; g1 = g + torch.einsum("bij,bj->bi", H, x1 - x)
Be1 = z(B) - 1e-3
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ops.rsprfo_step(H, x1, g1, st, method=23, saddle_order=1, x_prev=x, Bg_prev=g, pre_move=out["move"], Be=Be1, out=out)
    torch.cuda.synchronize()
tot = 0.0
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
    if e.device_time_total > 0:
        print(f"{e.device_time_total/1e3:10.3f} ms  x{e.count:<3d} {e.key[:90]}")
        tot += e.device_time_total
print(f"{tot/1e3:10.3f} ms  total device time")
