"""Timing of the P-RFO step at BASELINE config 5's shape (diagnostics, run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic

natoms = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
n = 3 * natoms
rng = np.random.default_rng(0)
Hs, xs, gs = [], [], []
for b in range(min(B, 8)):
    x0, H, g0, _ = synthetic.structure(5, b, natoms, saddle=True)
    Hs.append(H); xs.append(x0); gs.append(g0)
rep = (B + len(Hs) - 1) // len(Hs)
H = torch.from_numpy(np.stack(Hs)).repeat(rep, 1, 1)[:B].contiguous().cuda()
x = torch.from_numpy(np.stack(xs)).repeat(rep, 1)[:B].contiguous().cuda()
g = torch.from_numpy(np.stack(gs)).repeat(rep, 1)[:B].contiguous().cuda()
z = lambda *s: torch.zeros(*s, dtype=torch.float64, device="cuda")
def fresh():
    st = dict(state=z(B, ops.PRFO_STATE), prev_grad=z(B, n), prev_move=z(B, n), ts_vec=z(B, n))
    st["state"][:, 0] = 0.1
    return st
Be = z(B)
st = fresh()
out = ops.rsprfo_step(H.clone(), x, g, st, method=23, saddle_order=1, Be=Be)
x1 = x - out["move"]; g1 = g + torch.einsum("bij,bj->bi", H, x1 - x)
Hc = H.clone()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    Hc.copy_(H)
    ops.rsprfo_step(Hc, x1, g1, st, method=23, saddle_order=1, x_prev=x, Bg_prev=g, pre_move=out["move"], Be=Be - 1e-3, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"P-RFO + Bofill step N={natoms} (n={n}) B={B}: {ms:.2f} ms per batch = {B / ms * 1e3:.0f} steps/s; status bits {int(torch.bitwise_or(out['status'][0], out['status'][-1]))}")
