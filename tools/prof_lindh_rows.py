"""Kernel times of mop_lindh_hessian on the row-table geometry (8192 x N = 24 grid geometries)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from multioptpy_b200 import ops, synthetic
from multioptpy_b200.ModelHessian.lindh import lindh_atom_params
B, N = 8192, 24
el = synthetic.elements(N, all_sulfur=True)
xyz = np.stack([synthetic.grid_geometry(N, np.random.default_rng(500 + b), spacing=2.8, jitter=0.2) for b in range(B)])
xd = torch.from_numpy(xyz).cuda(); prm = lindh_atom_params(el)
H, kd, counts, st = ops.lindh_hessian(xd, prm); torch.cuda.synchronize()
print("counts (bonds, angles, dihedrals) mean", counts.double().mean(0).cpu().numpy(), "max", counts.max(0).values.cpu().numpy())
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ops.lindh_hessian(xd, prm); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:6]:
    print(f"{e.device_time_total/1e3:10.3f} ms  x{e.count:<3d} {e.key[:90]}")
