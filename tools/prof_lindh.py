import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from multioptpy_b200 import ops, synthetic
from multioptpy_b200.ModelHessian.lindh import lindh_atom_params
from multioptpy_b200.Utils.bond_connectivity import radii_array
B, N = 8192, 24
el = synthetic.elements(N, all_sulfur=True)
xyz = torch.from_numpy(np.stack([synthetic.grid_geometry(N, np.random.default_rng(500 + b), spacing=2.6, jitter=0.25) for b in range(B)])).cuda()
prm = lindh_atom_params(el)
ops.lindh_hessian(xyz, prm); ops.fischer_hessian(xyz, radii_array(el)); ops.connectivity(xyz, radii_array(el)); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ops.lindh_hessian(xyz, prm); ops.fischer_hessian(xyz, radii_array(el)); ops.connectivity(xyz, radii_array(el)); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f"{e.device_time_total/1e3:10.3f} ms  x{e.count:<3d} {e.key[:100]}")
