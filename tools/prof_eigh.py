import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from multioptpy_b200 import ops, synthetic
n = int(sys.argv[1]); B = int(sys.argv[2])
A = np.stack([synthetic.spd_hessian(n, np.random.default_rng(b)) for b in range(8)])
Ad = torch.from_numpy(np.tile(A, (B // 8, 1, 1))).cuda()
ops.eigh(Ad); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ops.eigh(Ad); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
    print(f"{e.device_time_total/1e3:10.3f} ms  x{e.count:<3d} {e.key[:90]}")
