"""CUDA-event timing of mop_project_trrot (one CTA per structure), C2 batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops
import bench
B = 1024
x0, H0, g0, rngs = bench.make_inputs(B, 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
Hd, xd, gd = T(H0), T(x0), T(g0)
for _ in range(3): Hp, gp, _ = ops.project_trrot(Hd, xd, g=gd)
torch.cuda.synchronize()
a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): Hp, gp, _ = ops.project_trrot(Hd, xd, g=gd)
b_.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b_) / 10
print(f"mop_project_trrot: {ms:.3f} ms per {B} structures, {B * 24 * 150 * 150 / ms / 1e6:.0f} GB/s algorithmic")
