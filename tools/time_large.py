"""Timing of the large-n eigensolver / P-RFO step (diagnostics, run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(0)
A = rng.standard_normal((B, n, n)); A = 0.5 * (A + A.transpose(0, 2, 1))
At = torch.from_numpy(A).cuda()
lib = _lib.load()
for cl in (8, 4, 2, 1):
    lib.mop_priv_large_cluster(cl)
    ops.eigh(At, "large"); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ev, V, st = ops.eigh(At, "large")
    e1.record(); torch.cuda.synchronize()
    print(f"n={n} B={B} cluster={cl}: {e0.elapsed_time(e1)/3:.2f} ms per batch, fallbacks {(st.cpu().numpy() & ops.ST_EIG_FALLBACK != 0).sum()}")
lib.mop_priv_large_cluster(0)
