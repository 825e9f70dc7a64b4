import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, _lib
n = int(sys.argv[1]); B = int(sys.argv[2])
rng = np.random.default_rng(0)
A = rng.standard_normal((B, n, n)); A = 0.5 * (A + A.transpose(0, 2, 1))
At = torch.from_numpy(A).cuda()
lib = _lib.load()
for cl in (8, 4, 2, 1):
    lib.mop_debug_large_cluster(cl)
    ops.eigh(At, "large"); torch.cuda.synchronize()
    dbg = torch.zeros(B, 4, dtype=torch.int64, device="cuda")
    lib.mop_debug_large_timing(dbg.data_ptr())
    ops.eigh(At, "large"); torch.cuda.synchronize()
    lib.mop_debug_large_timing(None)
    d = dbg.cpu().numpy().astype(float)
    print(f"n={n} B={B} cl={cl}: mean cycles per matrix  w={d[:,0].mean():.0f} house={d[:,1].mean():.0f} upd+symv={d[:,2].mean():.0f} barrier={d[:,3].mean():.0f}  total={d.sum(1).mean():.0f}  per column {d.sum(1).mean()/n:.0f}")
