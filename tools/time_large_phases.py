import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, _lib
n = int(sys.argv[1]); B = int(sys.argv[2])
rng = np.random.default_rng(0)
if len(sys.argv) > 3 and sys.argv[3] == "spd":
    from multioptpy_b200 import synthetic
    A = np.stack([synthetic.spd_hessian(n, rng, neg_lowest=True) for _ in range(min(B, 4))]); A = np.tile(A, (B // len(A) + 1, 1, 1))[:B].copy()
else:
    A = rng.standard_normal((B, n, n)); A = 0.5 * (A + A.transpose(0, 2, 1))
At = torch.from_numpy(A).cuda()
lib = _lib.load()
abl = int(os.environ.get('ABL', '0')); lib.mop_debug_large_ablate(abl)
for cl in (8,):
    lib.mop_debug_large_cluster(cl)
    ops.eigh(At, "large"); torch.cuda.synchronize()
    dbg = torch.zeros(2 * B, 4, dtype=torch.int64, device="cuda")
    lib.mop_debug_large_timing(dbg.data_ptr())
    ops.eigh(At, "large"); torch.cuda.synchronize()
    lib.mop_debug_large_timing(None)
    d2 = dbg.cpu().numpy().astype(float)[B:]
    print(f"   trieig cycles: bisection={d2[:,0].mean():.0f} vectors={d2[:,1].mean():.0f} cgs2={d2[:,2].mean():.0f}")
    d = dbg.cpu().numpy().astype(float)[:B]
    print(f"n={n} B={B} cl={cl}: mean cycles per matrix  w={d[:,0].mean():.0f} house={d[:,1].mean():.0f} upd+symv={d[:,2].mean():.0f} barrier={d[:,3].mean():.0f}  total={d.sum(1).mean():.0f}  per column {d.sum(1).mean()/n:.0f}")
