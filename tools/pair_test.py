import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from multioptpy_b200 import ops, synthetic, _lib
lib = _lib.load()
for n, B in ((600, 5), (201, 3), (384, 4), (1024, 2)):
    rng = np.random.default_rng(n)
    A = rng.standard_normal((B, n, n)); A = 0.5 * (A + A.transpose(0, 2, 1))
    A[1] = synthetic.spd_hessian(n, rng, neg_lowest=True)
    for mode in (1, -1):
        lib.mop_debug_large_pair(mode)
        ev, V, st = ops.eigh(torch.from_numpy(A).cuda(), "large"); torch.cuda.synchronize()
        ev, V = ev.cpu().numpy(), V.cpu().numpy()
        w = 0.0; r = 0.0
        for b in range(B):
            ref = np.linalg.eigvalsh(A[b]); sc = np.abs(ref).max()
            w = max(w, np.abs(ev[b] - ref).max() / sc)
            r = max(r, np.abs(A[b] @ V[b].T - V[b].T * ev[b]).max() / sc)
        print(f"n={n} B={B} pair={mode}: eigenvalue err {w:.2e} residual {r:.2e}")
lib.mop_debug_large_pair(0)
