"""Phase clocks of k_lg_trieig (eigenvalues | twisted vectors | cluster repair): python tools/trieig_phases.py [n] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.default_rng(0)
A = rng.standard_normal((B, n, n)); A = 0.5 * (A + A.transpose(0, 2, 1))
At = torch.from_numpy(A).cuda()
lib = _lib.load()
ops.eigh(At, "large"); torch.cuda.synchronize()
dbg = torch.zeros(2 * B * 4, dtype=torch.int64, device="cuda")
lib.mop_priv_large_timing(dbg.data_ptr())
ops.eigh(At, "large"); torch.cuda.synchronize()
lib.mop_priv_large_timing(0)
d = dbg.cpu().numpy().astype(float)[B * 4:].reshape(B, 4)
for q, nm in enumerate(["eigenvalues", "twisted vectors", "cluster repair"]):
    print(f"{nm:18s} mean {d[:, q].mean():12.0f} cycles  max {d[:, q].max():12.0f}")
