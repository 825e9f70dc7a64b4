"""Phase clocks of k_tridiag_blk<5, DBG> (non-fused, one launch for all columns): python tools/tb_segments.py [n] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
A = np.stack([synthetic.spd_hessian(n, np.random.default_rng(b)) for b in range(8)])
Ad = torch.from_numpy(np.tile(A, (B // 8, 1, 1))).cuda()
ops.eigh(Ad); torch.cuda.synchronize()
lib = _lib.load()
dbg = torch.zeros(B, 16, dtype=torch.int64, device="cuda")
lib.mop_priv_tridiag_blk_timing(dbg.data_ptr())
ops.eigh(Ad); torch.cuda.synchronize()
lib.mop_priv_tridiag_blk_timing(0)
d = dbg.cpu().numpy().astype(float)
names = ["(a) panel products", "(b) symv", "(c) reduction + barrier", "(d) scalars + barrier", "(e) trailing update"]
for w, off in (("warp 0", 0), ("warp 3", 8)):
    tot = d[:, off:off + 5].sum(1).mean()
    print(w, "total cycles", tot, "per column", tot / (n - 2))
    for q, nm in enumerate(names):
        print(f"   {nm:26s} {d[:, off + q].mean():12.0f}  {100 * d[:, off + q].mean() / tot:5.1f} %")
