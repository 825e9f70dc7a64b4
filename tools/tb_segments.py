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

if len(sys.argv) > 3 and sys.argv[3] == "front":   # phase clocks of the fused front end (C2 step with the update active)
    import bench
    B = 1024; dev = torch.device("cuda:0")
    x0, H0, g0, rngs = bench.make_inputs(B, 0)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    H = T(H0); st = ops.new_rsirfo_state(B, 0.5, dev); zero = torch.zeros(B, dtype=torch.float64, device=dev)
    m = ops.resolve_update_method("rsirfo_bfgs")
    out = ops.rsirfo_step(H, T(x0), T(g0), T(g0), st, method=m, Be=zero)
    mv0 = out["move"].cpu().numpy()
    x1 = np.empty_like(x0); g1 = np.empty_like(g0)
    for b in range(B):
        x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
    dbg.zero_()
    lib.mop_priv_tridiag_blk_timing(dbg.data_ptr())
    ops.rsirfo_step(H, T(x1), T(g1), T(g1), st, method=m, x_prev=T(x0), g_prev=T(g0), Be=zero - 1e-3)
    torch.cuda.synchronize()
    lib.mop_priv_tridiag_blk_timing(0)
    d = dbg.cpu().numpy().astype(float)
    names = ["s, y, guards", "read of H", "coefficients + update", "write-back", "basis + projected gradient", "W = S T, M, Y", "rank-12 update"]
    tot = d[:, :7].sum(1).mean()
    print("front end: total cycles", tot)
    for q, nm in enumerate(names):
        print(f"   {nm:28s} {d[:, q].mean():10.0f}  {100 * d[:, q].mean() / tot:5.1f} %")
