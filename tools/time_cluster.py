"""Phase cycles and time of the blocked cluster tridiagonalisation (tridiag_cluster.cu) inside mop_eigh(large)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.default_rng(0)
A = rng.standard_normal((B, n, n)); A = 0.5 * (A + A.transpose(0, 2, 1))
At = torch.from_numpy(A).cuda()
lib = _lib.load()
raw = ctypes.CDLL(lib._name) if hasattr(lib, "_name") else lib
for blocked, cl, sym in ((1, 8, 0), (1, 4, 1), (1, 4, 0), (1, 2, 1), (1, 2, 0), (1, 1, 1)):
    lib.mop_priv_large_cluster(cl); raw.mop_priv_tridiag_cluster_sym(sym)
    ops.eigh(At, "large"); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        ev, V, st = ops.eigh(At, "large")
    e1.record(); torch.cuda.synchronize()
    ref = np.linalg.eigvalsh(A[:2])
    err = np.abs(ev[:2].cpu().numpy() - ref).max() / np.abs(ref).max()
    print(f"n={n} B={B} blocked={blocked} cluster={cl} sym={sym}: {e0.elapsed_time(e1)/2:.2f} ms per mop_eigh batch, eig err {err:.2e}, fallbacks {(st.cpu().numpy() & ops.ST_EIG_FALLBACK != 0).sum()}")
lib.mop_priv_large_cluster(int(os.environ.get("CL", "8")))
dbg = torch.zeros(B, 16, dtype=torch.int64, device="cuda")
f = raw.mop_priv_tridiag_cluster_timing; f.argtypes = [ctypes.c_void_p]; f.restype = ctypes.c_int
raw.mop_priv_tridiag_cluster_sym(int(os.environ.get("SYM", "1")))
for abl in [int(x) for x in os.environ.get("ABL", "0").split(",")]:
  raw.mop_priv_tridiag_cluster_ablate(abl)
  f(dbg.data_ptr())
  ops.eigh(At, "large"); torch.cuda.synchronize()
  f(None)
  d = dbg.cpu().numpy().astype(float)
  print("   interior cycles/col", d[:, 8].mean() / n, "edge cycles/col", d[:, 9].mean() / n, "interior blocks/col", d[:, 10].mean() / n, "edge blocks/col", d[:, 11].mean() / n)
  print("ablate", abl, "symv cycles per column", d[:, 0].mean() / n, "wait", d[:, 7].mean() / n, "total", d[:, :8].sum(1).mean() / n)
raw.mop_priv_tridiag_cluster_ablate(0)
names = ["symv + push", "c, panel products", "block_sum16", "mbarrier wait", "reduction B", "scalars + writes + barrier", "trailing update + barrier", "sym: wait for the slowest warp"]
for half, nm in ((0, "CTA 0"),):
    tot = d[:, half:half + 8].sum(1).mean()
    print(f"{nm}: mean cycles per matrix {tot:.0f} ({tot / n:.0f} per column)")
    for q, s in enumerate(names):
        print(f"   {s:32s} {d[:, half + q].mean():12.0f}  {100 * d[:, half + q].mean() / tot:5.1f} %")
