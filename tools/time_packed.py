import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic, _lib
import bench
lib = _lib.load()
for B in (148, 296, 1024):
    x0, H0, g0, rngs = bench.make_inputs(B, 0)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    Hp, gp, _ = ops.project_trrot(T(H0), T(x0), g=T(g0))
    st = ops.new_rsirfo_state(B, 0.5, torch.device("cuda:0")); zero = torch.zeros(B, dtype=torch.float64, device="cuda")
    for thr in (256, 320, 512):
        lib.mop_debug_packed_threads(thr)
        ops.rsirfo_spectral_step(Hp, gp, T(g0), st.clone(), Be=zero); torch.cuda.synchronize()
        dbg = torch.zeros(B, 16, dtype=torch.int64, device="cuda")
        lib.mop_debug_packed_timing(dbg.data_ptr())
        ops.rsirfo_spectral_step(Hp, gp, T(g0), st.clone(), Be=zero); torch.cuda.synchronize()
        lib.mop_debug_packed_timing(None)
        d = dbg.cpu().numpy().astype(float)
        a = d[:, :6].mean(0) / 148; c = d[:, 8:14].mean(0) / 148
        print(f"B={B} thr={thr}: per column cycles thread0 v={a[0]:.0f} symv={a[1]:.0f} redC={a[2]:.0f} w={a[3]:.0f} upd={a[4]:.0f} redE={a[5]:.0f} total={a.sum():.0f} | thread96 symv={c[1]:.0f} upd={c[4]:.0f} total={c.sum():.0f}")
