"""Spectral step (tridiagonalisation + spectrum + RFO step) timing at other sizes: B x natoms from the command line,
all four kernel combinations (diagnostics for the dispatch by n)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic, _lib
lib = _lib.load()
for B, natoms in [(8192, 24), (8192, 8), (2048, 30), (1024, 50)]:
    n = 3 * natoms
    x0, H0, g0, rngs = synthetic.batch(4, min(B, 256), natoms)
    rep = B // x0.shape[0]
    x0, H0, g0 = np.tile(x0, (rep, 1)), np.tile(H0, (rep, 1, 1)), np.tile(g0, (rep, 1))
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    Hp, gp, _ = ops.project_trrot(T(H0), T(x0), g=T(g0))
    st = ops.new_rsirfo_state(B, 0.5, torch.device("cuda:0")); zero = torch.zeros(B, dtype=torch.float64, device="cuda")
    g0d = T(g0); ref = None
    for spectrum, rowwarp in ((1, 1), (1, 0), (0, 1), (0, 0)):
        lib.mop_debug_tri_spectrum(spectrum); lib.mop_debug_packed_rowwarp(rowwarp)
        out = None
        for _ in range(3): out = ops.rsirfo_spectral_step(Hp, gp, g0d, st.clone(), Be=zero, out=out)
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sts = [st.clone() for _ in range(5)]
        a.record()
        for j in range(5): out = ops.rsirfo_spectral_step(Hp, gp, g0d, sts[j], Be=zero, out=out)
        b_.record(); torch.cuda.synchronize()
        mv = out["move"].clone()
        if ref is None: ref = mv
        print(f"B={B} n={n}: spectrum kernel {spectrum} rowwarp {rowwarp}: {a.elapsed_time(b_)/5:.3f} ms, max rel diff vs default {float(((mv-ref).norm(dim=1)/ref.norm(dim=1)).max()):.1e}")
lib.mop_debug_tri_spectrum(1); lib.mop_debug_packed_rowwarp(1)
