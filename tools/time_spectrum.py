"""CUDA-event timing of the spectral step (packed tridiagonalisation + k_spectrum_step), C2 batch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, _lib
import bench
lib = _lib.load(); B = int(os.environ.get("DIAG_B", "1024"))
x0, H0, g0, rngs = bench.make_inputs(B, 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
Hp, gp, _ = ops.project_trrot(T(H0), T(x0), g=T(g0))
st = ops.new_rsirfo_state(B, 0.5, torch.device("cuda:0")); zero = torch.zeros(B, dtype=torch.float64, device="cuda")
g0d = T(g0)
for spectrum in (1, 0):
    lib.mop_debug_tri_spectrum(spectrum)
    out = None
    for _ in range(3): out = ops.rsirfo_spectral_step(Hp, gp, g0d, st.clone(), Be=zero, out=out)
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sts = [st.clone() for _ in range(10)]
    a.record()
    for j in range(10): out = ops.rsirfo_spectral_step(Hp, gp, g0d, sts[j], Be=zero, out=out)
    b_.record(); torch.cuda.synchronize()
    print(f"k_spectrum_step {'on' if spectrum else 'off (k_eigh_tridiag prefactored)'}: {a.elapsed_time(b_)/10:.3f} ms per spectral step")
