"""One launch of every producer kernel at its config size (ncu target): Swart / Fischer / fischerd3old (1024 x N=50),
Lindh / AFIR / bias restraints (8192 x N=24), RIC K matrix + back-transformation (1024 x N=24)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic
from multioptpy_b200.Utils.bond_connectivity import radii_array
from multioptpy_b200.ModelHessian.lindh import lindh_atom_params
from multioptpy_b200.ModelHessian.swart import swart_radii
from multioptpy_b200.ModelHessian.fischerd3old import d3_atom_params
from multioptpy_b200.Parameters.tables import covalent_radius

dev = torch.device("cuda:0")
T = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev) if dt is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
B, N = 1024, 50
elems = synthetic.elements(N)
xyz = np.stack([synthetic.grid_geometry(N, np.random.default_rng(100 + b), spacing=2.6, jitter=0.25) for b in range(B)])
xd = T(xyz)
ops.swart_hessian(xd, T(np.array(swart_radii(elems))))
ops.fischer_hessian(xd, T(np.array(radii_array(elems))))
ops.fischer_d3old_hessian(xd, d3_atom_params(elems))
B4, N4 = 8192, 24
x4, g4 = synthetic.conformer_batch(B4, N4, seed=4168)
el4 = synthetic.elements(N4, all_sulfur=True)
x4d = T(x4)
ops.lindh_hessian(x4d, lindh_atom_params(el4))
f1 = torch.arange(0, 12, dtype=torch.int32, device=dev); f2 = torch.arange(12, 24, dtype=torch.int32, device=dev)
rad = torch.tensor([covalent_radius(e) for e in el4], dtype=torch.float32, device=dev)
ops.afir(x4d, f1, f2, rad, torch.full((B4,), 100.0, dtype=torch.float64, device=dev))
terms = [(ops.BIAS_KEEP, [0], [5], 0.4, 2.1), (ops.BIAS_KEEP_ANGLE, [1, 0, 2], [], 0.3, 109.5),
         (ops.BIAS_KEEP_DIHEDRAL, [1, 0, 4, 5], [], 0.3, 1.0), (ops.BIAS_WELL, [0, 1, 2], [8, 9, 10], 0.01, 0.0, [1.0, 2.0, 9.0, 10.0])]
ops.bias_terms(x4d, ops.pack_bias_terms(terms, dev), len(terms))
torch.cuda.synchronize()
if os.environ.get("TIME"):   # CUDA-event medians per call (ms), host wrapper included
    def ms(f, reps=7):
        f(); torch.cuda.synchronize(); ts = []
        for _ in range(reps):
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        return float(np.median(ts))
    sr = T(np.array(swart_radii(elems))); fr = T(np.array(radii_array(elems))); dp = d3_atom_params(elems); lp = lindh_atom_params(el4)
    gam = torch.full((B4,), 100.0, dtype=torch.float64, device=dev)
    x24 = x4d[:1024].contiguous()
    kd = torch.rand(1024, N4 * (N4 - 1) // 2, dtype=torch.float64, device=dev)
    print(f"swart 1024xN50        {ms(lambda: ops.swart_hessian(xd, sr)):8.3f} ms")
    print(f"fischer 1024xN50      {ms(lambda: ops.fischer_hessian(xd, fr)):8.3f} ms")
    print(f"fischerd3old 1024xN50 {ms(lambda: ops.fischer_d3old_hessian(xd, dp)):8.3f} ms")
    print(f"lindh 8192xN24        {ms(lambda: ops.lindh_hessian(x4d, lp)):8.3f} ms")
    print(f"afir 8192xN24         {ms(lambda: ops.afir(x4d, f1, f2, rad, gam)):8.3f} ms")
print("ok")
