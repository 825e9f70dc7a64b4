"""How do the eigensolvers behave on TR/ROT-projected Hessians (exact six-fold zero cluster)? diagnostics"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic
from oracle import np_oracle as O
for natoms, B in ((50, 64), (200, 16)):
    n = 3 * natoms
    Hs = []
    for b in range(4):
        rng = np.random.default_rng(b)
        x = synthetic.grid_geometry(natoms, rng).reshape(-1)
        Hs.append(O.project_hessian_trrot(synthetic.spd_hessian(n, rng, neg_lowest=True), x))
    A = torch.from_numpy(np.tile(np.stack(Hs), (B // 4, 1, 1))).cuda()
    ops.eigh(A); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ev, V, st = ops.eigh(A); e1.record(); torch.cuda.synchronize()
    fb = int((st & ops.ST_EIG_FALLBACK).ne(0).sum())
    Vn = V[0].cpu().numpy(); evn = ev[0].cpu().numpy()
    res = np.abs(Hs[0] @ Vn.T - Vn.T * evn).max(); orth = np.abs(Vn @ Vn.T - np.eye(n)).max()
    print(f"n={n} B={B}: eigh {e0.elapsed_time(e1):.2f} ms, fallbacks {fb}/{B}, residual {res:.1e}, orthogonality {orth:.1e}")
