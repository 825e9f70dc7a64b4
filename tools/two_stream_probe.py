"""Probe: does splitting the C2 batch over two streams (each half a complete mop_rsirfo_step on its own workspace) beat
one launch sequence?  The halves' partial waves and under-filled late stages can fill each other's idle SMs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, synthetic, _lib
import bench
B = 1024; dev = torch.device("cuda:0")
x0, H0, g0, rngs = bench.make_inputs(B, 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
m = ops.resolve_update_method("rsirfo_bfgs")
zero = torch.zeros(B, dtype=torch.float64, device=dev)
H = T(H0); st = ops.new_rsirfo_state(B, 0.5, dev)
out = ops.rsirfo_step(H, T(x0), T(g0), T(g0), st, method=m, Be=zero)
mv0 = out["move"].cpu().numpy()
x1 = np.empty_like(x0); g1 = np.empty_like(g0)
for b in range(B):
    x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
X1, G1, X0, G0 = T(x1), T(g1), T(x0), T(g0)
Be = zero - 1e-3
ncopy = 6
Hs = [H.clone() for _ in range(ncopy)]; sts = [st.clone() for _ in range(ncopy)]

def run(parts, streams):
    main = torch.cuda.current_stream(dev)
    bounds = np.linspace(0, B, parts + 1).astype(int)
    def step(j):
        if parts == 1:
            return ops.rsirfo_step(Hs[j], X1, G1, G1, sts[j], method=m, x_prev=X0, g_prev=G0, Be=Be)
        ev0 = torch.cuda.Event(); ev0.record(main)
        for p in range(parts):
            s = streams[p]; sl = slice(int(bounds[p]), int(bounds[p + 1]))
            s.wait_event(ev0)
            with torch.cuda.stream(s):
                ops.rsirfo_step(Hs[j][sl], X1[sl], G1[sl], G1[sl], sts[j][sl], method=m, x_prev=X0[sl], g_prev=G0[sl], Be=Be[sl])
                e = torch.cuda.Event(); e.record(s)
            main.wait_event(e)
    for j in range(3):
        for k in range(ncopy):
            Hs[k].copy_(H); sts[k].copy_(st)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for k in range(ncopy):
            step(k)
        b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / ncopy

streams = [torch.cuda.Stream(dev) for _ in range(4)]
for parts in (1, 2, 3, 4):
    print(f"parts {parts}: {run(parts, streams):.3f} ms per 1024-structure step")
