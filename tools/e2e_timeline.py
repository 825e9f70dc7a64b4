"""Timeline of one e2e step of bench.py (torch.profiler): start / duration / stream of every copy and kernel."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
from multioptpy_b200 import ops, synthetic
import bench
B = 1024; n = 150; dev = torch.device("cuda:0"); f64 = torch.float64
x0, H0, g0, rngs = bench.make_inputs(B, 0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
st = ops.new_rsirfo_state(B, 0.5, dev); zero = torch.zeros(B, dtype=f64, device=dev)
m_id = ops.resolve_update_method("rsirfo_bfgs")
out = ops.rsirfo_step(T(H0), T(x0), T(g0), T(g0), st, method=m_id, Be=zero)
mv0 = out["move"].cpu().numpy()
x1 = np.empty_like(x0); g1 = np.empty_like(g0)
for b in range(B):
    x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
sizes = [int(v) for v in os.environ.get("SPLIT", "256,256,256,256").split(",")]
bounds = np.concatenate([[0], np.cumsum(sizes)]); cb = max(sizes); ns = min(len(sizes), int(os.environ.get("NS", "4")))
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
hH, hx1, hg1, hx0, hg0 = pin(H0), pin(x1), pin(g1), pin(x0), pin(g0)
hBe = pin(np.full(B, -1e-3)); hst = st.cpu().pin_memory()
h_move = torch.empty(B, n, dtype=f64).pin_memory(); h_Hout = torch.empty(B, n, n, dtype=f64).pin_memory()
streams = [torch.cuda.Stream(dev) for _ in range(ns)]
dbuf = [dict(H=torch.empty(cb, n, n, dtype=f64, device=dev), x1=torch.empty(cb, n, dtype=f64, device=dev),
             g1=torch.empty(cb, n, dtype=f64, device=dev), x0=torch.empty(cb, n, dtype=f64, device=dev),
             g0=torch.empty(cb, n, dtype=f64, device=dev), Be=torch.empty(cb, dtype=f64, device=dev),
             st=torch.empty(cb, ops.RSIRFO_STATE, dtype=f64, device=dev), outs={}) for _ in range(ns)]
def step():
    for c in range(len(sizes)):
        s = streams[c % ns]; d = dbuf[c % ns]; lo, hi = int(bounds[c]), int(bounds[c + 1]); m = hi - lo; sl = slice(lo, hi)
        with torch.cuda.stream(s):
            d["H"][:m].copy_(hH[sl], non_blocking=True); d["x1"][:m].copy_(hx1[sl], non_blocking=True)
            d["g1"][:m].copy_(hg1[sl], non_blocking=True); d["x0"][:m].copy_(hx0[sl], non_blocking=True)
            d["g0"][:m].copy_(hg0[sl], non_blocking=True); d["Be"][:m].copy_(hBe[sl], non_blocking=True)
            d["st"][:m].copy_(hst[sl], non_blocking=True)
            o = ops.rsirfo_step(d["H"][:m], d["x1"][:m], d["g1"][:m], d["g1"][:m], d["st"][:m], method=m_id,
                                x_prev=d["x0"][:m], g_prev=d["g0"][:m], Be=d["Be"][:m], out=d["outs"].get(m))
            d["outs"][m] = o
            h_move[sl].copy_(o["move"], non_blocking=True); h_Hout[sl].copy_(d["H"][:m], non_blocking=True)
    for s in streams: s.synchronize()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/e2e_trace.json")
tr = json.load(open("/tmp/e2e_trace.json"))["traceEvents"]
ev = [e for e in tr if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and e.get("dur", 0) > 20]
t0 = min(e["ts"] for e in ev)
for e in sorted(ev, key=lambda e: e["ts"]):
    print(f"{(e['ts']-t0)/1e3:8.3f} ms +{e['dur']/1e3:7.3f}  stream {e['args'].get('stream')}  {e['name'][:60]}")
print("span ms", (max(e["ts"] + e["dur"] for e in ev) - t0) / 1e3)
