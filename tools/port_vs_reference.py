"""Speed of the oracle port (oracle/np_oracle.py RSIRFOOracle.run) relative to the UNMODIFIED reference
(multioptpy.Optimizer.rsirfo.RSIRFO.run) on the configs[1] inputs, one core, step-1 calls (update active).
Needs /root/reference, so it runs in the build container; writes profiles/port_vs_reference.json, which bench.py
attaches to its cpu_baseline record (the GPU box only has the port).
    python tools/port_vs_reference.py [nstruct]"""
import contextlib, io, json, os, sys, time
os.environ.setdefault("OMP_NUM_THREADS", "1"); os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from multioptpy_b200 import synthetic
from oracle import np_oracle as O, ref_shim

nstruct = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rs = ref_shim.ref("Optimizer.rsirfo")
col = lambda a: np.asarray(a, float).reshape(-1, 1).copy()
t_ref = t_port = 0.0
worst = 0.0
for b in range(nstruct):
    x0, H0, g0, rng = synthetic.structure(2, b, 50)
    o = O.RSIRFOOracle(method="rsirfo_bfgs", saddle_order=0); o.set_hessian(H0.copy()); o.set_bias_hessian(None)
    m0 = o.run(x0, g0, g0, None, None, 0.0)
    x1, g1 = synthetic.second_point(x0, H0, g0, m0, rng)
    t0 = time.perf_counter(); m1 = o.run(x1, g1, g1, x0, g0, -1e-3); t_port += time.perf_counter() - t0
    with contextlib.redirect_stdout(io.StringIO()):
        r = rs.RSIRFO(method="rsirfo_bfgs", saddle_order=0)
        r.set_hessian(H0.copy()); r.set_bias_hessian(np.zeros_like(H0))
        r.run(col(x0), col(g0), [], [], 0.0, 0.0, [], col(x0), col(g0), [])
        t0 = time.perf_counter()
        mr = r.run(col(x1), col(g1), col(g0), col(x0), -1e-3, 0.0, col(m0), col(x0), col(g1), col(g0))
        t_ref += time.perf_counter() - t0
    worst = max(worst, float(np.linalg.norm(np.asarray(mr).ravel() - m1) / np.linalg.norm(m1)))
out = {"workload": "configs[1] inputs (N=50, rsirfo_bfgs), step-1 calls, one core, 1 BLAS thread", "structures": nstruct,
       "reference_ms_per_step": 1e3 * t_ref / nstruct, "port_ms_per_step": 1e3 * t_port / nstruct,
       "port_speed_over_reference": t_ref / t_port, "max_rel_diff_port_vs_reference": worst,
       "numpy": np.__version__}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "port_vs_reference.json"), "w"), indent=1)
print(json.dumps(out))
