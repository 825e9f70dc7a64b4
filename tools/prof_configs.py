"""Per-kernel device times (torch.profiler) of one iteration of the per-config chains of bench_configs.py:
   python tools/prof_configs.py c4 24 | c4 8 | c5 | c3"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import bench_configs as bc
from multioptpy_b200 import ops, synthetic

dev = "cuda:0"
which = sys.argv[1]


def report(prof):
    tot = 0.0
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total):
        if e.device_time_total > 0:
            tot += e.device_time_total
            print(f"{e.device_time_total/1e3:10.3f} ms  x{e.count:<3d} {e.key[:110]}")
    print(f"{tot/1e3:10.3f} ms  total")


if which == "c4":
    natoms = int(sys.argv[2]); B = int(os.environ.get("DIAG_B", "8192"))
    xyz, g = synthetic.conformer_batch(B, natoms, seed=4000 + 7 * natoms)
    ch = bc.C4Chain(xyz, g, dev)
    r = ch.two_iterations()
    x1g = r["x1"].reshape(B, natoms, 3).contiguous()
    H = r["H_model"].clone(); st0 = ops.new_rsirfo_state(B, 0.5, dev)
    ch.step(H, r["x0"], ch.g, r["E_afir"], r["g_afir"], r["H_afir"], st0)
    def it():
        H.copy_(r["H_model"]); st = st0.clone()
        Eb1, gb1, Hb1 = ch.afir(x1g)
        ch.step(H, r["x1"], r["g1"], Eb1, gb1, Hb1, st, x_prev=r["x0"], g_prev=ch.g, dE=1e-3)
    it(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        it(); ch.lindh(ch.xyz); torch.cuda.synchronize()
    report(prof)
elif which == "c5":
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        rec = bc.c5_record(dev, 0, 1, None, steps=1)
        torch.cuda.synchronize()
    report(prof)
    print(rec["ms_per_step"])
elif which == "c3":
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        rec = bc.c3_record(dev, 0, 1, None, steps=1)
        torch.cuda.synchronize()
    report(prof)
    print(rec["ms_per_iteration"])
elif which == "sp":   # phase cycles of k_spectrum_step on the C4 chain (iteration 1)
    from multioptpy_b200 import _lib
    lib = _lib.load()
    natoms = int(sys.argv[2]); B = int(os.environ.get("DIAG_B", "8192"))
    xyz, g = synthetic.conformer_batch(B, natoms, seed=4000 + 7 * natoms)
    ch = bc.C4Chain(xyz, g, dev)
    r = ch.two_iterations()
    x1g = r["x1"].reshape(B, natoms, 3).contiguous()
    H = r["H_model"].clone(); st0 = ops.new_rsirfo_state(B, 0.5, dev)
    ch.step(H, r["x0"], ch.g, r["E_afir"], r["g_afir"], r["H_afir"], st0)
    Eb1, gb1, Hb1 = ch.afir(x1g)
    dbg = torch.zeros(B, 16, dtype=torch.int64, device=dev)
    lib.mop_priv_spectrum_timing(dbg.data_ptr())
    H.copy_(r["H_model"]); st = st0.clone()
    o = ch.step(H, r["x1"], r["g1"], Eb1, gb1, Hb1, st, x_prev=r["x0"], g_prev=ch.g, dE=1e-3)
    torch.cuda.synchronize()
    lib.mop_priv_spectrum_timing(0)
    d = dbg.cpu().numpy().astype(float)
    names = ["load/scale/split", "eigenvalues", "twisted vectors", "cluster CGS2", "gamma + rfo_core", "y = Z c", "Q y"]
    tot = d[:, :7].sum(1)
    print("mean cycles per structure", tot.mean(), "max", tot.max())
    for q, nm in enumerate(names):
        print(f"  {nm:20s} mean {d[:, q].mean():10.0f}  p50 {np.median(d[:, q]):10.0f}  p99 {np.percentile(d[:, q], 99):10.0f}  max {d[:, q].max():10.0f}")
    stt = o["status"].cpu().numpy()
    print("alpha search fraction", np.mean((stt & ops.ST_ALPHA_SEARCH) != 0), "unstable", np.mean((stt & ops.ST_ALPHA_UNSTABLE) != 0))
