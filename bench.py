#!/usr/bin/env python
"""bench.py — batched RS-I-RFO + Hessian-update steps/s (FP64) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): B = 1024 independent structures per GPU,
N = 50 atoms (n = 150), `rsirfo_bfgs`, FP64; the timed unit is "step 1" of a
two-step sequence (history present, Hessian update active, s.y > 0), SURVEY §8d.
One bench "step" = one pass of the hot path over the whole batch; `value` counts
structure-steps per second over all ranks (weak scaling: per-GPU batch fixed).

`e2e` goes through the C ABI with HOST buffers every step: the Hessian batch travels as packed lower triangles
(mop_rsirfo_step_packed, 92.6 MB instead of 184 MB per 1024 structures), geometry / gradients / state beside it, the
move vectors and status words come back; the updated Hessians stay on the device (RSIRFO keeps its Hessian between
run() calls; get_hessian() unpacks on demand) and are read back ONCE after the timed loop for the check.
`per_config` carries the other named shapes of BASELINE.json (configs[2] NEB 64 x 30, configs[3] 8192 x N=24 / N=8,
configs[4] 256 x 3N=600), each with value, parity_vs_oracle and roofline (bench_configs.py); --no-per-config skips them.

Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "batched RFO+Hessian-update steps/sec (FP64)"
UNIT = "structure-steps/s"
NATOMS = 50
BATCH = 1024
METHOD = "rsirfo_bfgs"
CONFIG_ID = 2


def workload_config(B, extra=None):
    cfg = {"workload": f"configs[1]: synthetic batch {B} independent RFO-BFGS minimisation steps, "
                       f"N={NATOMS} atoms (3N={3 * NATOMS}), FP64, step 1 of 2 (update active)",
           "batch_per_gpu": B, "natoms": NATOMS, "n": 3 * NATOMS, "method": METHOD,
           "saddle_order": 0, "parallelism": "independent structures sharded by rank, no collective",
           "l2": "inputs larger than L2 (per-step Hessian batch 184 MB, rotating pristine copies)"}
    if extra:
        cfg.update(extra)
    return cfg


# --------------------------------------------------------------------------- inputs
def make_inputs(B, rank):
    """Step-1 inputs for B structures (seed = 1000*config + global index).  Step 0 is
    run through the NumPy oracle on a few structures and through the CUDA path on all
    of them by the caller; here only the seeded raw data is produced."""
    from multioptpy_b200 import synthetic
    x0 = np.empty((B, 3 * NATOMS)); H0 = np.empty((B, 3 * NATOMS, 3 * NATOMS)); g0 = np.empty((B, 3 * NATOMS))
    rngs = []
    for b in range(B):
        x0[b], H0[b], g0[b], rng = synthetic.structure(CONFIG_ID, rank * B + b, NATOMS)
        rngs.append(rng)
    return x0, H0, g0, rngs


# ---------------------------------------------------------------------- CPU baseline
def _cpu_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    import contextlib, io
    from oracle import np_oracle as O
    from multioptpy_b200 import synthetic
    idx_list = args
    prepared = []
    for gi in idx_list:
        x0, H0, g0, rng = synthetic.structure(CONFIG_ID, gi, NATOMS)
        o = O.RSIRFOOracle(method=METHOD, saddle_order=0)
        o.set_hessian(H0.copy()); o.set_bias_hessian(None)
        m = o.run(x0, g0, g0, None, None, 0.0)
        x1, g1 = synthetic.second_point(x0, H0, g0, m, rng)
        prepared.append((o, x0, g0, x1, g1))
    t0 = time.perf_counter()
    for o, x0, g0, x1, g1 in prepared:
        o.run(x1, g1, g1, x0, g0, -1e-3)
    return time.perf_counter() - t0, len(prepared)


def cpu_baseline(sample, cores):
    """Oracle port (NumPy restatement of RSIRFO.run) on `cores` host processes, one
    BLAS thread each; returns structure-steps/s over the timed step-1 calls."""
    import multiprocessing as mp
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    chunks = [list(range(w, sample, cores)) for w in range(cores)]
    chunks = [c for c in chunks if c]
    ctx = mp.get_context("spawn")
    with ctx.Pool(len(chunks)) as pool:
        res = pool.map(_cpu_worker, chunks)
    wall = max(t for t, _ in res)          # workers run concurrently
    done = sum(k for _, k in res)
    return done / wall, len(chunks)


# ------------------------------------------------------------------------ clocks
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop_flag = [], set(), False
        self.max_mhz = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False
        self.t = None

    def _loop(self):
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap", 0x80: "hw_power_brake", 0x2: "applications_clocks", 0x100: "display"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=1)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = min(BATCH, max(cores * 8, 64))
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(min(sample, cores * 2), cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, used = cpu_baseline(sample, cores)
        vals.append(v)
    v = statistics.median(vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sample / v,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(BATCH),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": used, "kind": "port",
                             "sample": f"{sample} structures of the same seeded batch per step, step-1 calls timed, "
                                       f"one process per core, 1 BLAS thread each (oracle/np_oracle.py; "
                                       f"/root/reference is not present on the GPU box)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    from multioptpy_b200 import _lib, ops
    from multioptpy_b200.Optimizer.rsirfo import RSIRFO

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    lib = _lib.load()
    B, n = BATCH, 3 * NATOMS
    K, W = args.steps, args.warmup
    f64 = torch.float64

    # ---- inputs: step 0 on the device for every structure, then the step-1 points ----
    x0, H0, g0, rngs = make_inputs(B, rank)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    from multioptpy_b200 import synthetic
    opt0 = RSIRFO(method=METHOD, saddle_order=0, device=dev)
    H_d0 = T(H0)
    opt0.set_hessian(H_d0); opt0.set_bias_hessian(None)
    zero = torch.zeros(B, dtype=f64, device=dev)
    mv0 = opt0.run(T(x0), T(g0), B_e=zero, g=T(g0)).cpu().numpy().copy()
    state1 = opt0.state_tensor.clone()              # RSIRFO state after step 0
    x1 = np.empty_like(x0); g1 = np.empty_like(g0)
    for b in range(B):
        x1[b], g1[b] = synthetic.second_point(x0[b], H0[b], g0[b], mv0[b], rngs[b])
    x0_d, g0_d, x1_d, g1_d = T(x0), T(g0), T(x1), T(g1)
    Be1 = zero - 1e-3

    # parity spot check against the oracle (checker only, not timed, not shipped)
    from oracle import np_oracle as O
    chk = min(4, B)
    Hc = H_d0[:chk].clone(); stc = state1[:chk].clone()
    outc = ops.rsirfo_step(Hc, x1_d[:chk].contiguous(), g1_d[:chk].contiguous(), g1_d[:chk].contiguous(), stc,
                           method=ops.resolve_update_method(METHOD), x_prev=x0_d[:chk].contiguous(),
                           g_prev=g0_d[:chk].contiguous(), Be=Be1[:chk].contiguous())
    worst = 0.0
    for b in range(chk):
        o = O.RSIRFOOracle(method=METHOD, saddle_order=0)
        o.set_hessian(H0[b].copy()); o.set_bias_hessian(None)
        o.run(x0[b], g0[b], g0[b], None, None, 0.0)
        m = o.run(x1[b], g1[b], g1[b], x0[b], g0[b], -1e-3)
        worst = max(worst, float(np.linalg.norm(outc["move"][b].cpu().numpy() - m) / np.linalg.norm(m)))
    if not worst < 1e-10:
        raise SystemExit(f"bench.py: parity check failed before timing ({worst:.3e})")

    # ---- rotating pristine copies (each timed step sees an un-updated Hessian batch) ----
    ncopy = max(2, K + W)     # every step of the run sees a pristine (not yet updated) Hessian batch
    if ncopy * B * n * n * 8 > 60e9:
        raise SystemExit(f"bench.py: {K} + {W} pristine Hessian copies do not fit; use fewer steps")
    Hs = [H_d0.clone() for _ in range(ncopy)]
    sts = [state1.clone() for _ in range(ncopy)]
    method_id = ops.resolve_update_method(METHOD)
    out = None

    def one_step(i):
        nonlocal out
        j = i % ncopy
        out = ops.rsirfo_step(Hs[j], x1_d, g1_d, g1_d, sts[j], method=method_id, x_prev=x0_d,
                              g_prev=g0_d, Be=Be1, out=out)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        one_step(i)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        one_step(W + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    if dist is not None:
        t = torch.tensor([ms], dtype=f64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * K / (ms * 1e-3)

    # ---- e2e: host buffers -> public host-buffer API (multioptpy_b200.host_pipeline.HostStepPipeline: chunked
    # cudaMemcpyAsync + mop_rsirfo_step_packed_begin per chunk, mop_rsirfo_step_packed_finish once) -> host buffers;
    # every copy is inside the timed region (MOP_BENCH_E2E_SPLIT / _STREAMS override the chunking)
    from multioptpy_b200.host_pipeline import HostStepPipeline, pack_lower_host
    split_env = os.environ.get("MOP_BENCH_E2E_SPLIT", "")
    sizes = [int(v) for v in split_env.split(",")] if split_env else None
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    hH = pin(pack_lower_host(H0))                    # packed lower triangles, row i at i (i + 1) / 2
    hx1, hg1, hx0, hg0 = pin(x1), pin(g1), pin(x0), pin(g0)
    hBe = pin(np.full(B, -1e-3)); hst = state1.cpu().pin_memory()
    h_move = torch.empty(B, n, dtype=f64).pin_memory()
    h_stat = torch.empty(B, dtype=torch.int32).pin_memory()
    pipe = HostStepPipeline(B, n, method_id, device=dev, chunks=sizes,
                            nstream=int(os.environ.get("MOP_BENCH_E2E_STREAMS", "4")))
    sizes = [int(v) for v in np.diff(pipe.bounds)]
    nstream = len(pipe.streams)
    d2h = h_move.numel() * 8 + h_stat.numel() * 4

    def e2e_step():
        pipe.step(hx1, hg1, hg1, hst, h_move, h_stat, hH=hH, hx_prev=hx0, hg_prev=hg0, hBe=hBe)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    Ke = max(3, min(K, 10))
    for _ in range(Ke):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=f64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = world * B * Ke / e2e_s
    # the host-buffer path must reproduce the device-resident result of the timed steps
    ref_mv = out["move"].cpu().numpy()
    e2e_diff = float(np.max(np.linalg.norm(h_move.numpy() - ref_mv, axis=1) / np.linalg.norm(ref_mv, axis=1)))
    # lazy read-back of the updated Hessians (once, outside the timed loop): must equal the resident path's
    H_e2e = pipe.hessians()
    h2d = pipe.h2d_bytes
    H_res = Hs[(W + K - 1) % ncopy]
    e2e_hdiff = float(((H_e2e - H_res).flatten(1).norm(dim=1) / H_res.flatten(1).norm(dim=1)).max())
    e2e_ok = bool(np.isfinite(h_move.numpy()).all() and e2e_diff < 1e-12 and e2e_hdiff < 1e-12)

    # ---- e2e with the Hessian batch resident on the device (steady-state drop-in) -------
    d_x1, d_g1 = torch.empty_like(x1_d), torch.empty_like(g1_d)
    h_mv2 = torch.empty(B, n, dtype=f64).pin_memory()

    def e2e_resident(i):
        nonlocal out
        j = i % ncopy
        d_x1.copy_(hx1, non_blocking=True); d_g1.copy_(hg1, non_blocking=True)
        out = ops.rsirfo_step(Hs[j], d_x1, d_g1, d_g1, sts[j], method=method_id, x_prev=x0_d, g_prev=g0_d,
                              Be=Be1, out=out)
        h_mv2.copy_(out["move"], non_blocking=True)
        torch.cuda.synchronize()

    e2e_resident(0)
    barrier()
    t0 = time.perf_counter()
    for i in range(Ke):
        e2e_resident(i + 1)
    res_s = time.perf_counter() - t0
    e2e_res_val = world * B * Ke / res_s

    line = None
    if rank == 0:
        # ---- FP64 peak probe + roofline of the dominant kernel (eigensolver) -------------
        probe = torch.empty(148 * 16 * 256, dtype=f64, device=dev)
        iters = 4096
        for _ in range(2):
            _lib.check(lib.mop_priv_bench_dfma(148 * 16, iters, probe.data_ptr(), torch.cuda.current_stream().cuda_stream))
        best = 1e30
        for _ in range(5):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.mop_priv_bench_dfma(148 * 16, iters, probe.data_ptr(), torch.cuda.current_stream().cuda_stream))
            b_.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b_))
        fp64_peak = 148 * 16 * 256 * iters * 64 * 2 / (best * 1e-3) / 1e12   # TFLOP/s

        reps = max(3, min(K, 10))
        # The dominant kernel pair IS the step: k_tridiag_blk<5, fused> (update + write-back + projection +
        # tridiagonalisation) and k_spectrum_step (spectrum, eigenvectors of T, RFO step) are the only launches of
        # the timed region that do work (the three fallback launches that follow are empty), so its duration is the
        # device time of the timed loop itself, ms / K, and its algorithmic work W_F = 9 n^3 + 40 n^2 per structure
        # (one eigendecomposition with vectors + the O(n^2) update / projection / step algebra).
        eig_ms = ms / K
        WF = 9.0 * n ** 3 + 40.0 * n * n
        eig_tflops = B * WF / (eig_ms * 1e-3) / 1e12
        executed_flops = B * (4.0 / 3.0 * n ** 3 + 60.0 * n * n)   # Householder reduction + the O(n^2) rest
        # streaming update kernel against the HBM roofline
        sd = (x1_d - x0_d).contiguous(); yd = (g1_d - g0_d).contiguous()
        Hu = H_d0.clone()
        ops.hessian_update(Hu, sd, yd, method_id, inplace=True, rsirfo_guards=True)
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for j in range(reps):
            ops.hessian_update(Hs[j % ncopy], sd, yd, method_id, inplace=True, rsirfo_guards=True)
        b_.record(); torch.cuda.synchronize()
        upd_ms = a.elapsed_time(b_) / reps
        upd_gbs = B * (16.0 * n * n + 32.0 * n) / (upd_ms * 1e-3) / 1e9
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"

        cores = os.cpu_count() or 1
        sample = min(B, max(64, cores * 8))      # >= 8 structures per process: the 64-structure sample was noisy
        cpu_val, used = cpu_baseline(sample, cores)
        traffic = None          # DRAM bytes of the dominant kernel pair per launch, from the committed ncu capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
            if tr.get("batch") == B and tr.get("n") == n:
                traffic = tr["dominant_pair_bytes_per_launch"]
        except Exception:
            pass

        step_tflops_alg = value / world * (9.0 * n ** 3 + 40.0 * n * n) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(B),
            "parity_vs_oracle": worst, "eigh": "auto",
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": f"pinned host buffers through multioptpy_b200.host_pipeline.HostStepPipeline.step: Hessian batch in as "
                            f"packed lower triangles every step (chunks of {sizes} structures on {nstream} streams, "
                            f"mop_rsirfo_step_packed_begin per chunk while the next chunk is in flight, "
                            f"mop_rsirfo_step_packed_finish once), moves + status out; updated Hessians stay on the device "
                            f"(read back once after the loop for the check)",
                    "matches_resident_path": e2e_ok, "max_rel_diff_vs_resident": e2e_diff,
                    "hessian_max_rel_diff_vs_resident": e2e_hdiff},
            "e2e_hessian_resident": {"value": e2e_res_val, "unit": UNIT,
                                     "h2d_bytes_per_step": 2 * B * n * 8, "d2h_bytes_per_step": B * n * 8},
            # fused update + projection + staged tridiagonalisation (5 launches at n = 150), spectrum + step, 3 (empty) fallbacks
            "gpu_launches": ((int(_lib.load().mop_tridiag_stage_count(n)) if B > 296 else 1) + 4) * K,
            "roofline": {"bound": "fp64",
                         "kernel": "k_tridiag_blk<5, fused> (+ its denser continuation stages <4>, <3>, <2>, <1>) + k_spectrum_step = the "
                                   "timed step: Hessian update, write-back, TR/ROT projection and blocked DMMA tridiagonalisation, then spectrum, "
                                   "eigenvectors of T and the RFO step in the eigenbasis (the other launches of the step are "
                                   "empty fallbacks); duration = ms_per_step of the timed region",
                         "achieved": eig_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": eig_tflops / fp64_peak, "traffic": traffic,
                         "traffic_source": "profiles/r2_traffic.json (ncu --set full, dram__bytes_read + dram__bytes_write)",
                         "algorithmic_flops_per_launch": B * WF, "kernel_ms": eig_ms,
                         "executed_flops_per_launch_estimate": executed_flops,
                         "executed_tflops_estimate": executed_flops / (eig_ms * 1e-3) / 1e12,
                         "peak_source": "in-run DFMA probe (mop_priv_bench_dfma); MEASURED_PEAKS.json has no FP64 figure",
                         "whole_step_algorithmic_tflops": step_tflops_alg,
                         "whole_step_frac": step_tflops_alg / fp64_peak},
            "roofline_hbm": {"bound": "hbm", "kernel": "Hessian update, multi-CTA path (k_upd_matvec + k_upd_scalars + k_upd_apply)",
                             "achieved": upd_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": upd_gbs / hbm_peak,
                             "traffic": None, "kernel_ms": upd_ms, "peak_source": hbm_src},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": used, "kind": "port",
                             "sample": f"{sample} structures of the same batch, step-1 calls timed, one process per "
                                       f"core, 1 BLAS thread each (oracle/np_oracle.py)",
                             "port_vs_reference": port_factor()},
        }
    if not args.no_per_config:
        import bench_configs as bc
        recs = {}
        fpk = fp64_peak if rank == 0 else None
        for key, fn in (("configs[2]", lambda: bc.c3_record(dev, rank, world, dist)),
                        ("configs[3] N=24", lambda: bc.c4_record(dev, rank, world, dist, 24)),
                        ("configs[3] N=8", lambda: bc.c4_record(dev, rank, world, dist, 8)),
                        ("configs[4]", lambda: bc.c5_record(dev, rank, world, dist, fp64_peak=fpk))):
            if key == "configs[2]" and 64 % world:
                continue
            r = fn()
            if rank == 0:
                recs[key] = r
        if line is not None:
            line["per_config"] = recs
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def port_factor():
    """Speed of the oracle port relative to the UNMODIFIED reference on the same inputs, measured where /root/reference
    exists by tools/port_vs_reference.py (committed result: profiles/port_vs_reference.json)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "port_vs_reference.json")))
    except Exception:
        return None


# ------------------------------------------------------------------- config 5 (non-default)
def _cpu_worker_c5(idx_list):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import np_oracle as O
    from multioptpy_b200 import synthetic
    prepared = []
    for gi in idx_list:
        x0, H0, g0, rng = synthetic.structure(5, gi, 200, saddle=True)
        o = O.RSPRFOOracle(method="rsprfo_bofill", saddle_order=1)
        o.set_hessian(H0)
        m = o.run(x0, g0, None, None, 0.0, None)
        x1, g1 = synthetic.second_point(x0, H0, g0, m, rng)
        prepared.append((o, x0, g0, x1, g1, m))
    t0 = time.perf_counter()
    for o, x0, g0, x1, g1, m in prepared:
        o.run(x1, g1, x0, g0, -1e-3, m)
    return time.perf_counter() - t0, len(prepared)


def run_b200_c5(args, rank, world, local_rank):
    """BASELINE configs[4]: P-RFO saddle search with Bofill update, N = 200 atoms (3N = 600), batch 256
    sharded over the ranks (strong scaling).  Selected with --workload c5; the default line is configs[1]."""
    import torch
    from multioptpy_b200 import _lib, ops, synthetic
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the hot path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    natoms, total = 200, 256
    n = 3 * natoms
    B = total // world
    K, W = args.steps, args.warmup
    f64 = torch.float64
    nuniq = min(B, 16)      # distinct seeded structures, tiled over the shard (generation cost on the host)
    xs, Hs_, gs, rngs = [], [], [], []
    for b in range(nuniq):
        x0, H0, g0, rng = synthetic.structure(5, rank * nuniq + b, natoms, saddle=True)
        xs.append(x0); Hs_.append(H0); gs.append(g0); rngs.append(rng)
    rep = (B + nuniq - 1) // nuniq
    tile = lambda a: torch.from_numpy(np.stack(a)).repeat(rep, *([1] * (np.stack(a).ndim - 1)))[:B].contiguous().to(dev)
    H_d0, x0_d, g0_d = tile(Hs_), tile(xs), tile(gs)
    z = lambda *sh: torch.zeros(*sh, dtype=f64, device=dev)
    def fresh_state():
        st = dict(state=z(B, ops.PRFO_STATE), prev_grad=z(B, n), prev_move=z(B, n), ts_vec=z(B, n))
        st["state"][:, 0] = 0.1
        return st
    st0 = fresh_state()
    out0 = ops.rsprfo_step(H_d0.clone(), x0_d, g0_d, st0, method=23, saddle_order=1, Be=z(B))
    mv0 = out0["move"].clone()
    x1_d = x0_d - mv0
    g1_d = g0_d + torch.einsum("bij,bj->bi", H_d0, x1_d - x0_d)
    Be1 = z(B) - 1e-3
    # parity spot check against the oracle (checker only)
    from oracle import np_oracle as O
    o = O.RSPRFOOracle(method="rsprfo_bofill", saddle_order=1); o.set_hessian(Hs_[0])
    m0 = o.run(xs[0], gs[0], None, None, 0.0, None)
    worst = float(np.linalg.norm(mv0[0].cpu().numpy() - m0) / np.linalg.norm(m0))
    m1 = o.run(x1_d[0].cpu().numpy(), g1_d[0].cpu().numpy(), xs[0], gs[0], -1e-3, m0)
    ncopy = 3
    Hc = [H_d0.clone() for _ in range(ncopy)]
    stc = [{k: v.clone() for k, v in st0.items()} for _ in range(ncopy)]
    st_ref = {k: v.clone() for k, v in st0.items()}
    out = None

    def one_step(i):
        nonlocal out
        j = i % ncopy
        Hc[j].copy_(H_d0)
        for k in st_ref:
            stc[j][k].copy_(st_ref[k])
        out = ops.rsprfo_step(Hc[j], x1_d, g1_d, stc[j], method=23, saddle_order=1, x_prev=x0_d, Bg_prev=g0_d,
                              pre_move=mv0, Be=Be1, out=out)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        one_step(i)
    worst = max(worst, float(np.linalg.norm(out["move"][0].cpu().numpy() - m1) / np.linalg.norm(m1)))
    if not worst < 1e-10:
        raise SystemExit(f"bench.py: parity check failed before timing ({worst:.3e})")
    barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(K):
        one_step(W + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # e2e: geometry / gradient from pinned host memory, step back to the host; Hessians stay resident
    hx1, hg1 = x1_d.cpu().pin_memory(), g1_d.cpu().pin_memory()
    h_mv = torch.empty(B, n, dtype=f64).pin_memory()
    dx, dg = torch.empty_like(x1_d), torch.empty_like(g1_d)
    def e2e(i):
        nonlocal out
        j = i % ncopy
        Hc[j].copy_(H_d0)
        for k in st_ref:
            stc[j][k].copy_(st_ref[k])
        dx.copy_(hx1, non_blocking=True); dg.copy_(hg1, non_blocking=True)
        out = ops.rsprfo_step(Hc[j], dx, dg, stc[j], method=23, saddle_order=1, x_prev=x0_d, Bg_prev=g0_d,
                              pre_move=mv0, Be=Be1, out=out)
        h_mv.copy_(out["move"], non_blocking=True)
        torch.cuda.synchronize()
    e2e(0); barrier()
    Ke = max(3, min(K, 5))
    t0 = time.perf_counter()
    for i in range(Ke):
        e2e(i + 1)
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([ms, e2e_s], dtype=f64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
    value = world * B * K / (ms * 1e-3)
    line = None
    if rank == 0:
        lib = _lib.load()
        probe = torch.empty(148 * 16 * 256, dtype=f64, device=dev)
        best = 1e30
        for _ in range(6):
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.mop_priv_bench_dfma(148 * 16, 4096, probe.data_ptr(), torch.cuda.current_stream().cuda_stream))
            b_.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b_))
        fp64_peak = 148 * 16 * 256 * 4096 * 64 * 2 / (best * 1e-3) / 1e12
        WF = 9.0 * n ** 3 + 4.0 / 3.0 * n ** 3 + 40.0 * n * n     # SURVEY §8d, RSPRFO
        step_tf = value / world * WF / 1e12
        cores = os.cpu_count() or 1
        sample = max(cores, 8)
        import multiprocessing as mp
        for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
            os.environ[k] = "1"          # before the workers import NumPy
        chunks = [c for c in ([list(range(w, sample, cores)) for w in range(cores)]) if c]
        with mp.get_context("spawn").Pool(len(chunks)) as pool:
            res = pool.map(_cpu_worker_c5, chunks)
        cpu_val = sum(k for _, k in res) / max(t for t, _ in res)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "configs[4]: P-RFO saddle search with Bofill update, N=200 atoms (3N=600), "
                                       f"batch 256 sharded over {world} GPU(s), step 1 of 2 (update active)",
                           "batch_per_gpu": B, "natoms": natoms, "n": n, "method": "rsprfo_bofill", "saddle_order": 1,
                           "distinct_structures_per_gpu": nuniq, "parity_vs_oracle": worst,
                           "l2": "inputs larger than L2 (Hessian batch 737 MB / world, fresh copy every step)"},
                "clocks": clocks,
                "e2e": {"value": world * B * Ke / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * B * n * 8,
                        "d2h_bytes_per_step": B * n * 8, "note": "Hessian batch resident on the device"},
                "gpu_launches": 16 * K,
                "roofline": {"bound": "fp64", "kernel": "whole P-RFO step (dominant: k_lg_tridiag2 cluster tridiagonalisation)",
                             "achieved": step_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": step_tf / fp64_peak,
                             "traffic": None, "algorithmic_flops_per_structure": WF,
                             "peak_source": "in-run DFMA probe (mop_priv_bench_dfma)"},
                "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": len(chunks), "kind": "port",
                                 "sample": f"{sample} structures, step-1 calls of oracle RSPRFOOracle timed, one process "
                                           "per core, 1 BLAS thread each"}}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-per-config", action="store_true",
                    help="skip the per_config records (configs[2], [3], [4]) of the default line")
    ap.add_argument("--workload", default="c2", choices=["c2", "c5"],
                    help="c2 (default): BASELINE configs[1], the headline line; c5: configs[4], P-RFO at 3N = 600")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.workload == "c5":
        run_b200_c5(args, rank, world, local_rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
