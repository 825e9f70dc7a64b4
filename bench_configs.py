"""Named shapes of BASELINE.json beside the headline line (configs[1]): configs[2] NEB 64 x 30, configs[3] conformer /
AFIR batch 8192 x N (N = 24 and N = 8), configs[4] P-RFO at 3N = 600.  bench.py attaches one record per config to its
JSON line (`per_config`), so the driver's 1 / 2 / 4 / 8-GPU runs witness them; tests/test_config_shapes.py checks the
same chains against the oracle at these shapes.

Every record: value (whole job, all ranks), unit, ms per unit of work (device time, CUDA events, max over ranks),
parity_vs_oracle (max relative error of a strided oracle sample, checked on rank 0 before timing) and the roofline
figure that bounds it.  The oracle is the checker only; nothing here is a CPU fallback."""
from __future__ import annotations

import numpy as np


def _ev(torch):
    return torch.cuda.Event(enable_timing=True)


def _max_over_ranks(vals, dist, dev, torch):
    if dist is None:
        return [float(v) for v in vals]
    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def _rel(a, b):
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0))


# --------------------------------------------------------------------------------------------------- configs[2]: NEB
def neb_reference_two_iterations(nimg, natoms, seed=3000):
    """Inputs + oracle results of two NEB iterations at the config shape (checker; tests and bench parity)."""
    from multioptpy_b200 import synthetic
    from oracle import np_oracle as O
    X, E, G, H = synthetic.neb_chain(nimg, natoms, seed)
    orc = O.NEBRFOOracle(H)
    F0, T0, gam0, d0, dl0 = orc.step(X, E, G)
    X1, G1 = synthetic.neb_next(X, G, H, dl0, np.random.default_rng(seed + 1))
    E1 = E - 1e-3
    F1, T1, gam1, d1, dl1 = orc.step(X1, E1, G1)
    return dict(X=X, E=E, G=G, H=H, X1=X1, E1=E1, G1=G1, delta0=dl0, delta1=dl1, force1=F1, H_after=np.stack(orc.H))


def c3_record(dev, rank, world, dist, steps=10):
    """configs[2]: one NEB iteration = halo exchange (NCCL when sharded), BNEB tangent force, Ayala curvature update,
    one RS-I-RFO step per image (update active), step limits.  Reported as time per iteration."""
    import torch
    from multioptpy_b200.Optimizer.rfo_neb import RFOOptimizer
    from multioptpy_b200.neb_halo import image_partition
    nimg, natoms = 64, 30
    n = 3 * natoms
    ref = neb_reference_two_iterations(nimg, natoms) if rank == 0 else None
    from multioptpy_b200 import synthetic
    X, E, G, H = synthetic.neb_chain(nimg, natoms)
    first, nloc = image_partition(nimg, world)[rank]
    sl = slice(first, first + nloc)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    # iteration 1 inputs come from the oracle-free device run of iteration 0 (identical on every rank)
    opt = RFOOptimizer(nimg, natoms, first=first, nloc=nloc, device=dev)
    opt.set_hessians(T(H[sl]))
    d0 = opt.rfo_move_vectors(T(X[sl]), T(E[sl]), T(G[sl]))
    if world > 1:
        parts = [torch.zeros(image_partition(nimg, world)[r][1], n, dtype=torch.float64, device=dev) for r in range(world)]
        dist.all_gather(parts, d0.contiguous()) if len({p.shape[0] for p in parts}) == 1 else None
        d0_all = torch.cat(parts).cpu().numpy()
    else:
        d0_all = d0.cpu().numpy()
    X1, G1 = synthetic.neb_next(X, G, H, d0_all, np.random.default_rng(3001))
    E1 = E - 1e-3
    x1, e1, g1 = T(X1[sl]), T(E1[sl]), T(G1[sl])
    H_after0 = opt.hessian.clone()
    px, pg = opt.prev_x.clone(), opt.prev_g.clone()
    d1 = opt.rfo_move_vectors(x1, e1, g1)
    parity = None
    if rank == 0:
        parity = max(_rel(d0_all[:nloc], ref["delta0"][:nloc]), _rel(d1.cpu().numpy(), ref["delta1"][sl]))
        if not parity < 1e-10:
            raise SystemExit(f"bench: configs[2] parity vs oracle failed ({parity:.3e})")

    def iteration():
        opt.hessian.copy_(H_after0)
        opt.prev_x, opt.prev_g = px, pg
        return opt.rfo_move_vectors(x1, e1, g1)

    for _ in range(3):
        iteration()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = _ev(torch), _ev(torch)
    a.record()
    for _ in range(steps):
        iteration()
    b.record(); torch.cuda.synchronize()
    it_ms = a.elapsed_time(b) / steps
    halo_ms = opt.time_halo(x1, e1, g1, reps=steps) if world > 1 else 0.0
    it_ms, halo_ms = _max_over_ranks([it_ms, halo_ms], dist, dev, torch)
    if rank != 0:
        return None
    return {"workload": "configs[2]: NEB 64 images of a 30-atom path (3N=90): per-image quasi-Newton step + tangent / "
                        "spring force with neighbour halo, iteration 2 (update active)",
            "value": nimg / (it_ms * 1e-3), "unit": "image-steps/s", "ms_per_iteration": it_ms,
            "halo_exchange_ms": halo_ms, "images_per_gpu": nloc, "n_gpus": world, "scaling": "strong",
            "parity_vs_oracle": parity,
            "roofline": {"bound": "launch latency", "note": "64 small images: one iteration is ~10 dependent launches; "
                         "time per iteration and halo latency are the figures of merit (SURVEY 8e)"}}


# -------------------------------------------------------------------------------- configs[3]: conformer / AFIR batch
def c4_chain_reference(xyz, g, elems, frag1, frag2, gamma, method="rsirfo_block_fsb", first_only=False):
    """Oracle chain for ONE structure: AFIR E/g/H, Lindh model Hessian (without the ill-posed K term),
    RS-I-RFO step 0 clamped to the outer trust radius by the caller, then AFIR at the moved geometry and step 1 with
    the update."""
    from multioptpy_b200.ModelHessian.lindh import lindh_atom_params
    from multioptpy_b200.Parameters.tables import covalent_radius
    from oracle import np_oracle as O
    radii = [covalent_radius(e) for e in elems]
    n = xyz.size
    Eb, gb, Hb = O.afir_egh(xyz, frag1, frag2, radii, gamma)
    Hm = O.lindh_hessian_bkb(xyz, lindh_atom_params(elems))
    o = O.RSIRFOOracle(method=method, saddle_order=0)
    o.set_hessian(Hm.copy()); o.set_bias_hessian(Hb)
    x0 = xyz.reshape(-1)
    m0 = o.run(x0, g + gb.reshape(-1), g, None, None, Eb)
    if first_only:
        return dict(move0_raw=m0)
    _, m0 = O.clamp_and_move(x0, m0, 0.5)          # the caller's clamp to the outer trust radius (optimizer.py:792-798)
    x1 = x0 - m0
    g1 = g + 0.5 * (x1 - x0)                       # isotropic PES curvature 0.5: s.y > 0, the update is active
    Eb1, gb1, Hb1 = O.afir_egh(x1.reshape(-1, 3), frag1, frag2, radii, gamma)
    o.set_bias_hessian(Hb1)
    m1 = o.run(x1, g1 + gb1.reshape(-1), g1, x0, g, Eb1 - 1e-3)
    return dict(E_afir=Eb, g_afir=gb.reshape(-1), H_afir=Hb, H_model=Hm, move0=m0, x1=x1, g1=g1, move1=m1, H_after=o.hessian)


class C4Chain:
    """The device chain of configs[3] for a batch: AFIR bias (E, gradient, Hessian), Lindh model Hessian at iteration 0,
    RS-I-RFO (`rsirfo_block_fsb`) steps with the AFIR Hessian as bias Hessian."""

    def __init__(self, xyz, g, dev, gamma=100.0):
        import torch
        from multioptpy_b200 import ops, synthetic
        from multioptpy_b200.ModelHessian.lindh import lindh_atom_params
        from multioptpy_b200.Parameters.tables import covalent_radius
        self.torch, self.ops = torch, ops
        B, N, _ = xyz.shape
        self.B, self.N, self.n, self.dev = B, N, 3 * N, dev
        self.elems = synthetic.elements(N, all_sulfur=True)
        T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        self.xyz, self.g = T(xyz), T(g)
        self.frag1 = list(range(N // 2)); self.frag2 = list(range(N // 2, N))      # "-ma gamma 1-12 13-24"
        self.f1 = torch.tensor(self.frag1, dtype=torch.int32, device=dev)
        self.f2 = torch.tensor(self.frag2, dtype=torch.int32, device=dev)
        self.radii = torch.tensor([covalent_radius(e) for e in self.elems], dtype=torch.float32, device=dev)
        self.gamma = torch.full((B,), float(gamma), dtype=torch.float64, device=dev)
        self.prm = lindh_atom_params(self.elems)
        self.method = ops.resolve_update_method("rsirfo_block_fsb")

    def afir(self, xyz):
        return self.ops.afir(xyz, self.f1, self.f2, self.radii, self.gamma)

    def lindh(self, xyz):
        return self.ops.lindh_hessian(xyz, self.prm)[0]

    def step(self, H, x, g, Eb, gb, Hb, state, x_prev=None, g_prev=None, out=None, dE=0.0):
        return self.ops.rsirfo_step(H, x, (g + gb).contiguous(), g, state, method=self.method, Hbias=Hb,
                                    x_prev=x_prev, g_prev=g_prev, Be=(Eb - dE).contiguous(), out=out)

    def two_iterations(self):
        """Iteration 0 (model Hessian, no history) and iteration 1 (update active); returns everything."""
        torch, ops = self.torch, self.ops
        x0 = self.xyz.reshape(self.B, self.n).contiguous()
        Eb, gb, Hb = self.afir(self.xyz)
        Hm = self.lindh(self.xyz)
        H = Hm.clone()
        st = ops.new_rsirfo_state(self.B, 0.5, self.dev)
        o0 = self.step(H, x0, self.g, Eb, gb, Hb, st)
        m0 = o0["move"].clone()
        status0 = o0["status"].clone()
        ops.clamp_and_move(x0, m0, torch.full((self.B,), 0.5, dtype=torch.float64, device=self.dev), want_geometry=False)
        x1 = (x0 - m0).contiguous()
        g1 = (self.g + 0.5 * (x1 - x0)).contiguous()
        Eb1, gb1, Hb1 = self.afir(x1.reshape(self.B, self.N, 3))
        o1 = self.step(H, x1, g1, Eb1, gb1, Hb1, st, x_prev=x0, g_prev=self.g, dE=1e-3)
        return dict(E_afir=Eb, g_afir=gb, H_afir=Hb, H_model=Hm, move0=m0, x0=x0, x1=x1, g1=g1, move1=o1["move"].clone(),
                    H_after=H, state=st, status0=status0, status1=o1["status"].clone())


def c4_record(dev, rank, world, dist, natoms, B=8192, steps=5, sample=4):
    """configs[3]: per structure and iteration AFIR E/g/H + one `rsirfo_block_fsb` step with the update; the Lindh
    model Hessian of iteration 0 is timed beside it.  Weak scaling: B structures per GPU."""
    import torch
    from multioptpy_b200 import ops, synthetic
    xyz, g = synthetic.conformer_batch(B, natoms, seed=4000 + 100000 * rank + 7 * natoms)
    ch = C4Chain(xyz, g, dev)
    r = ch.two_iterations()
    parity = None
    # structures whose alpha loop left its rounding-free exits (MOP_ST_ALPHA_UNSTABLE: the reference's own step depends
    # on the summation order of its BLAS there) are not comparable to 1e-10; the strict sample skips them
    unstable = ((r["status0"] | r["status1"]) & ops.ST_ALPHA_UNSTABLE) != 0
    unstable_frac = float(unstable.double().mean())
    if rank == 0:
        parity = 0.0
        unst = unstable.cpu().numpy()
        picks = [b for b in range(0, B, max(1, B // (2 * sample))) if not unst[b]][:sample]
        if len(picks) < sample:
            raise SystemExit(f"bench: configs[3] (N={natoms}): too many alpha-unstable structures ({unstable_frac:.3f})")
        for b in picks:
            ref = c4_chain_reference(xyz[b], g[b], ch.elems, ch.frag1, ch.frag2, 100.0)
            parity = max(parity, _rel(r["H_afir"][b].cpu().numpy(), ref["H_afir"]), _rel(r["H_model"][b].cpu().numpy(), ref["H_model"]),
                         _rel(r["move0"][b].cpu().numpy(), ref["move0"]), _rel(r["move1"][b].cpu().numpy(), ref["move1"]),
                         _rel(r["H_after"][b].cpu().numpy(), ref["H_after"]))
        if not parity < 1e-10:
            raise SystemExit(f"bench: configs[3] (N={natoms}) parity vs oracle failed ({parity:.3e})")
    n = ch.n
    x1g = r["x1"].reshape(B, natoms, 3).contiguous()
    ncopy = 3
    Hs = [r["H_model"].clone() for _ in range(ncopy)]
    st0 = ops.new_rsirfo_state(B, 0.5, dev)
    o = ch.step(Hs[0], r["x0"], ch.g, r["E_afir"], r["g_afir"], r["H_afir"], st0)   # state after iteration 0
    sts = [st0.clone() for _ in range(ncopy)]
    out = None

    def iteration(i):
        nonlocal out
        j = i % ncopy
        Hs[j].copy_(r["H_model"]); sts[j].copy_(st0)
        Eb1, gb1, Hb1 = ch.afir(x1g)
        out = ch.step(Hs[j], r["x1"], r["g1"], Eb1, gb1, Hb1, sts[j], x_prev=r["x0"], g_prev=ch.g, out=out, dE=1e-3)

    for i in range(3):
        iteration(i)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    a, b_ = _ev(torch), _ev(torch)
    a.record()
    for i in range(steps):
        iteration(i)
    b_.record(); torch.cuda.synchronize()
    it_ms = a.elapsed_time(b_) / steps
    a, b_ = _ev(torch), _ev(torch)
    a.record()
    for i in range(steps):
        ch.lindh(ch.xyz)
    b_.record(); torch.cuda.synchronize()
    lindh_ms = a.elapsed_time(b_) / steps
    a, b_ = _ev(torch), _ev(torch)
    a.record()
    for i in range(steps):
        ch.afir(x1g)
    b_.record(); torch.cuda.synchronize()
    afir_ms = a.elapsed_time(b_) / steps
    it_ms, lindh_ms, afir_ms = _max_over_ranks([it_ms, lindh_ms, afir_ms], dist, dev, torch)
    if rank != 0:
        return None
    wb = 16.0 * n * n + 8.0 * n * n + 8.0 * n * n + 80.0 * n       # H in/out, bias Hessian read, AFIR Hessian written
    return {"workload": f"configs[3]: conformer / AFIR search batch {B} x N={natoms} (3N={n}, all S): AFIR bias E + "
                        "gradient + Hessian and one rsirfo_block_fsb step (update active) per structure; Lindh model "
                        "Hessian (iteration 0) timed beside it",
            "value": world * B / (it_ms * 1e-3), "unit": "structure-steps/s", "ms_per_iteration": it_ms,
            "afir_ms": afir_ms, "lindh_model_hessian_ms": lindh_ms, "batch_per_gpu": B, "n_gpus": world,
            "scaling": "weak", "parity_vs_oracle": parity, "alpha_unstable_fraction": unstable_frac,
            "roofline": {"bound": "hbm", "achieved": B * wb / (it_ms * 1e-3) / 1e9, "unit": "GB/s",
                         "algorithmic_bytes_per_structure": wb}}


# --------------------------------------------------------------------------------------- configs[4]: P-RFO, 3N = 600
def c5_record(dev, rank, world, dist, steps=3, total=256, fp64_peak=None):
    import torch
    from multioptpy_b200 import ops, synthetic
    from oracle import np_oracle as O
    natoms = 200
    n = 3 * natoms
    B = total // world
    f64 = torch.float64
    nuniq = min(B, 8)
    xs, Hs_, gs = [], [], []
    for b in range(nuniq):
        x0, H0, g0, _ = synthetic.structure(5, rank * nuniq + b, natoms, saddle=True)
        xs.append(x0); Hs_.append(H0); gs.append(g0)
    rep = (B + nuniq - 1) // nuniq
    tile = lambda a: torch.from_numpy(np.stack(a)).repeat(rep, *([1] * (np.stack(a).ndim - 1)))[:B].contiguous().to(dev)
    H_d0, x0_d, g0_d = tile(Hs_), tile(xs), tile(gs)
    z = lambda *sh: torch.zeros(*sh, dtype=f64, device=dev)
    st0 = dict(state=z(B, ops.PRFO_STATE), prev_grad=z(B, n), prev_move=z(B, n), ts_vec=z(B, n))
    st0["state"][:, 0] = 0.1
    out0 = ops.rsprfo_step(H_d0.clone(), x0_d, g0_d, st0, method=23, saddle_order=1, Be=z(B))
    mv0 = out0["move"].clone()
    x1_d = (x0_d - mv0).contiguous()
    g1_d = (g0_d + torch.einsum("bij,bj->bi", H_d0, x1_d - x0_d)).contiguous()
    Be1 = z(B) - 1e-3
    parity = None
    m1 = None
    if rank == 0:
        o = O.RSPRFOOracle(method="rsprfo_bofill", saddle_order=1); o.set_hessian(Hs_[0])
        m0 = o.run(xs[0], gs[0], None, None, 0.0, None)
        parity = _rel(mv0[0].cpu().numpy(), m0)
        m1 = o.run(x1_d[0].cpu().numpy(), g1_d[0].cpu().numpy(), xs[0], gs[0], -1e-3, m0)
    ncopy = 2
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

    for i in range(2):
        one_step(i)
    if rank == 0:
        parity = max(parity, _rel(out["move"][0].cpu().numpy(), m1))
        if not parity < 1e-10:
            raise SystemExit(f"bench: configs[4] parity vs oracle failed ({parity:.3e})")
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    a, b_ = _ev(torch), _ev(torch)
    a.record()
    for i in range(steps):
        one_step(i)
    b_.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b_) / steps
    (ms,) = _max_over_ranks([ms], dist, dev, torch)
    if rank != 0:
        return None
    WF = 9.0 * n ** 3 + 4.0 / 3.0 * n ** 3 + 40.0 * n * n
    val = world * B / (ms * 1e-3)
    roof = {"bound": "fp64", "achieved": val / world * WF / 1e12, "unit": "TFLOP/s", "algorithmic_flops_per_structure": WF}
    if fp64_peak:
        roof.update(peak=fp64_peak, frac=roof["achieved"] / fp64_peak)
    return {"workload": f"configs[4]: P-RFO saddle search with Bofill update, N=200 atoms (3N=600), batch {total} sharded "
                        f"over {world} GPU(s), step 1 of 2 (update active)",
            "value": val, "unit": "structure-steps/s", "ms_per_step": ms, "batch_per_gpu": B, "n_gpus": world,
            "scaling": "strong", "parity_vs_oracle": parity, "roofline": roof}
