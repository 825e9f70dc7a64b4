"""NEB (config 3 shape): tangent projection, Ayala curvature update, per-image RFO step and
step limits vs the golden trace recorded from the reference; halo exchange over gloo."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import np_oracle as O

RTOL = 1e-10


def rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / (nb if nb > 0 else 1.0)


def test_oracle_replays_reference_neb_trace(golden_dir):
    z = np.load(os.path.join(golden_dir, "neb_rfo.npz"))
    neb = O.NEBRFOOracle(z["H_init"])
    for it in range(z["X"].shape[0]):
        F, T, gam, delta, mv = neb.step(z["X"][it], z["E"][it], z["G"][it])
        assert rel(F, z["force"][it]) < RTOL, it
        assert rel(T, z["tau"][it]) < RTOL, it
        assert np.abs(gam - z["gamma"][it]).max() <= 1e-8 * max(1.0, np.abs(z["gamma"][it]).max()), it
        assert rel(np.stack(neb.H), z["H_after"][it]) < 1e-9, it
        assert rel(mv, z["rfo_move"][it]) < 1e-9, it


@pytest.mark.gpu
def test_gpu_neb_rfo_vs_reference_trace(golden_dir):
    from multioptpy_b200.Optimizer.rfo_neb import RFOOptimizer
    z = np.load(os.path.join(golden_dir, "neb_rfo.npz"))
    nimg, natoms = [int(v) for v in z["meta"]]
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    opt = RFOOptimizer(nimg, natoms, device=dev)
    opt.set_hessians(T(z["H_init"]))
    for it in range(z["X"].shape[0]):
        mv = opt.rfo_move_vectors(T(z["X"][it]), T(z["E"][it]), T(z["G"][it])).cpu().numpy()
        assert rel(opt.last["force"].cpu().numpy(), z["force"][it]) < RTOL, it
        assert rel(opt.last["tau"].cpu().numpy(), z["tau"][it]) < RTOL, it
        g = opt.last["gamma"].cpu().numpy()
        assert np.abs(g - z["gamma"][it]).max() <= 1e-8 * max(1.0, np.abs(z["gamma"][it]).max()), it
        assert rel(opt.hessian.cpu().numpy(), z["H_after"][it]) < 1e-9, it
        assert rel(mv, z["rfo_move"][it]) < 1e-9, it


@pytest.mark.gpu
def test_gpu_bneb_dropin_class(golden_dir):
    from multioptpy_b200.MEP.pathopt_bneb_force import CaluculationBNEB
    z = np.load(os.path.join(golden_dir, "neb_rfo.npz"))
    nimg, natoms = [int(v) for v in z["meta"]]
    calc = CaluculationBNEB(device="cuda:0")
    F = calc.calc_force(z["X"][0].reshape(nimg, natoms, 3), z["E"][0], z["G"][0].reshape(nimg, natoms, 3), 0, ["C"] * natoms)
    assert F.shape == (nimg, natoms, 3)
    assert rel(F.reshape(nimg, -1), z["force"][0]) < RTOL
    assert rel(calc.get_tau(3).ravel(), z["tau"][0][3]) < RTOL


# ---------------------------------------------------------------- halo over gloo (world_size 2, 3)
def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _halo_worker(rank, world, port, nimg, n, q):
    import torch.distributed as dist
    from multioptpy_b200.neb_halo import exchange_halo, image_partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    X = torch.from_numpy(rng.standard_normal((nimg, n))); E = torch.from_numpy(rng.standard_normal(nimg))
    G = torch.from_numpy(rng.standard_normal((nimg, n)))
    first, nloc = image_partition(nimg, world)[rank]
    xh, Eh, gh = exchange_halo(X[first:first + nloc].contiguous(), E[first:first + nloc].contiguous(),
                               G[first:first + nloc].contiguous())
    ok = True
    for l in range(nloc + 2):
        gi = first + l - 1
        if 0 <= gi < nimg:
            ok &= bool(torch.equal(xh[l], X[gi]) and torch.equal(gh[l], G[gi]) and Eh[l] == E[gi])
        else:
            ok &= bool((xh[l] == 0).all())
    q.put((rank, ok, first, nloc))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, 11, 12, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res)
    assert sorted(f for _, _, f, _ in res)[0] == 0 and sum(nl for _, _, _, nl in res) == 11


# ------------------------------------------------------------------ FIRE optimizer of the NEB driver
def _fire_oracle(z):
    c = z["cfg"]
    return O.FIRENEBOracle(dt=c[0], a=c[1], n_reset=int(c[2]), N_accelerate=int(c[3]), f_inc=c[4], f_decelerate=c[5],
                           a_start=c[6], dt_max=c[7])


def test_oracle_fire_vs_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "fire_neb.npz"))
    o = _fire_oracle(z)
    for it in range(len(z["X"])):
        _, _, mv, _ = o.step(z["X"][it], z["F"][it], z["V"][it], z["Vprev"][it] if z["have_prev"][it] else None, it)
        ref = z["move"][it]          # recovered from the reference's Angstrom output: a few 1e-11 of round trip
        assert np.abs(mv - ref).max() <= 2e-10 * np.abs(ref).max(), it
        assert (o.dt, o.a, o.n_reset) == tuple(z["state"][it]), it


@pytest.mark.gpu
def test_gpu_fire_vs_golden(golden_dir):
    import types
    from multioptpy_b200.Optimizer.fire_neb import FIREOptimizer
    z = np.load(os.path.join(golden_dir, "fire_neb.npz"))
    c = z["cfg"]
    cfg = types.SimpleNamespace(dt=c[0], a=c[1], n_reset=int(c[2]), FIRE_N_accelerate=int(c[3]), FIRE_f_inc=c[4],
                                FIRE_f_accelerate=0.99, FIRE_f_decelerate=c[5], FIRE_a_start=c[6], FIRE_dt_max=c[7])
    opt = FIREOptimizer(cfg, device="cuda:0")
    o = _fire_oracle(z)
    for it in range(len(z["X"])):
        pre = z["Vprev"][it] if z["have_prev"][it] else []
        new_ang = opt.optimize(z["X"][it], z["F"][it], pre, it, z["V"][it])
        mv = new_ang / 0.52917721067 - z["X"][it]
        Vn, _, mo, _ = o.step(z["X"][it], z["F"][it], z["V"][it], z["Vprev"][it] if z["have_prev"][it] else None, it)
        assert np.abs(mv - mo).max() <= 1e-9 * np.abs(mo).max(), it       # Angstrom round trip of the return value
        assert np.abs(opt.total_velocity - Vn).max() <= 1e-12 * max(np.abs(Vn).max(), 1e-300), it
        assert (opt.dt, opt.a, opt.n_reset) == (o.dt, o.a, o.n_reset) == tuple(z["state"][it]), it


# ------------------------------------------- the whole RFOOptimizer.optimize incl. the RFO / FIRE combine (rfo_neb.py:186-206)
def test_oracle_full_neb_optimize_vs_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "neb_full.npz"))
    nimg, natoms = [int(v) for v in z["meta"]]
    orc = O.NEBRFOOracle(z["H_init"])
    for it in range(z["X"].shape[0]):
        new = O.neb_optimize_step(orc, z["X"][it], z["E"][it], z["G"][it], z["V"][it], z["Vprev"][it], it)
        ref = z["new_geom_ang"][it].reshape(nimg, -1) / 0.52917721067
        assert np.abs(new - ref).max() <= 1e-10 * np.abs(ref).max(), it
    assert rel(np.stack(orc.H), z["H_final"]) < 1e-9


@pytest.mark.gpu
def test_gpu_full_neb_optimize_vs_reference(golden_dir):
    import types
    from multioptpy_b200.Optimizer.rfo_neb import RFOOptimizer
    from multioptpy_b200.Optimizer.fire_neb import FIREOptimizer
    z = np.load(os.path.join(golden_dir, "neb_full.npz"))
    nimg, natoms = [int(v) for v in z["meta"]]
    dev = "cuda:0"
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cfg = types.SimpleNamespace(dt=0.5, a=0.10, n_reset=0, FIRE_N_accelerate=5, FIRE_f_inc=1.10, FIRE_f_accelerate=0.99,
                                FIRE_f_decelerate=0.5, FIRE_a_start=0.1, FIRE_dt_max=3.0)
    opt = RFOOptimizer(nimg, natoms, device=dev)
    opt.set_hessians(T(z["H_init"]))
    for it in range(z["X"].shape[0]):
        fire = FIREOptimizer(cfg, device=dev)          # the reference builds a fresh one per call (rfo_neb.py:185)
        x = T(z["X"][it])
        move, _ = opt.optimize_step(x, T(z["E"][it]), T(z["G"][it]), fire, T(z["V"][it]), T(z["Vprev"][it]), it)
        new = (x + move).cpu().numpy()
        ref = z["new_geom_ang"][it].reshape(nimg, -1) / 0.52917721067
        assert np.abs(new - ref).max() <= 1e-10 * np.abs(ref).max(), it
    assert rel(opt.hessian.cpu().numpy(), z["H_final"]) < 1e-9


def test_neb_on_disk_formats_roundtrip(tmp_path):
    """tmp_hessian_<i>.npy and the 12-decimal xyz samples of the reference (rfo_neb.py:18-25,175; fileio.py:441-446)."""
    from multioptpy_b200 import fileio
    rng = np.random.default_rng(3)
    H = rng.standard_normal((3, 9, 9))
    fileio.save_neb_hessians(str(tmp_path), H, first=2)
    assert sorted(p.name for p in tmp_path.iterdir()) == ["tmp_hessian_2.npy", "tmp_hessian_3.npy", "tmp_hessian_4.npy"]
    back = fileio.load_neb_hessians(str(tmp_path), nimg=6, natoms=3, first=1, nloc=5)
    assert np.array_equal(back[1:4], H) and np.array_equal(back[0], np.eye(9)) and np.array_equal(back[4], np.eye(9))
    G = rng.standard_normal((2, 3, 3)) * 3
    paths = fileio.write_xyz_samples(str(tmp_path / "s"), "aldol", ["C", "H", "Cl"], G, (0, 1))
    txt = open(paths[0]).read().splitlines()
    assert txt[0] == "3" and txt[1] == "0 1"
    x = G[0][2]
    assert txt[4] == f"Cl   {x[0]:>17.12f}   {x[1]:>17.12f}   {x[2]:>17.12f}"      # the reference's f-string, verbatim layout
    e, xyz, cm = fileio.read_xyz_sample(paths[1])
    assert e == ["C", "H", "Cl"] and cm == (0, 1) and np.abs(xyz - G[1]).max() < 5e-13


# ---- image redistribution (align_distances): Interpolation/linear_interpolation.py:308-336 -------------------------
def test_oracle_redistribute_vs_reference(golden_dir):
    z = np.load(os.path.join(golden_dir, "neb_redistribute.npz"))
    for name in z["names"]:
        X = z[f"{name}/X"]
        assert rel(O.path_length_list(X), z[f"{name}/path_length"]) < 1e-14, name
        assert rel(O.distribute_geometry(X), z[f"{name}/out"]) < 1e-13, name


@pytest.mark.gpu
def test_gpu_redistribute_vs_reference(golden_dir):
    from multioptpy_b200 import ops
    from multioptpy_b200.Interpolation.linear_interpolation import distribute_geometry
    z = np.load(os.path.join(golden_dir, "neb_redistribute.npz"))
    for name in z["names"]:
        X = z[f"{name}/X"]
        xd = torch.from_numpy(X).to("cuda:0")
        out, pl = ops.neb_redistribute(xd, want_path_length=True)
        assert rel(pl.cpu().numpy(), z[f"{name}/path_length"]) < 1e-13, name
        assert rel(out.cpu().numpy(), z[f"{name}/out"]) < RTOL, name
        # a rank's slice of the chain equals the slice of the full result
        part = ops.neb_redistribute(xd, 1, len(X) - 2)
        assert torch.equal(part, out[1:-1]), name
        lst = distribute_geometry([x for x in X])           # list in -> list out, like the reference
        assert isinstance(lst, list) and rel(np.array(lst), z[f"{name}/out"]) < RTOL, name
    # a straight, equally spaced chain is a fixed point (a bent chain is not: the chords cut its corners)
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=(30, 3)), rng.normal(size=(30, 3))
    X = np.stack([a + t * (b - a) for t in np.linspace(0.0, 1.0, 64)])
    again = ops.neb_redistribute(torch.from_numpy(X).to("cuda:0")).cpu().numpy()
    assert rel(again, X) < 1e-12


def _gather_worker(rank, world, port, nimg, q):
    import torch.distributed as dist
    from multioptpy_b200.neb_halo import gather_chain, image_partition
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    X = torch.arange(nimg * 4 * 3, dtype=torch.float64).reshape(nimg, 4, 3)
    first, nloc = image_partition(nimg, world)[rank]
    chain = gather_chain(X[first:first + nloc].clone(), nimg)
    q.put((rank, bool(torch.equal(chain, X))))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gather_chain_gloo(world):
    """The all-gather that precedes a redistribution of a sharded chain (uneven blocks at world 3)."""
    import torch.multiprocessing as mp
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_gather_worker, args=(r, world, port, 7, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
