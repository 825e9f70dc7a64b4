"""Config 3 on several GPUs (torchrun): 64 images x 30 atoms, contiguous image blocks per rank, one NCCL
halo exchange per iteration, then the path kernels (tangent force, Ayala curvature, step limits) and one
RS-I-RFO step per image.  Checks the sharded result against the single-rank one computed on rank 0 and
times the halo exchange and the whole iteration (device time, max over ranks).
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/neb_multigpu.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from multioptpy_b200 import ops, synthetic
from multioptpy_b200.neb_halo import exchange_halo, image_partition

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
nimg, N = 64, 30; n = 3 * N
rng = np.random.default_rng(3)
xa = synthetic.grid_geometry(N, rng).reshape(-1); xb = xa + rng.normal(0, 0.3, n)
X = np.stack([xa + (xb - xa) * t for t in np.linspace(0, 1, nimg)]) + rng.normal(0, 0.02, (nimg, n))
E = -np.sin(np.linspace(0, np.pi, nimg)) * 0.05; G = rng.normal(0, 1e-2, (nimg, n))
Hs = np.stack([synthetic.spd_hessian(n, np.random.default_rng(40 + i)) for i in range(nimg)])
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def iteration(first, nloc, x, e, g, H, timed=None):
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True); t2 = torch.cuda.Event(enable_timing=True)
    t0.record()
    xh, Eh, gh = exchange_halo(x, e, g)
    t1.record()
    f, tau = ops.bneb_force(nimg, first, xh, Eh, g)
    Hc = H.clone()
    ops.neb_ayala(nimg, first, xh, Eh, gh, tau, Hc)
    st = ops.new_rsirfo_state(nloc, 0.5, dev)
    zero = torch.zeros(nloc, dtype=torch.float64, device=dev)
    out = ops.rsirfo_step(Hc, x, -f, -f, st, method=0, neb_mode=True, Be=zero)   # per-image RFO step on the NEB force
    d = out["move"].clone()
    ops.neb_limit_tr(nimg, first, xh, g, d)
    t2.record(); torch.cuda.synchronize()
    if timed is not None:
        timed.append((t0.elapsed_time(t1), t0.elapsed_time(t2)))
    return d


first, nloc = image_partition(nimg, world)[rank]
sl = slice(first, first + nloc)
args = (T(X[sl]), T(E[sl]), T(G[sl]), T(Hs[sl]))
times = []
for _ in range(3):
    d_loc = iteration(first, nloc, *args)
for _ in range(10):
    if world > 1:
        dist.barrier()
    d_loc = iteration(first, nloc, *args, timed=times)
halo_ms = float(np.median([t[0] for t in times])); it_ms = float(np.median([t[1] for t in times]))
if world > 1:
    t = torch.tensor([halo_ms, it_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    halo_ms, it_ms = float(t[0]), float(t[1])
    parts = [torch.zeros(image_partition(nimg, world)[r][1], n, dtype=torch.float64, device=dev) for r in range(world)]
    if len({p.shape[0] for p in parts}) == 1:
        dist.all_gather(parts, d_loc.contiguous())
    else:
        raise SystemExit("uneven partition: pick a world size dividing 64")
    d_all = torch.cat(parts)
else:
    d_all = d_loc
if rank == 0:
    # single-rank reference of the same iteration
    world_bak = world
    ref = None
    if world > 1:
        xh = T(np.concatenate([X[:1], X, X[-1:]])); Eh = T(np.concatenate([E[:1], E, E[-1:]])); gh = T(np.concatenate([G[:1], G, G[-1:]]))
        xh[0] = 0; xh[-1] = 0; Eh[0] = 0; Eh[-1] = 0; gh[0] = 0; gh[-1] = 0
        g = T(G); f, tau = ops.bneb_force(nimg, 0, xh, Eh, g)
        Hc = T(Hs).clone(); ops.neb_ayala(nimg, 0, xh, Eh, gh, tau, Hc)
        st = ops.new_rsirfo_state(nimg, 0.5, dev); zero = torch.zeros(nimg, dtype=torch.float64, device=dev)
        out = ops.rsirfo_step(Hc, T(X), -f, -f, st, method=0, neb_mode=True, Be=zero)
        ref = out["move"].clone(); ops.neb_limit_tr(nimg, 0, xh, g, ref)
        err = float((d_all - ref).abs().max() / ref.abs().max())
    else:
        err = 0.0
    print(json.dumps({"config": "configs[2]: NEB 64 images x 30 atoms, per-image RFO step + tangent/spring force with neighbour halo",
                      "n_gpus": world, "images_per_gpu": nloc, "halo_exchange_ms": halo_ms, "iteration_ms": it_ms,
                      "images_per_s": nimg / (it_ms * 1e-3), "max_rel_diff_vs_single_rank": err,
                      "halo_bytes_per_side": (2 * n + 1) * 8}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
