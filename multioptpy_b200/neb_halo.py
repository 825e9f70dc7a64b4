"""Neighbour-image halo for a NEB chain sharded over ranks (the only collective of the path,
SURVEY §8e): every rank owns a contiguous block of images and needs the coordinates, energy
and gradient of the image just before and just after its block; image redistribution (every
`align_distances` iterations) needs the whole chain once: ``gather_chain``.  Implemented with
``torch.distributed`` point-to-point batches — NCCL over NVLink on GPUs, gloo on CPU tensors
for the host-side tests.  Message size: (2 n + 1) doubles per side (~1.5 KB at N = 30)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def image_partition(nimg: int, world: int):
    """Contiguous blocks, sizes differing by at most one: [(first, nloc)] per rank."""
    base, rem = divmod(nimg, world)
    out, first = [], 0
    for r in range(world):
        nloc = base + (1 if r < rem else 0)
        out.append((first, nloc))
        first += nloc
    return out


def exchange_halo(x, E, g, group=None):
    """x (nloc, n), E (nloc,), g (nloc, n) of this rank's images -> (x_halo (nloc+2, n),
    E_halo (nloc+2,), g_halo (nloc+2, n)).  Halo slots at the chain ends stay zero (unused)."""
    nloc, n = x.shape
    xh = torch.zeros(nloc + 2, n, dtype=x.dtype, device=x.device); xh[1:-1] = x
    Eh = torch.zeros(nloc + 2, dtype=E.dtype, device=E.device); Eh[1:-1] = E
    gh = torch.zeros(nloc + 2, n, dtype=g.dtype, device=g.device); gh[1:-1] = g
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return xh, Eh, gh
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    pack = lambda i: torch.cat([x[i], g[i], E[i:i + 1]]).contiguous()
    ops = []
    recv_left = torch.empty(2 * n + 1, dtype=x.dtype, device=x.device)
    recv_right = torch.empty(2 * n + 1, dtype=x.dtype, device=x.device)
    if rank > 0:
        ops += [dist.P2POp(dist.isend, pack(0), rank - 1, group), dist.P2POp(dist.irecv, recv_left, rank - 1, group)]
    if rank < world - 1:
        ops += [dist.P2POp(dist.isend, pack(nloc - 1), rank + 1, group), dist.P2POp(dist.irecv, recv_right, rank + 1, group)]
    for req in dist.batch_isend_irecv(ops) if ops else []:
        req.wait()
    if rank > 0:
        xh[0], gh[0], Eh[0] = recv_left[:n], recv_left[n:2 * n], recv_left[2 * n]
    if rank < world - 1:
        xh[-1], gh[-1], Eh[-1] = recv_right[:n], recv_right[n:2 * n], recv_right[2 * n]
    return xh, Eh, gh


def gather_chain(x, nimg, group=None):
    """x (nloc, ...) of this rank's contiguous image block -> the whole chain (nimg, ...) on every rank (all_gather;
    blocks as in ``image_partition``).  46 KB at config 3: one NCCL all-gather per redistribution."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    world = dist.get_world_size(group)
    blocks = image_partition(nimg, world)
    maxloc = max(nl for _, nl in blocks)          # blocks differ by at most one image: pad to equal messages
    send = torch.zeros((maxloc,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    send[: x.shape[0]] = x
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    return torch.cat([recv[r][:nl] for r, (_, nl) in enumerate(blocks)], dim=0)
