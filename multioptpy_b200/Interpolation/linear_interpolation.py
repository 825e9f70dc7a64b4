"""Device form of ``multioptpy.Interpolation.linear_interpolation.distribute_geometry``
(Interpolation/linear_interpolation.py:308-336): images at equal intervals of the centroid-free
path length (Utils/calc_tools.py:853-862) — the ``align_distances`` strategy of
``NEB._align_geometries`` (neb.py:649-760).  The other strategies of that table (energy-weighted,
Bernstein, spline, Savitzky-Golay, geodesic) are not on the hot path and raise in the NEB mirror."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..neb_halo import gather_chain


def distribute_geometry(geometry_list, device="cuda"):
    """geometry_list: sequence of (natoms, 3) arrays, an (nimg, natoms, 3) array or a CUDA tensor.
    Returns the same kind (a list of arrays for a list, like the reference)."""
    if isinstance(geometry_list, torch.Tensor):
        return ops.neb_redistribute(geometry_list.contiguous())
    X = np.ascontiguousarray(np.asarray(geometry_list, dtype=np.float64))
    out = ops.neb_redistribute(torch.from_numpy(X).to(device)).cpu().numpy()
    return list(out) if isinstance(geometry_list, (list, tuple)) else out


def distribute_geometry_sharded(x_local, nimg, first, group=None):
    """x_local (nloc, natoms, 3): this rank's contiguous image block of a chain of nimg images.
    One all-gather of the chain, then every rank interpolates its own images."""
    chain = gather_chain(x_local, nimg, group)
    return ops.neb_redistribute(chain.contiguous(), first, x_local.shape[0])
