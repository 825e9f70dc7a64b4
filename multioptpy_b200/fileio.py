"""On-disk formats either side of the NEB quasi-Newton step, so that a run can hand its state to
the reference driver and back: the per-image Hessian files the reference round-trips between iterations
(``tmp_hessian_<i>.npy``, Optimizer/rfo_neb.py:18-25,175) and the per-iteration geometry files with twelve decimals
in Angstrom (fileio.py:418-447, ``make_psi4_input_file``).  Host code only; nothing here is on the hot path."""
from __future__ import annotations

import os

import numpy as np


def save_neb_hessians(folder, hessians, first=0):
    """hessians: (nloc, 3N, 3N) array or CUDA tensor of this rank's images (global image index = first + i) ->
    ``<folder>/tmp_hessian_<index>.npy`` as np.save writes them in the reference (rfo_neb.py:175)."""
    os.makedirs(folder, exist_ok=True)
    H = hessians.detach().cpu().numpy() if hasattr(hessians, "detach") else np.asarray(hessians)
    for i in range(H.shape[0]):
        np.save(os.path.join(folder, f"tmp_hessian_{first + i}.npy"), np.ascontiguousarray(H[i], dtype=np.float64))


def load_neb_hessians(folder, nimg, natoms, first=0, nloc=None):
    """_load_or_init_hessian (rfo_neb.py:18-25): the saved Hessian of every image, the identity where no file exists."""
    nloc = nimg - first if nloc is None else nloc
    n = 3 * natoms
    out = np.empty((nloc, n, n))
    for i in range(nloc):
        path = os.path.join(folder, f"tmp_hessian_{first + i}.npy")
        out[i] = np.load(path) if os.path.exists(path) else np.eye(n)
    return out


def write_xyz_samples(folder, stem, element_list, geometries_ang, charge_and_multiplicity=(0, 1)):
    """One ``<stem>_<image>.xyz`` per image: atom count, "charge multiplicity", then ``El   x   y   z`` with the
    reference's ``{:2}   {:>17.12f}`` layout (fileio.py:441-446).  geometries_ang: (nimg, N, 3) in Angstrom."""
    os.makedirs(folder, exist_ok=True)
    G = np.asarray(geometries_ang, dtype=np.float64)
    paths = []
    for y in range(G.shape[0]):
        path = os.path.join(folder, f"{stem}_{y}.xyz")
        with open(path, "w") as w:
            w.write(str(len(element_list)) + "\n")
            w.write(str(charge_and_multiplicity[0]) + " " + str(charge_and_multiplicity[1]) + "\n")
            for e, row in zip(element_list, G[y]):
                w.write(f"{e:2}   {float(row[0]):>17.12f}   {float(row[1]):>17.12f}   {float(row[2]):>17.12f}\n")
        paths.append(path)
    return paths


def read_xyz_sample(path):
    """-> (elements, (N, 3) Angstrom, (charge, multiplicity)) of a file written by write_xyz_samples / the reference."""
    with open(path) as f:
        lines = f.read().splitlines()
    n = int(lines[0].split()[0])
    cm = tuple(int(v) for v in lines[1].split()[:2])
    elems, xyz = [], []
    for ln in lines[2:2 + n]:
        p = ln.split()
        elems.append(p[0]); xyz.append([float(v) for v in p[1:4]])
    return elems, np.array(xyz), cm
