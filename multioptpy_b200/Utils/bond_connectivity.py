"""Drop-in for ``multioptpy.Utils.bond_connectivity.BondConnectivity`` (coordinates in Bohr).

The tables come from the CUDA kernel (csrc/connectivity.cuh) and are bit-exact with the
reference's nested Python loops (Utils/bond_connectivity.py:34-134): same indices, same order.
Batched use: ``connectivity_tables_batched(xyz (B,N,3) tensor, element_list)``.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..Parameters.tables import covalent_radius


def radii_array(element_list):
    return np.array([covalent_radius(e) for e in element_list], dtype=np.float64)


class BondConnectivity:
    def __init__(self, device="cuda"):
        self.covalent_radii_lib_func = covalent_radius
        self.covalent_radii_threshold = 1.1
        self.device = torch.device(device)

    def connectivity_tables_batched(self, xyz: torch.Tensor, element_list, caps=None):
        """xyz: (B, N, 3) float64 CUDA tensor.  Returns (bonds, angles, dihedrals, counts)."""
        bonds, angles, dihs, counts, status = ops.connectivity(
            xyz, radii_array(element_list), self.covalent_radii_threshold, caps)
        if bool((status != 0).any()):
            raise ops.MopError("connectivity table capacity exceeded")
        return bonds, angles, dihs, counts

    def connectivity_table(self, coord, element_list):
        """[bond_table, angle_table, dihedral_table] as lists of index lists
        (bond_connectivity.py:130-134)."""
        xyz = torch.as_tensor(np.ascontiguousarray(np.asarray(coord, dtype=np.float64)).reshape(1, -1, 3)).to(self.device)
        bonds, angles, dihs, counts = self.connectivity_tables_batched(xyz, element_list)
        c = counts[0].cpu().numpy()
        return [bonds[0, :c[0]].cpu().numpy().tolist(), angles[0, :c[1]].cpu().numpy().tolist(),
                dihs[0, :c[2]].cpu().numpy().tolist()]

    def bond_connect_matrix(self, element_list, coord):
        """0/1 matrix (bond_connectivity.py:34-41)."""
        n = len(element_list)
        table = self.connectivity_table(coord, element_list)[0]
        m = np.zeros((n, n), dtype=int)
        for i, j in table:
            m[i, j] = m[j, i] = 1
        return m
