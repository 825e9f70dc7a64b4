"""Drop-in for ``multioptpy.ModelHessian.lindh.LindhApproxHessian`` (ModelHessian/lindh.py:11-165).

Implemented on the device: the connectivity tables, the diagonal redundant-internal force
constants (bond / angle / dihedral decay terms, reduced-mass scaling of bonds, Lennard-Jones and
electrostatic terms of non-bonded pairs) and ``B^T diag(k) B`` with the all-pairs distance
B matrix, then the TR/ROT projection.  The reference's ``K`` term multiplies second derivatives by
an internal-coordinate gradient obtained from ``np.linalg.solve`` on the singular ``B B^T`` and
indexes it by a bond/angle/dihedral counter although its rows are atom pairs — a 1e-13 shift of the
coordinates changes it by O(1), so that gradient has no reproducible value.  ``main``
therefore adds ``K`` only when the caller supplies the internal gradient (``int_grad=``, e.g. the
reference's own ``cartgrad2RICgrad`` output); without it ``main`` equals the reference for a zero
gradient.  When a NON-ZERO ``cart_gradient`` arrives without ``int_grad`` the omission is reported: a
``UserWarning`` and ``ops.ST_LINDH_NO_K`` in ``last_status`` (one int32 per structure, OR-ed with the
kernel's table-overflow bit 1)."""
from __future__ import annotations

import warnings

import numpy as np
import torch

from .. import ops
from ..Parameters.tables import (ATOMIC_MASS, UFF_EFFECTIVE_CHARGE, UFF_VDW_DISTANCE, UFF_VDW_WELL_DEPTH,
                                 covalent_radius)

_FIRST = {"H", "He"}
_SECOND = {"Li", "Be", "B", "C", "N", "O", "F", "Ne"}


def lindh_atom_params(element_list):
    """(N, 6): covalent radius, period index, mass, UFF distance, UFF well depth, UFF charge."""
    rows = []
    for e in element_list:
        per = 0 if e in _FIRST else (1 if e in _SECOND else 2)
        rows.append([covalent_radius(e), float(per), ATOMIC_MASS[e], UFF_VDW_DISTANCE[e], UFF_VDW_WELL_DEPTH[e],
                     UFF_EFFECTIVE_CHARGE[e]])
    return np.array(rows, dtype=np.float64)


class LindhApproxHessian:
    def __init__(self, device="cuda"):
        self.force_const_list = [0.45, 0.15, 0.005]
        self.device = torch.device(device)
        self.last_status = None

    def _k_omitted(self, cart_gradient):
        if cart_gradient is None:
            return False
        if isinstance(cart_gradient, torch.Tensor):
            nz = bool((cart_gradient != 0).any().item())
        else:
            nz = bool(np.any(np.asarray(cart_gradient, dtype=np.float64) != 0.0))
        if nz:
            warnings.warn("LindhApproxHessian(B200): non-zero cart_gradient without int_grad - the reference's K term "
                          "(second derivatives times an internal gradient from a singular solve, SURVEY H2) is "
                          "omitted; pass int_grad= to add it", UserWarning, stacklevel=3)
        return nz

    def guess_lindh_diagonal(self, coord, element_list):
        """Diagonal of guess_lindh_hessian (lindh.py:79-143): force constant per atom pair."""
        xyz = torch.as_tensor(np.ascontiguousarray(np.asarray(coord, dtype=np.float64)).reshape(1, -1, 3)).to(self.device)
        _, kd, _, _ = ops.lindh_hessian(xyz, lindh_atom_params(element_list), want_kdiag=True)
        return kd[0].cpu().numpy()

    def main(self, coord, element_list, cart_gradient=None, int_grad=None):
        prm = lindh_atom_params(element_list)
        if int_grad is not None:
            return self._main_with_int_grad(coord, element_list, prm, int_grad)
        no_k = ops.ST_LINDH_NO_K if self._k_omitted(cart_gradient) else 0
        if isinstance(coord, torch.Tensor):
            H, _, _, status = ops.lindh_hessian(coord, prm)
            self.last_status = status | no_k       # bit 0: table capacity exceeded (caller's to check, no host sync here)
            return H
        xyz = torch.as_tensor(np.ascontiguousarray(np.asarray(coord, dtype=np.float64)).reshape(1, -1, 3)).to(self.device)
        H, _, _, status = ops.lindh_hessian(xyz, prm)
        self.last_status = status | no_k
        if int(status[0].item()) != 0:
            raise ops.MopError("Lindh model Hessian: connectivity table capacity exceeded")
        return H[0].cpu().numpy()

    def _main_with_int_grad(self, coord, element_list, prm, int_grad):
        """B^T diag(k) B + K(int_grad), nan_to_num, TR/ROT projection (lindh.py:153-164)."""
        from ..Utils.bond_connectivity import radii_array
        tensor = isinstance(coord, torch.Tensor)
        if tensor:
            xyz, q = coord.contiguous(), int_grad.contiguous()
        else:
            xyz = torch.as_tensor(np.ascontiguousarray(np.asarray(coord, dtype=np.float64)).reshape(1, -1, 3)).to(self.device)
            q = torch.as_tensor(np.asarray(int_grad, dtype=np.float64).reshape(1, -1)).to(self.device).contiguous()
        B, N, _ = xyz.shape
        _, kd, _, _ = ops.lindh_hessian(xyz, prm, want_kdiag=True)
        bonds, angles, dihs, counts, _ = ops.connectivity(xyz, radii_array(element_list))
        K = ops.ric_kmatrix(xyz, bonds, angles, dihs, counts, q)
        raw = torch.nan_to_num(ops.ric_hess_to_cart(xyz, kd, K), nan=0.0)
        H, _, _ = ops.project_trrot(raw.contiguous(), xyz.reshape(B, 3 * N))
        return H if tensor else H[0].cpu().numpy()
