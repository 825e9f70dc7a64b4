"""Drop-in for ``multioptpy.MEP.pathopt_bneb_force.CaluculationBNEB`` (default NEB force)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import MopError


class CaluculationBNEB:
    def __init__(self, APPLY_CI_NEB=99999, device="cuda"):
        self.APPLY_CI_NEB = APPLY_CI_NEB
        self.device = torch.device(device)
        self.tau_list = []

    def calc_force(self, geometry_num_list, energy_list, gradient_list, optimize_num, element_list):
        """(nimg, N, 3) forces; stores the tangents for get_tau (pathopt_bneb_force.py:33-65)."""
        if optimize_num > self.APPLY_CI_NEB:
            raise MopError("CI-NEB branches (optimize_num > APPLY_CI_NEB) are not implemented on the device")
        X = np.asarray(geometry_num_list, dtype=np.float64)
        nimg, N, _ = X.shape
        n = 3 * N
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(self.device)
        xh = torch.zeros(nimg + 2, n, dtype=torch.float64, device=self.device); xh[1:-1] = t(X.reshape(nimg, n))
        Eh = torch.zeros(nimg + 2, dtype=torch.float64, device=self.device); Eh[1:-1] = t(np.asarray(energy_list, dtype=np.float64))
        g = t(np.asarray(gradient_list, dtype=np.float64).reshape(nimg, n))
        force, tau = ops.bneb_force(nimg, 0, xh, Eh, g)
        self.tau_list = [r.reshape(N, 3) for r in tau.cpu().numpy()]
        return force.cpu().numpy().reshape(nimg, N, 3)

    def get_tau(self, node_num):
        if len(self.tau_list) == 0:
            raise ValueError("Tangent list is empty. Calculate forces first.")
        return self.tau_list[node_num]
