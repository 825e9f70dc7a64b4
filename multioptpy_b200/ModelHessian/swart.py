"""Drop-in for ``multioptpy.ModelHessian.swart.SwartApproxHessian``
(ModelHessian/swart.py:27-355) on the CUDA model-Hessian kernel."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..Parameters.tables import SWART_RADII


def swart_radii(element_list) -> np.ndarray:
    """_get_radii_array (swart.py:60-62): the model's own Bohr table, 1.0 for unknown elements."""
    return np.array([SWART_RADII.get(e.capitalize(), 1.0) for e in element_list], dtype=np.float64)


class SwartApproxHessian:
    def __init__(self, device="cuda"):
        self.wthr = 0.3
        self.f = 0.12
        self.tolth = 0.2
        self.device = torch.device(device)
        self.cart_hess = None

    def main(self, coord, element_list, cart_gradient=None):
        """coord: (N,3) Bohr NumPy array -> (3N,3N) NumPy array; or (B,N,3) CUDA tensor ->
        (B,3N,3N) tensor.  cart_gradient is unused by the reference model as well."""
        radii = swart_radii(element_list)
        if isinstance(coord, torch.Tensor):
            H, _, _ = ops.swart_hessian(coord, radii)
            return H
        xyz = torch.as_tensor(np.ascontiguousarray(np.asarray(coord, dtype=np.float64)).reshape(1, -1, 3)).to(self.device)
        H, Hraw, _ = ops.swart_hessian(xyz, radii, want_raw=True)
        self.cart_hess = Hraw[0].cpu().numpy()
        return H[0].cpu().numpy()
