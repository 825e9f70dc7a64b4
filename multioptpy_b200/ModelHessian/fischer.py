"""Drop-in for ``multioptpy.ModelHessian.fischer.FischerApproxHessian``
(ModelHessian/fischer.py:9-236) on the CUDA model-Hessian kernel."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..Utils.bond_connectivity import radii_array


class FischerApproxHessian:
    def __init__(self, device="cuda"):
        self.bond_factor = 1.3
        self.device = torch.device(device)
        self.last_status = None   # [B] int32 device tensor of the last call: 1 = connectivity table capacity exceeded

    def main(self, coord, element_list, cart_gradient=None):
        """coord: (N,3) Bohr NumPy array -> (3N,3N) NumPy array; or (B,N,3) CUDA tensor ->
        (B,3N,3N) tensor.  cart_gradient is unused by the reference model as well."""
        if isinstance(coord, torch.Tensor):
            H, _, status = ops.fischer_hessian(coord, radii_array(element_list))
            self.last_status = status   # tensor mode stays asynchronous: the caller checks it (no host sync here)
            return H
        xyz = torch.as_tensor(np.ascontiguousarray(np.asarray(coord, dtype=np.float64)).reshape(1, -1, 3)).to(self.device)
        H, _, status = ops.fischer_hessian(xyz, radii_array(element_list))
        self.last_status = status
        if int(status[0].item()) != 0:
            raise ops.MopError("Fischer model Hessian: connectivity table capacity exceeded")
        return H[0].cpu().numpy()
