"""Drop-in for ``multioptpy.Optimizer.trust_radius.TrustRadius`` (composite outer trust radius,
Optimizer/trust_radius.py:3-206) on the CUDA kernel ``mop_outer_trust_radius``; works for one
structure (NumPy, reference signature) or a batch (CUDA tensors)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


class TrustRadius:
    def __init__(self, initial_trust_radius=0.3, min_trust_radius=0.01, max_trust_radius=0.5, history_size=5,
                 adaptive_factor_scale=0.8, energy_precision_threshold=1e-8, quality="Normal", device="cuda"):
        if history_size != 5 or adaptive_factor_scale != 0.8 or energy_precision_threshold != 1e-8:
            raise ops.MopError("TrustRadius(B200): history_size / adaptive_factor_scale / "
                               "energy_precision_threshold are baked into the kernel")
        self.trust_radius = initial_trust_radius
        self.min_trust_radius = min_trust_radius
        self.max_trust_radius = max_trust_radius
        self.device = torch.device(device)
        self._state = None

    def set_min_trust_radius(self, v):
        self.min_trust_radius = v

    def set_max_trust_radius(self, v):
        self.max_trust_radius = v

    @property
    def iteration_count(self):
        return 0 if self._state is None else int(self._state[0, 0].item())

    def update_trust_radii(self, B_e, pre_B_e, pre_B_g, pre_move_vector, model_hess, geom_num_list, trust_radii,
                           atom_types=None, constraints=None):
        """Reference signature (:120-146).  Tensors with a leading batch dimension update a
        (B,) trust tensor in place; NumPy inputs return a float."""
        if isinstance(model_hess, torch.Tensor):
            B = model_hess.shape[0]
            if self._state is None:
                self._state = torch.zeros(B, ops.TR_STATE, dtype=torch.float64, device=model_hess.device)
            return ops.outer_trust_radius(model_hess, None, pre_B_g.reshape(B, -1).contiguous(),
                                          pre_move_vector.reshape(B, -1).contiguous(), B_e, pre_B_e, trust_radii,
                                          self._state, self.min_trust_radius, self.max_trust_radius)
        n = np.asarray(model_hess).shape[0]
        dev = self.device
        if self._state is None:
            self._state = torch.zeros(1, ops.TR_STATE, dtype=torch.float64, device=dev)
        t = lambda a, shape: torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(shape)).to(dev)
        trust = torch.tensor([float(trust_radii)], dtype=torch.float64, device=dev)
        ops.outer_trust_radius(t(model_hess, (1, n, n)), None, t(pre_B_g, (1, n)), t(pre_move_vector, (1, n)),
                               torch.tensor([float(B_e)], dtype=torch.float64, device=dev),
                               torch.tensor([float(pre_B_e)], dtype=torch.float64, device=dev), trust, self._state,
                               self.min_trust_radius, self.max_trust_radius)
        return float(trust.item())
