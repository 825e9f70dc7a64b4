"""Drop-in for ``multioptpy.Potential.AFIR_potential.AFIRPotential`` with analytic gradient and
Hessian from the CUDA kernel (csrc/afir.cu) instead of torch.func autograd on the CPU
(Potential/potential.py:127-152)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..Parameters.tables import covalent_radius


class AFIRPotential:
    def __init__(self, **kwarg):
        self.config = kwarg
        self.device = torch.device(kwarg.get("device", "cuda"))
        elems = self.config["element_list"]
        # the reference builds the radii as float32 tensors (AFIR_potential.py:43-44)
        self.radii_f32 = torch.tensor([covalent_radius(e) for e in elems], dtype=torch.float32, device=self.device)
        self.frag1 = torch.tensor([int(i) - 1 for i in self.config["AFIR_Fragm_1"]], dtype=torch.int32, device=self.device)
        self.frag2 = torch.tensor([int(i) - 1 for i in self.config["AFIR_Fragm_2"]], dtype=torch.int32, device=self.device)

    def _xyz(self, geom):
        if isinstance(geom, torch.Tensor) and geom.is_cuda:
            x = geom.detach().to(torch.float64)
        else:
            x = torch.as_tensor(np.asarray(geom.detach().cpu() if isinstance(geom, torch.Tensor) else geom,
                                           dtype=np.float64)).to(self.device)
        return (x.reshape(1, -1, 3) if x.dim() == 2 else x).contiguous()

    def _gamma(self, bias_pot_params, B):
        g = bias_pot_params[0] if not isinstance(bias_pot_params, (int, float)) else bias_pot_params
        if isinstance(g, torch.Tensor) and g.numel() == B and B > 1:
            return g.detach().to(self.device, torch.float64).reshape(B).contiguous()
        return torch.full((B,), float(g), dtype=torch.float64, device=self.device)

    def calc_energy(self, geom_num_list, bias_pot_params):
        """Energy in Hartree: 0-d tensor for an (N,3) geometry, (B,) for a (B,N,3) batch."""
        x = self._xyz(geom_num_list)
        E, _, _ = ops.afir(x, self.frag1, self.frag2, self.radii_f32, self._gamma(bias_pot_params, x.shape[0]),
                           want_grad=False, want_hess=False)
        single = not (isinstance(geom_num_list, torch.Tensor) and geom_num_list.dim() == 3)
        return E[0] if single else E

    def calc_energy_grad_hess(self, geom_num_list, bias_pot_params):
        """(E, grad (.., N, 3), hess (.., 3N, 3N)) — what jacrev / hessian give in the reference."""
        x = self._xyz(geom_num_list)
        B, N, _ = x.shape
        E, g, H = ops.afir(x, self.frag1, self.frag2, self.radii_f32, self._gamma(bias_pot_params, B))
        g = g.reshape(B, N, 3)
        single = not (isinstance(geom_num_list, torch.Tensor) and geom_num_list.dim() == 3)
        return (E[0], g[0], H[0]) if single else (E, g, H)
