"""Drop-ins for ``multioptpy.Potential.keep_potential.StructKeepPotential`` / ``StructKeepPotentialv2``,
``multioptpy.Potential.keep_angle_potential.StructKeepAnglePotential`` and
``multioptpy.Potential.keep_dihedral_angle_potential.StructKeepDihedralAnglePotential`` on the CUDA restraint kernel
(csrc/bias.cu).  ``calc_energy`` keeps the reference signature; ``calc_energy_grad_hess`` returns what the
aggregator obtains from ``torch.func.jacrev`` / ``hessian`` (Potential/potential.py:127-137)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


class _Restraint:
    kind = 0

    def __init__(self, device="cuda", **kwarg):
        self.config = kwarg
        self.device = torch.device(device)

    def _term(self, bias_pot_params):
        raise NotImplementedError

    def calc_energy_grad_hess(self, geom_num_list, bias_pot_params=[]):
        """geom (N,3) NumPy / tensor or (B,N,3) CUDA tensor -> (E (B,), grad (B,3N), hess (B,3N,3N)) tensors."""
        if isinstance(geom_num_list, torch.Tensor):
            xyz = geom_num_list.to(self.device, torch.float64)
        else:
            xyz = torch.as_tensor(np.ascontiguousarray(np.asarray(geom_num_list, dtype=np.float64))).to(self.device)
        if xyz.dim() == 2:
            xyz = xyz.unsqueeze(0)
        packed = ops.pack_bias_terms([self._term(bias_pot_params)], self.device)
        return ops.bias_terms(xyz.contiguous(), packed, 1)

    def calc_energy(self, geom_num_list, bias_pot_params=[]):
        E, _, _ = self.calc_energy_grad_hess(geom_num_list, bias_pot_params)
        return E[0] if E.numel() == 1 else E


def _pair(params, config, kkey, pkey):
    if len(params) == 0:
        return float(config[kkey]), float(config[pkey])
    return float(params[0]), float(params[1])


class StructKeepPotential(_Restraint):
    def _term(self, params):
        k, r0 = _pair(params, self.config, "keep_pot_spring_const", "keep_pot_distance")
        i, j = self.config["keep_pot_atom_pairs"]
        return (ops.BIAS_KEEP, [i - 1], [j - 1], k, r0)


class StructKeepPotentialv2(_Restraint):
    def _term(self, params):
        k, r0 = _pair(params, self.config, "keep_pot_v2_spring_const", "keep_pot_v2_distance")
        return (ops.BIAS_KEEP_V2, [a - 1 for a in self.config["keep_pot_v2_fragm1"]],
                [a - 1 for a in self.config["keep_pot_v2_fragm2"]], k, r0)


class StructKeepAnglePotential(_Restraint):
    def _term(self, params):
        k, th = _pair(params, self.config, "keep_angle_spring_const", "keep_angle_angle")
        return (ops.BIAS_KEEP_ANGLE, [a - 1 for a in self.config["keep_angle_atom_pairs"]], [], k, th)


def dihedral_phi0(params, config):
    """(k, phi0 in radians) exactly as keep_dihedral_angle_potential.py:76-91 obtains them: without parameters
    the configured angle becomes a float32 tensor before ``torch.deg2rad``; with parameters a tensor entry keeps
    its own dtype (the aggregator passes float64) and a Python number becomes float32."""
    if len(params) == 0:
        return float(config["keep_dihedral_angle_spring_const"]), float(torch.deg2rad(torch.tensor(config["keep_dihedral_angle_angle"])))
    deg = params[1]
    if not isinstance(deg, torch.Tensor):
        deg = torch.tensor(deg)
    return float(params[0]), float(torch.deg2rad(deg.detach()))


class StructKeepDihedralAnglePotential(_Restraint):
    def _term(self, params):
        k, phi0 = dihedral_phi0(params, self.config)
        return (ops.BIAS_KEEP_DIHEDRAL, [a - 1 for a in self.config["keep_dihedral_angle_atom_pairs"]], [], k, phi0)
