"""FischerD3ApproxHessianOld (ModelHessian/fischerd3old.py): Fischer-Almloef model Hessian with the reference's
simplified D3(BJ) pair term for non-bonded pairs, the linear-angle skips and the sin^2 damping of the torsion force
constants - the model Hessian a bare `-modelhess` selects (interface.py:184-191).  The work is in csrc/model_hessian.cu
(k_model_hessian, kind 2); this is the host mirror: NumPy (N, 3) input -> NumPy (3N, 3N) like the reference, CUDA
tensors with a leading batch dimension -> one launch for the batch."""
from __future__ import annotations

import numpy as np

from .. import ops
from ..Parameters import tables


def d3_atom_params(element_list) -> np.ndarray:
    """(N, 4): covalent radius (Bohr), D2 C6 (hartree bohr^6), D3 r4r2, D2 van der Waals radius (Bohr) - the per-atom
    data of get_c6_coefficient / get_c8_coefficient / get_r0_value (fischerd3old.py:48-72).  Elements outside the
    reference's D2 table raise, as the reference's dictionary lookup does."""
    out = np.empty((len(element_list), 4))
    for a, e in enumerate(element_list):
        out[a, 0] = tables.covalent_radius(e)
        out[a, 1] = tables.D2_C6[e]
        out[a, 2] = tables.D3_R4R2.get(e, tables.D3_R4R2_DEFAULT)
        out[a, 3] = tables.D2_VDW_RADII[e]
    return out


def d3_dynamic_atom_params(element_list) -> np.ndarray:
    """(N, 5): d3_atom_params + the reference coordination number of FischerD3ApproxHessian (fischerd3.py:27-44;
    unknown elements: 4, :222)."""
    base = d3_atom_params(element_list)
    ref = np.array([[float(tables.D3_REF_CN.get(e, tables.D3_REF_CN_DEFAULT))] for e in element_list])
    return np.concatenate([base, ref], axis=1)


class FischerD3ApproxHessianOld:
    def __init__(self, device=None):
        self.device = device
        self.bond_factor = 1.3
        self.cart_hess = None
        self.last_status = None

    def main(self, coord, element_list, cart_gradient=None):
        import torch
        prm = d3_atom_params(element_list)
        if isinstance(coord, torch.Tensor):
            H, counts, status = ops.fischer_d3old_hessian(coord, prm)
            self.last_status = status
            return H
        dev = self.device or "cuda:0"
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(coord, dtype=np.float64)[None])).to(dev)
        H, counts, status = ops.fischer_d3old_hessian(x, prm)
        self.last_status = status
        self.cart_hess = H[0].cpu().numpy()
        return self.cart_hess


class FischerD3ApproxHessian:
    """ModelHessian/fischerd3.py: the 'dynamic D3' variant the AutoTS configurations select (`fischerd3`)."""

    def __init__(self, device=None):
        self.device = device
        self.cart_hess = None
        self.last_status = None

    def main(self, coord, element_list, cart_gradient=None):
        import torch
        prm = d3_dynamic_atom_params(element_list)
        if isinstance(coord, torch.Tensor):
            H, counts, status = ops.fischer_d3old_hessian(coord, prm, dynamic=True)
            self.last_status = status
            return H
        dev = self.device or "cuda:0"
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(coord, dtype=np.float64)[None])).to(dev)
        H, counts, status = ops.fischer_d3old_hessian(x, prm, dynamic=True)
        self.last_status = status
        self.cart_hess = H[0].cpu().numpy()
        return self.cart_hess
