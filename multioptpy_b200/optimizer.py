"""Drop-in for ``multioptpy.optimizer.CalculateMoveVector`` restricted to the quasi-Newton RFO
families of the hot path (``rsirfo_*``; ``rsprfo_*`` when the P-RFO kernel is present).

``calc_move_vector`` keeps the reference signature and return convention
(optimizer.py:740-818): ``(new_geometry [Angstrom], move_vector [Bohr], optimizer_instances)``
with ``new_geometry = (geom - move) * 0.52917721067``.  Inputs may be NumPy ``(N,3)`` arrays (one
structure, the reference calling convention) or CUDA tensors ``(B,N,3)`` (B structures per call,
everything stays on the device).  ``crsirfo_*`` with a ``projection_constraint`` object builds the constrained
RS-I-RFO drop-in (optimizer.py:450-451).  Enhancement chains (lookahead, DIIS, ...), first-order optimizers and the
mode-following RSIRFO subclasses are outside the scope and raise.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._lib import MopError
from .Optimizer.rsirfo import RSIRFO
from .Optimizer.trust_radius import TrustRadius
from .Parameters.tables import BOHR2ANG

_OUT_OF_SCOPE = ["mf_rsirfo", "lookahead", "lars", "linesearch", "diis", "coordinate_locking",
                 "component_wise_scaling", "gpr_step", "gan_step", "rl_step", "geodesic_step", "trim"]


class CalculateMoveVector:
    def __init__(self, DELTA, element_list, saddle_order=0, FC_COUNT=-1, temperature=0.0, model_hess_flag=None,
                 max_trust_radius=None, min_trust_radius=None, device="cuda", **kwargs):
        self.DELTA = DELTA
        self.temperature = temperature
        self.FC_COUNT = FC_COUNT
        self.device = torch.device(device)
        self.max_trust_radius = max_trust_radius
        self.min_trust_radius = min_trust_radius
        self.CALC_TRUST_RADII = TrustRadius(device=device)
        if self.max_trust_radius is not None:
            if self.max_trust_radius <= 0.0:
                raise ValueError("max_trust_radius must be greater than 0.0")
            self.CALC_TRUST_RADII.set_max_trust_radius(self.max_trust_radius)
        if self.max_trust_radius is None:                       # optimizer.py:281-287
            self.max_trust_radius = 0.1 if saddle_order > 0 else 0.5
            self.trust_radii = self.max_trust_radius
        else:                                                   # :288-292
            if saddle_order > 0:
                self.trust_radii = min(self.max_trust_radius, 0.1)
            else:
                self.trust_radii = self.max_trust_radius if type(self.max_trust_radius) is float else 0.5
        if self.min_trust_radius is not None:
            if self.min_trust_radius <= 0.0:
                raise ValueError("min_trust_radius must be greater than 0.0")
            self.CALC_TRUST_RADII.set_min_trust_radius(self.min_trust_radius)
        self.min_trust_radius = min_trust_radius if min_trust_radius is not None else 0.01
        self.saddle_order = saddle_order
        self.iter = 0
        self.element_list = element_list
        self.model_hess_flag = model_hess_flag
        self.projection_constraint = kwargs.get("projection_constraint", None)    # optimizer.py:270
        self.newton_tag = []
        self._trust_t = None

    def initialization(self, method):
        """Name -> optimizer instances (optimizer.py:310-532), RFO families only."""
        instances = []
        self.newton_tag = []
        for m in method:
            low = m.lower()
            for bad in _OUT_OF_SCOPE:
                if bad in low:
                    raise MopError(f"optimizer option '{bad}' in '{m}' is outside the B200 hot-path scope")
            if "rsprfo" in low:
                try:
                    from .Optimizer.rsprfo import EnhancedRSPRFO
                except ImportError as exc:
                    raise MopError("rsprfo_* needs the P-RFO kernel") from exc
                opt = EnhancedRSPRFO(method=m, saddle_order=self.saddle_order, element_list=self.element_list,
                                     trust_radius_max=self.max_trust_radius, trust_radius_min=self.min_trust_radius,
                                     device=self.device)
            elif "crsirfo" in low and self.projection_constraint:   # optimizer.py:450-451 (falls through to RSIRFO otherwise)
                from .Optimizer.crsirfo import CRSIRFO
                opt = CRSIRFO(method=m, constraints=self.projection_constraint, saddle_order=self.saddle_order,
                              element_list=self.element_list, trust_radius_max=self.max_trust_radius,
                              trust_radius_min=self.min_trust_radius, device=self.device)
            elif "rsirfo" in low:
                opt = RSIRFO(method=m, saddle_order=self.saddle_order, element_list=self.element_list,
                             trust_radius_max=self.max_trust_radius, trust_radius_min=self.min_trust_radius,
                             device=self.device)
            else:
                raise MopError(f"optimizer '{m}' is outside the B200 hot-path scope (rsirfo_* / rsprfo_*)")
            opt.DELTA = 0.50                                     # quasi_newton_mapping[...]["delta"]
            instances.append(opt)
            self.newton_tag.append(True)
        if len(instances) != 1:
            raise MopError("exactly one optimizer method is supported (no force-switching blends)")
        self.method = method
        return instances

    # ---- outer trust radius (optimizer.py:534-568) -----------------------------------------
    def update_trust_radius_conditionally(self, optimizer_instances, B_e, pre_B_e, pre_B_g, pre_move_vector, geom):
        if self.FC_COUNT == -1 and self.model_hess_flag is None:
            return
        opt = optimizer_instances[0]
        H, Hb = opt.hessian, opt.bias_hessian
        if isinstance(H, torch.Tensor):
            B = H.shape[0]
            if self._trust_t is None:
                self._trust_t = torch.full((B,), float(self.trust_radii), dtype=torch.float64, device=H.device)
            tr = self.CALC_TRUST_RADII
            if tr._state is None:
                tr._state = torch.zeros(B, ops.TR_STATE, dtype=torch.float64, device=H.device)
            ops.outer_trust_radius(H, Hb, pre_B_g.reshape(B, -1).contiguous(), pre_move_vector.reshape(B, -1).contiguous(),
                                   B_e, pre_B_e, self._trust_t, tr._state, tr.min_trust_radius, tr.max_trust_radius)
        else:
            model_hess = np.asarray(H) + (np.asarray(Hb) if Hb is not None else 0.0)
            self.trust_radii = self.CALC_TRUST_RADII.update_trust_radii(
                B_e, pre_B_e, pre_B_g, pre_move_vector, model_hess, geom, self.trust_radii)

    # ---- the step (optimizer.py:740-818) ------------------------------------------------------
    def calc_move_vector(self, iter, geom_num_list, B_g, pre_B_g, pre_geom, B_e, pre_B_e, pre_move_vector,
                         initial_geom_num_list, g, pre_g, optimizer_instances, projection_constrain=False,
                         print_flag=True):
        if projection_constrain:
            raise MopError("projection constraints are outside the B200 hot-path scope")
        self.iter = iter
        opt = optimizer_instances[0]
        if isinstance(geom_num_list, torch.Tensor):
            return self._calc_batched(geom_num_list, B_g, pre_B_g, pre_geom, B_e, pre_B_e, pre_move_vector, g, pre_g,
                                      optimizer_instances)
        natom = len(geom_num_list)
        col = lambda a: np.asarray(a, dtype=np.float64).reshape(natom * 3, 1)
        geom = col(geom_num_list)
        self.geom_num_list = geom
        self.update_trust_radius_conditionally(optimizer_instances, B_e, pre_B_e, col(pre_B_g), col(pre_move_vector), geom)
        move = np.array(opt.run(geom, col(B_g), col(pre_B_g), col(pre_geom), B_e, pre_B_e, col(pre_move_vector),
                                col(initial_geom_num_list), col(g), col(pre_g)), dtype="float64")
        nrm = np.linalg.norm(move)
        if nrm > self.trust_radii:                               # :792-793
            move = self.trust_radii * move / nrm
        new_geometry = (geom - move).reshape(natom, 3) * BOHR2ANG
        return new_geometry, np.array(move.reshape(natom, 3), dtype="float64"), optimizer_instances

    def _calc_batched(self, geom, B_g, pre_B_g, pre_geom, B_e, pre_B_e, pre_move, g, pre_g, optimizer_instances):
        opt = optimizer_instances[0]
        B, natom = geom.shape[0], geom.shape[1]
        flat = lambda a: a.reshape(B, natom * 3).contiguous() if isinstance(a, torch.Tensor) and a.numel() else a
        x = flat(geom)
        if self._trust_t is None:
            self._trust_t = torch.full((B,), float(self.trust_radii), dtype=torch.float64, device=geom.device)
        if isinstance(pre_move, torch.Tensor) and pre_move.numel():
            self.update_trust_radius_conditionally(optimizer_instances, B_e, pre_B_e, flat(pre_B_g), flat(pre_move), x)
        move = opt.run(x, flat(B_g), flat(pre_B_g), flat(pre_geom), B_e, pre_B_e, flat(pre_move), x, flat(g), flat(pre_g))
        move = move.clone()
        new_geom, move = ops.clamp_and_move(x, move, self._trust_t)
        return new_geom.reshape(B, natom, 3), move.reshape(B, natom, 3), optimizer_instances
