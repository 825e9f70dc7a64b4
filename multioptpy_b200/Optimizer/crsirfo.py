"""Drop-in for ``multioptpy.Optimizer.crsirfo.CRSIRFO`` (Optimizer/crsirfo.py:5-170): RS-I-RFO in the null space of a
set of constraint vectors.

The reference builds a null-space basis U from a full SVD, steps in the subspace and lifts the step with U; the
device path projects in the full space (``mop_constraint_project``: same span rule, same step, see csrc/crsirfo.cu),
runs the unchanged spectrum / step kernels (``mop_rsirfo_spectral_step``) and applies the reference's explicit
convergence test (``mop_crsirfo_finalize``).  The Hessian update is RSIRFO's own (raw gradients, rsirfo.py:1316-1372).

``constraints`` is the reference's constraint object: ``_get_all_constraint_vectors(geom (N, 3)) -> (k, 3N)`` and
``adjust_init_coord(geom (N, 3)) -> (N, 3)`` (SHAKE-like correction) are called on the HOST per structure, exactly as
the reference calls them; in tensor mode precomputed rows can be passed instead (``constraint_vectors=``, and the
corrected geometry as ``geom_num_list`` with ``shake_displacement=``).  At most 12 constraint rows; 3N <= 160 (the
shared-memory spectrum / step kernels - larger systems raise ``MopError``).

Reproduced quirks: the bias Hessian is added INTO ``self.hessian`` (crsirfo.py:76,86 ``+=`` on an alias) - once per
call, twice when the SHAKE correction fires; ``eigvals`` of the last step are the subspace spectrum (the ``rank``
largest entries of the full-space spectrum belong to the constrained directions and are dropped).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import MopError
from .rsirfo import RSIRFO


class CRSIRFO(RSIRFO):
    def __init__(self, constraints=None, **config):
        super().__init__(**config)
        self.constraints_obj = constraints
        self.null_space_basis = None
        self.svd_threshold = config.get("svd_threshold", 1e-5)
        self.proj_grad_converged = False
        self.last_rank = None

    # ---- host-side calls into the reference's constraint object (one structure at a time, as the reference) --------
    def _constraint_rows(self, x):
        """x (B, n) device tensor -> (B, k, n) device tensor of raw constraint rows, or None (unconstrained)."""
        if self.constraints_obj is None:
            return None
        xs = x.detach().cpu().numpy()
        rows = []
        for b in range(xs.shape[0]):
            Bm = self.constraints_obj._get_all_constraint_vectors(xs[b].reshape(-1, 3))
            if Bm is None or len(Bm) == 0:
                return None
            rows.append(np.asarray(Bm, dtype=np.float64).reshape(len(Bm), -1))
        if len({r.shape for r in rows}) != 1:
            raise MopError("CRSIRFO: every structure of a batch must carry the same number of constraint rows")
        return torch.from_numpy(np.ascontiguousarray(np.stack(rows))).to(x.device)

    def _shake(self, x):
        """adjust_init_coord per structure: (corrected x (B, n), displacement (B, n)) on the device."""
        xs = x.detach().cpu().numpy()
        out = np.stack([np.asarray(self.constraints_obj.adjust_init_coord(xs[b].reshape(-1, 3)), dtype=np.float64).ravel()
                        for b in range(xs.shape[0])])
        xc = torch.from_numpy(np.ascontiguousarray(out)).to(x.device)
        return xc, xc - x

    def run(self, geom_num_list, B_g, pre_B_g=[], pre_geom=[], B_e=0.0, pre_B_e=0.0, pre_move_vector=[],
            initial_geom_num_list=[], g=[], pre_g=[], constraint_vectors=None, shake_displacement=None):
        if self.hessian is None:
            raise ValueError("Hessian matrix must be set before running optimization")
        if isinstance(geom_num_list, torch.Tensor):
            return self._run_constrained(geom_num_list, B_g, pre_geom, B_e, g, pre_g, constraint_vectors, shake_displacement)
        # NumPy mode: one structure through the tensor path; the Hessian is written back into the caller's array
        x = np.asarray(geom_num_list, dtype=np.float64).reshape(1, -1)
        n = x.shape[1]
        H_host = self.hessian
        if not isinstance(H_host, np.ndarray):
            raise MopError("NumPy-mode run() needs a NumPy Hessian (set_hessian)")
        Hb_host = self.bias_hessian
        dev = self.device
        T = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(1, -1))).to(dev)
        self.hessian = torch.from_numpy(np.ascontiguousarray(H_host.reshape(1, n, n))).to(dev)
        self.bias_hessian = None if Hb_host is None else torch.from_numpy(np.ascontiguousarray(np.asarray(Hb_host).reshape(1, n, n))).to(dev)
        have = len(pre_g) > 0 and len(pre_geom) > 0
        try:
            mv = self._run_constrained(T(x), T(B_g), T(pre_geom) if have else [], float(B_e), T(g), T(pre_g) if have else [],
                                       constraint_vectors, shake_displacement)
            H_host[...] = self.hessian[0].cpu().numpy()
        finally:
            self.hessian, self.bias_hessian = H_host, Hb_host
        return mv[0].cpu().numpy().reshape(-1, 1)

    def _run_constrained(self, x, Bg, x_prev, Be, g, g_prev, rows, shake):
        if x.dim() == 3:
            x, Bg, g = x.squeeze(-1), Bg.squeeze(-1), g.squeeze(-1)
        B, n = x.shape
        self._ensure_state(B)
        if not isinstance(Be, torch.Tensor):
            Be = torch.full((B,), float(Be), dtype=torch.float64, device=x.device)
        # 0. SHAKE-like correction and gradient transport (crsirfo.py:60-82)
        if shake is None and self.constraints_obj is not None:
            x, shake = self._shake(x)
        if shake is not None and self.bias_hessian is not None and bool((shake.norm(dim=1) > 1e-6).any()):
            if not bool((shake.norm(dim=1) > 1e-6).all()):
                raise MopError("CRSIRFO: a bias Hessian with a SHAKE correction on part of the batch only is not supported")
            ops.add_inplace(self.hessian, self.bias_hessian)             # H_eff += bias (alias of self.hessian)
        # (the transport term H delta is formed inside mop_constraint_project - but with the Hessian BEFORE the update,
        # as the reference orders it: take it now)
        gfull = Bg
        if shake is not None:
            none_rows = torch.zeros(B, 1, n, dtype=torch.float64, device=x.device)
            _, gfull, _ = ops.constraint_project(none_rows, self.hessian, Bg.contiguous(), shake=shake.contiguous(), want_hessian=False)
        # 1. Hessian update with the raw gradients (crsirfo.py:85-86 -> rsirfo.py:1316-1372)
        have_hist = (self.prev_geometry is not None and self.prev_gradient is not None and isinstance(x_prev, torch.Tensor)
                     and isinstance(g_prev, torch.Tensor) and x_prev.numel() > 0 and g_prev.numel() > 0)
        if have_hist:
            if x_prev.dim() == 3:
                x_prev, g_prev = x_prev.squeeze(-1), g_prev.squeeze(-1)
            ops.hessian_update(self.hessian, (x - x_prev).contiguous(), (g - g_prev).contiguous(), self._method_id,
                               inplace=True, rsirfo_guards=True)
        if self.bias_hessian is not None:
            ops.add_inplace(self.hessian, self.bias_hessian)             # hessian_full += bias (crsirfo.py:88-90)
        # 2. projection
        if rows is None:
            rows = self._constraint_rows(x)
        if rows is None:                                                 # no constraints: U = identity
            rows = torch.zeros(B, 1, n, dtype=torch.float64, device=x.device)
        Hp, gp, rank = ops.constraint_project(rows.contiguous(), self.hessian, gfull.contiguous(), svd_threshold=self.svd_threshold)
        self.last_rank = rank
        # 3. RFO in the subspace = the spectrum / step kernels on (Hp, gp); |gp| is the subspace gradient norm
        state_before = self._state.clone()
        if self._out is not None and tuple(self._out["move"].shape) != (B, n):
            self._out = None
        self._out = ops.rsirfo_spectral_step(Hp, gp, gp, self._state, saddle_order=self.saddle_order, neb_mode=self.NEB_mode,
                                             Be=Be, trust_min=self.trust_radius_min, trust_max=self.trust_radius_max,
                                             out=self._out)
        ops.crsirfo_finalize(gp, Be, state_before, self._state, self._out, self.gradient_norm_threshold)
        self.last_status = self._out["status"]
        self.proj_grad_converged = bool(((self.last_status & ops.ST_CONSTR_CONVERGED) != 0).all())
        self.prev_geometry, self.prev_gradient, self.prev_energy = x, Bg, Be
        self.iteration += 1
        return self._out["move"]
