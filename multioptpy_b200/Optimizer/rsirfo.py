"""Drop-in for ``multioptpy.Optimizer.rsirfo.RSIRFO`` backed by the B200 kernels.

Same constructor keywords, attributes and ``run`` signature as the reference
(Optimizer/rsirfo.py:9-126,285-490,1734-1754), with one extension: every array
argument may be a float64 CUDA tensor with a leading batch dimension, in which
case B independent structures are stepped by one kernel sequence and all state
stays on the device.

* NumPy mode (reference calling convention): ``(n,1)`` / ``(n,)`` arrays in,
  ``(n,1)`` array out; the Hessian given to ``set_hessian`` is kept BY REFERENCE
  and overwritten in place by the update, as the reference's aliasing does
  (SURVEY H4).
* Tensor mode: ``(B,n)`` tensors in, ``(B,n)`` tensor out; ``set_hessian`` takes a
  ``(B,n,n)`` tensor that is updated in place.

There is no CPU fallback: construction succeeds without a GPU, ``run`` raises.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import MopError

_DEFAULTS = dict(alpha0=1.0, max_micro_cycles=40, small_eigval_thresh=1e-6, alpha_max=1000.0,
                 alpha_step_max=10.0, good_step_threshold=0.75, poor_step_threshold=0.25,
                 trust_radius_increase_factor=1.2, trust_radius_decrease_factor=0.5,
                 step_norm_tolerance=1e-3, use_adaptive_trust_radius=True,
                 adaptive_trust_gradient_norm_threshold=1e-2, max_curvature_factor=2.5,
                 negative_curvature_safety=0.8, use_level_shift=False, level_shift_value=1e-5,
                 auto_level_shift=True, condition_number_threshold=1e8)


class RSIRFO:
    def __init__(self, **config):
        for key, default in _DEFAULTS.items():
            if key in config and config[key] != default:
                raise MopError(f"RSIRFO(B200): non-default {key}={config[key]!r} is baked into the "
                               f"kernels (default {default!r}) and cannot be changed")
        self.saddle_order = config.get("saddle_order", 1)
        self.hessian_update_method = config.get("method", "auto")
        if self.saddle_order == 0:                                  # rsirfo.py:36-41
            self.trust_radius_initial = config.get("trust_radius", 0.5)
            self.trust_radius_max = config.get("trust_radius_max", 0.5)
        else:
            self.trust_radius_initial = config.get("trust_radius", 0.1)
            self.trust_radius_max = config.get("trust_radius_max", 0.1)
        self.trust_radius = self.trust_radius_initial
        self.trust_radius_min = config.get("trust_radius_min", 0.01)
        self.energy_change_threshold = config.get("energy_change_threshold", 1e-6)
        self.gradient_norm_threshold = config.get("gradient_norm_threshold", 1e-4)
        self.debug_mode = config.get("debug_mode", False)
        self.display_flag = config.get("display_flag", True)
        self.device = torch.device(config.get("device", "cuda"))
        self.eigh_algo = config.get("eigh_algo", "auto")
        self.Initialization = True
        self.hessian = None
        self.bias_hessian = None
        self.predicted_energy_changes = []
        self.actual_energy_changes = []
        self.prev_geometry = None
        self.prev_gradient = None
        self.prev_energy = None
        self.converged = False
        self.iteration = 0
        self.roots = list(range(self.saddle_order))
        self.NEB_mode = False
        self.level_shift_applied = False
        self.last_status = None
        self._state = None     # (B, 16) device tensor
        self._out = None
        self._packed = False
        self._method_id = ops.resolve_update_method(self.hessian_update_method)

    # ---- reference API ------------------------------------------------------------
    def switch_NEB_mode(self):
        self.NEB_mode = not self.NEB_mode

    def log(self, message, force=False):
        if self.display_flag and (force or self.debug_mode):
            print(message)

    def set_hessian(self, hessian):
        self.hessian = hessian
        self._packed = False

    def set_hessian_packed(self, packed):
        """Tensor mode: keep the (symmetric) Hessian batch as packed lower triangles (B, n (n + 1) / 2) - half the
        bytes in HBM and over PCIe (n <= 160).  ``get_hessian()`` unpacks on demand."""
        self.hessian = packed
        self._packed = True

    def set_bias_hessian(self, bias_hessian):
        self.bias_hessian = bias_hessian

    def get_hessian(self):
        if getattr(self, "_packed", False) and self.hessian is not None:
            np_ = self.hessian.shape[1]
            n = int(round((np.sqrt(8.0 * np_ + 1.0) - 1.0) / 2.0))
            return ops.unpack_lower(self.hessian, n)
        return self.hessian

    def get_bias_hessian(self):
        return self.bias_hessian

    def is_converged(self):
        return self.converged

    def get_predicted_energy_changes(self):
        return self.predicted_energy_changes

    def get_actual_energy_changes(self):
        return self.actual_energy_changes

    def reset_trust_radius(self):
        self.trust_radius = self.trust_radius_initial
        if self._state is not None:
            self._state[:, ops.RS_TRUST] = self.trust_radius_initial

    # ---- helpers --------------------------------------------------------------------
    def _dev(self, a, shape):
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(shape))
        return t.to(self.device, non_blocking=False)

    def _ensure_state(self, B):
        if self.Initialization or self._state is None or self._state.shape[0] != B:
            self._state = ops.new_rsirfo_state(B, self.trust_radius, self.device)
            self._out = None     # result buffers belong to the batch shape of the state
            self.predicted_energy_changes = []
            self.actual_energy_changes = []
            self.prev_geometry = None
            self.prev_gradient = None
            self.prev_energy = None
            self.converged = False
            self.iteration = 0
            self.Initialization = False

    def run(self, geom_num_list, B_g, pre_B_g=[], pre_geom=[], B_e=0.0, pre_B_e=0.0,
            pre_move_vector=[], initial_geom_num_list=[], g=[], pre_g=[]):
        """One RS-I-RFO step (Optimizer/rsirfo.py:285-490).  Returns minus the step."""
        if self.hessian is None:
            raise ValueError("Hessian matrix must be set before running optimization")
        if isinstance(geom_num_list, torch.Tensor):
            return self._run_tensor(geom_num_list, B_g, pre_geom, B_e, g, pre_g)
        return self._run_numpy(geom_num_list, B_g, pre_geom, B_e, g, pre_g)

    # ---- tensor mode -------------------------------------------------------------------
    def _run_tensor(self, x, Bg, x_prev, Be, g, g_prev):
        if x.dim() == 3:
            x, Bg, g = x.squeeze(-1), Bg.squeeze(-1), g.squeeze(-1)
        B, n = x.shape
        self._ensure_state(B)
        have_hist = (isinstance(x_prev, torch.Tensor) and isinstance(g_prev, torch.Tensor)
                     and x_prev.numel() > 0 and g_prev.numel() > 0)
        if have_hist and x_prev.dim() == 3:
            x_prev, g_prev = x_prev.squeeze(-1), g_prev.squeeze(-1)
        if not isinstance(Be, torch.Tensor):
            Be = torch.full((B,), float(Be), dtype=torch.float64, device=x.device)
        if self._out is not None and tuple(self._out["move"].shape) != (B, n):
            self._out = None
        self._out = ops.rsirfo_step(
            self.hessian, x, Bg, g, self._state, method=self._method_id,
            saddle_order=self.saddle_order, neb_mode=self.NEB_mode, Hbias=self.bias_hessian,
            x_prev=x_prev if have_hist else None, g_prev=g_prev if have_hist else None, Be=Be,
            trust_min=self.trust_radius_min, trust_max=self.trust_radius_max,
            eigh_algo=self.eigh_algo, out=self._out, packed=getattr(self, "_packed", False))
        self.last_status = self._out["status"]
        self.prev_geometry, self.prev_gradient, self.prev_energy = x, Bg, Be
        self.iteration += 1
        return self._out["move"]

    @property
    def state_tensor(self):
        return self._state

    # ---- NumPy mode (reference calling convention, one structure) ---------------------------
    def _run_numpy(self, geom, B_g, pre_geom, B_e, g, pre_g):
        x = np.asarray(geom, dtype=np.float64).reshape(-1)
        n = x.size
        self._ensure_state(1)
        H_host = self.hessian
        if not isinstance(H_host, np.ndarray):
            raise MopError("NumPy-mode run() needs a NumPy Hessian (set_hessian)")
        Hd = self._dev(H_host, (1, n, n))
        Hb = None
        if self.bias_hessian is not None:
            Hb = self._dev(self.bias_hessian, (1, n, n))
        have_hist = (self.prev_geometry is not None and self.prev_gradient is not None
                     and len(pre_g) > 0 and len(pre_geom) > 0)
        out = ops.rsirfo_step(
            Hd, self._dev(x, (1, n)), self._dev(B_g, (1, n)), self._dev(g, (1, n)), self._state,
            method=self._method_id, saddle_order=self.saddle_order, neb_mode=self.NEB_mode, Hbias=Hb,
            x_prev=self._dev(pre_geom, (1, n)) if have_hist else None,
            g_prev=self._dev(pre_g, (1, n)) if have_hist else None,
            Be=torch.tensor([float(B_e)], dtype=torch.float64, device=self.device),
            trust_min=self.trust_radius_min, trust_max=self.trust_radius_max,
            eigh_algo=self.eigh_algo)
        status = int(out["status"].item())
        self.last_status = status
        if status & ops.ST_UPDATED:
            # aliasing of the reference (rsirfo.py:1368): the caller's array
            # carries the updated Hessian into the next iteration
            H_host[...] = Hd[0].cpu().numpy()
        st = self._state[0].cpu().numpy()
        self.trust_radius = float(st[ops.RS_TRUST])
        self.predicted_energy_changes = [float(v) for v in st[ops.RS_PRED0:ops.RS_PRED0 + int(st[ops.RS_NPRED])]]
        self.actual_energy_changes = [float(v) for v in st[ops.RS_ACT0:ops.RS_ACT0 + int(st[ops.RS_NACT])]]
        self.level_shift_applied = bool(status & ops.ST_LEVEL_SHIFT)
        gnorm = float(np.linalg.norm(np.asarray(B_g, dtype=np.float64)))
        if gnorm < self.gradient_norm_threshold:
            self.converged = True
        if self.actual_energy_changes and abs(self.actual_energy_changes[-1]) < self.energy_change_threshold:
            self.converged = True
        self.eigvals = out["eigvals"][0].cpu().numpy()
        self.prev_geometry = geom
        self.prev_gradient = B_g
        self.prev_energy = B_e
        self.iteration += 1
        return out["move"][0].cpu().numpy().reshape(-1, 1)
