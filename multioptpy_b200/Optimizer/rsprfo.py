"""Drop-in for ``multioptpy.Optimizer.rsprfo.EnhancedRSPRFO`` (P-RFO saddle search,
Optimizer/rsprfo.py:10-1362) on the CUDA path.  Same constructor keywords and ``run`` signature;
NumPy ``(n,1)`` arrays for one structure or CUDA ``(B,n)`` tensors for a batch.  ``set_hessian``
copies and symmetrises, as the reference does (rsprfo.py:1317-1318)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import MopError

_DEFAULTS = dict(alpha0=1.0, max_micro_cycles=50, alpha_max=1e8, alpha_min=1e-8, alpha_step_max=100.0,
                 micro_cycle_rtol=1e-3, micro_cycle_atol=1e-6, eta_1=0.1, eta_2=0.25, eta_3=0.75, gamma_1=0.25,
                 gamma_2=2.0, step_rejection=True, rejection_threshold=-0.5, max_consecutive_rejections=3,
                 hessian_shift_enabled=True, min_positive_eigval=0.001, min_negative_eigval=-0.001,
                 gradient_scaling_enabled=True, gradient_scaling_threshold=0.001, min_step_scale=0.1,
                 adaptive_trust_enabled=True, gradient_trust_coupling=0.5, adapt_trust_radius=True,
                 mode_following=True, eigvec_following=True, overlap_threshold=0.5, mixing_threshold=0.3, maxstep=None)


class EnhancedRSPRFO:
    def __init__(self, **config):
        for key, default in _DEFAULTS.items():
            if key in config and config[key] != default:
                raise MopError(f"EnhancedRSPRFO(B200): non-default {key}={config[key]!r} is baked into the kernels")
        self.config = config
        self.saddle_order = config.get("saddle_order", 1)
        self.hessian_update_method = config.get("method", "auto")
        self.display_flag = config.get("display_flag", True)
        if self.saddle_order == 0:
            self.trust_radius_initial = config.get("trust_radius", 0.5)
            self.trust_radius_max = config.get("trust_radius_max", 0.5)
        else:
            self.trust_radius_initial = config.get("trust_radius", 0.1)
            self.trust_radius_max = config.get("trust_radius_max", 0.3)
        self.trust_radius = self.trust_radius_initial
        self.trust_radius_min = config.get("trust_radius_min", 0.01)
        self.device = torch.device(config.get("device", "cuda"))
        self.eigh_algo = config.get("eigh_algo", "auto")
        self.Initialization = True
        self.iter = 0
        self.hessian = None
        self.bias_hessian = None
        self.predicted_energy_changes = []
        self.roots = list(range(self.saddle_order))
        self._method_id = ops.resolve_update_method(self.hessian_update_method)
        self._st = None
        self._out = None
        self.last_status = None

    def log(self, message, force=False):
        if self.display_flag or force:
            print(message)

    def set_hessian(self, hessian):
        if isinstance(hessian, torch.Tensor):
            self.hessian = (0.5 * (hessian + hessian.transpose(-1, -2))).contiguous()
        else:
            h = np.asarray(hessian, dtype=np.float64).copy()
            self.hessian = 0.5 * (h + h.T)

    def set_bias_hessian(self, bias_hessian):
        if bias_hessian is None:
            self.bias_hessian = None
        elif isinstance(bias_hessian, torch.Tensor):
            self.bias_hessian = bias_hessian.clone()
        else:
            self.bias_hessian = np.asarray(bias_hessian, dtype=np.float64).copy()

    def get_hessian(self):
        return self.hessian

    def get_bias_hessian(self):
        return self.bias_hessian

    def _ensure(self, B, n, dev):
        if self.Initialization or self._st is None or self._st["state"].shape[0] != B:
            z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=dev)
            self._st = dict(state=z(B, ops.PRFO_STATE), prev_grad=z(B, n), prev_move=z(B, n), ts_vec=z(B, n))
            self._st["state"][:, 0] = self.trust_radius_initial
            self._out = None     # result buffers belong to the batch shape of the state
            self.trust_radius = self.trust_radius_initial
            self.predicted_energy_changes = []
            self.iter = 0
            self.Initialization = False
            return True
        return False

    def run(self, geom_num_list, B_g, pre_B_g=[], pre_geom=[], B_e=0.0, pre_B_e=0.0, pre_move_vector=[],
            initial_geom_num_list=[], g=[], pre_g=[]):
        if self.hessian is None:
            raise ValueError("Hessian matrix must be set before running optimization")
        tensor_mode = isinstance(geom_num_list, torch.Tensor)
        if tensor_mode:
            x = geom_num_list.reshape(geom_num_list.shape[0], -1).contiguous()
            B, n = x.shape
            dev = x.device
            fl = lambda a: a.reshape(B, n).contiguous()
            H, Hb = self.hessian, self.bias_hessian
            Bg = fl(B_g)
            have = lambda a: isinstance(a, torch.Tensor) and a.numel() > 0
            Be = B_e if isinstance(B_e, torch.Tensor) else torch.full((B,), float(B_e), dtype=torch.float64, device=dev)
        else:
            xn = np.asarray(geom_num_list, dtype=np.float64).reshape(-1)
            B, n, dev = 1, xn.size, self.device
            t = lambda a: torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(1, -1)).to(dev)
            fl = t
            x, Bg = t(xn), t(B_g)
            H = torch.as_tensor(self.hessian).reshape(1, n, n).to(dev).contiguous()
            Hb = None if self.bias_hessian is None else torch.as_tensor(self.bias_hessian).reshape(1, n, n).to(dev).contiguous()
            have = lambda a: a is not None and len(a) > 0
            Be = torch.tensor([float(B_e)], dtype=torch.float64, device=dev)
        first = self._ensure(B, n, dev)
        if tensor_mode and self._out is not None and tuple(self._out["move"].shape) != (B, n):
            self._out = None
        hist = (not first) and have(pre_B_g) and have(pre_geom)
        self._out = ops.rsprfo_step(
            H, x, Bg, self._st, method=self._method_id, saddle_order=self.saddle_order, Hbias=Hb,
            x_prev=fl(pre_geom) if hist else None, Bg_prev=fl(pre_B_g) if hist else None,
            pre_move=fl(pre_move_vector) if ((not first) and have(pre_move_vector)) else None, Be=Be,
            trust_min=self.trust_radius_min, trust_max=self.trust_radius_max, eigh_algo=self.eigh_algo,
            out=self._out if tensor_mode else None)
        self.last_status = self._out["status"]
        self.iter += 1
        if tensor_mode:
            return self._out["move"]
        self.hessian = H[0].cpu().numpy()
        self.trust_radius = float(self._st["state"][0, 0].item())
        self.predicted_energy_changes.append(float(self._out["pred"][0].item()))
        self.eigvals = self._out["eigvals"][0].cpu().numpy()
        return self._out["move"][0].cpu().numpy().reshape(-1, 1)

    @property
    def state_tensor(self):
        return None if self._st is None else self._st["state"]
