"""Drop-in for the step post-processing helpers of ``multioptpy.Utils.calc_tools.Calculationtools`` and
``multioptpy.optimization.ConvergenceChecker`` that sit either side of the optimizer step:
``kabsch_algorithm`` (calc_tools.py:412-425; mutates its arguments in place, as the reference does) and
``check_convergence`` (optimization.py:1244-1289)."""
from __future__ import annotations

import numpy as np
import torch

from .. import ops


class Calculationtools:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)

    def kabsch_algorithm(self, P, Q):
        if isinstance(P, torch.Tensor):
            Pa, Qc, _ = ops.kabsch(P.contiguous(), Q.contiguous())
            P.copy_(P - P.mean(dim=-2, keepdim=True)); Q.copy_(Qc)      # the reference centres both in place
            return Pa, Q
        t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))[None]).to(self.device)
        Pa, Qc, _ = ops.kabsch(t(P), t(Q))
        P -= P.mean(axis=0)
        Q[...] = Qc[0].cpu().numpy()
        return Pa[0].cpu().numpy(), Q


class ConvergenceChecker:
    """config needs MAX_FORCE_THRESHOLD, RMS_FORCE_THRESHOLD, MAX_DISPLACEMENT_THRESHOLD, RMS_DISPLACEMENT_THRESHOLD."""

    def __init__(self, config, device="cuda"):
        self.config = config
        self.device = torch.device(device)

    def check_convergence(self, state, displacement_vector, optimizer_instances=()):
        c = self.config
        g = state.effective_gradient
        if isinstance(g, torch.Tensor):
            B = g.shape[0]
            conv, out = ops.check_convergence(g.reshape(B, -1).contiguous(), displacement_vector.reshape(B, -1).contiguous(),
                                              c.MAX_FORCE_THRESHOLD, c.RMS_FORCE_THRESHOLD, c.MAX_DISPLACEMENT_THRESHOLD,
                                              c.RMS_DISPLACEMENT_THRESHOLD)
            return conv, out[:, 1], out[:, 2]
        t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(1, -1)).to(self.device)
        conv, out = ops.check_convergence(t(g), t(displacement_vector), c.MAX_FORCE_THRESHOLD, c.RMS_FORCE_THRESHOLD,
                                          c.MAX_DISPLACEMENT_THRESHOLD, c.RMS_DISPLACEMENT_THRESHOLD)
        o = out[0].cpu().numpy()
        ok = bool(int(conv[0]))
        if not ok:
            ok = any(getattr(opt, "proj_grad_converged", False) for opt in optimizer_instances)
        return ok, float(o[1]), float(o[2])
