"""Drop-ins for ``multioptpy.Optimizer.hessian_update.ModelHessianUpdate`` and
``multioptpy.Optimizer.block_hessian_update.BlockHessianUpdate``.

Operator contract of the reference (Optimizer/hessian_update.py:248-433,
Optimizer/block_hessian_update.py:443-709):
``f(hess (n,n), displacement (n,1), delta_grad (n,1)) -> delta_hess (n,n)``.
Here every argument may also carry a leading batch dimension as a float64 CUDA
tensor (``(B,n,n)``, ``(B,n)``), and the result is then a ``(B,n,n)`` tensor.
The arithmetic runs in the fused rank-k CUDA kernel (csrc/hessian_update.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from .._lib import MopError


def _delta(method_id, hess, displacement, delta_grad, device="cuda"):
    if isinstance(hess, torch.Tensor):
        s = displacement.reshape(hess.shape[0], -1)
        y = delta_grad.reshape(hess.shape[0], -1)
        d, _ = ops.hessian_update(hess, s.contiguous(), y.contiguous(), method_id)
        return d
    H = np.ascontiguousarray(np.asarray(hess, dtype=np.float64))
    n = H.shape[0]
    dev = torch.device(device)
    Hd = torch.from_numpy(H).reshape(1, n, n).to(dev)
    sd = torch.from_numpy(np.ascontiguousarray(np.asarray(displacement, dtype=np.float64)).reshape(1, n)).to(dev)
    yd = torch.from_numpy(np.ascontiguousarray(np.asarray(delta_grad, dtype=np.float64)).reshape(1, n)).to(dev)
    d, _ = ops.hessian_update(Hd, sd, yd, method_id)
    return d[0].cpu().numpy()


class ModelHessianUpdate:
    def __init__(self, device="cuda"):
        self.Initialization = True
        self.denom_threshold = 1e-10
        self.dd_mu1 = 0.2
        self.dd_mu2 = 0.2
        self.device = device

    def flowchart_hessian_update(self, hess, displacement, delta_grad, method="auto"):
        return _delta(1, hess, displacement, delta_grad, self.device)

    def BFGS_hessian_update(self, hess, displacement, delta_grad):
        return _delta(15, hess, displacement, delta_grad, self.device)

    def SR1_hessian_update(self, hess, displacement, delta_grad):
        return _delta(16, hess, displacement, delta_grad, self.device)

    def PSB_hessian_update(self, hess, displacement, delta_grad):
        return _delta(24, hess, displacement, delta_grad, self.device)

    def FSB_hessian_update(self, hess, displacement, delta_grad):
        return _delta(22, hess, displacement, delta_grad, self.device)

    def CFD_FSB_hessian_update(self, hess, displacement, delta_grad):
        return _delta(19, hess, displacement, delta_grad, self.device)

    def Bofill_hessian_update(self, hess, displacement, delta_grad):
        return _delta(23, hess, displacement, delta_grad, self.device)

    def CFD_Bofill_hessian_update(self, hess, displacement, delta_grad):
        return _delta(20, hess, displacement, delta_grad, self.device)

    def MSP_hessian_update(self, hess, displacement, delta_grad):
        return _delta(25, hess, displacement, delta_grad, self.device)

    def BFGS_hessian_update_dd(self, hess, displacement, delta_grad):
        return _delta(14, hess, displacement, delta_grad, self.device)

    def FSB_hessian_update_dd(self, hess, displacement, delta_grad):
        return _delta(21, hess, displacement, delta_grad, self.device)

    def CFD_FSB_hessian_update_dd(self, hess, displacement, delta_grad):
        return _delta(18, hess, displacement, delta_grad, self.device)

    def pCFD_Bofill_hessian_update(self, hess, displacement, delta_grad):
        raise MopError("pCFD_Bofill (O(n^4) null-space perturbation, hessian_update.py:309-343) "
                       "is not implemented on the device")


class BlockHessianUpdate:
    """History depth is always 1 in the reference (push, assemble, pop:
    block_hessian_update.py:447-450), so no history is kept here either."""

    def __init__(self, block_size=4, max_window=8, denom_threshold=1e-12, inv_reg=1e-10, device="cuda"):
        if denom_threshold != 1e-12 or inv_reg != 1e-10:
            raise MopError("BlockHessianUpdate(B200): thresholds are baked into the kernel")
        self.block_size = int(block_size)
        self.max_window = int(max_window)
        self.denom_threshold = denom_threshold
        self.inv_reg = inv_reg
        self.S_list = []
        self.Y_list = []
        self.device = device

    def block_BFGS_hessian_update(self, B, displacement, delta_grad):
        return _delta(8, B, displacement, delta_grad, self.device)

    def block_FSB_hessian_update(self, B, displacement, delta_grad):
        return _delta(11, B, displacement, delta_grad, self.device)

    def block_CFD_FSB_hessian_update(self, B, displacement, delta_grad):
        return _delta(4, B, displacement, delta_grad, self.device)

    def block_Bofill_hessian_update(self, B, displacement, delta_grad):
        return _delta(13, B, displacement, delta_grad, self.device)

    def block_CFD_Bofill_hessian_update(self, B, displacement, delta_grad):
        return _delta(6, B, displacement, delta_grad, self.device)

    def block_FSB_hessian_update_weighted(self, B, displacement, delta_grad):
        return _delta(10, B, displacement, delta_grad, self.device)

    def block_CFD_FSB_hessian_update_weighted(self, B, displacement, delta_grad):
        return _delta(3, B, displacement, delta_grad, self.device)

    def block_Bofill_hessian_update_weighted(self, B, displacement, delta_grad):
        return _delta(12, B, displacement, delta_grad, self.device)

    def block_CFD_Bofill_hessian_update_weighted(self, B, displacement, delta_grad):
        return _delta(5, B, displacement, delta_grad, self.device)

    def block_BFGS_hessian_update_dd(self, B, displacement, delta_grad):
        return _delta(7, B, displacement, delta_grad, self.device)

    def block_FSB_hessian_update_dd(self, B, displacement, delta_grad):
        return _delta(9, B, displacement, delta_grad, self.device)

    def block_CFD_FSB_hessian_update_dd(self, B, displacement, delta_grad):
        return _delta(2, B, displacement, delta_grad, self.device)
