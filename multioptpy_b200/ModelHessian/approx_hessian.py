"""Drop-in for ``multioptpy.ModelHessian.approx_hessian.ApproxHessian``: the same substring dispatch order
(ModelHessian/approx_hessian.py:34-93) and the same modifier pass (:95-110, "ts" then "sr" then "clip", each a substring
test on the whole type string).  Models outside the hot-path scope (D4 / GFN variants, Schlegel, Morse, the
Lindh-2007 and Swart dispersion variants, the "sr" short-range correction) raise instead of silently falling back."""
from __future__ import annotations

import numpy as np

from .. import ops
from .._lib import MopError
from .fischer import FischerApproxHessian

_UNSUPPORTED = ["gfnff", "gfn0xtb", "fischerd4", "schlegeld3", "schlegeld4",
                "schlegel", "swartd3", "swartd4", "swartd2", "lindh2007d3", "lindh2007d4", "lindh2007"]


class ApproxHessian:
    def __init__(self, device="cuda"):
        self.device = device

    def _base(self, coord, element_list, cart_gradient, t):
        if "fischerd3old" in t:     # (before "fischerd3", as in approx_hessian.py:40-45)
            from .fischerd3old import FischerD3ApproxHessianOld
            return FischerD3ApproxHessianOld(device=self.device).main(coord, element_list, cart_gradient)
        if "fischerd3" in t:
            from .fischerd3old import FischerD3ApproxHessian
            return FischerD3ApproxHessian(device=self.device).main(coord, element_list, cart_gradient)
        for key in _UNSUPPORTED:
            if key in t:
                raise MopError(f"model Hessian '{key}' is outside the B200 hot-path scope")
        if "swart" in t:
            from .swart import SwartApproxHessian
            return SwartApproxHessian(device=self.device).main(coord, element_list, cart_gradient)
        if "lindh" in t:
            from .lindh import LindhApproxHessian
            return LindhApproxHessian(device=self.device).main(coord, element_list, cart_gradient)
        if "fischer" in t:
            return FischerApproxHessian(device=self.device).main(coord, element_list, cart_gradient)
        raise MopError(f"model Hessian type '{t}' is not implemented on the device")

    def main(self, coord, element_list, cart_gradient, approx_hess_type="lindh2007d3"):
        import torch
        t = approx_hess_type.lower()
        if "sr" in t:
            raise MopError("model Hessian modifier 'sr' (ModelHessian/shortrange.py) is not implemented on the device")
        H = self._base(coord, element_list, cart_gradient, t)
        if "ts" not in t and "clip" not in t:
            return H
        as_numpy = not isinstance(H, torch.Tensor)
        Hd = torch.from_numpy(np.ascontiguousarray(H)[None]).to(self.device if self.device != "cuda" else "cuda:0") if as_numpy else H
        if "ts" in t:       # approx_hessian.py:95-98
            Hd, _ = ops.hessian_ts_modify(Hd.contiguous())
        if "clip" in t:     # approx_hessian.py:103-110
            Hd = ops.hessian_clip_eigvals(Hd.contiguous())
        return Hd[0].cpu().numpy() if as_numpy else Hd
