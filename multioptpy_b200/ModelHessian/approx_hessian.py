"""Drop-in for ``multioptpy.ModelHessian.approx_hessian.ApproxHessian``: the same substring
dispatch order (ModelHessian/approx_hessian.py:34-93); models outside the hot-path scope
(D3/D4/GFN variants, Schlegel, Morse, ...) raise instead of silently falling back."""
from __future__ import annotations

from .._lib import MopError
from .fischer import FischerApproxHessian

_UNSUPPORTED = ["gfnff", "gfn0xtb", "fischerd3old", "fischerd3", "fischerd4", "schlegeld3", "schlegeld4",
                "schlegel", "swartd3", "swartd4", "swartd2", "lindh2007d3", "lindh2007d4", "lindh2007"]


class ApproxHessian:
    def __init__(self, device="cuda"):
        self.device = device

    def main(self, coord, element_list, cart_gradient, approx_hess_type="lindh2007d3"):
        t = approx_hess_type.lower()
        for key in _UNSUPPORTED:
            if key in t:
                raise MopError(f"model Hessian '{key}' is outside the B200 hot-path scope")
        for mod in ("ts", "sr", "clip"):
            if mod in t.replace("fischer", "").replace("swart", "").replace("lindh", ""):
                raise MopError(f"model Hessian modifier '{mod}' is not implemented on the device")
        if "swart" in t:
            from .swart import SwartApproxHessian
            return SwartApproxHessian(device=self.device).main(coord, element_list, cart_gradient)
        if "lindh" in t:
            from .lindh import LindhApproxHessian
            return LindhApproxHessian(device=self.device).main(coord, element_list, cart_gradient)
        if "fischer" in t:
            return FischerApproxHessian(device=self.device).main(coord, element_list, cart_gradient)
        raise MopError(f"model Hessian type '{approx_hess_type}' is not implemented on the device")
