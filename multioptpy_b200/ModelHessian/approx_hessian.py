"""Drop-in for ``multioptpy.ModelHessian.approx_hessian.ApproxHessian``: the same substring dispatch order
(ModelHessian/approx_hessian.py:34-93) and the same modifier pass (:95-110, "ts" then "sr" then "clip", each a substring
test on the whole type string).  Models outside the hot-path scope (D4 / GFN variants, Schlegel, Morse, the
Lindh-2007 and Swart dispersion variants) raise instead of silently falling back."""
from __future__ import annotations

import numpy as np

from .. import ops
from .._lib import MopError
from .fischer import FischerApproxHessian

# Pauling electronegativities of ShortRangeCorrectionHessian.estimate_atomic_charges (ModelHessian/shortrange.py:160-172);
# elements outside the table count 2.0
_PAULING_EN = {'H': 2.20, 'He': 0.00, 'Li': 0.98, 'Be': 1.57, 'B': 2.04, 'C': 2.55, 'N': 3.04, 'O': 3.44, 'F': 3.98, 'Ne': 0.00,
               'Na': 0.93, 'Mg': 1.31, 'Al': 1.61, 'Si': 1.90, 'P': 2.19, 'S': 2.58, 'Cl': 3.16, 'Ar': 0.00, 'K': 0.82,
               'Ca': 1.00, 'Sc': 1.36, 'Ti': 1.54, 'V': 1.63, 'Cr': 1.66, 'Mn': 1.55, 'Fe': 1.83, 'Co': 1.88, 'Ni': 1.91,
               'Cu': 1.90, 'Zn': 1.65, 'Ga': 1.81, 'Ge': 2.01, 'As': 2.18, 'Se': 2.55, 'Br': 2.96, 'Kr': 0.00}


def sr_charges(element_list):
    en = [_PAULING_EN.get(e, 2.0) for e in element_list]
    avg = sum(en) / len(en)
    return np.array([0.2 * (avg - v) for v in en])


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
        H = self._base(coord, element_list, cart_gradient, t)
        if "ts" not in t and "clip" not in t and "sr" not in t:
            return H
        as_numpy = not isinstance(H, torch.Tensor)
        dev = self.device if self.device != "cuda" else "cuda:0"
        Hd = torch.from_numpy(np.ascontiguousarray(H)[None]).to(dev) if as_numpy else H
        if "ts" in t:       # approx_hessian.py:95-98
            Hd, _ = ops.hessian_ts_modify(Hd.contiguous())
        if "sr" in t:       # approx_hessian.py:100-102
            from ..Parameters.tables import covalent_radius
            xd = coord if isinstance(coord, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(coord, float))[None]).to(dev)
            rad = np.array([covalent_radius(e) for e in element_list], dtype=np.float64)
            Hd = ops.hessian_sr_correction(Hd.contiguous(), xd.contiguous(), rad, sr_charges(element_list))
        if "clip" in t:     # approx_hessian.py:103-110
            Hd = ops.hessian_clip_eigvals(Hd.contiguous())
        return Hd[0].cpu().numpy() if as_numpy else Hd
