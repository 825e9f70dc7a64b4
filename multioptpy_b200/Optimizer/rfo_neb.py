"""Batched, device-resident form of ``multioptpy.Optimizer.rfo_neb.RFOOptimizer`` for the NEB
quasi-Newton step (Optimizer/rfo_neb.py:86-208), FIRE blend excluded.

Per NEB iteration and for this rank's contiguous block of images: halo exchange (NCCL),
BNEB tangents, Ayala curvature update of the per-image Hessians, one RS-I-RFO step per image
(`rsirfo_block_fsb`, trust 0.5 at the chain ends; `rsirfo_block_bofill`, trust 0.2 inside, all
with saddle_order 0 and B_e = pre_B_e = 0 as the reference passes), step limits and the
neighbour-distance trust radius.  The Hessians stay in HBM between iterations (the reference
round-trips them through tmp_hessian_<i>.npy, rfo_neb.py:18-25,175).
"""
from __future__ import annotations

import torch

from .. import ops
from ..neb_halo import exchange_halo, image_partition
from .rsirfo import RSIRFO


class RFOOptimizer:
    def __init__(self, nimg, natoms, first=0, nloc=None, device="cuda", fix_init_edge=False, fix_end_edge=False):
        self.nimg, self.natoms, self.n = nimg, natoms, 3 * natoms
        self.first = first
        self.nloc = nimg if nloc is None else nloc
        self.device = torch.device(device)
        self.fix_init_edge, self.fix_end_edge = fix_init_edge, fix_end_edge
        n = self.n
        self.hessian = torch.eye(n, dtype=torch.float64, device=self.device).repeat(self.nloc, 1, 1).contiguous()
        idx = torch.arange(self.first, self.first + self.nloc)
        self.end_mask = (idx == 0) | (idx == nimg - 1)
        self.end_idx = torch.nonzero(self.end_mask).flatten().to(self.device)
        self.mid_idx = torch.nonzero(~self.end_mask).flatten().to(self.device)
        self.opt_end = RSIRFO(method="rsirfo_block_fsb", saddle_order=0, trust_radius=0.5, device=self.device)
        self.opt_mid = RSIRFO(method="rsirfo_block_bofill", saddle_order=0, trust_radius=0.2, device=self.device)
        self.opt_mid.switch_NEB_mode()
        self.prev_x = None
        self.prev_g = None
        self.last = {}

    def set_hessians(self, H):
        self.hessian.copy_(H)

    def _run_group(self, opt, idx, x, g):
        if idx.numel() == 0:
            return None
        H = self.hessian.index_select(0, idx).contiguous()
        opt.set_hessian(H)
        opt.set_bias_hessian(None)
        xs, gs = x.index_select(0, idx).contiguous(), g.index_select(0, idx).contiguous()
        zero = torch.zeros(idx.numel(), dtype=torch.float64, device=self.device)
        if self.prev_x is None:
            mv = opt.run(xs, gs, B_e=zero, g=gs)
        else:
            mv = opt.run(xs, gs, pre_geom=self.prev_x.index_select(0, idx).contiguous(), B_e=zero, g=gs,
                         pre_g=self.prev_g.index_select(0, idx).contiguous())
        self.hessian.index_copy_(0, idx, H)
        return mv

    def rfo_move_vectors(self, x, E, g):
        """x (nloc, n) Bohr, E (nloc,), g (nloc, n) raw gradients of this rank's images.
        Returns the RFO move vectors after TR_calc (rfo_move_vector_list, rfo_neb.py:178-182)."""
        xh, Eh, gh = exchange_halo(x, E, g)
        force, tau = ops.bneb_force(self.nimg, self.first, xh, Eh, g)
        gamma = ops.neb_ayala(self.nimg, self.first, xh, Eh, gh, tau, self.hessian)
        delta = torch.empty_like(x)
        mv = self._run_group(self.opt_end, self.end_idx, x, g)
        if mv is not None:
            delta.index_copy_(0, self.end_idx, mv)
        mv = self._run_group(self.opt_mid, self.mid_idx, x, g)
        if mv is not None:
            delta.index_copy_(0, self.mid_idx, mv)
        ops.neb_limit_tr(self.nimg, self.first, xh, g, delta, self.fix_init_edge, self.fix_end_edge)
        self.prev_x, self.prev_g = x.clone(), g.clone()
        self.last = dict(force=force, tau=tau, gamma=gamma)
        return delta
