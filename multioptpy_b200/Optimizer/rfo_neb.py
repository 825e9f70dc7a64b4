"""Batched, device-resident form of ``multioptpy.Optimizer.rfo_neb.RFOOptimizer`` for the NEB
quasi-Newton step (Optimizer/rfo_neb.py:86-208).

Per NEB iteration and for this rank's contiguous block of images: halo exchange (NCCL),
BNEB tangents, Ayala curvature update of the per-image Hessians, ONE launch of the fused RS-I-RFO
step for all local images (`rsirfo_block_fsb` with trust radius 0.5 at the chain ends,
`rsirfo_block_bofill` with 0.2 inside — per-image method array and per-image state, all with
saddle_order 0 and B_e = pre_B_e = 0 as the reference passes; with saddle_order 0 the NEB-mode
switch of the interior optimizers, rsirfo.py:416-419, never acts), step limits and the
neighbour-distance trust radius.  ``optimize_step`` adds the FIRE move and the RFO / FIRE combine
of rfo_neb.py:186-206.  The Hessians stay in HBM between iterations (the reference round-trips
them through tmp_hessian_<i>.npy, rfo_neb.py:18-25,175).

What cannot overlap the halo: the Ayala update puts gamma t t^T into every interior Hessian BEFORE
the per-image step (rfo_neb.py:150-153), and t and gamma need the neighbour images, so update,
projection and eigensolve all depend on the exchange; only the buffer packing is independent.
"""
from __future__ import annotations

import torch

from .. import ops
from ..neb_halo import exchange_halo


class RFOOptimizer:
    def __init__(self, nimg, natoms, first=0, nloc=None, device="cuda", fix_init_edge=False, fix_end_edge=False,
                 ratio_of_rfo_step=0.5):
        self.nimg, self.natoms, self.n = nimg, natoms, 3 * natoms
        self.first = first
        self.nloc = nimg if nloc is None else nloc
        self.device = torch.device(device)
        self.fix_init_edge, self.fix_end_edge = fix_init_edge, fix_end_edge
        self.ratio_of_rfo_step = ratio_of_rfo_step          # rfo_neb.py:97
        n = self.n
        self.hessian = torch.eye(n, dtype=torch.float64, device=self.device).repeat(self.nloc, 1, 1).contiguous()
        idx = torch.arange(self.first, self.first + self.nloc)
        self.end_mask = ((idx == 0) | (idx == nimg - 1)).to(self.device)
        m_end = ops.resolve_update_method("rsirfo_block_fsb")
        m_mid = ops.resolve_update_method("rsirfo_block_bofill")
        self.method = torch.where(self.end_mask, torch.tensor(m_end), torch.tensor(m_mid)).to(torch.int32).to(self.device)
        # RSIRFO(trust_radius=0.5) at the ends, RSIRFO(trust_radius=0.2) inside (rfo_neb.py:118-120)
        self.state = ops.new_rsirfo_state(self.nloc, 0.2, self.device)
        self.state[:, ops.RS_TRUST] = torch.where(self.end_mask, torch.tensor(0.5, dtype=torch.float64),
                                                  torch.tensor(0.2, dtype=torch.float64)).to(self.device)
        self.zero = torch.zeros(self.nloc, dtype=torch.float64, device=self.device)
        self.prev_x = None
        self.prev_g = None
        self._out = None
        self.last = {}

    def set_hessians(self, H):
        self.hessian.copy_(H)

    def rfo_move_vectors(self, x, E, g):
        """x (nloc, n) Bohr, E (nloc,), g (nloc, n) raw gradients of this rank's images.
        Returns the RFO move vectors after TR_calc (rfo_move_vector_list, rfo_neb.py:178-182)."""
        xh, Eh, gh = exchange_halo(x, E, g)
        force, tau = ops.bneb_force(self.nimg, self.first, xh, Eh, g)
        gamma = ops.neb_ayala(self.nimg, self.first, xh, Eh, gh, tau, self.hessian)
        self._out = ops.rsirfo_step(self.hessian, x, g, g, self.state, method=self.method, saddle_order=0,
                                    x_prev=self.prev_x, g_prev=self.prev_g, Be=self.zero, trust_min=0.01,
                                    trust_max=0.5, out=self._out)
        delta = self._out["move"].clone()
        ops.neb_limit_tr(self.nimg, self.first, xh, g, delta, self.fix_init_edge, self.fix_end_edge)
        self.prev_x, self.prev_g = x.clone(), g.clone()
        self.last = dict(force=force, tau=tau, gamma=gamma, x_halo=xh, status=self._out["status"])
        return delta

    def optimize_step(self, x, E, g, fire, velocity, prev_velocity, optimize_num):
        """RFO move, FIRE move and their combination (rfo_neb.py:104-206) for this rank's images:
        ends  -> -rfo_move;  interior -> (1 - r) fire_move - r rfo_move.  Returns (move (nloc, n) to ADD to the
        geometry in Bohr, new velocity)."""
        rfo = self.rfo_move_vectors(x, E, g)
        force = self.last["force"].reshape(self.nloc, self.natoms, 3).contiguous()
        fire_move, vnew = fire.step(self.nimg, self.first, self.last["x_halo"], force, velocity, prev_velocity,
                                    optimize_num)
        r = self.ratio_of_rfo_step
        move = torch.where(self.end_mask[:, None], -rfo, (1.0 - r) * fire_move - r * rfo)
        return move, vnew

    def align_geometries(self, x, optimize_num, align_distances=0, **other_strategies):
        """NEB._align_geometries (neb.py:649-760) for this rank's images x (nloc, n): every `align_distances`
        iterations the chain is redistributed at equal arc length (one all-gather of the chain + one kernel).  The
        other strategies of the reference's table are not built: a non-zero interval for one of them raises."""
        on = [k for k, v in other_strategies.items() if v]
        if on:
            raise ops.MopError(f"NEB alignment strategies {on} are not implemented on the device (only align_distances)")
        if optimize_num <= 0 or align_distances < 1 or optimize_num % align_distances != 0:
            return x
        from ..Interpolation.linear_interpolation import distribute_geometry_sharded
        xg = x.reshape(self.nloc, self.natoms, 3).contiguous()
        return distribute_geometry_sharded(xg, self.nimg, self.first).reshape(self.nloc, self.n)

    def time_halo(self, x, E, g, reps=10):
        """Device time of the halo exchange alone (ms, median-free mean over reps)."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        exchange_halo(x, E, g)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            exchange_halo(x, E, g)
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
