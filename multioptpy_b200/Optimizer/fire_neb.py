"""Drop-in for ``multioptpy.Optimizer.fire_neb.FIREOptimizer`` (Optimizer/fire_neb.py:14-92) on the CUDA
kernels.  ``optimize`` keeps the reference signature for a whole chain held on one GPU (NumPy in, new geometry in
Angstrom out); ``step`` is the sharded form: this rank's images with their halo, the power sum all-reduced over
the ranks, the move vector returned as a tensor."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .. import ops
from ..Parameters.tables import BOHR2ANG


class FIREOptimizer:
    def __init__(self, config, device="cuda"):
        self.config = config
        self.dt = config.dt
        self.a = config.a
        self.n_reset = config.n_reset
        self.FIRE_N_accelerate = config.FIRE_N_accelerate
        self.FIRE_f_inc = config.FIRE_f_inc
        self.FIRE_f_accelerate = config.FIRE_f_accelerate
        self.FIRE_f_decelerate = config.FIRE_f_decelerate
        self.FIRE_a_start = config.FIRE_a_start
        self.FIRE_dt_max = config.FIRE_dt_max
        self.fix_init_edge = getattr(config, "fix_init_edge", False)
        self.fix_end_edge = getattr(config, "fix_end_edge", False)
        self.device = torch.device(device)

    def step(self, nimg, first, x_halo, force, velocity, prev_velocity, optimize_num, group=None):
        """x_halo (nloc+2, natoms*3) from neb_halo.exchange_halo; force / velocity / prev_velocity (nloc, natoms, 3)
        tensors (prev_velocity None on the first iteration) -> (move (nloc, natoms*3), new velocity)."""
        nloc, natoms, _ = force.shape
        have_prev = prev_velocity is not None and optimize_num != 0
        power = torch.zeros(1, dtype=torch.float64, device=force.device)
        vneb = ops.neb_fire_blend(force, velocity, prev_velocity if have_prev else None, self.a, power)
        if have_prev and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(power, group=group)
        p = float(power.item()) if have_prev else 0.0
        if optimize_num > 0 and p > 0 and have_prev:
            if self.n_reset > self.FIRE_N_accelerate:
                self.dt = min(self.dt * self.FIRE_f_inc, self.FIRE_dt_max)
                self.a *= self.FIRE_f_inc
            self.n_reset += 1
            reset = False
        else:
            self.a = self.FIRE_a_start
            self.dt *= self.FIRE_f_decelerate
            self.n_reset = 0
            reset = True
        vnew, delta = ops.neb_fire_advance(vneb, force, prev_velocity if have_prev else None, self.dt, reset)
        move = delta.reshape(nloc, natoms * 3).clone()
        ops.neb_limit_tr(nimg, first, x_halo, force.reshape(nloc, natoms * 3).contiguous(), move, self.fix_init_edge,
                         self.fix_end_edge, step_limit=False)
        return move, vnew

    def optimize(self, geometry_num_list, total_force_list, pre_total_velocity, optimize_num, total_velocity,
                 cos_list=None, biased_energy_list=None, pre_biased_energy_list=None, pre_geom=None):
        X = np.asarray(geometry_num_list, dtype=np.float64)
        nimg, natoms, _ = X.shape
        t = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(self.device)
        xh = torch.zeros(nimg + 2, natoms * 3, dtype=torch.float64, device=self.device)
        xh[1:-1] = t(X.reshape(nimg, -1))
        prev = t(pre_total_velocity) if (pre_total_velocity is not None and len(pre_total_velocity) > 1) else None
        move, vnew = self.step(nimg, 0, xh, t(total_force_list), t(total_velocity), prev, optimize_num)
        self.total_velocity = vnew.cpu().numpy()
        return (X + move.reshape(nimg, natoms, 3).cpu().numpy()) * BOHR2ANG
