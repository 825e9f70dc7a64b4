"""RS-I-RFO steps for a batch that lives in HOST memory (the reference's calling convention: NumPy arrays in,
NumPy arrays out, ``RSIRFO.run`` of Optimizer/rsirfo.py:285-490 for every structure).

The Hessian batch is the only large operand (92.6 MB as packed lower triangles at 1024 x 3N = 150), so the
pipeline is built around its trip over PCIe:

* the batch is cut into chunks of whole waves of the fused front-end / tridiagonalisation kernel (2 CTAs x 148
  SMs = 296 structures), the short remainder chunk first so that compute starts early;
* the vectors of the whole batch (geometry, gradients, state: 2.5 MB) go first in one copy each, then chunk c of
  the Hessians travels on stream c mod S (pinned host memory, ``cudaMemcpyAsync``) and is reduced by
  ``mop_rsirfo_step_packed_begin`` on the same stream while chunk c + 1 is still in flight;
* ``mop_rsirfo_step_packed_finish`` then runs the spectrum / step kernel ONCE for the whole batch (it hides its
  dependent chains behind seven resident CTAs per SM and cannot share an SM with the reduction, whose two CTAs
  take 218 KB of shared memory: per-chunk spectrum launches ran at two CTAs per SM and cost 0.8 ms per step);
* the steps and status words go back in one copy; the updated Hessians stay on the device (``hessians()`` reads them
  back on demand, ``upload_hessians=False`` steps on the resident copy).

Nothing here computes on the CPU.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from .ops import MopError, _ptr


def pack_lower_host(H: np.ndarray) -> np.ndarray:
    """(B, n, n) symmetric -> (B, n (n + 1) / 2) packed lower triangles, row i at i (i + 1) / 2 (host side)."""
    n = H.shape[-1]
    il = np.tril_indices(n)
    return np.ascontiguousarray(H[:, il[0], il[1]])


class HostStepPipeline:
    WAVE = 296   # structures per full wave of k_tridiag_blk at 3N = 150 (2 CTAs per SM)

    def __init__(self, B: int, n: int, method: int, device="cuda:0", saddle_order: int = 0, neb_mode: bool = False,
                 trust_min: float = 0.01, trust_max: float = 0.5, chunks=None, nstream: int = 4):
        self.B, self.n, self.method = int(B), int(n), int(method)
        self.saddle_order, self.neb_mode = int(saddle_order), bool(neb_mode)
        self.trust_min, self.trust_max = float(trust_min), float(trust_max)
        self.dev = torch.device(device)
        if chunks is None:
            chunks = [B % self.WAVE] * (1 if B % self.WAVE else 0) + [self.WAVE] * (B // self.WAVE)
        if sum(chunks) != B or any(c <= 0 for c in chunks):
            raise MopError(f"chunks {chunks} do not partition a batch of {B}")
        self.bounds = np.concatenate([[0], np.cumsum(chunks)]).astype(int)
        ns = max(1, min(len(chunks), nstream))
        # earlier chunks get the higher priority: their CTAs are scheduled first
        self.streams = [torch.cuda.Stream(self.dev, priority=-min(5, ns - 1 - i)) for i in range(ns)]
        self.events = [torch.cuda.Event() for _ in chunks]
        f64, ntri = torch.float64, n * (n + 1) // 2
        d = lambda *s, dt=f64: torch.empty(*s, dtype=dt, device=self.dev)
        self.dH = d(B, ntri)
        self.dx, self.dg, self.dBg, self.dxp, self.dgp = d(B, n), d(B, n), d(B, n), d(B, n), d(B, n)
        self.dBe, self.dstate = d(B), d(B, ops.RSIRFO_STATE)
        self.move, self.eigvals, self.pred = d(B, n), d(B, n), d(B)
        self.status = d(B, dt=torch.int32)
        lib = _lib.load()
        self.nbytes = lib.mop_rsirfo_workspace_bytes(B, n, ops.EIGH_TRIDIAG)
        self.work = torch.empty(self.nbytes, dtype=torch.uint8, device=self.dev)
        self.h2d_bytes = 0

    @staticmethod
    def _pinned(name, t, shape, dtype=torch.float64):
        if not isinstance(t, torch.Tensor) or t.is_cuda or not t.is_pinned():
            raise MopError(f"{name}: expected a pinned host tensor")
        if t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise MopError(f"{name}: expected contiguous {dtype} of shape {tuple(shape)}")
        return t

    def step(self, hx, hBg, hg, hstate, h_move, h_status, hH=None, hx_prev=None, hg_prev=None, hBe=None,
             state_back=False):
        """One step.  All arguments are PINNED host tensors: hx, hBg, hg (B, n); hstate (B, 16) optimizer state
        (read; written back when state_back); hH (B, n (n + 1) / 2) packed Hessians or None (resident copy);
        hx_prev / hg_prev (B, n) or None (no update); hBe (B,) or None.  h_move (B, n) and h_status (B,) int32
        receive the results; the call returns when they are complete."""
        B, n, lib = self.B, self.n, _lib.load()
        ntri = n * (n + 1) // 2
        P = self._pinned
        P("hx", hx, (B, n)); P("hBg", hBg, (B, n)); P("hg", hg, (B, n)); P("hstate", hstate, (B, ops.RSIRFO_STATE))
        P("h_move", h_move, (B, n)); P("h_status", h_status, (B,), torch.int32)
        if hH is not None:
            P("hH", hH, (B, ntri))
        if (hx_prev is None) != (hg_prev is None):
            raise MopError("hx_prev and hg_prev must be given together")
        if hx_prev is not None:
            P("hx_prev", hx_prev, (B, n)); P("hg_prev", hg_prev, (B, n))
        if hBe is not None:
            P("hBe", hBe, (B,))
        same_g = hBg is hg or hBg.data_ptr() == hg.data_ptr()
        main = torch.cuda.current_stream(self.dev)
        start = torch.cuda.Event(); start.record(main)
        nb = 0
        with torch.cuda.device(self.dev):
            # the vectors of the WHOLE batch first (2.5 MB, seven copies instead of seven per chunk), then the Hessian
            # chunks back to back on the copy engine
            s0 = self.streams[0]
            s0.wait_event(start)
            with torch.cuda.stream(s0):
                pairs = [(self.dx, hx), (self.dg, hg), (self.dstate, hstate)]
                if not same_g:
                    pairs.append((self.dBg, hBg))
                if hx_prev is not None:
                    pairs += [(self.dxp, hx_prev), (self.dgp, hg_prev)]
                if hBe is not None:
                    pairs.append((self.dBe, hBe))
                for dst, src in pairs:
                    dst.copy_(src, non_blocking=True)
                    nb += src.numel() * src.element_size()
                vec_ready = torch.cuda.Event(); vec_ready.record(s0)
            dBg = self.dg if same_g else self.dBg
            for c in range(len(self.bounds) - 1):
                s = self.streams[c % len(self.streams)]
                lo, hi = int(self.bounds[c]), int(self.bounds[c + 1]); sl = slice(lo, hi)
                s.wait_event(vec_ready)
                with torch.cuda.stream(s):
                    if hH is not None:
                        self.dH[sl].copy_(hH[sl], non_blocking=True)
                        nb += hH[sl].numel() * 8
                    rc = lib.mop_rsirfo_step_packed_begin(
                        B, lo, hi - lo, n, self.method, _ptr(self.dH), None, _ptr(self.dx), _ptr(dBg), _ptr(self.dg),
                        _ptr(self.dxp) if hx_prev is not None else None, _ptr(self.dgp) if hx_prev is not None else None,
                        _ptr(self.dstate), _ptr(self.status), _ptr(self.work), self.nbytes, s.cuda_stream)
                    _lib.check(rc, "mop_rsirfo_step_packed_begin")
                    self.events[c].record(s)
            for ev in self.events:
                main.wait_event(ev)
            rc = lib.mop_rsirfo_step_packed_finish(
                B, n, self.saddle_order, int(self.neb_mode), self.trust_min, self.trust_max, _ptr(self.dH), None,
                _ptr(self.dx), _ptr(dBg), _ptr(self.dBe) if hBe is not None else None, _ptr(self.dstate), _ptr(self.move),
                _ptr(self.eigvals), _ptr(self.pred), _ptr(self.status), _ptr(self.work), self.nbytes, main.cuda_stream)
            _lib.check(rc, "mop_rsirfo_step_packed_finish")
            h_move.copy_(self.move, non_blocking=True)
            h_status.copy_(self.status, non_blocking=True)
            if state_back:
                hstate.copy_(self.dstate, non_blocking=True)
            main.synchronize()
        self.h2d_bytes = nb
        return h_move, h_status

    def hessians(self) -> torch.Tensor:
        """The (updated) Hessians as full squares (B, n, n) on the device (lazy read-back)."""
        return ops.unpack_lower(self.dH, self.n)
