"""Thin torch-facing wrappers over the C ABI (device memory + streams only).

Every function takes float64 CUDA tensors (batch-major, contiguous) and
launches the hand-written sm_100a kernels asynchronously on the current torch
stream.  Nothing here computes on the CPU and nothing falls back to torch
linear algebra.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import MopError

# ---- ids shared with include/mop_b200.h -----------------------------------
UPDATE_DISPATCH = [
    ("flowchart", 1), ("block_cfd_fsb_dd", 2), ("block_cfd_fsb_weighted", 3), ("block_cfd_fsb", 4),
    ("block_cfd_bofill_weighted", 5), ("block_cfd_bofill", 6), ("block_bfgs_dd", 7), ("block_bfgs", 8),
    ("block_fsb_dd", 9), ("block_fsb_weighted", 10), ("block_fsb", 11), ("block_bofill_weighted", 12),
    ("block_bofill", 13), ("bfgs_dd", 14), ("bfgs", 15), ("sr1", 16), ("pcfd_bofill", 17),
    ("cfd_fsb_dd", 18), ("cfd_fsb", 19), ("cfd_bofill", 20), ("fsb_dd", 21), ("fsb", 22),
    ("bofill", 23), ("psb", 24), ("msp", 25),
]
UPD_NONE, UPD_FLOWCHART = 0, 1
EIGH_AUTO, EIGH_JACOBI, EIGH_TRIDIAG = 0, 1, 2
EIGH_LARGE = 3
EIGH_ALGOS = {"auto": EIGH_AUTO, "jacobi": EIGH_JACOBI, "tridiag": EIGH_TRIDIAG, "large": EIGH_LARGE}
RSIRFO_STATE = 16
RS_TRUST, RS_HAVE_PREV, RS_PREV_ENERGY, RS_HAVE_ENERGY, RS_NPRED, RS_PRED0, RS_NACT, RS_ACT0, RS_ITER = (
    0, 1, 2, 3, 4, 5, 8, 9, 12)

ST_UPDATED = 1 << 0
ST_UPD_SKIP_SMALL = 1 << 1
ST_UPD_SKIP_CURV = 1 << 2
ST_UPD_TERM_ZEROED = 1 << 3
ST_LEVEL_SHIFT = 1 << 4
ST_EIG_NONFINITE = 1 << 5
ST_ALPHA_SEARCH = 1 << 6
ST_STEP_NAN_SD = 1 << 7
ST_HARD_CASE = 1 << 8
ST_TRROT_RANKDEF = 1 << 9
ST_BRENT_BRACKET = 1 << 10
ST_ALPHA_UNSTABLE = 1 << 16
ST_EIG_NOCONV = 1 << 11
ST_EIG_FALLBACK = 1 << 12
ST_NO_HISTORY = 1 << 13
ST_UPD_REJECTED = 1 << 14
ST_LINDH_NO_K = 1 << 15
ST_CONSTR_CONVERGED = 1 << 17


def resolve_update_method(name: str) -> int:
    """Prioritised substring dispatch of Optimizer/rsirfo.py:208-251,1341-1356."""
    low = name.lower()
    for key, mid in UPDATE_DISPATCH:
        if key in low:
            return mid
    return UPD_FLOWCHART


# ---- helpers -----------------------------------------------------------------
def _chk(t: torch.Tensor, name: str, shape=None, dtype=torch.float64) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise MopError(f"{name}: expected a torch tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise MopError(f"{name}: must be a CUDA tensor (the hot path has no CPU fallback)")
    if t.dtype != dtype:
        raise MopError(f"{name}: dtype {t.dtype}, expected {dtype}")
    if not t.is_contiguous():
        raise MopError(f"{name}: must be contiguous")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise MopError(f"{name}: shape {tuple(t.shape)}, expected {tuple(shape)}")
    return t


def _ptr(t):
    return None if t is None else t.data_ptr()


def _chk_out(out, B, n):
    """A caller-supplied result dict is written by the kernels with B * n (move, eigvals) / B (pred, status)
    elements: validate every entry so a stale dict from a smaller batch can never be overrun."""
    _chk(out["move"], "out['move']", (B, n)); _chk(out["eigvals"], "out['eigvals']", (B, n))
    _chk(out["pred"], "out['pred']", (B,)); _chk(out["status"], "out['status']", (B,), torch.int32)


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


_WORK: dict = {}


def workspace(dev: torch.device, nbytes: int) -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream): the library's kernels of one call run on the
    caller's current stream, so calls issued on different streams must not share scratch."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(dev).cuda_stream)
    buf = _WORK.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=dev)
        _WORK[key] = buf
    return buf


# ---- (1) Hessian update --------------------------------------------------------
def hessian_update(H, s, y, method: int, *, inplace: bool = False, rsirfo_guards: bool = False,
                   status=None, multi_cta: bool = True):
    """delta_hess of one quasi-Newton update for each structure of the batch, or the
    in-place symmetrised update (RSIRFO.update_hessian) when ``inplace``."""
    lib = _lib.load()
    B, n, _ = H.shape
    _chk(H, "H", (B, n, n)); _chk(s, "s", (B, n)); _chk(y, "y", (B, n))
    if status is None:
        status = torch.zeros(B, dtype=torch.int32, device=H.device)
    _chk(status, "status", (B,), torch.int32)
    delta = None if inplace else torch.zeros_like(H)   # skipped structures (RSIRFO guards) leave their delta at zero
    nbytes = lib.mop_hessian_update_workspace_bytes(B, n) if multi_cta else 0
    work = workspace(H.device, nbytes) if nbytes else None
    with torch.cuda.device(H.device):
        rc = lib.mop_hessian_update(B, n, int(method), 1 if inplace else 0, int(rsirfo_guards),
                                    _ptr(H), _ptr(s), _ptr(y), _ptr(delta), _ptr(status), _ptr(work), nbytes,
                                    _stream(H.device))
    _lib.check(rc, "mop_hessian_update")
    return (H if inplace else delta), status


# ---- (2a) TR/ROT projection ------------------------------------------------------
def project_trrot(H, x, Hbias=None, g=None, status=None):
    lib = _lib.load()
    B, n = x.shape
    _chk(x, "x", (B, n))
    Hp = gp = None
    if H is not None:
        _chk(H, "H", (B, n, n))
        Hp = torch.empty_like(H)
    if Hbias is not None:
        _chk(Hbias, "Hbias", (B, n, n))
    if g is not None:
        _chk(g, "g", (B, n))
        gp = torch.empty_like(g)
    if status is None:
        status = torch.zeros(B, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.mop_project_trrot(B, n, _ptr(H), _ptr(Hbias), _ptr(x), _ptr(g), _ptr(Hp), _ptr(gp),
                                   _ptr(status), _stream(x.device))
    _lib.check(rc, "mop_project_trrot")
    return Hp, gp, status


# ---- (2b) eigensolver ---------------------------------------------------------------
def eigh(A, algo="auto", status=None):
    """Batched symmetric eigendecomposition.  Returns (evals [B,n] ascending,
    evecs [B,n,n] with ROW k = eigenvector k, status)."""
    lib = _lib.load()
    B, n, _ = A.shape
    _chk(A, "A", (B, n, n))
    algo_id = EIGH_ALGOS[algo] if isinstance(algo, str) else int(algo)
    evals = torch.empty(B, n, dtype=torch.float64, device=A.device)
    evecs = torch.empty_like(A)
    if status is None:
        status = torch.zeros(B, dtype=torch.int32, device=A.device)
    nbytes = lib.mop_eigh_workspace_bytes(B, n, algo_id)
    work = workspace(A.device, nbytes)
    with torch.cuda.device(A.device):
        rc = lib.mop_eigh(B, n, algo_id, _ptr(A), _ptr(evals), _ptr(evecs), _ptr(status), _ptr(work),
                          nbytes, _stream(A.device))
    _lib.check(rc, "mop_eigh")
    return evals, evecs, status


# ---- (2c) RS-I-RFO step ---------------------------------------------------------------
def new_rsirfo_state(B: int, trust0: float, device) -> torch.Tensor:
    st = torch.zeros(B, RSIRFO_STATE, dtype=torch.float64, device=device)
    st[:, RS_TRUST] = trust0
    return st


def pack_lower(H):
    """(B, n, n) symmetric -> (B, n (n + 1) / 2) packed lower triangle (row i at i (i + 1) / 2), on the device."""
    lib = _lib.load()
    B, n, _ = H.shape
    _chk(H, "H", (B, n, n))
    P = torch.empty(B, n * (n + 1) // 2, dtype=torch.float64, device=H.device)
    with torch.cuda.device(H.device):
        _lib.check(lib.mop_pack_lower(B, n, _ptr(H), _ptr(P), _stream(H.device)), "mop_pack_lower")
    return P


def unpack_lower(P, n: int):
    """(B, n (n + 1) / 2) packed lower triangle -> (B, n, n) full symmetric matrix, on the device."""
    lib = _lib.load()
    B = P.shape[0]
    _chk(P, "P", (B, n * (n + 1) // 2))
    H = torch.empty(B, n, n, dtype=torch.float64, device=P.device)
    with torch.cuda.device(P.device):
        _lib.check(lib.mop_unpack_lower(B, n, _ptr(P), _ptr(H), _stream(P.device)), "mop_unpack_lower")
    return H


def rsirfo_step(H, x, Bg, g, state, *, method: int, saddle_order: int = 0, neb_mode: bool = False,
                Hbias=None, x_prev=None, g_prev=None, Be=None, trust_min: float = 0.01,
                trust_max: float = 0.5, eigh_algo="auto", out=None, packed: bool = False):
    """One RSIRFO.run for every structure of the batch.  H is updated in place.
    ``packed``: H (and Hbias) are packed lower triangles (B, n (n + 1) / 2) - half the bytes; n <= 160.
    Returns dict(move, eigvals, pred, status)."""
    lib = _lib.load()
    B, n = x.shape
    dev = x.device
    hshape = (B, n * (n + 1) // 2) if packed else (B, n, n)
    _chk(H, "H", hshape); _chk(x, "x", (B, n)); _chk(Bg, "Bg", (B, n)); _chk(g, "g", (B, n))
    _chk(state, "state", (B, RSIRFO_STATE))
    if Hbias is not None:
        _chk(Hbias, "Hbias", hshape)
    if (x_prev is None) != (g_prev is None):
        raise MopError("x_prev and g_prev must be given together")
    if x_prev is not None:
        _chk(x_prev, "x_prev", (B, n)); _chk(g_prev, "g_prev", (B, n))
    if Be is not None:
        _chk(Be, "Be", (B,))
    algo_id = EIGH_ALGOS[eigh_algo] if isinstance(eigh_algo, str) else int(eigh_algo)
    if out is None:
        out = {
            "move": torch.empty(B, n, dtype=torch.float64, device=dev),
            "eigvals": torch.empty(B, n, dtype=torch.float64, device=dev),
            "pred": torch.empty(B, dtype=torch.float64, device=dev),
            "status": torch.empty(B, dtype=torch.int32, device=dev),
        }
    else:
        _chk_out(out, B, n)
    if isinstance(method, torch.Tensor):   # per-structure update methods (NEB chains): one launch for the batch
        if packed:
            raise MopError("rsirfo_step: per-structure methods and packed storage are not combined")
        _chk(method, "method", (B,), torch.int32)
        nbytes = lib.mop_rsirfo_workspace_bytes(B, n, EIGH_TRIDIAG)
        work = workspace(dev, nbytes)
        with torch.cuda.device(dev):
            rc = lib.mop_rsirfo_step_mixed(B, n, _ptr(method), int(saddle_order), int(bool(neb_mode)),
                                           float(trust_min), float(trust_max), _ptr(H), _ptr(Hbias), _ptr(x),
                                           _ptr(Bg), _ptr(g), _ptr(x_prev), _ptr(g_prev), _ptr(Be), _ptr(state),
                                           _ptr(out["move"]), _ptr(out["eigvals"]), _ptr(out["pred"]),
                                           _ptr(out["status"]), _ptr(work), nbytes, _stream(dev))
        _lib.check(rc, "mop_rsirfo_step_mixed")
        return out
    if packed:
        nbytes = lib.mop_rsirfo_workspace_bytes(B, n, EIGH_TRIDIAG)
        work = workspace(dev, nbytes)
        with torch.cuda.device(dev):
            rc = lib.mop_rsirfo_step_packed(B, n, int(method), int(saddle_order), int(bool(neb_mode)),
                                            float(trust_min), float(trust_max), _ptr(H), _ptr(Hbias), _ptr(x),
                                            _ptr(Bg), _ptr(g), _ptr(x_prev), _ptr(g_prev), _ptr(Be), _ptr(state),
                                            _ptr(out["move"]), _ptr(out["eigvals"]), _ptr(out["pred"]),
                                            _ptr(out["status"]), _ptr(work), nbytes, _stream(dev))
        _lib.check(rc, "mop_rsirfo_step_packed")
        return out
    nbytes = lib.mop_rsirfo_workspace_bytes(B, n, algo_id)
    work = workspace(dev, nbytes)
    with torch.cuda.device(dev):
        rc = lib.mop_rsirfo_step(B, n, int(method), int(saddle_order), int(bool(neb_mode)), algo_id,
                                 float(trust_min), float(trust_max), _ptr(H), _ptr(Hbias), _ptr(x),
                                 _ptr(Bg), _ptr(g), _ptr(x_prev), _ptr(g_prev), _ptr(Be), _ptr(state),
                                 _ptr(out["move"]), _ptr(out["eigvals"]), _ptr(out["pred"]),
                                 _ptr(out["status"]), _ptr(work), nbytes, _stream(dev))
    _lib.check(rc, "mop_rsirfo_step")
    return out


def rsirfo_spectral_step(Hp, gp, Bg, state, *, saddle_order: int = 0, neb_mode: bool = False, Be=None,
                         trust_min: float = 0.01, trust_max: float = 0.5, out=None):
    """RSIRFO.run after the projections: fused tridiagonal eigensolve + RFO step (n <= 160)."""
    lib = _lib.load()
    B, n = gp.shape
    dev = gp.device
    _chk(Hp, "Hp", (B, n, n)); _chk(gp, "gp", (B, n)); _chk(Bg, "Bg", (B, n))
    _chk(state, "state", (B, RSIRFO_STATE))
    if Be is not None:
        _chk(Be, "Be", (B,))
    if out is None:
        out = {
            "move": torch.empty(B, n, dtype=torch.float64, device=dev),
            "eigvals": torch.empty(B, n, dtype=torch.float64, device=dev),
            "pred": torch.empty(B, dtype=torch.float64, device=dev),
            "status": torch.zeros(B, dtype=torch.int32, device=dev),
        }
    else:
        _chk_out(out, B, n)
    nbytes = lib.mop_rsirfo_spectral_workspace_bytes(B, n)
    work = workspace(dev, nbytes)
    with torch.cuda.device(dev):
        rc = lib.mop_rsirfo_spectral_step(B, n, int(saddle_order), int(bool(neb_mode)), float(trust_min),
                                          float(trust_max), _ptr(Hp), _ptr(gp), _ptr(Bg), _ptr(Be),
                                          _ptr(state), _ptr(out["move"]), _ptr(out["eigvals"]),
                                          _ptr(out["pred"]), _ptr(out["status"]), _ptr(work), nbytes,
                                          _stream(dev))
    _lib.check(rc, "mop_rsirfo_spectral_step")
    return out


def clamp_and_move(x, move, trust_outer, want_geometry: bool = True):
    """optimizer.py:792-798,812: clamp ||move|| to trust_outer (in place) and return
    the new geometry in Angstrom."""
    lib = _lib.load()
    B, n = move.shape
    _chk(move, "move", (B, n)); _chk(trust_outer, "trust_outer", (B,))
    xnew = None
    if want_geometry:
        _chk(x, "x", (B, n))
        xnew = torch.empty_like(x)
    with torch.cuda.device(move.device):
        rc = lib.mop_clamp_and_move(B, n, _ptr(x), _ptr(move), _ptr(trust_outer), _ptr(xnew),
                                    _stream(move.device))
    _lib.check(rc, "mop_clamp_and_move")
    return xnew, move


# ---- (3) producers -------------------------------------------------------------------------
def _radii_arg(radii, B, N, dev):
    """radii: (N,) shared by the batch or (B, N); returns (tensor, stride)."""
    r = radii if isinstance(radii, torch.Tensor) else torch.as_tensor(radii, dtype=torch.float64)
    r = r.to(dev, torch.float64).contiguous()
    if r.dim() == 1:
        _chk(r, "radii", (N,))
        return r, 0
    _chk(r, "radii", (B, N))
    return r, N


def connectivity(xyz, radii, factor: float = 1.1, caps=None):
    """Bond / angle / dihedral tables of every structure.  xyz: (B, N, 3) Bohr.
    Returns (bonds, angles, dihedrals, counts, status) int32 tensors (padded to the capacity)."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3))
    r, stride = _radii_arg(radii, B, N, xyz.device)
    capB, capA, capD = caps if caps else (min(8 * N, N * (N - 1) // 2 + 1), 28 * N, 64 * N)
    i32 = dict(dtype=torch.int32, device=xyz.device)
    bonds = torch.full((B, capB, 2), -1, **i32); angles = torch.full((B, capA, 3), -1, **i32)
    dihs = torch.full((B, capD, 4), -1, **i32); counts = torch.zeros(B, 3, **i32)
    status = torch.zeros(B, **i32)
    with torch.cuda.device(xyz.device):
        rc = lib.mop_connectivity(B, N, _ptr(xyz), _ptr(r), stride, float(factor), capB, capA, capD,
                                  _ptr(bonds), _ptr(angles), _ptr(dihs), _ptr(counts), _ptr(status),
                                  _stream(xyz.device))
    _lib.check(rc, "mop_connectivity")
    return bonds, angles, dihs, counts, status


def fischer_hessian(xyz, radii):
    """FischerApproxHessian.main for every structure: (B, 3N, 3N) projected model Hessians."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3))
    r, stride = _radii_arg(radii, B, N, xyz.device)
    H = torch.empty(B, 3 * N, 3 * N, dtype=torch.float64, device=xyz.device)
    counts = torch.zeros(B, 3, dtype=torch.int32, device=xyz.device)
    status = torch.zeros(B, dtype=torch.int32, device=xyz.device)
    nbytes = lib.mop_fischer_workspace_bytes(B, N)
    work = workspace(xyz.device, nbytes)
    with torch.cuda.device(xyz.device):
        rc = lib.mop_fischer_hessian(B, N, _ptr(xyz), _ptr(r), stride, _ptr(H), _ptr(counts), _ptr(status),
                                     _ptr(work), nbytes, _stream(xyz.device))
    _lib.check(rc, "mop_fischer_hessian")
    return H, counts, status


def fischer_d3old_hessian(xyz, atom_params, d3=None, dynamic=False):
    """FischerD3ApproxHessianOld.main (dynamic=True: FischerD3ApproxHessian.main) for every structure: (B, 3N, 3N)
    projected model Hessians.  atom_params (N, 4) or (B, N, 4): covalent radius, D2 C6, D3 r4r2, D2 vdW radius
    (ModelHessian.fischerd3old.d3_atom_params; dynamic: 5 columns, + the reference coordination number);
    d3 = (s6, s8, a1, a2), default: the reference's PBE0 values."""
    from .Parameters import tables
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3))
    if not isinstance(atom_params, torch.Tensor):
        atom_params = torch.from_numpy(np.ascontiguousarray(np.asarray(atom_params, dtype=np.float64))).to(xyz.device)
    npar = 5 if dynamic else 4
    if atom_params.dim() == 2:
        _chk(atom_params, "atom_params", (N, npar)); stride = 0
    else:
        _chk(atom_params, "atom_params", (B, N, npar)); stride = N
    s6, s8, a1, a2 = d3 if d3 is not None else (tables.D3_S6, tables.D3_S8, tables.D3_A1, tables.D3_A2)
    H = torch.empty(B, 3 * N, 3 * N, dtype=torch.float64, device=xyz.device)
    counts = torch.zeros(B, 3, dtype=torch.int32, device=xyz.device)
    status = torch.zeros(B, dtype=torch.int32, device=xyz.device)
    nbytes = lib.mop_fischer_workspace_bytes(B, N)
    work = workspace(xyz.device, nbytes)
    with torch.cuda.device(xyz.device):
        fn = lib.mop_fischer_d3_hessian if dynamic else lib.mop_fischer_d3old_hessian
        rc = fn(B, N, _ptr(xyz), _ptr(atom_params), stride, float(s6), float(s8), float(a1), float(a2), _ptr(H),
                _ptr(counts), _ptr(status), _ptr(work), nbytes, _stream(xyz.device))
    _lib.check(rc, "mop_fischer_d3old_hessian")
    return H, counts, status


def fix_atoms_effective_hessian(H, fix_atoms):
    """HessianManager.calc_eff_hess_for_fix_atoms_and_set_hess (optimization.py:1325-1343): in place
    H -= H[:, f] pinv(H[f, f] + 1e-10 I) H[f, :], f = the coordinates of the 1-based atoms in fix_atoms."""
    lib = _lib.load()
    B, n, _ = H.shape
    _chk(H, "H", (B, n, n))
    fix = []
    for a in fix_atoms:
        fix.extend([3 * (a - 1), 3 * (a - 1) + 1, 3 * (a - 1) + 2])
    m = len(fix)
    if m == 0:
        return H
    if max(fix) >= n or min(fix) < 0:
        raise MopError("fix_atoms: atom index out of range")
    fd = torch.tensor(fix, dtype=torch.int32, device=H.device)
    blocks = torch.empty(B, m, m, dtype=torch.float64, device=H.device)
    with torch.cuda.device(H.device):
        _lib.check(lib.mop_fix_atoms_gather(B, n, m, _ptr(fd), _ptr(H), _ptr(blocks), _stream(H.device)), "mop_fix_atoms_gather")
    evals, evecs, st = eigh(blocks, "jacobi" if m <= 64 else "auto")
    with torch.cuda.device(H.device):
        _lib.check(lib.mop_fix_atoms_schur(B, n, m, _ptr(fd), _ptr(evals), _ptr(evecs), _ptr(H), _stream(H.device)),
                   "mop_fix_atoms_schur")
    return H


def hessian_ts_modify(H):
    """TransitionStateHessian.create_ts_hessian (ModelHessian/tshess.py) for a batch -> (H_ts, modified [B])."""
    lib = _lib.load()
    B, n, _ = H.shape
    _chk(H, "H", (B, n, n))
    evals, evecs, st = eigh(H)
    out = torch.empty_like(H)
    mod = torch.zeros(B, dtype=torch.int32, device=H.device)
    with torch.cuda.device(H.device):
        rc = lib.mop_hessian_ts_modify(B, n, _ptr(H), _ptr(evals), _ptr(evecs), _ptr(out), _ptr(mod), _stream(H.device))
    _lib.check(rc, "mop_hessian_ts_modify")
    return out, mod


def hessian_clip_eigvals(H):
    """The "clip" modifier of ApproxHessian.main (approx_hessian.py:103-126): V diag(smooth(lambda)) V^T."""
    lib = _lib.load()
    B, n, _ = H.shape
    _chk(H, "H", (B, n, n))
    evals, evecs, st = eigh(H)
    out = torch.empty_like(H)
    with torch.cuda.device(H.device):
        rc = lib.mop_hessian_clip_eigvals(B, n, _ptr(evals), _ptr(evecs), _ptr(out), _stream(H.device))
    _lib.check(rc, "mop_hessian_clip_eigvals")
    return out


def hessian_sr_correction(H, xyz, radii, charges, omega: float = 0.2, cx_sr: float = 0.78, scaling_factor: float = 0.5):
    """The "sr" modifier of ApproxHessian.main (ModelHessian/shortrange.py): out = sym(H + P C P).  radii (covalent,
    Bohr) and charges (N,) or (B, N)."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3)); _chk(H, "H", (B, 3 * N, 3 * N))
    r, rs = _radii_arg(radii, B, N, xyz.device)
    c, cs = _radii_arg(charges, B, N, xyz.device)
    out = torch.empty_like(H)
    nbytes = lib.mop_hessian_sr_workspace_bytes(B, N)
    work = workspace(xyz.device, nbytes)
    with torch.cuda.device(xyz.device):
        rc = lib.mop_hessian_sr_correction(B, N, _ptr(xyz), _ptr(r), rs, _ptr(c), cs, float(omega), float(cx_sr),
                                           float(scaling_factor), _ptr(H), _ptr(out), _ptr(work), nbytes, _stream(xyz.device))
    _lib.check(rc, "mop_hessian_sr_correction")
    return out


def afir(xyz, frag1, frag2, radii_f32, gamma, want_grad: bool = True, want_hess: bool = True):
    """AFIR energy / gradient / Hessian.  xyz (B, N, 3); frag1/frag2 0-based int32 index tensors;
    radii_f32 (N,) float32 Bohr; gamma (B,) kJ/mol.  Returns (E (B,), grad (B, 3N), H (B, 3N, 3N))."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    dev = xyz.device
    _chk(xyz, "xyz", (B, N, 3)); _chk(gamma, "gamma", (B,))
    _chk(radii_f32, "radii_f32", (N,), torch.float32)
    _chk(frag1, "frag1", None, torch.int32); _chk(frag2, "frag2", None, torch.int32)
    E = torch.empty(B, dtype=torch.float64, device=dev)
    g = torch.empty(B, 3 * N, dtype=torch.float64, device=dev) if want_grad else None
    H = torch.empty(B, 3 * N, 3 * N, dtype=torch.float64, device=dev) if want_hess else None
    with torch.cuda.device(dev):
        rc = lib.mop_afir(B, N, _ptr(xyz), frag1.numel(), _ptr(frag1), frag2.numel(), _ptr(frag2),
                          _ptr(radii_f32), _ptr(gamma), _ptr(E), _ptr(g), _ptr(H), _stream(dev))
    _lib.check(rc, "mop_afir")
    return E, g, H


TR_STATE = 12


def outer_trust_radius(H, Hbias, pre_Bg, pre_move, Be, pre_Be, trust, state, trust_min=0.01, trust_max=0.5):
    """TrustRadius.update_trust_radii for a batch (trust and state updated in place)."""
    lib = _lib.load()
    B, n = pre_move.shape
    _chk(H, "H", (B, n, n)); _chk(pre_Bg, "pre_Bg", (B, n)); _chk(pre_move, "pre_move", (B, n))
    _chk(Be, "Be", (B,)); _chk(pre_Be, "pre_Be", (B,)); _chk(trust, "trust", (B,)); _chk(state, "state", (B, TR_STATE))
    if Hbias is not None:
        _chk(Hbias, "Hbias", (B, n, n))
    with torch.cuda.device(H.device):
        rc = lib.mop_outer_trust_radius(B, n, _ptr(H), _ptr(Hbias), _ptr(pre_Bg), _ptr(pre_move), _ptr(Be),
                                        _ptr(pre_Be), _ptr(trust), _ptr(state), float(trust_min),
                                        float(trust_max), _stream(H.device))
    _lib.check(rc, "mop_outer_trust_radius")
    return trust


# ---- (5) NEB -------------------------------------------------------------------------------
def bneb_force(nimg, first, x_halo, E_halo, g):
    """force, tau for the local images (x_halo (nloc+2, n), E_halo (nloc+2,), g (nloc, n))."""
    lib = _lib.load()
    nloc, n = g.shape
    _chk(x_halo, "x_halo", (nloc + 2, n)); _chk(E_halo, "E_halo", (nloc + 2,)); _chk(g, "g", (nloc, n))
    force = torch.empty_like(g); tau = torch.empty_like(g)
    with torch.cuda.device(g.device):
        rc = lib.mop_bneb_force(int(nimg), int(first), nloc, n, _ptr(x_halo), _ptr(E_halo), _ptr(g), _ptr(force),
                                _ptr(tau), _stream(g.device))
    _lib.check(rc, "mop_bneb_force")
    return force, tau


def neb_ayala(nimg, first, x_halo, E_halo, g_halo, tau, H):
    """H (nloc, n, n) += gamma t t^T in place; returns gamma (nloc,)."""
    lib = _lib.load()
    nloc, n = tau.shape
    _chk(x_halo, "x_halo", (nloc + 2, n)); _chk(E_halo, "E_halo", (nloc + 2,)); _chk(g_halo, "g_halo", (nloc + 2, n))
    _chk(tau, "tau", (nloc, n)); _chk(H, "H", (nloc, n, n))
    gamma = torch.zeros(nloc, dtype=torch.float64, device=tau.device)
    with torch.cuda.device(tau.device):
        rc = lib.mop_neb_ayala(int(nimg), int(first), nloc, n, _ptr(x_halo), _ptr(E_halo), _ptr(g_halo), _ptr(tau),
                               _ptr(H), _ptr(gamma), _stream(tau.device))
    _lib.check(rc, "mop_neb_ayala")
    return gamma


def neb_limit_tr(nimg, first, x_halo, g, delta, fix_init_edge=False, fix_end_edge=False, step_limit=True):
    """_limit_step_size (unless step_limit is False: the FIRE optimizer) + TR_calc on delta (nloc, n), in place."""
    lib = _lib.load()
    nloc, n = delta.shape
    _chk(x_halo, "x_halo", (nloc + 2, n)); _chk(g, "g", (nloc, n)); _chk(delta, "delta", (nloc, n))
    with torch.cuda.device(delta.device):
        rc = lib.mop_neb_limit_tr(int(nimg), int(first), nloc, n, int(bool(fix_init_edge)), int(bool(fix_end_edge)),
                                  int(bool(step_limit)), _ptr(x_halo), _ptr(g), _ptr(delta), _stream(delta.device))
    _lib.check(rc, "mop_neb_limit_tr")
    return delta


def constraint_project(C, H, g, shake=None, svd_threshold: float = 1e-5, want_hessian: bool = True):
    """CRSIRFO's subspace projection in the full space: C (B, k, n) raw constraint rows, H (B, n, n), g (B, n)
    -> (Hp (B, n, n) or None, gp (B, n), rank (B,) int32).  See mop_constraint_project."""
    lib = _lib.load()
    B, k, n = C.shape
    _chk(C, "C", (B, k, n)); _chk(H, "H", (B, n, n)); _chk(g, "g", (B, n))
    if shake is not None:
        _chk(shake, "shake", (B, n))
    dev = g.device
    Hp = torch.empty_like(H) if want_hessian else None
    gp = torch.empty_like(g)
    rank = torch.zeros(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mop_constraint_project(B, n, k, float(svd_threshold), _ptr(C), _ptr(H), _ptr(g), _ptr(shake), _ptr(Hp),
                                        _ptr(gp), _ptr(rank), _stream(dev))
    _lib.check(rc, "mop_constraint_project")
    return Hp, gp, rank


def crsirfo_finalize(gp, Be, state_before, state, out, grad_threshold: float):
    lib = _lib.load()
    B, n = gp.shape
    _chk(gp, "gp", (B, n)); _chk(state_before, "state_before", (B, RSIRFO_STATE)); _chk(state, "state", (B, RSIRFO_STATE))
    _chk_out(out, B, n)
    with torch.cuda.device(gp.device):
        rc = lib.mop_crsirfo_finalize(B, n, float(grad_threshold), _ptr(gp), _ptr(Be), _ptr(state_before), _ptr(state),
                                      _ptr(out["move"]), _ptr(out["pred"]), _ptr(out["status"]), _stream(gp.device))
    _lib.check(rc, "mop_crsirfo_finalize")


def add_inplace(dst, src):
    """dst += src on the device (same shape, float64)."""
    lib = _lib.load()
    _chk(dst, "dst"); _chk(src, "src", tuple(dst.shape))
    with torch.cuda.device(dst.device):
        _lib.check(lib.mop_add_inplace(dst.numel(), _ptr(dst), _ptr(src), _stream(dst.device)), "mop_add_inplace")
    return dst


def neb_redistribute(x_chain, first=0, nloc=None, want_path_length=False):
    """distribute_geometry on the whole chain x_chain (nimg, natoms, 3): images first .. first+nloc-1 at equal arc
    length (new tensor); optionally the running path length (nimg,)."""
    lib = _lib.load()
    nimg, natoms, _ = x_chain.shape
    _chk(x_chain, "x_chain", (nimg, natoms, 3))
    nloc = nimg - first if nloc is None else nloc
    out = torch.empty(nloc, natoms, 3, dtype=torch.float64, device=x_chain.device)
    pl = torch.empty(nimg, dtype=torch.float64, device=x_chain.device) if want_path_length else None
    with torch.cuda.device(x_chain.device):
        rc = lib.mop_neb_redistribute(int(nimg), int(natoms), int(first), int(nloc), _ptr(x_chain), _ptr(out), _ptr(pl),
                                      _stream(x_chain.device))
    _lib.check(rc, "mop_neb_redistribute")
    return (out, pl) if want_path_length else out


def lindh_hessian(xyz, atom_params, want_kdiag: bool = False):
    """Lindh model Hessian without the ill-posed K term: (B, 3N, 3N) projected; atom_params (N, 6) or
    (B, N, 6).  Returns (H, kdiag or None, counts, status)."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3))
    p = atom_params if isinstance(atom_params, torch.Tensor) else torch.as_tensor(atom_params, dtype=torch.float64)
    p = p.to(xyz.device, torch.float64).contiguous()
    stride = 0 if p.dim() == 2 else N
    _chk(p, "atom_params", (N, 6) if p.dim() == 2 else (B, N, 6))
    H = torch.empty(B, 3 * N, 3 * N, dtype=torch.float64, device=xyz.device)
    kd = torch.empty(B, N * (N - 1) // 2, dtype=torch.float64, device=xyz.device) if want_kdiag else None
    counts = torch.zeros(B, 3, dtype=torch.int32, device=xyz.device)
    status = torch.zeros(B, dtype=torch.int32, device=xyz.device)
    nbytes = lib.mop_lindh_workspace_bytes(B, N)
    work = workspace(xyz.device, nbytes)
    with torch.cuda.device(xyz.device):
        rc = lib.mop_lindh_hessian(B, N, _ptr(xyz), _ptr(p), stride, _ptr(H), _ptr(kd), _ptr(counts), _ptr(status),
                                   _ptr(work), nbytes, _stream(xyz.device))
    _lib.check(rc, "mop_lindh_hessian")
    return H, kd, counts, status


PRFO_STATE = 8


def rsprfo_step(H, x, Bg, st, *, method: int, saddle_order: int = 1, Hbias=None, x_prev=None, Bg_prev=None,
                pre_move=None, Be=None, trust_min=0.01, trust_max=0.3, eigh_algo="auto", out=None):
    """One EnhancedRSPRFO.run for a batch.  st: dict(state (B,8), prev_grad, prev_move, ts_vec (B,n))."""
    lib = _lib.load()
    B, n = x.shape
    dev = x.device
    _chk(H, "H", (B, n, n)); _chk(x, "x", (B, n)); _chk(Bg, "Bg", (B, n)); _chk(Be, "Be", (B,))
    _chk(st["state"], "state", (B, PRFO_STATE))
    for k in ("prev_grad", "prev_move", "ts_vec"):
        _chk(st[k], k, (B, n))
    if Hbias is not None:
        _chk(Hbias, "Hbias", (B, n, n))
    if x_prev is not None:
        _chk(x_prev, "x_prev", (B, n)); _chk(Bg_prev, "Bg_prev", (B, n))
    if pre_move is not None:
        _chk(pre_move, "pre_move", (B, n))
    algo_id = EIGH_ALGOS[eigh_algo] if isinstance(eigh_algo, str) else int(eigh_algo)
    if out is None:
        out = {"move": torch.empty(B, n, dtype=torch.float64, device=dev),
               "eigvals": torch.empty(B, n, dtype=torch.float64, device=dev),
               "pred": torch.empty(B, dtype=torch.float64, device=dev),
               "status": torch.empty(B, dtype=torch.int32, device=dev)}
    else:
        _chk_out(out, B, n)
    nbytes = lib.mop_rsprfo_workspace_bytes(B, n, algo_id)
    work = workspace(dev, nbytes)
    with torch.cuda.device(dev):
        rc = lib.mop_rsprfo_step(B, n, int(method), int(saddle_order), algo_id, float(trust_min), float(trust_max),
                                 _ptr(H), _ptr(Hbias), _ptr(x), _ptr(Bg), _ptr(x_prev), _ptr(Bg_prev), _ptr(pre_move),
                                 _ptr(Be), _ptr(st["state"]), _ptr(st["prev_grad"]), _ptr(st["prev_move"]),
                                 _ptr(st["ts_vec"]), _ptr(out["move"]), _ptr(out["eigvals"]), _ptr(out["pred"]),
                                 _ptr(out["status"]), _ptr(work), nbytes, _stream(dev))
    _lib.check(rc, "mop_rsprfo_step")
    return out


def swart_hessian(xyz, radii, want_raw: bool = False):
    """SwartApproxHessian.main for every structure: (B, 3N, 3N) projected model Hessians.
    radii: (N,) or (B, N) Swart-table radii in Bohr.  Returns (H, Hraw or None, status)."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3))
    r, stride = _radii_arg(radii, B, N, xyz.device)
    H = torch.empty(B, 3 * N, 3 * N, dtype=torch.float64, device=xyz.device)
    Hraw = torch.empty_like(H) if want_raw else None
    status = torch.zeros(B, dtype=torch.int32, device=xyz.device)
    nbytes = 0 if want_raw else lib.mop_swart_workspace_bytes(B, N)
    work = workspace(xyz.device, nbytes) if nbytes else None
    with torch.cuda.device(xyz.device):
        rc = lib.mop_swart_hessian(B, N, _ptr(xyz), _ptr(r), stride, _ptr(H), _ptr(Hraw), _ptr(status),
                                   _ptr(work), nbytes, _stream(xyz.device))
    _lib.check(rc, "mop_swart_hessian")
    return H, Hraw, status


# ------------------------------------------------------------------ redundant internal coordinates
def _f64(B, *shape, dev):
    return torch.empty(B, *shape, dtype=torch.float64, device=dev)


def ric_bmatrix(xyz):
    """All-pairs distance B matrix (B, M, 3N), rows in itertools.combinations order."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3))
    out = _f64(B, N * (N - 1) // 2, 3 * N, dev=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(lib.mop_ric_bmatrix(B, N, _ptr(xyz), _ptr(out), _stream(xyz.device)), "mop_ric_bmatrix")
    return out


def ric_partial_rows(xyz, labels):
    """Stretch / bend / torsion Wilson rows.  labels: (nrows, 4) int32, 1-based, 0 = unused."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3)); _chk(labels, "labels", None, torch.int32)
    nrows = labels.shape[0]
    out = _f64(B, nrows, 3 * N, dev=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(lib.mop_ric_partial_rows(B, N, _ptr(xyz), nrows, _ptr(labels), _ptr(out), _stream(xyz.device)),
                   "mop_ric_partial_rows")
    return out


def ric_grad_to_cart(xyz, ric_grad):
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3)); _chk(ric_grad, "ric_grad", (B, N * (N - 1) // 2))
    out = _f64(B, 3 * N, dev=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(lib.mop_ric_grad_to_cart(B, N, _ptr(xyz), _ptr(ric_grad), _ptr(out), _stream(xyz.device)),
                   "mop_ric_grad_to_cart")
    return out


def ric_hess_to_cart(xyz, ric_hess, K=None):
    """B^T H B + K.  ric_hess: (B, M, M) dense or (B, M) diagonal."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    M = N * (N - 1) // 2
    _chk(xyz, "xyz", (B, N, 3))
    diag = ric_hess.dim() == 2
    _chk(ric_hess, "ric_hess", (B, M) if diag else (B, M, M))
    if K is not None:
        _chk(K, "K", (B, 3 * N, 3 * N))
    out = _f64(B, 3 * N, 3 * N, dev=xyz.device)
    nbytes = lib.mop_ric_hess_workspace_bytes(B, N, int(diag))
    work = workspace(xyz.device, nbytes) if nbytes else None
    with torch.cuda.device(xyz.device):
        _lib.check(lib.mop_ric_hess_to_cart(B, N, _ptr(xyz), _ptr(ric_hess), int(diag), _ptr(K), _ptr(out), _ptr(work),
                                            nbytes, _stream(xyz.device)), "mop_ric_hess_to_cart")
    return out


def ric_kmatrix(xyz, bonds, angles, dihedrals, counts, ric_grad):
    """K of RIChess2carthess.  Tables: int32 (capacity, 2/3/4) shared by the batch, or (B, capacity, 2/3/4)."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3))
    per = bonds.dim() == 3
    for t, nm in ((bonds, "bonds"), (angles, "angles"), (dihedrals, "dihedrals"), (counts, "counts")):
        _chk(t, nm, None, torch.int32)
    _chk(ric_grad, "ric_grad", None)
    out = _f64(B, 3 * N, 3 * N, dev=xyz.device)
    with torch.cuda.device(xyz.device):
        rc = lib.mop_ric_kmatrix(B, N, _ptr(xyz), _ptr(bonds), _ptr(angles), _ptr(dihedrals), _ptr(counts),
                                 bonds.shape[-2], angles.shape[-2], dihedrals.shape[-2], int(per), _ptr(ric_grad),
                                 ric_grad.shape[-1], _ptr(out), _stream(xyz.device))
    _lib.check(rc, "mop_ric_kmatrix")
    return out


def ric_pb_int_grad(pB, cart_grad):
    """calc_int_grad_from_pBmat for a batch: pB (B, m, n), cart_grad (B, n) -> (B, m)."""
    lib = _lib.load()
    B, m, n = pB.shape
    _chk(pB, "pB", (B, m, n)); _chk(cart_grad, "cart_grad", (B, n))
    out = _f64(B, m, dev=pB.device)
    status = torch.zeros(B, dtype=torch.int32, device=pB.device)
    nbytes = lib.mop_ric_pb_workspace_bytes(B, n)
    work = workspace(pB.device, nbytes)
    with torch.cuda.device(pB.device):
        _lib.check(lib.mop_ric_pb_int_grad(B, n, m, _ptr(pB), _ptr(cart_grad), _ptr(out), _ptr(status), _ptr(work),
                                           nbytes, _stream(pB.device)), "mop_ric_pb_int_grad")
    return out


def ric_pb_cart_grad(pB, int_grad):
    lib = _lib.load()
    B, m, n = pB.shape
    _chk(pB, "pB", (B, m, n)); _chk(int_grad, "int_grad", (B, m))
    out = _f64(B, n, dev=pB.device)
    with torch.cuda.device(pB.device):
        _lib.check(lib.mop_ric_pb_cart_grad(B, n, m, _ptr(pB), _ptr(int_grad), _ptr(out), _stream(pB.device)),
                   "mop_ric_pb_cart_grad")
    return out


# ------------------------------------------------------------------ step post-processing
def kabsch(P, Q):
    """Calculationtools.kabsch_algorithm for a batch: P, Q (B, N, 3) -> (P aligned, Q centred, status)."""
    lib = _lib.load()
    B, N, _ = P.shape
    _chk(P, "P", (B, N, 3)); _chk(Q, "Q", (B, N, 3))
    Pa, Qc = torch.empty_like(P), torch.empty_like(Q)
    status = torch.zeros(B, dtype=torch.int32, device=P.device)
    with torch.cuda.device(P.device):
        _lib.check(lib.mop_kabsch(B, N, _ptr(P), _ptr(Q), _ptr(Pa), _ptr(Qc), _ptr(status), _stream(P.device)), "mop_kabsch")
    return Pa, Qc, status


def check_convergence(grad, disp, max_force_thr, rms_force_thr, max_disp_thr, rms_disp_thr):
    """ConvergenceChecker.check_convergence for a batch: grad, disp (B, n) -> (converged int32 (B,), out (B, 8))."""
    lib = _lib.load()
    B, n = grad.shape
    _chk(grad, "grad", (B, n)); _chk(disp, "disp", (B, n))
    out = torch.empty(B, 8, dtype=torch.float64, device=grad.device)
    conv = torch.empty(B, dtype=torch.int32, device=grad.device)
    with torch.cuda.device(grad.device):
        rc = lib.mop_check_convergence(B, n, _ptr(grad), _ptr(disp), float(max_force_thr), float(rms_force_thr),
                                       float(max_disp_thr), float(rms_disp_thr), _ptr(out), _ptr(conv), _stream(grad.device))
    _lib.check(rc, "mop_check_convergence")
    return conv, out


def neb_fire_blend(force, velocity, prev_velocity, a, power_accum):
    """FIRE velocity / force blend; force, velocity (nloc, natoms, 3); adds sum v_prev . F to power_accum (1,)."""
    lib = _lib.load()
    nloc, natoms, _ = force.shape
    _chk(force, "force", (nloc, natoms, 3)); _chk(velocity, "velocity", (nloc, natoms, 3)); _chk(power_accum, "power_accum", (1,))
    if prev_velocity is not None:
        _chk(prev_velocity, "prev_velocity", (nloc, natoms, 3))
    out = torch.empty_like(velocity)
    with torch.cuda.device(force.device):
        rc = lib.mop_neb_fire_blend(nloc, natoms, float(a), _ptr(force), _ptr(velocity), _ptr(prev_velocity), _ptr(out),
                                    _ptr(power_accum), _stream(force.device))
    _lib.check(rc, "mop_neb_fire_blend")
    return out


def neb_fire_advance(vneb, force, prev_velocity, dt, reset):
    """(velocity_new, delta) of the FIRE step; arrays (nloc, natoms, 3)."""
    lib = _lib.load()
    nloc = vneb.shape[0]
    n = vneb[0].numel()
    vnew, delta = torch.empty_like(vneb), torch.empty_like(vneb)
    with torch.cuda.device(vneb.device):
        rc = lib.mop_neb_fire_advance(nloc, n, float(dt), int(bool(reset)), _ptr(vneb), _ptr(force), _ptr(prev_velocity),
                                      _ptr(vnew), _ptr(delta), _stream(vneb.device))
    _lib.check(rc, "mop_neb_fire_advance")
    return vnew, delta


# ------------------------------------------------------------------ restraint bias potentials
BIAS_KEEP, BIAS_KEEP_V2, BIAS_KEEP_ANGLE, BIAS_KEEP_DIHEDRAL = 1, 2, 3, 4   # kind 4: p = phi0 in RADIANS
# kind 5: k = eps (Hartree), p = sigma (Bohr); kind 6: k, p = r0 (Angstrom), q[0] = well depth; kind 7: k = wall energy
# (Hartree), q = the four limits (Bohr); kind 8: atoms centre, probe, plane 1, plane 2; p = phi0 in RADIANS
BIAS_LJ_PAIR, BIAS_ANHARMONIC_KEEP, BIAS_WELL, BIAS_KEEP_OOP = 5, 6, 7, 8
BIAS_KEEP_ANGLE_V2, BIAS_KEEP_DIHEDRAL_V2, BIAS_KEEP_OOP_V2 = 9, 10, 11   # fragment centroids: q = fragment sizes
BIAS_WELL_POINT, BIAS_WELL_WALL = 12, 13   # well in the distance to a fixed point (r3) / in |x_axis| (frag2 = [axis])
BIAS_MAXA = 64


def pack_bias_terms(terms, device):
    """terms: list of (kind, frag1 (0-based atoms), frag2, k, p[, q (up to four extra parameters)]) -> uint8 device
    tensor for bias_terms."""
    import numpy as _np
    lib = _lib.load()
    rec = int(lib.mop_bias_term_bytes())
    dt = _np.dtype([("kind", "<i4"), ("n1", "<i4"), ("n2", "<i4"), ("atoms", "<i4", (BIAS_MAXA,)), ("pad", "<i4"), ("k", "<f8"), ("p", "<f8"), ("q", "<f8", (4,)),
                    ("r3", "<f8", (3,)), ("pad2", "<f8")])
    if dt.itemsize != rec:
        raise MopError(f"bias term layout mismatch ({dt.itemsize} != {rec})")
    buf = _np.zeros(len(terms), dtype=dt)
    for i, term in enumerate(terms):
        kind, f1, f2, k, p = term[:5]
        q = list(term[5]) if len(term) > 5 else []
        buf[i]["q"][: min(len(q), 4)] = q[:4]
        if len(term) > 6:                      # kind 12: the point as a seventh entry, or behind the four limits in q
            buf[i]["r3"][:] = term[6]
        elif len(q) >= 7:
            buf[i]["r3"][:] = q[4:7]
        if kind == BIAS_WELL_WALL:            # the axis travels in n2, no second fragment
            f1, f2 = list(f1), []
            buf[i]["kind"], buf[i]["n1"], buf[i]["n2"] = kind, 1, int(term[2][0])
            buf[i]["atoms"][:1] = f1
            buf[i]["k"], buf[i]["p"] = k, p
            continue
        at = list(f1) + list(f2)
        if len(at) > BIAS_MAXA:
            raise MopError("bias term: more than 64 atoms")
        buf[i]["kind"], buf[i]["n1"], buf[i]["n2"] = kind, len(f1), len(f2)
        buf[i]["atoms"][: len(at)] = at
        buf[i]["k"], buf[i]["p"] = k, p
    return torch.from_numpy(buf.view(_np.uint8).reshape(-1).copy()).to(device)


def bias_terms(xyz, packed, nterm, E=None, grad=None, hess=None):
    """Add E / gradient / Hessian of the packed restraint terms to E (B,), grad (B, 3N), hess (B, 3N, 3N)."""
    lib = _lib.load()
    B, N, _ = xyz.shape
    _chk(xyz, "xyz", (B, N, 3))
    dev = xyz.device
    if E is None:
        E = torch.zeros(B, dtype=torch.float64, device=dev)
    if grad is None:
        grad = torch.zeros(B, 3 * N, dtype=torch.float64, device=dev)
    if hess is None:
        hess = torch.zeros(B, 3 * N, 3 * N, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mop_bias_terms(B, N, int(nterm), _ptr(packed), _ptr(xyz), _ptr(E), _ptr(grad), _ptr(hess), _stream(dev))
    _lib.check(rc, "mop_bias_terms")
    return E, grad, hess
