"""NumPy restatement of the reference optimizer-step hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY — see ``oracle/__init__.py``.  Never imported by the
product path.  Parity status: PINNED against outputs of the unmodified
reference (``tests/golden/*.npz`` produced by ``oracle/gen_golden.py``).

Every function cites the reference file:line it follows (paths relative to
``/root/reference/multioptpy``).  Third-party arithmetic the reference calls
and which is therefore also called here: ``numpy.linalg.eigh`` / ``qr``
(LAPACK; reference pins numpy~=2.2, this image has 2.3) and
``scipy.optimize.brentq`` (reference pins scipy~=1.13, image has 1.18).

All arrays are float64, flat ``(n,)`` vectors and ``(n, n)`` row-major
matrices, ``n = 3 * natoms``; geometry is in Bohr.
"""
from __future__ import annotations

import math

import numpy as np

# --------------------------------------------------------------------------
# Hessian-update method ids — shared with include/mop_b200.h (MOP_UPD_*).
# Order = the prioritised substring list of Optimizer/rsirfo.py:208-251.
# --------------------------------------------------------------------------
UPDATE_DISPATCH = [
    ("flowchart", 1),
    ("block_cfd_fsb_dd", 2),
    ("block_cfd_fsb_weighted", 3),
    ("block_cfd_fsb", 4),
    ("block_cfd_bofill_weighted", 5),
    ("block_cfd_bofill", 6),
    ("block_bfgs_dd", 7),
    ("block_bfgs", 8),
    ("block_fsb_dd", 9),
    ("block_fsb_weighted", 10),
    ("block_fsb", 11),
    ("block_bofill_weighted", 12),
    ("block_bofill", 13),
    ("bfgs_dd", 14),
    ("bfgs", 15),
    ("sr1", 16),
    ("pcfd_bofill", 17),
    ("cfd_fsb_dd", 18),
    ("cfd_fsb", 19),
    ("cfd_bofill", 20),
    ("fsb_dd", 21),
    ("fsb", 22),
    ("bofill", 23),
    ("psb", 24),
    ("msp", 25),
]
UPD_NONE = 0
UPD_FLOWCHART = 1
UPDATE_NAMES = {mid: key for key, mid in UPDATE_DISPATCH}


def resolve_update_method(name: str) -> int:
    """First substring hit in priority order, else the 'auto' flowchart
    (Optimizer/rsirfo.py:1341-1356)."""
    low = name.lower()
    for key, mid in UPDATE_DISPATCH:
        if key in low:
            return mid
    return UPD_FLOWCHART


TAU = 1e-10        # ModelHessianUpdate.denom_threshold (hessian_update.py:25)
TAU_BLOCK = 1e-12  # BlockHessianUpdate.denom_threshold (block_hessian_update.py:24)


def _outer(a, b):
    return np.outer(a, b)


def _bfgs(H, s, y, u):
    """hessian_update.py:35-65.  (Hs) s^T H^T == u u^T."""
    n = s.size
    d = np.zeros((n, n))
    sy = s @ y
    if abs(sy) >= TAU:
        d = d + _outer(y, y) / sy
    shs = s @ u
    if abs(shs) >= TAU:
        d = d - _outer(u, u) / shs
    return d


def _sr1(A, s):
    """hessian_update.py:67-85."""
    den = A @ s
    if abs(den) >= TAU:
        return _outer(A, A) / den
    return np.zeros((s.size, s.size))


def _psb(s, y, u):
    """hessian_update.py:87-104 (always uses r = y - Hs, never the CFD factor)."""
    r = y - u
    ss = s @ s
    if abs(ss) >= TAU:
        return -(r @ s) * _outer(s, s) / ss ** 2 + (_outer(r, s) + _outer(s, r)) / ss
    return np.zeros((s.size, s.size))


def _phi2(A, s):
    """hessian_update.py:106-130."""
    num = (A @ s) * (A @ s)
    den = (A @ A) * (s @ s)
    return num / den if abs(den) >= TAU else 0.0


def _dd(s, y, thr):
    """Powell damping with B = I ("double damping step 2"),
    hessian_update.py:200-242 / block_hessian_update.py:565-595."""
    sy = s @ y
    ss = s @ s
    if sy < 0.2 * ss:
        den = ss - sy
        theta = 0.1 if abs(den) < thr else 0.8 * ss / den
        theta = min(1.0, max(0.0, theta))
        y = theta * y + (1.0 - theta) * s
    return y


def _symm(A):
    return 0.5 * (A + A.T)


def _inv1(x, reg=1e-10):
    """1x1 numpy.linalg.inv with the safe_inv fallback
    (block_hessian_update.py:12-21): singular only when exactly zero."""
    if x == 0.0:
        return 1.0 / (x + reg)
    return 1.0 / x


def _blk_bfgs(B, s, y, u, curvature_guard=True):
    """block_hessian_update.py:75-118 with q = 1 (history is popped every call,
    :447-450).  Returns the updated matrix."""
    if not (np.linalg.norm(s) > 1e-8):
        return B.copy()
    if curvature_guard and (y @ s) <= TAU_BLOCK:
        return B.copy()
    t1 = _outer(u, u) * _inv1(s @ u)
    t2 = _outer(y, y) * _inv1(s @ y)
    return _symm(B - t1 + t2)


def _blk_sr1(B, s, y, u, c):
    """block_hessian_update.py:159-184 with q = 1."""
    R = c * (y - u)
    return _symm(B + _outer(R, R) * _inv1(s @ R))


def _blk_psb(B, s, y, u, thr=1e-8):
    """block_hessian_update.py:120-157 with q = 1."""
    if not (np.linalg.norm(s) > thr):
        return B.copy()
    ss = float(s @ s)
    if abs(ss) >= thr:
        r = y - u
        return B + (-(r @ s) * _outer(s, s) / ss ** 2 + (_outer(r, s) + _outer(s, r)) / ss)
    return B.copy()


def _blk_weight(s, y, u, cfd):
    """block_hessian_update.py:190-231 with q = 1 -> phi^2 clipped to [0, 1]."""
    A = y - u
    if cfd:
        A = 2.0 * A
    num = (A @ s) ** 2
    den = (A @ A) * (s @ s)
    c = num / den if abs(den) > TAU_BLOCK else 0.0
    if math.isnan(c):
        c = 0.0
    return float(max(0.0, min(1.0, c)))


def hessian_update_delta(method: int, H: np.ndarray, s: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Delta-Hessian of every supported update (the operator contract
    ``f(hess, displacement, delta_grad) -> delta_hess`` of
    Optimizer/hessian_update.py:248-433 and block_hessian_update.py:443-709)."""
    n = s.size
    u = H @ s
    if method == 1:  # flowchart, hessian_update.py:163-194 (z = y - H y, sic)
        z = y - H @ y
        zs_den = np.linalg.norm(s) * np.linalg.norm(z)
        if abs(zs_den) < TAU:
            zs_den += TAU
        zs = (z @ s) / zs_den
        ys_den = np.linalg.norm(s) * np.linalg.norm(y)
        if abs(ys_den) < TAU:
            ys_den += TAU
        ys = (y @ s) / ys_den
        if zs < -0.1:
            return hessian_update_delta(16, H, s, y)
        if ys > 0.1:
            return hessian_update_delta(15, H, s, y)
        return hessian_update_delta(22, H, s, y)
    if method in (14, 15):  # bfgs(_dd)
        if method == 14:
            y = _dd(s, y, TAU)
        return _bfgs(H, s, y, u)
    if method == 16:
        return _sr1(y - u, s)
    if method == 24:
        return _psb(s, y, u)
    if method in (18, 19, 21, 22):  # (cfd_)fsb(_dd)
        if method in (18, 21):
            y = _dd(s, y, TAU)
        c = 2.0 if method in (18, 19) else 1.0
        A = c * (y - u)
        phi = math.sqrt(_phi2(A, s))
        return (1.0 - phi) * _bfgs(H, s, y, u) + phi * _sr1(A, s)
    if method in (20, 23):  # (cfd_)bofill
        c = 2.0 if method == 20 else 1.0
        A = c * (y - u)
        p2 = _phi2(A, s)
        return (1.0 - p2) * _psb(s, y, u) + p2 * _sr1(A, s)
    if method == 25:  # msp, hessian_update.py:345-368
        A = y - u
        den = np.linalg.norm(A) * np.linalg.norm(s)
        arg = 0.0
        if den >= TAU:
            arg = min(1.0, max(-1.0, (s @ A) / den))
        phi = 1.0 - arg ** 2
        return phi * _psb(s, y, u) + (1.0 - phi) * _sr1(A, s)
    # ---- block family (q = 1) -------------------------------------------
    B = H
    if method == 8:
        return _blk_bfgs(B, s, y, u) - B
    if method == 7:  # block_bfgs_dd: rank guard, damping, no curvature guard (:619-641)
        if not (np.linalg.norm(s) > 1e-8):
            return np.zeros((n, n))
        yt = _dd(s, y, TAU_BLOCK)
        return _blk_bfgs(B, s, yt, u, curvature_guard=False) - B
    if method in (11, 9, 4, 2):  # block_fsb / _dd / block_cfd_fsb / _dd
        if method in (9, 2):
            y = _dd(s, y, TAU_BLOCK)
        cfd = method in (4, 2)
        d_sr1 = _blk_sr1(B, s, y, u, 2.0 if cfd else 1.0) - B
        d_bfgs = _blk_bfgs(B, s, y, u) - B
        c = _blk_weight(s, y, u, cfd)
        w = c if cfd else math.sqrt(c)  # CFD-FSB uses c, FSB sqrt(c) (:249-251,:269-271)
        return _symm(B + w * d_sr1 + (1.0 - w) * d_bfgs) - B
    if method in (13, 6):  # block_bofill / block_cfd_bofill
        cfd = method == 6
        d_psb = _blk_psb(B, s, y, u) - B
        d_sr1 = _blk_sr1(B, s, y, u, 2.0 if cfd else 1.0) - B
        w = _blk_weight(s, y, u, cfd)
        return _symm(B + w * d_sr1 + (1.0 - w) * d_psb) - B
    if method in (10, 3, 12, 5):  # "weighted subspace" variants (:319-437)
        cfd = method in (3, 5)
        c = _blk_weight(s, y, u, cfd)
        w = c if (cfd or method in (12, 5)) else math.sqrt(c)
        d_sr1 = _blk_sr1(B, w * s, w * y, w * u, 2.0 if cfd else 1.0) - B
        a = 1.0 - w
        if method in (10, 3):
            d_other = _blk_bfgs(B, a * s, a * y, a * u) - B
        else:
            d_other = _blk_psb(B, a * s, a * y, a * u) - B
        return _symm(B + d_sr1 + d_other) - B
    raise NotImplementedError(f"update method id {method} ({UPDATE_NAMES.get(method)})")


def rsirfo_update_hessian(H, x, g, x_prev, g_prev, method: int):
    """RSIRFO.update_hessian, Optimizer/rsirfo.py:1316-1372.
    Returns (H_new, updated: bool).  H_new is symmetrised only when updated."""
    s = (x - x_prev).ravel()
    y = (g - g_prev).ravel()
    if np.linalg.norm(s) < 1e-10 or np.linalg.norm(y) < 1e-10:
        return H, False
    if (s @ y) <= 0:
        return H, False
    Hn = H + hessian_update_delta(method, H, s, y)
    return 0.5 * (Hn + Hn.T), True


# --------------------------------------------------------------------------
# translation / rotation projection
# --------------------------------------------------------------------------
def trrot_vectors(x):
    """Six un-normalised TR/ROT vectors about the plain mean of the coordinates
    (Utils/calc_tools.py:261-288, Optimizer/rsirfo.py:134-167)."""
    c = x.reshape(-1, 3)
    c = c - _plain_mean(c)
    N = c.shape[0]
    v = np.zeros((6, 3 * N))
    for k in range(3):
        v[k, k::3] = 1.0
    v[3, 1::3] = -c[:, 2]
    v[3, 2::3] = c[:, 1]
    v[4, 0::3] = c[:, 2]
    v[4, 2::3] = -c[:, 0]
    v[5, 0::3] = -c[:, 1]
    v[5, 1::3] = c[:, 0]
    return v


def _plain_mean(c):
    """calc_center: sequential sum then divide (Utils/calc_tools.py:138-145)."""
    acc = np.zeros(3)
    for row in c:
        acc = acc + row
    return acc / c.shape[0]


def gram_schmidt_cgs(vectors, drop=1e-10):
    """Classical Gram-Schmidt with drop threshold (Utils/calc_tools.py:250-259):
    projections use the ORIGINAL vector v, not the running w."""
    basis = []
    for v in vectors:
        w = v.copy()
        for b in basis:
            w = w - (v @ b) * b
        nrm = np.linalg.norm(w)
        if nrm > drop:
            basis.append(w / nrm)
    return np.array(basis)


def project_hessian_trrot(H, x):
    """Calculationtools.project_out_hess_tr_and_rot_for_coord,
    Utils/calc_tools.py:249-304 (dense P^T H P, then symmetrise)."""
    T = gram_schmidt_cgs(trrot_vectors(x))
    n = H.shape[0]
    P = np.eye(n)
    for t in T:
        P = P - np.outer(t, t)
    Hp = P.T @ H @ P
    return (Hp + Hp.T) / 2


def project_grad_trrot(g, x):
    """RSIRFO._project_grad_tr_rot, Optimizer/rsirfo.py:128-190 (reduced QR)."""
    A = trrot_vectors(x).T
    Q, _ = np.linalg.qr(A, mode="reduced")
    return g - Q @ (Q.T @ g)


# --------------------------------------------------------------------------
# eigendecomposition with conditional level shift
# --------------------------------------------------------------------------
def hessian_is_ill_conditioned(lam, thresh=1e8):
    """check_hessian_conditioning, Optimizer/rsirfo.py:492-551 -> bool."""
    if lam.size < 2:
        return False
    nz = lam[np.abs(lam) > 1e-10]
    if nz.size < 2:
        return True
    mx = np.max(np.abs(nz))
    mn = np.min(np.abs(nz))
    if mn < 1e-15:
        return True
    return (mx / mn) > thresh


def eigh_with_shift(H, shift=1e-5):
    """compute_eigendecomposition_with_shift (auto_level_shift=True default),
    Optimizer/rsirfo.py:553-657.  Returns (lam, V, shifted: bool)."""
    lam, V = np.linalg.eigh(H)
    if hessian_is_ill_conditioned(lam):
        lam_s, V = np.linalg.eigh(H + shift * np.eye(H.shape[0]))
        return lam_s - shift, V, True
    return lam, V, False


# --------------------------------------------------------------------------
# RFO secular equation
# --------------------------------------------------------------------------
def _f_secular(lmd, lam_p, g2):
    den = lam_p - lmd
    safe = np.where(np.abs(den) < 1e-30, np.sign(den) * 1e-30, den)
    safe[safe == 0] = 1e-30
    return lmd + np.sum(g2 / safe)


def _fp_secular(lmd, lam_p, g2):
    den = lam_p - lmd
    safe = np.where(np.abs(den) < 1e-30, np.sign(den) * 1e-30, den)
    safe[safe == 0] = 1e-30
    return 1.0 + np.sum(g2 / safe ** 2)


def secular_safeguarded(lam_p, g2, pole, guess):
    """_solve_secular_safeguarded, Optimizer/rsirfo.py:1374-1503."""
    b = pole
    a = guess
    fa = _f_secular(a, lam_p, g2)
    gnorm = math.sqrt(np.sum(g2))
    limit = 10
    while fa > 0 and limit > 0:
        a = a - max(gnorm, abs(a) * 0.1, 1e-8)
        fa = _f_secular(a, lam_p, g2)
        limit -= 1
    if fa > 0:
        return guess
    lk = guess
    if lk <= a or lk >= b:
        lk = (a + b) / 2.0
    tol = 1e-10 * abs(pole) + 1e-12
    for _ in range(250):
        f = _f_secular(lk, lam_p, g2)
        if abs(f) < tol:
            return lk
        fp = _fp_secular(lk, lam_p, g2)
        dn = -f / fp if abs(fp) > 1e-20 else 0.0
        ln = lk + dn
        lb = (a + b) / 2.0
        nxt = ln if (dn != 0.0 and a < ln < b) else lb
        if f > 0:
            b = lk
        else:
            a = lk
        lk = nxt
        if abs(b - a) < tol:
            return (a + b) / 2.0
    return (a + b) / 2.0


def secular_root(lam, gam, alpha):
    """_solve_secular_more_sorensen, Optimizer/rsirfo.py:1505-1575 (the code
    after the first ``return`` at :1575 is only reached on an exception)."""
    lam_p = lam / alpha
    gp = gam / alpha
    g2 = gp ** 2
    pole = None
    gsum = 0.0
    for i in range(lam_p.size):
        gsum += g2[i]
        if pole is None and g2[i] > 1e-20:
            pole = lam_p[i]
    if pole is None:
        return lam_p[0]
    guess = 0.5 * (pole - math.sqrt(max(0.0, pole ** 2 + 4 * gsum)))
    return secular_safeguarded(lam_p, g2, pole, guess)


def solve_rfo(lam, gam, alpha):
    """RSIRFO.solve_rfo, Optimizer/rsirfo.py:1688-1715 -> (step, lambda_aug)."""
    mu = secular_root(lam, gam, alpha)
    den = lam / alpha - mu
    safe = np.where(np.abs(den) < 1e-20, np.sign(den) * 1e-20, den)
    safe[safe == 0] = 1e-20
    return -(gam / alpha) / safe, mu


def step_derivative(alpha, lam, gam, mu):
    """get_step_derivative, Optimizer/rsirfo.py:1250-1313."""
    den = lam - mu * alpha
    small = np.abs(den) < 1e-8
    if np.any(small):
        den = den.copy()
        den[small] = np.sign(den[small]) * np.maximum(1e-8, np.abs(den[small]))
        # (:1275-1277 writes to a temporary: exact zeros stay zero)
    num = gam ** 2
    d3 = den ** 3
    valid = np.abs(d3) > 1e-10
    if not np.any(valid):
        return 1e-8
    terms = np.zeros_like(num)
    terms[valid] = num[valid] / d3[valid]
    big = np.abs(terms) > 1e20
    if np.any(big):
        terms[big] = np.sign(terms[big]) * 1e20
    d = 2.0 * mu * np.sum(terms)
    if not np.isfinite(d) or abs(d) > 1e20:
        d = np.sign(d) * 1e20 if d != 0 else 1e-8
    return d


def alpha_search(lam, gam, trust, alpha0=1.0, alpha_max=1000.0, alpha_step_max=10.0,
                 max_micro=40, step_tol=1e-3):
    """compute_rsprfo_step, Optimizer/rsirfo.py:986-1248.  Because solve_rfo
    scales eigenvalues AND gradient by 1/alpha the step does not depend on
    alpha analytically; the loop is restated iteration for
    iteration because its exit decides which (rounding-different) step is
    returned.  Returns (step, info) with info in {'brent','newton'}."""
    r2 = trust ** 2
    alpha = alpha0
    try:
        s_lo, _ = solve_rfo(lam, gam, 1e-6)
        s_hi, _ = solve_rfo(lam, gam, alpha_max)
        o_lo = np.linalg.norm(s_lo) ** 2 - r2
        o_hi = np.linalg.norm(s_hi) ** 2 - r2
        if o_lo * o_hi < 0:
            from scipy.optimize import brentq

            def obj(a):
                st, _ = solve_rfo(lam, gam, a)
                return st @ st - r2

            a_b = brentq(obj, 1e-6, alpha_max, xtol=1e-6, rtol=1e-6, maxiter=50)
            st, _ = solve_rfo(lam, gam, a_b)
            if abs(np.linalg.norm(st) - trust) < step_tol:
                return st, "brent"
            alpha = a_b
    except Exception:
        alpha = alpha0
    hist = []
    best = None
    best_diff = float("inf")
    a_left = a_right = None
    step = None
    for _mu in range(max_micro):
        step, mu_aug = solve_rfo(lam, gam, alpha)
        nrm = np.linalg.norm(step)
        diff = abs(nrm - trust)
        if diff < best_diff:
            best = step.copy()
            best_diff = diff
        obj = nrm ** 2 - r2
        if obj < 0 and (a_left is None or alpha > a_left):
            a_left = alpha
        elif obj > 0 and (a_right is None or alpha < a_right):
            a_right = alpha
        if abs(obj) < 1e-8 or diff < step_tol:
            return step, "newton"
        hist.append(nrm)
        d = step_derivative(alpha, lam, gam, mu_aug)
        if abs(d) < 1e-10:
            if a_left is not None and a_right is not None:
                a_new = (a_left + a_right) / 2
            elif obj > 0:
                a_new = max(alpha / 2, 1e-6)
            else:
                a_new = min(alpha * 2, alpha_max)
        else:
            a_step = min(alpha_step_max, max(-alpha_step_max, -obj / d))
            a_new = alpha + a_step
            if a_left is not None and a_right is not None:
                a_new = max(min(a_new, a_right * 0.99), a_left * 1.01)
        alpha = min(max(a_new, 1e-6), alpha_max)
        if alpha == alpha_max or alpha == 1e-6:
            return step, "newton"
        if len(hist) >= 3 and abs(hist[-1] - hist[-2]) < 1e-6 and abs(hist[-2] - hist[-3]) < 1e-6:
            return step, "newton"
    # micro-cycles exhausted (:1213-1246)
    if best is not None and abs(np.linalg.norm(best) - trust) < step_tol * 1.1:
        return best, "newton"
    sd = -gam
    nrm = np.linalg.norm(sd)
    sd = sd / nrm * trust if nrm > 1e-10 else np.zeros_like(gam)
    return sd, "newton"


def rs_step(lam, V, g, trust):
    """RSIRFO.get_rs_step, Optimizer/rsirfo.py:924-985 (no-exception path)."""
    gam = V.T @ g
    step0, _ = solve_rfo(lam, gam, 1.0)
    if np.linalg.norm(step0) <= trust:
        return V @ step0, False
    step, _ = alpha_search(lam, gam, trust)
    return V @ step, True


# --------------------------------------------------------------------------
# inner trust radius of RSIRFO (state only)
# --------------------------------------------------------------------------
def adjust_trust_radius(trust, actual, predicted, min_eig, gnorm, saddle_order,
                        trust_min, trust_max):
    """RSIRFO.adjust_trust_radius(+_adaptive), Optimizer/rsirfo.py:660-887."""
    if gnorm < 1e-2:  # adaptive rule (:835)
        if abs(predicted) < 1e-10:
            return trust
        ratio = actual / predicted
        a = abs(min_eig)
        cf = min(2.5, 1.0 / max(a, 0.1)) if a > 1e-6 else 1.5
        if saddle_order > 0 and min_eig < -1e-6:
            cf *= 0.8
        if ratio > 0.75:
            trust = min(trust * min(1.5 * cf, 2.5), trust_max)
        elif ratio > 0.5:
            trust = min(trust * min(1.1 * cf, 1.5), trust_max)
        elif ratio > 0.25:
            if cf > 1.2:
                trust = min(trust * 1.05, trust_max)
        elif ratio > 0.1:
            trust = max(trust * 0.5, trust_min)
        else:
            trust = max(trust * 0.25, trust_min)
        return float(min(max(trust, trust_min), trust_max))
    if abs(predicted) < 1e-10:
        return trust
    ratio = actual / predicted
    if ratio > 0.75:
        trust = min(trust * 1.2, trust_max)
    elif ratio < 0.25:
        trust = max(trust * 0.5, trust_min)
    return trust


class RSIRFOOracle:
    """State + ``run`` of the reference RSIRFO (Optimizer/rsirfo.py:9-490) for
    ONE structure, default configuration as constructed by
    CalculateMoveVector.initialization (optimizer.py:452-453)."""

    def __init__(self, method="rsirfo_bofill", saddle_order=0, trust_radius_max=None,
                 trust_radius_min=0.01, **_):
        self.method_id = resolve_update_method(method)
        self.saddle_order = saddle_order
        default = 0.5 if saddle_order == 0 else 0.1
        self.trust_radius = default                      # (:36-43)
        self.trust_radius_max = default if trust_radius_max is None else trust_radius_max
        self.trust_radius_min = 0.01 if trust_radius_min is None else trust_radius_min
        self.hessian = None
        self.bias_hessian = None
        self.have_prev = False
        self.prev_energy = None
        self.pred = []
        self.act = []
        self.NEB_mode = False
        self.iteration = 0
        self.last = {}

    def set_hessian(self, H):
        self.hessian = H

    def set_bias_hessian(self, H):
        self.bias_hessian = H

    def run(self, x, Bg, g, x_prev=None, g_prev=None, Be=0.0):
        """Returns the reference's return value (= minus the RFO step, (n,))."""
        x = np.asarray(x, float).ravel()
        Bg = np.asarray(Bg, float).ravel()
        g = np.asarray(g, float).ravel()
        info = {"updated": False, "shift1": False, "shift2": False, "alpha_search": False,
                "nan_fallback": False}
        if self.have_prev and x_prev is not None and g_prev is not None and len(x_prev) > 0 and len(g_prev) > 0:
            self.hessian, info["updated"] = rsirfo_update_hessian(
                self.hessian, x, g, np.asarray(x_prev, float).ravel(), np.asarray(g_prev, float).ravel(),
                self.method_id)
        gnorm = np.linalg.norm(Bg)
        gp = project_grad_trrot(Bg, x)
        Hsum = self.hessian + self.bias_hessian if self.bias_hessian is not None else self.hessian
        Hp = project_hessian_trrot(Hsum, x)
        Hp = 0.5 * (Hp + Hp.T)
        lam, V, info["shift1"] = eigh_with_shift(Hp)
        if not (np.all(np.isfinite(lam)) and np.all(np.isfinite(V))):
            lam = np.ones_like(lam)
            V = np.eye(lam.size)
        if self.prev_energy is not None:                   # (:381-398)
            actual = Be - self.prev_energy
            if len(self.act) >= 3:
                self.act.pop(0)
            self.act.append(actual)
            if self.pred:
                self.trust_radius = adjust_trust_radius(
                    self.trust_radius, actual, self.pred[-1], lam[0], gnorm, self.saddle_order,
                    self.trust_radius_min, self.trust_radius_max)
        n = lam.size
        P = np.eye(n)                                       # (:408-421)
        found = 0
        i = 0
        while found < self.saddle_order:
            if abs(lam[i]) > 1e-10:
                f = 1.0 if self.NEB_mode else 2.0
                P = P - f * np.outer(V[:, i], V[:, i])
                found += 1
            i += 1
        Hs = P @ Hp
        Hs = 0.5 * (Hs + Hs.T)
        gs = P @ gp
        lam_s, V_s, info["shift2"] = eigh_with_shift(Hs)
        if not (np.all(np.isfinite(lam_s)) and np.all(np.isfinite(V_s))):
            lam_s = np.ones_like(lam_s)
            V_s = np.eye(lam_s.size)
        keep = ~(np.abs(lam_s) < 1e-6)                      # (:265-283)
        lam_k = lam_s[keep]
        V_k = V_s[:, keep]
        step, info["alpha_search"] = rs_step(lam_k, V_k, gs, self.trust_radius)
        if not np.all(np.isfinite(step)):                   # (:456-462)
            info["nan_fallback"] = True
            step = -gp
            nrm = np.linalg.norm(step)
            if nrm > self.trust_radius:
                step = step * (self.trust_radius / nrm)
        pred = gp @ step + 0.5 * (step @ Hp @ step)         # (:469, :1717-1720)
        if len(self.pred) >= 3:
            self.pred.pop(0)
        self.pred.append(pred)
        self.have_prev = True
        self.prev_energy = Be
        self.iteration += 1
        info.update(eigvals=lam, pred=pred, trust=self.trust_radius, gproj=gp, Hproj=Hp)
        self.last = info
        return -step


# --------------------------------------------------------------------------
# caller side: CalculateMoveVector.calc_move_vector + TrustRadius
# --------------------------------------------------------------------------
BOHR2ANG = 0.52917721067  # Parameters/unit_values.py:26


def clamp_and_move(x, move, trust_outer):
    """optimizer.py:792-798,812 -> (new_geometry[Angstrom], move[Bohr])."""
    nrm = np.linalg.norm(move)
    if nrm > trust_outer:
        move = trust_outer * move / nrm
    return (x - move) * BOHR2ANG, move


class TrustRadiusOracle:
    """Optimizer/trust_radius.py:120-206 (composite outer trust radius)."""

    def __init__(self, min_trust_radius=0.01, max_trust_radius=0.5):
        self.min = min_trust_radius
        self.max = max_trust_radius
        self.ratios = []
        self.changes = []
        self.count = 0

    def _adaptive_factor(self):
        if not self.ratios:
            return 2.0
        rec = self.ratios[-min(5, len(self.ratios)):]
        var = float(np.var(rec)) if len(rec) > 1 else 0.0
        f = 2.0 * math.exp(-var)
        if len(self.changes) >= 2:
            ch = np.abs(self.changes[-min(3, len(self.changes)):])
            if np.all(ch < 0.01) and np.mean(ch) < 0.005:
                f *= 0.8
        return max(1.1, min(f, 3.0))

    def update(self, Be, pre_Be, pre_Bg, pre_move, H, trust):
        if self.count == 0:
            self.count += 1
            return trust
        Ce = float(pre_Bg @ pre_move + 0.5 * (pre_move @ H @ pre_move))
        eps = 1e-8
        if abs(Ce) < eps:
            Ce += np.sign(Ce) * eps
            if abs(Ce) < eps:
                Ce = eps
        r = (pre_Be - Be) / Ce
        self.ratios.append(float(r))
        self.changes.append(float(pre_Be - Be))
        f = self._adaptive_factor()
        if r <= 0.25 or r >= 1.75:
            trust /= f
        elif 0.75 <= r <= 1.25:
            if abs(np.linalg.norm(pre_move) - trust) < eps:
                trust *= f ** 0.5
        self.count += 1
        return float(np.clip(trust, self.min, self.max))


# --------------------------------------------------------------------------
# connectivity tables (integer, bit-exact part of the contract)
# --------------------------------------------------------------------------
def bond_matrix(coord, radii, factor=1.1):
    """BondConnectivity.bond_connect_matrix, Utils/bond_connectivity.py:13-41."""
    n = len(coord)
    m = np.zeros((n, n), dtype=int)
    for i in range(n):
        d = np.linalg.norm(coord - coord[i], axis=1)
        d[i] = 0.0
        thr = (radii + radii[i]) * factor
        thr[i] = -1.0
        m[i] = np.where(d <= thr, 1, 0)
    return m


def connectivity_tables(coord, radii, factor=1.1):
    """bond / angle / dihedral tables, Utils/bond_connectivity.py:43-134."""
    m = bond_matrix(coord, radii, factor)
    n = len(m)
    bonds = [[i, j] for i in range(n) for j in range(n) if i <= j and m[i, j] == 1]
    angles = []
    for i in range(n):
        for j in range(n):
            if m[i][j] == 1:
                for k in range(j + 1, n):
                    if m[i][k] == 1 and m[j][k] == 0:
                        angles.append([j, i, k])
    dihs = []
    for i in range(len(angles)):
        a = angles[i]
        for j in range(i + 1, len(angles)):
            b = angles[j]
            hit = None
            if (a[1] == b[1] and a[2] == b[2]) or (a[1] == b[2] and a[2] == b[1]):
                c = [a[0], a[1], a[2], b[0]]
                if m[c[2]][c[3]] == 1:
                    hit = c
                else:
                    c = [b[0], a[0], a[1], a[2]]
                    if m[c[1]][c[0]] == 1:
                        hit = c
            if hit is None and ((a[1] == b[1] and a[0] == b[0]) or (a[1] == b[0] and a[0] == b[1])):
                c = [b[2], a[0], a[1], a[2]]
                if m[c[1]][c[0]] == 1:
                    hit = c
                else:
                    c = [a[0], a[1], a[2], b[2]]
                    if m[c[2]][c[3]] == 1:
                        hit = c
            if hit is None and ((a[1] == b[0] and a[2] == b[1]) or (a[1] == b[1] and a[2] == b[0])):
                c = [a[0], a[1], a[2], b[2]]
                if m[c[2]][c[3]] == 1:
                    hit = c
                else:
                    c = [b[2], a[0], a[1], a[2]]
                    if m[c[1]][c[0]] == 1:
                        hit = c
            if hit is None and ((a[0] == b[1] and a[1] == b[2]) or (a[0] == b[2] and a[1] == b[1])):
                c = [b[0], a[0], a[1], a[2]]
                if m[c[1]][c[0]] == 1:
                    hit = c
                else:
                    c = [a[0], a[1], a[2], b[0]]
                    if m[c[2]][c[3]] == 1:
                        hit = c
            if hit is not None:
                dihs.append(hit)
    return bonds, angles, dihs


# --------------------------------------------------------------------------
# Wilson vectors (ModelHessian/calc_params.py) and the Fischer model Hessian
# --------------------------------------------------------------------------
def w_stretch(x0, x1):
    """stretch2, calc_params.py:220-227."""
    d = x0 - x1
    r = np.linalg.norm(d)
    return r, np.array([-1 * d / r, d / r])


def w_bend(x0, x1, x2):
    """bend2, calc_params.py:183-218."""
    r1, b1 = w_stretch(x0, x1)
    r2, b2 = w_stretch(x1, x2)
    co = float(b1[0] @ b2[1])
    crap = float(b1[0] @ b1[0] + b2[1] @ b2[1])
    if math.sqrt(crap) < 1e-12:
        fir = math.pi - math.asin(math.sqrt(crap))
        si = math.sqrt(crap)
    else:
        fir = math.acos(co)
        si = math.sqrt(1 - co ** 2)
    if abs(fir - math.pi) < 1e-12:
        fir = math.pi
    bf = np.zeros((3, 3))
    d1, d2 = r1 * si, r2 * si
    for i in range(3):
        bf[0][i] = 0.0 if d1 < 1e-12 else (co * b1[0][i] - b2[1][i]) / d1
        bf[2][i] = 0.0 if d2 < 1e-12 else (co * b2[1][i] - b1[0][i]) / d2
        bf[1][i] = -1 * (bf[0][i] + bf[2][i])
    return fir, bf


def w_torsion(x0, x1, x2, x3):
    """torsion2 b-vectors, calc_params.py:137-181."""
    r1, bij = w_stretch(x0, x1)
    r2, bjk = w_stretch(x1, x2)
    r3, bkl = w_stretch(x2, x3)
    f2, _ = w_bend(x0, x1, x2)
    f3, _ = w_bend(x1, x2, x3)
    s2, s3, c2, c3 = math.sin(f2), math.sin(f3), math.cos(f2), math.cos(f3)
    bt = np.zeros((4, 3))
    for ix in range(3):
        iy = (ix + 1) % 3
        iz = (iy + 1) % 3
        bt[0][ix] = (bij[1][iy] * bjk[1][iz] - bij[1][iz] * bjk[1][iy]) / (r1 * s2 ** 2)
        bt[3][ix] = (bkl[0][iy] * bjk[0][iz] - bkl[0][iz] * bjk[0][iy]) / (r3 * s3 ** 2)
        bt[1][ix] = -1 * ((r2 - r1 * c2) * bt[0][ix] + r3 * c3 * bt[3][ix]) / r2
        bt[2][ix] = -1 * (bt[0][ix] + bt[1][ix] + bt[3][ix])
    return bt


def fischer_hessian(coord, radii):
    """FischerApproxHessian.main, ModelHessian/fischer.py:212-236 (radii: covalent, Bohr)."""
    coord = np.asarray(coord, float)
    N = len(coord)
    H = np.zeros((3 * N, 3 * N))
    bonds, angles, dihs = connectivity_tables(coord, radii, 1.1)
    bm13 = np.zeros((N, N), bool)                          # second connectivity, factor 1.3 (:43-66)
    for i in range(N):
        for j in range(i + 1, N):
            bm13[i, j] = bm13[j, i] = np.linalg.norm(coord[i] - coord[j]) <= (radii[i] + radii[j]) * 1.3

    def add(atoms, k, b):
        for a in range(len(atoms)):
            for c in range(len(atoms)):
                for p in range(3):
                    for q in range(3):
                        H[3 * atoms[a] + p, 3 * atoms[c] + q] += k * b[a][p] * b[c][q]

    for i, j in bonds:
        r = np.linalg.norm(coord[i] - coord[j])
        k = 0.3601 * math.exp(-1.944 * (r - (radii[i] + radii[j])))
        add([i, j], k, w_stretch(coord[i], coord[j])[1])
    for i, j, k_ in angles:
        r1 = np.linalg.norm(coord[i] - coord[j]); r2 = np.linalg.norm(coord[j] - coord[k_])
        c1 = radii[i] + radii[j]; c2 = radii[j] + radii[k_]
        val = c1 * c2
        k = 0.0 if abs(val) < 1e-10 else 0.089 + 0.11 / val ** (-0.42) * math.exp(-0.44 * (r1 + r2 - c1 - c2))
        add([i, j, k_], k, w_bend(coord[i], coord[j], coord[k_])[1])

    def sin_sq(a, b, c):
        v1 = coord[a] - coord[b]; v2 = coord[c] - coord[b]
        cr = np.cross(v1, v2)
        n1 = v1 @ v1; n2 = v2 @ v2
        return 0.0 if n1 * n2 < 1e-12 else (cr @ cr) / (n1 * n2)

    for i, j, k_, l in dihs:
        if sin_sq(i, j, k_) < 1e-3 or sin_sq(j, k_, l) < 1e-3:
            continue
        r = np.linalg.norm(coord[j] - coord[k_]); rc = radii[j] + radii[k_]
        bond_sum = int(bm13[j].sum() + bm13[k_].sum() - 2)
        val = r * rc
        k = 0.0 if abs(val) < 1e-10 else 0.0015 + 14.0 * max(bond_sum, 0) ** 0.57 / val ** 4.0 * math.exp(-2.85 * (r - rc))
        add([i, j, k_, l], k, w_torsion(coord[i], coord[j], coord[k_], coord[l]))
    for i in range(3 * N):
        for j in range(i):
            H[i, j] = H[j, i]
    return project_hessian_trrot(H, coord.reshape(-1))


def d3_pair_block(r_vec, r, c6i, c6j, r4r2i, r4r2j, r0, s6, s8, a1, a2):
    """FischerD3ApproxHessianOld.d3_hessian_contribution, ModelHessian/fischerd3old.py:85-128: the 3 x 3 block
    h_proj P + h_perp (1 - P) of one non-bonded pair (BJ damping; the 'simplified' second derivative of the
    reference, not the analytic one)."""
    c6 = math.sqrt(c6i * c6j)
    c8 = 3.0 * c6 * math.sqrt(r4r2i * r4r2j)
    d6 = r ** 6 + (a1 * r0 + a2) ** 6
    d8 = r ** 8 + (a1 * r0 + (a2 + 2.0)) ** 8
    f6 = r ** 6 / d6
    f8 = r ** 8 / d8
    df6 = 6 * r ** 5 / d6 - 6 * r ** 12 / d6 ** 2
    df8 = 8 * r ** 7 / d8 - 8 * r ** 16 / d8 ** 2
    g6 = -s6 * c6 * ((-6.0 / r ** 7) * f6 + (1.0 / r ** 6) * df6)
    g8 = -s8 * c8 * ((-8.0 / r ** 9) * f8 + (1.0 / r ** 8) * df8)
    u = r_vec / r
    P = np.outer(u, u)
    h_proj = s6 * c6 / r ** 8 * (42.0 * f6 - r * df6) + s8 * c8 / r ** 10 * (72.0 * f8 - r * df8)
    h_perp = (g6 + g8) / r
    return h_proj * P + h_perp * (np.eye(3) - P)


def fischerd3old_hessian(coord, prm, d3=(1.0, 0.7875, 0.4289, 4.4407)):
    """FischerD3ApproxHessianOld.main, ModelHessian/fischerd3old.py:355-381 - the model Hessian a bare `-modelhess`
    selects (interface.py:184-191).  prm (N, 4): covalent radius, D2 C6 (hartree bohr^6), D3 r4r2, D2 vdW radius
    (Bohr); d3 = (s6, s8, a1, a2), PBE0 defaults of Parameters/d3.py."""
    coord = np.asarray(coord, float)
    prm = np.asarray(prm, float)
    radii = prm[:, 0]
    N = len(coord)
    H = np.zeros((3 * N, 3 * N))
    bonds, angles, dihs = connectivity_tables(coord, radii, 1.1)
    dist = np.linalg.norm(coord[:, None, :] - coord[None, :, :], axis=-1)
    bm13 = dist <= (radii[:, None] + radii[None, :]) * 1.3             # get_bond_connectivity (:131-150)
    np.fill_diagonal(bm13, False)

    def add(atoms, k, b):
        for a in range(len(atoms)):
            for c in range(len(atoms)):
                H[3 * atoms[a]:3 * atoms[a] + 3, 3 * atoms[c]:3 * atoms[c] + 3] += k * np.outer(b[a], b[c])

    for i, j in bonds:                                                    # fischer_bond (:153-183)
        r = np.linalg.norm(coord[i] - coord[j])
        k = 0.3601 * math.exp(-1.944 * (r - (radii[i] + radii[j])))
        add([i, j], k, w_stretch(coord[i], coord[j])[1])
    for i, j, k_ in angles:                                               # fischer_angle (:186-233)
        v1 = coord[i] - coord[j]; v2 = coord[k_] - coord[j]
        r1 = np.linalg.norm(v1); r2 = np.linalg.norm(v2)
        if r1 < 0.1 or r2 < 0.1:
            continue
        if abs(np.dot(v1, v2) / (r1 * r2)) > 0.9999:
            continue
        c1 = radii[i] + radii[j]; c2 = radii[j] + radii[k_]
        val = c1 * c2
        k = 0.0 if abs(val) < 1e-10 else 0.089 + 0.11 / val ** (-0.42) * math.exp(-0.44 * (r1 + r2 - c1 - c2))
        add([i, j, k_], k, w_bend(coord[i], coord[j], coord[k_])[1])
    for i, j, k_, l in dihs:                                              # fischer_dihedral (:236-319)
        vji = coord[i] - coord[j]; vjk = coord[k_] - coord[j]; vkl = coord[l] - coord[k_]
        r = np.linalg.norm(vjk); rc = radii[j] + radii[k_]
        bond_sum = int(bm13[j].sum() + bm13[k_].sum() - 2)
        val = r * rc
        k = 0.0 if abs(val) < 1e-10 else 0.0015 + 14.0 * max(bond_sum, 0) ** 0.57 / val ** 4.0 * math.exp(-2.85 * (r - rc))
        nji = np.linalg.norm(vji)
        if nji < 1e-8 or r < 1e-8:
            continue
        c1 = np.dot(vji, vjk) / (nji * r)
        nkl = np.linalg.norm(vkl)
        if nkl < 1e-8:
            continue
        c2 = np.dot(-vjk, vkl) / (r * nkl)
        s1 = 1.0 - min(c1 ** 2, 1.0); s2 = 1.0 - min(c2 ** 2, 1.0)
        if s1 < 1e-4 or s2 < 1e-4:
            continue
        add([i, j, k_, l], k * (s1 * s2), w_torsion(coord[i], coord[j], coord[k_], coord[l]))
    s6, s8, a1, a2 = d3
    for i in range(N):                                                    # d3_dispersion_hessian (:322-352)
        for j in range(i):
            if bm13[i, j]:
                continue
            rv = coord[i] - coord[j]
            r = np.linalg.norm(rv)
            if r < 0.1:
                continue
            blk = d3_pair_block(rv, r, prm[i, 1], prm[j, 1], prm[i, 2], prm[j, 2], prm[i, 3] + prm[j, 3], s6, s8, a1, a2)
            H[3 * i:3 * i + 3, 3 * i:3 * i + 3] += blk
            H[3 * j:3 * j + 3, 3 * j:3 * j + 3] += blk
            H[3 * i:3 * i + 3, 3 * j:3 * j + 3] -= blk
            H[3 * j:3 * j + 3, 3 * i:3 * i + 3] -= blk
    H = (H + H.T) / 2.0
    return project_hessian_trrot(H, coord.reshape(-1))


def fischerd3_hessian(coord, prm, d3=(1.0, 0.7875, 0.4289, 4.4407)):
    """FischerD3ApproxHessian.main, ModelHessian/fischerd3.py:186-304 (the finite-value fallbacks at :290-302 are not
    restated: they only fire on NaN input).  prm (N, 5): covalent radius, D2 C6, D3 r4r2, D2 vdW radius, reference
    coordination number."""
    coord = np.asarray(coord, float)
    prm = np.asarray(prm, float)
    radii = prm[:, 0]
    N = len(coord)
    H = np.zeros((3 * N, 3 * N))
    bm = bond_matrix(coord, radii, 1.1).astype(bool)
    bonds, angles, dihs = connectivity_tables(coord, radii, 1.1)
    dist = np.linalg.norm(coord[:, None, :] - coord[None, :, :], axis=-1)
    rm = dist.copy(); np.fill_diagonal(rm, np.inf)                        # calc_coordination_numbers (:47-62)
    with np.errstate(over="ignore"):
        term = np.clip(-16.0 * ((4.0 / 3.0) * (rm / (radii[:, None] + radii[None, :])) - 1.0), -100, 100)
    cn = np.sum(1.0 / (1.0 + np.exp(term)), axis=1)

    def add(atoms, k, b):
        for a in range(len(atoms)):
            for c in range(len(atoms)):
                H[3 * atoms[a]:3 * atoms[a] + 3, 3 * atoms[c]:3 * atoms[c] + 3] += k * np.outer(b[a], b[c])

    for i, j in bonds:                                                    # fischer_bond (:83-100)
        rv = coord[i] - coord[j]
        r = np.linalg.norm(rv)
        if r < 0.1:
            continue
        k = 0.3601 * math.exp(-1.944 * (r - (radii[i] + radii[j])))
        u = rv / r
        add([i, j], k, [u, -u])
    for i, j, k_ in angles:                                               # fischer_angle (:102-135)
        v1 = coord[i] - coord[j]; v2 = coord[k_] - coord[j]
        r1 = np.linalg.norm(v1); r2 = np.linalg.norm(v2)
        if r1 < 0.1 or r2 < 0.1:
            continue
        if abs(np.dot(v1, v2) / (r1 * r2)) > 0.9999:
            continue
        c1 = radii[i] + radii[j]; c2 = radii[j] + radii[k_]
        val = c1 * c2
        k = 0.0 if abs(val) < 1e-10 else 0.089 + 0.11 / val ** (-0.42) * math.exp(-0.44 * (r1 + r2 - c1 - c2))
        add([i, j, k_], k, w_bend(coord[i], coord[j], coord[k_])[1])
    ncount = bm.sum(axis=1)
    for i, j, k_, l in dihs:                                              # fischer_dihedral (:137-184)
        vji = coord[i] - coord[j]; vjk = coord[k_] - coord[j]; vkl = coord[l] - coord[k_]
        nji, njk, nkl = np.linalg.norm(vji), np.linalg.norm(vjk), np.linalg.norm(vkl)
        if min(nji, njk, nkl) < 0.1:
            continue
        c1 = np.dot(vji, vjk) / (nji * njk)
        c2 = np.dot(-vjk, vkl) / (njk * nkl)
        s1 = 1.0 - min(c1 ** 2, 1.0); s2 = 1.0 - min(c2 ** 2, 1.0)
        if s1 < 1e-3 or s2 < 1e-3:
            continue
        bond_sum = int(ncount[j] + ncount[k_] - 2)
        rc = radii[j] + radii[k_]
        val = njk * rc
        k = 0.0 if abs(val) < 1e-10 else 0.0015 + 14.0 * max(bond_sum, 0) ** 0.57 / val ** 4.0 * math.exp(-2.85 * (njk - rc))
        b = w_torsion(coord[i], coord[j], coord[k_], coord[l])
        if not np.all(np.isfinite(b)):
            continue
        add([i, j, k_, l], k * (s1 * s2), b)
    s6, s8, a1, a2 = d3
    scale = np.clip(1.0 - 0.05 * (cn - prm[:, 4]), 0.75, 1.25)            # dynamic C6 (:232-237)
    for i in range(N):
        for j in range(i):
            if bm[i, j] or not dist[i, j] > 0.1:
                continue
            blk = d3_pair_block(coord[i] - coord[j], dist[i, j], prm[i, 1] * scale[i], prm[j, 1] * scale[j], prm[i, 2],
                                prm[j, 2], prm[i, 3] + prm[j, 3], s6, s8, a1, a2)
            H[3 * i:3 * i + 3, 3 * i:3 * i + 3] += blk
            H[3 * j:3 * j + 3, 3 * j:3 * j + 3] += blk
            H[3 * i:3 * i + 3, 3 * j:3 * j + 3] -= blk
            H[3 * j:3 * j + 3, 3 * i:3 * i + 3] -= blk
    H = (H + H.T) / 2.0
    return project_hessian_trrot(H, coord.reshape(-1))


def fix_atoms_effective_hessian(H, fix_atoms):
    """HessianManager.calc_eff_hess_for_fix_atoms_and_set_hess, optimization.py:1325-1343 (one Hessian)."""
    fix = []
    for a in fix_atoms:
        fix.extend([3 * (a - 1), 3 * (a - 1) + 1, 3 * (a - 1) + 2])
    fix = np.array(fix, dtype="int64")
    inv = np.linalg.pinv(H[np.ix_(fix, fix)] + np.eye(len(fix)) * 1e-10)
    return H - np.dot(H[:, fix], np.dot(inv, H[fix, :]))


def ts_hessian(H):
    """TransitionStateHessian.create_ts_hessian, ModelHessian/tshess.py:14-40."""
    lam, V = np.linalg.eigh(H)
    if np.any(lam < -1e-8):
        return H
    count = 0
    for x in lam:
        if abs(x) < 1e-8:
            count += 1
        else:
            break
    v = V[:, count]
    M = (np.eye(len(lam)) - 2.0 * np.outer(v, v)) @ H
    return 0.5 * (M + M.T)


def clip_hessian(H, alpha=0.1):
    """The "clip" modifier, ModelHessian/approx_hessian.py:103-126."""
    lam, V = np.linalg.eigh(H)
    out = lam.astype(float, copy=True)
    m = np.abs(lam) >= 1.0
    out[m] = np.sign(lam[m]) * (2.0 - 1.0 / (np.abs(lam)[m] ** alpha))
    return V @ (np.diag(out) @ V.T)


# --------------------------------------------------------------------------
# AFIR bias potential (Potential/AFIR_potential.py:18-55 + autograd, potential.py:130-135)
# --------------------------------------------------------------------------
def afir_energy_torch(geom, frag1, frag2, radii_f32, gamma):
    """Energy as a torch expression (float64 geometry, FLOAT32 radii added in float32)."""
    import torch
    hartree2kjmol, bohr2ang = 2625.5, 0.52917721067
    R0 = 3.8164 / bohr2ang
    EPS = 1.0061 / hartree2kjmol
    if gamma != 0.0:
        gh = gamma / hartree2kjmol
        alpha = gh / ((2 ** (-1 / 6) - (1 + math.sqrt(1 + abs(gh) / EPS)) ** (-1 / 6)) * R0)
    else:
        alpha = 0.0
    i_idx = torch.tensor(frag1); j_idx = torch.tensor(frag2)
    Ri = radii_f32[i_idx]; Rj = radii_f32[j_idx]
    vec = torch.linalg.norm(geom[i_idx].unsqueeze(1) - geom[j_idx].unsqueeze(0), dim=2)
    omega = ((Ri.unsqueeze(1) + Rj.unsqueeze(0)) / vec) ** 6.0
    return alpha * ((omega * vec).sum() / omega.sum())


def afir_egh(coord, frag1, frag2, radii, gamma):
    """(E, grad (N,3), hess (3N,3N)); frag indices 0-based; radii float64 Bohr (rounded here)."""
    import torch
    geom = torch.tensor(np.asarray(coord, float), dtype=torch.float64, requires_grad=True)
    rf = torch.tensor([float(r) for r in radii])          # float32, as in the reference
    f = lambda x: afir_energy_torch(x, list(frag1), list(frag2), rf, float(gamma))
    E = f(geom)
    g = torch.func.jacrev(f)(geom)
    H = torch.func.hessian(f)(geom).reshape(geom.numel(), geom.numel())
    return float(E), g.detach().numpy(), H.detach().numpy()


class CalcMoveVectorOracle:
    """CalculateMoveVector.calc_move_vector with one RSIRFO instance (optimizer.py:259-309,
    534-553, 740-818): outer trust radius (only when a model Hessian / FC_COUNT is configured),
    RSIRFO.run, norm clamp, geometry update in Angstrom."""

    def __init__(self, method, saddle_order=0, model_hess_flag=None, FC_COUNT=-1):
        self.trust = 0.1 if saddle_order > 0 else 0.5
        self.max_trust = self.trust
        self.opt = RSIRFOOracle(method=method, saddle_order=saddle_order, trust_radius_max=self.max_trust,
                                trust_radius_min=0.01)
        self.tr = TrustRadiusOracle(0.01, 0.5)
        self.update_outer = not (FC_COUNT == -1 and model_hess_flag is None)

    def step(self, x, Bg, g, Be, pre=None):
        """pre: dict(x, g, Bg, Be, move) of the previous call or None.  Returns (x_new_ang, move)."""
        if self.update_outer:
            Hm = self.opt.hessian + (self.opt.bias_hessian if self.opt.bias_hessian is not None else 0.0)
            if pre is None:
                n = x.size
                self.trust = self.tr.update(Be, 0.0, np.zeros(n), np.zeros(n), Hm, self.trust)
            else:
                self.trust = self.tr.update(Be, pre["Be"], pre["Bg"], pre["move"], Hm, self.trust)
        if pre is None:
            move = self.opt.run(x, Bg, g, None, None, Be)
        else:
            move = self.opt.run(x, Bg, g, pre["x"], pre["g"], Be)
        return clamp_and_move(x, move, self.trust)


# --------------------------------------------------------------------------
# NEB: BNEB tangent projection, Ayala curvature, step limits (config 3)
# --------------------------------------------------------------------------
def _bneb_unit_projection(xa, xb, g, w):
    """-w * sum_atoms u_hat (u_hat . g) through the SVD pseudo-inverse of B^T B
    (MEP/pathopt_bneb_force.py:104-117, Coordinate/redundant_coordinate.py:377-439)."""
    out = np.zeros_like(g)
    A = xa.reshape(-1, 3); Bc = xb.reshape(-1, 3); G = g.reshape(-1, 3)
    for i in range(A.shape[0]):
        u = (Bc[i] - A[i]) / (np.linalg.norm(A[i] - Bc[i]) + 1e-15)
        s = u @ u
        if s > 1e-6:
            out[3 * i:3 * i + 3] = -w * ((u @ G[i]) / s) * u
    return out


def bneb_force(X, E, G):
    """CaluculationBNEB.calc_force without the CI branches -> (force, tau), (nimg, n) each."""
    nimg = X.shape[0]
    F = np.zeros_like(G); T = np.zeros_like(G)
    for i in range(nimg):
        if i == 0 or i == nimg - 1:
            F[i] = -G[i]
            continue
        e = E[i - 1:i + 2]
        if e[0] < e[1] < e[2]:
            proj = _bneb_unit_projection(X[i], X[i + 1], G[i], 1.0)
        elif e[0] > e[1] > e[2]:
            proj = _bneb_unit_projection(X[i - 1], X[i], G[i], 1.0)
        else:
            mx = max(abs(e[2] - e[1]), abs(e[1] - e[0])); mn = min(abs(e[2] - e[1]), abs(e[1] - e[0]))
            a = mx / (mx + mn + 1e-8); b = mn / (mx + mn + 1e-8)
            wp, wm = (a, b) if e[0] < e[2] else (b, a)
            proj = _bneb_unit_projection(X[i], X[i + 1], G[i], wp) + _bneb_unit_projection(X[i - 1], X[i], G[i], wm)
        F[i] = -(G[i] + proj)
        T[i] = proj
    return F, T


def ayala_gamma(qp, qc, qn, Ep, Ec, En, gp, gc, gn, tangent):
    """calculate_gamma, MEP/pathopt_bneb_force.py:161-222."""
    dp = np.linalg.norm(qc - qp); dn = np.linalg.norm(qn - qc)
    if dp < 1e-6 or dn < 1e-6:
        return 0.0
    s = [-dp, 0.0, dn]
    A = np.array([[1, v, v ** 2, v ** 3, v ** 4, v ** 5] for v in s] +
                 [[0, 1, 2 * v, 3 * v ** 2, 4 * v ** 3, 5 * v ** 4] for v in s], float)
    b = np.array([Ep, Ec, En, gp @ ((qc - qp) / dp), gc @ tangent, gn @ ((qn - qc) / dn)])
    try:
        return 2.0 * np.linalg.solve(A, b)[2]
    except np.linalg.LinAlgError:
        return 0.0


def neb_limit_tr(X, G, delta, step_limit=True):
    """_limit_step_size (Optimizer/rfo_neb.py:76-83; skipped with step_limit=False, as the FIRE
    optimizer does) + TR_NEB.TR_calc (Optimizer/trust_radius_neb.py:17-98), free end images."""
    nimg = X.shape[0]
    out = np.zeros_like(delta)
    for i in range(nimg):
        d = delta[i].copy()
        end = i == 0 or i == nimg - 1
        nrm = np.linalg.norm(d)
        if step_limit and nrm > 1e-8:
            d = d / nrm * min(0.2 if end else 0.1, nrm)
        nrm = np.linalg.norm(d)
        if end:
            out[i] = 0.0 if nrm < 1e-15 else min(0.5, nrm) * d / nrm
            continue
        t1 = np.linalg.norm(X[i] - X[i - 1]) / 2.0; t2 = np.linalg.norm(X[i] - X[i + 1]) / 2.0
        v1 = (X[i - 1] - X[i]) / (np.linalg.norm(X[i - 1] - X[i]) + 1e-15)
        v2 = (X[i + 1] - X[i]) / (np.linalg.norm(X[i + 1] - X[i]) + 1e-15)
        with np.errstate(all="ignore"):
            nd = d / nrm
            c1 = np.sum(v1 * nd); c2 = np.sum(v2 * nd)
            fc = np.sum(G[i] * d) / (np.linalg.norm(G[i]) * nrm)
        if fc >= 0:
            if (c1 > 0 and c2 < 0) or (c1 < 0 and c2 > 0):
                if nrm > t1 and c1 > 0:
                    d = d * t1 / nrm
                elif nrm > t2 and c2 > 0:
                    d = d * t2 / nrm
            elif c1 < 0 and c2 < 0:
                pass
            else:
                if nrm > t1:
                    d = d * t1 / nrm
                elif nrm > t2:
                    d = d * t2 / nrm
            out[i] = d
        else:
            out[i] = 0.0
    return out


class NEBRFOOracle:
    """RFOOptimizer.optimize steps 1-3 (Optimizer/rfo_neb.py:104-182): tangents, Ayala update,
    per-image RSIRFO.run with B_e = pre_B_e = 0, step limits, TR_calc.  FIRE blend excluded."""

    def __init__(self, H_init):
        nimg = H_init.shape[0]
        self.H = [h.copy() for h in H_init]
        self.opts = []
        for i in range(nimg):
            if i == 0 or i == nimg - 1:
                o = RSIRFOOracle(method="rsirfo_block_fsb", saddle_order=0)
                o.trust_radius = 0.5
            else:
                o = RSIRFOOracle(method="rsirfo_block_bofill", saddle_order=0)
                o.trust_radius = 0.2
                o.NEB_mode = True
            self.opts.append(o)
        self.prev = None

    def step(self, X, E, G):
        nimg = X.shape[0]
        F, T = bneb_force(X, E, G)
        gam = np.zeros(nimg); delta = np.zeros_like(X)
        for i in range(nimg):
            if 0 < i < nimg - 1:
                gam[i] = ayala_gamma(X[i - 1], X[i], X[i + 1], E[i - 1], E[i], E[i + 1], G[i - 1], G[i], G[i + 1], T[i])
                self.H[i] = self.H[i] + gam[i] * np.outer(T[i], T[i])
            o = self.opts[i]
            o.set_hessian(self.H[i]); o.set_bias_hessian(np.zeros_like(self.H[i]))
            if self.prev is None:
                delta[i] = o.run(X[i], G[i], G[i], None, None, 0.0)
            else:
                delta[i] = o.run(X[i], G[i], G[i], self.prev[0][i], self.prev[1][i], 0.0)
            self.H[i] = o.hessian
        self.prev = (X.copy(), G.copy())
        return F, T, gam, delta, neb_limit_tr(X, G, delta)


# --------------------------------------------------------------------------
# Lindh model Hessian, decomposed (ModelHessian/lindh.py:79-165)
# --------------------------------------------------------------------------
LINDH_ALPHA = [[1.0000, 0.3949, 0.3949], [0.3949, 0.2800, 0.2800], [0.3949, 0.2800, 0.2800]]


def lindh_kdiag(coord, prm):
    """Diagonal RIC force constants of guess_lindh_hessian (lindh.py:79-143).
    prm (N, 6): covalent radius, period index, mass, UFF distance, UFF well depth, UFF charge."""
    coord = np.asarray(coord, float)
    N = len(coord)
    rad = prm[:, 0]
    tabs = connectivity_tables(coord, rad, 1.1)
    pairs = [(i, j) for i in range(N) for j in range(i + 1, N)]
    index = {p: k for k, p in enumerate(pairs)}
    kd = [0.0] * len(pairs)
    base = [0.45, 0.15, 0.005]
    for tab in tabs:
        for idx in tab:
            f = base[len(idx) - 2]
            for q in range(len(idx) - 1):
                i, j = idx[q], idx[q + 1]
                cR = rad[i] + rad[j]
                al = LINDH_ALPHA[int(prm[i, 1])][int(prm[j, 1])]
                R = np.linalg.norm(coord[i] - coord[j])
                f *= math.exp(al * (cR ** 2 - R ** 2))
            if len(idx) == 2:
                lo, hi = sorted(idx)
                m1, m2 = prm[lo, 2], prm[hi, 2]
                kd[index[(lo, hi)]] += f / ((m1 * m2) / (m1 + m2))
            else:
                for q in range(len(idx) - 1):
                    kd[index[tuple(sorted((idx[q], idx[q + 1])))]] += f
    bonded = {tuple(b) for b in tabs[0]}
    for k, (i, j) in enumerate(pairs):
        if (i, j) in bonded:
            continue
        d = np.linalg.norm(coord[i] - coord[j])
        eps = math.sqrt(prm[i, 4] * prm[j, 4]); sig = math.sqrt(prm[i, 3] * prm[j, 3])
        kd[k] += -12 * eps * (-7 * (sig ** 6 / d ** 8) + 13 * (sig ** 12 / d ** 14))
        kd[k] += 664.12 * (prm[i, 5] * prm[j, 5] / d ** 3) * (0.52917721067 ** 2 / 627.509)
    return np.array(kd)


def lindh_hessian_bkb(coord, prm):
    """project(B^T diag(k) B) with the all-pairs distance B matrix
    (Coordinate/redundant_coordinate.py:15-43,145 without the K term)."""
    coord = np.asarray(coord, float)
    N = len(coord)
    kd = lindh_kdiag(coord, prm)
    pairs = [(i, j) for i in range(N) for j in range(i + 1, N)]
    B = np.zeros((len(pairs), 3 * N))
    for k, (i, j) in enumerate(pairs):
        e = (coord[i] - coord[j]) / np.linalg.norm(coord[i] - coord[j])
        B[k, 3 * i:3 * i + 3] = e
        B[k, 3 * j:3 * j + 3] = -e
    return project_hessian_trrot(B.T @ np.diag(kd) @ B, coord.reshape(-1))


# --------------------------------------------------------------------------
# EnhancedRSPRFO (Optimizer/rsprfo.py) — config 5
# --------------------------------------------------------------------------
def project_grad_trrot_qr_valid(g, x):
    """EnhancedRSPRFO._project_grad_tr_rot, rsprfo.py:224-285 (drops |diag R| <= 1e-10 columns)."""
    coords = x.reshape(-1, 3)
    if coords.shape[0] < 3:
        return g
    A = trrot_vectors(x).T
    Q, R = np.linalg.qr(A, mode="reduced")
    Q = Q[:, np.abs(np.diag(R)) > 1e-10]
    return g - Q @ (Q.T @ g)


def arrowhead_extreme(lam, gam, mode):
    """solve_rfo on the augmented (arrowhead) Hessian, rsprfo.py:1097-1168, alpha = 1:
    step = -v[:-1] / v[-1] of the smallest ('min') or largest ('max') eigenpair."""
    k = lam.size
    Haug = np.zeros((k + 1, k + 1))
    Haug[np.arange(k), np.arange(k)] = lam
    Haug[:k, k] = gam
    Haug[k, :k] = gam
    w, V = np.linalg.eigh(Haug)
    idx = int(np.argmin(w)) if mode == "min" else int(np.argmax(w))
    v = V[:, idx]
    nu = v[-1]
    if abs(nu) < 1e-12:
        nu = np.sign(nu) * 1e-12 if nu != 0 else 1e-12
    return -v[:-1] / nu, w[idx]


class RSPRFOOracle:
    """EnhancedRSPRFO.run for one structure (default configuration).  The alpha micro-cycles are a
    no-op analytically (eigenvalues and gradient are both divided by alpha, rsprfo.py:1116-1121,
    SURVEY H3): every cycle reproduces the alpha = 1 step, the loop leaves through the
    trust-radius / stagnation / bound exits and returns that step scaled to the effective trust
    radius; this restatement evaluates the alpha = 1 step once."""

    def __init__(self, method="rsprfo_bofill", saddle_order=1, trust_radius_max=None, trust_radius_min=0.01, **_):
        self.method_id = resolve_update_method(method)
        self.saddle_order = saddle_order
        if saddle_order == 0:
            self.trust0 = 0.5; self.trust_max = 0.5 if trust_radius_max is None else trust_radius_max
        else:
            self.trust0 = 0.1; self.trust_max = 0.3 if trust_radius_max is None else trust_radius_max
        self.trust = self.trust0
        self.trust_min = 0.01 if trust_radius_min is None else trust_radius_min
        self.hessian = None
        self.bias_hessian = None
        self.first = True
        self.prev_energy = None
        self.pred = []
        self.prev_gradient = None
        self.prev_move = None
        self.ts_vec = None
        self.rejections = 0
        self.last = {}

    def set_hessian(self, H):
        self.hessian = 0.5 * (np.asarray(H, float) + np.asarray(H, float).T)     # copies (rsprfo.py:1317-1318)

    def set_bias_hessian(self, H):
        self.bias_hessian = None if H is None else np.asarray(H, float).copy()

    def _eff_trust(self, gnorm):
        if gnorm < 1e-3:                                           # rsprfo.py:392-419
            r = 0.5 * gnorm / 1e-3 * self.trust_max
            return min(max(r, self.trust_min), self.trust)
        return self.trust

    def run(self, x, Bg, x_prev=None, Bg_prev=None, Be=0.0, pre_move=None):
        x = np.asarray(x, float).ravel(); Bg = np.asarray(Bg, float).ravel()
        info = {"updated": False, "shifted": False}
        if self.first:
            self.first = False
        elif self.prev_energy is not None and self.pred:            # _process_previous_step (:908-962)
            actual = Be - self.prev_energy
            psn = np.linalg.norm(pre_move) if (pre_move is not None and len(pre_move) > 0) else np.linalg.norm(self.prev_move)
            Hs = self.hessian + self.bias_hessian if self.bias_hessian is not None else self.hessian
            Hp = project_hessian_trrot(Hs, x)
            s = self.prev_move
            pred_red = -(self.prev_gradient @ s + 0.5 * s @ (Hp @ s))
            if abs(pred_red) < 1e-14:
                ratio = 1.0 if abs(actual) < 1e-14 else 0.0
            else:
                ratio = actual / pred_red
                if not np.isfinite(ratio):
                    ratio = 0.0
            at_boundary = psn >= self.trust * 0.95
            if ratio < 0.25:
                self.trust = max(0.25 * psn, self.trust_min)
            elif ratio > 0.75 and at_boundary:
                self.trust = min(2.0 * self.trust, self.trust_max)
            info["ratio"] = ratio
        # Hessian update with the BIASED gradients, no curvature sign test (:1190-1260)
        if self.prev_gradient is not None and x_prev is not None and Bg_prev is not None and len(x_prev) > 0 and len(Bg_prev) > 0:
            s = x - np.asarray(x_prev, float).ravel(); y = Bg - np.asarray(Bg_prev, float).ravel()
            if not (np.linalg.norm(s) < 1e-10 or np.linalg.norm(y) < 1e-10):
                Hn = self.hessian + hessian_update_delta(self.method_id, self.hessian, s, y)
                Hn = 0.5 * (Hn + Hn.T)
                if not np.max(np.abs(np.linalg.eigvalsh(Hn))) > 1e6:
                    self.hessian = Hn
                    info["updated"] = True
        g = project_grad_trrot_qr_valid(Bg, x)
        gnorm = np.linalg.norm(g)
        H = self.hessian + self.bias_hessian if self.bias_hessian is not None else self.hessian
        lam, V = np.linalg.eigh(H)
        if not (np.all(np.isfinite(lam)) and np.all(np.isfinite(V))):
            lam = np.ones_like(lam); V = np.eye(lam.size)
        # eigenvalue shifting (:287-355)
        so = self.saddle_order
        lam2 = lam.copy(); shifted = False
        if so == 0:
            if lam.min() < 0.001:
                lam2 = lam + (0.001 - lam.min()); shifted = True
        else:
            order = np.argsort(lam)
            for i in range(so):
                if lam[order[i]] > -0.001:
                    lam2[order[i]] = -0.001; shifted = True
            for i in range(so, lam.size):
                if lam[order[i]] < 1e-6:
                    lam2[order[i]] = 0.001; shifted = True
        if shifted:
            Hsh = V @ np.diag(lam2) @ V.T
            H = 0.5 * (Hsh + Hsh.T)
            lam, V = np.linalg.eigh(H)
        info["shifted"] = shifted
        n = lam.size
        # mode selection (:964-1071)
        if so == 0:
            max_idx = []
        else:
            order = np.argsort(lam)
            if self.ts_vec is None:
                self.ts_vec = V[:, order[0]].copy()
                max_idx = order[:so].tolist()
            else:
                ov = np.abs(V.T @ self.ts_vec)
                best = int(np.argmax(ov))
                if ov[best] > 0.5:
                    self.ts_vec = V[:, best].copy()
                    max_idx = [best] + [i for i in order if i != best][:so - 1]
                else:
                    sig = np.where(ov > 0.3)[0]
                    if len(sig) == 0:
                        self.ts_vec = V[:, order[0]].copy()
                        max_idx = order[:so].tolist()
                    else:
                        wts = [ov[i] ** 2 * (1.0 if lam[i] < 0 else 0.1) for i in sig]
                        best = int(sig[int(np.argmax(wts))])
                        self.ts_vec = V[:, best].copy()
                        max_idx = [best] + [i for i in order if i != best][:so - 1]
        min_idx = [i for i in range(n) if i not in max_idx]
        gt = V.T @ g
        step = np.zeros(n)
        if max_idx:
            step[max_idx], _ = arrowhead_extreme(lam[max_idx], gt[max_idx], "max")
        if min_idx:
            step[min_idx], _ = arrowhead_extreme(lam[min_idx], gt[min_idx], "min")
        eff = self._eff_trust(gnorm)
        nrm = np.linalg.norm(step)
        if nrm > eff:
            step = step * (eff / nrm)
        if not np.all(np.isfinite(step)):                          # (:833-846)
            sd = -gt; sdn = np.linalg.norm(sd); tgt = min(sdn, self.trust)
            step = sd * (tgt / sdn) if sdn > 1e-12 else np.zeros(n)
        move = V @ step
        snorm = np.linalg.norm(move)
        if not (gnorm < 1e-10 or snorm < 1e-10):                   # gradient-based scaling (:357-390)
            r = snorm / gnorm
            if r > 50.0:
                sc = max(50.0 / r, 0.1)
                move = move * sc; snorm = snorm * sc
        eff = self._eff_trust(gnorm)
        if snorm > eff * 1.01:
            move = move * (eff / snorm)
        pred = g @ move + 0.5 * move @ (H @ move)
        self.pred.append(pred)
        self.prev_gradient = Bg.copy(); self.prev_energy = Be; self.prev_move = move.copy()
        info.update(eigvals=lam, pred=pred, trust=self.trust, max_idx=list(max_idx))
        self.last = info
        return move


# ---------------------------------------------------------------------------------------------
# Swart model Hessian: ModelHessian/swart.py:58-355
# ---------------------------------------------------------------------------------------------
def swart_geometry(xyz, radii):
    """_precompute_geometry, swart.py:64-82."""
    diff = xyz[:, None, :] - xyz[None, :, :]
    d = np.sqrt((diff * diff).sum(axis=2))
    d = np.maximum(d, 1e-8)
    np.fill_diagonal(d, 1.0)
    cs = np.maximum(radii[:, None] + radii[None, :], 1e-8)
    screen = np.exp(1.0 - d / cs)
    np.fill_diagonal(screen, 0.0)
    return diff, d, screen


def swart_angle_rows(v1, v2, l1, l2):
    """_calculate_batch_angle_B for one triple, swart.py:109-133: (row (9,), cos, sin^2)."""
    l1 = max(l1, 1e-8); l2 = max(l2, 1e-8)
    n1, n2 = v1 / l1, v2 / l2
    c = min(max(float(n1 @ n2), -1.0), 1.0)
    s2 = max(1e-12, 1.0 - c * c)
    den = max(np.sqrt(s2), 1e-6)
    bi = (c * n1 - n2) / (l1 * den)
    bk = (c * n2 - n1) / (l2 * den)
    return np.concatenate([bi, -(bi + bk), bk]), c, s2


def swart_linear_rows(v1, v2, l1, l2):
    """_calculate_batch_linear_B for one triple, swart.py:135-190: (2, 9)."""
    l1 = max(l1, 1e-8); l2 = max(l2, 1e-8)
    vn = np.cross(v1, v2)
    nvn = np.linalg.norm(vn)
    if nvn < 1e-12:
        ref = np.array([1.0, 0.0, 0.0])
        cand = ref - (ref @ v1) / l1 ** 2 * v1
        cn = np.linalg.norm(cand)
        if cn >= 1e-12:
            vn, nvn = cand, cn
        else:
            ref = np.array([0.0, 1.0, 0.0])
            cand = ref - (ref @ v1) / l1 ** 2 * v1
            vn, nvn = cand, max(np.linalg.norm(cand), 1e-12)
    vnn = vn / max(nvn, 1e-12)
    vn2 = np.cross(v1 - v2, vnn)
    vn2 = vn2 / max(np.linalg.norm(vn2), 1e-12)
    Bm = np.zeros((2, 9))
    for row, u in ((1, vnn), (0, vn2)):
        Bm[row, 0:3] = u / l1
        Bm[row, 6:9] = u / l2
        Bm[row, 3:6] = -Bm[row, 0:3] - Bm[row, 6:9]
    return Bm


def swart_raw_hessian(xyz, radii, angles=True):
    """Bond + angle terms before the TR/ROT projection (swart.py:83-107,192-315)."""
    xyz = np.asarray(xyz, float); N = len(xyz); n = 3 * N
    diff, d, screen = swart_geometry(xyz, np.asarray(radii, float))
    H = np.zeros((n, n))
    for i in range(N):
        for j in range(i + 1, N):
            e = diff[i, j] / d[i, j]
            b = np.concatenate([e, -e])
            idx = np.r_[3 * i:3 * i + 3, 3 * j:3 * j + 3]
            H[np.ix_(idx, idx)] += 0.35 * screen[i, j] ** 3 * np.outer(b, b)
    if not angles:
        return H
    f, tolth, eps1 = 0.12, 0.2, 0.3 ** 2
    eps2 = 0.3 ** 2 / np.exp(1)
    for j in range(N):
        nb = np.where(screen[j] >= eps2)[0]
        for a in range(len(nb)):
            for c in range(a + 1, len(nb)):
                i, k = nb[a], nb[c]
                ss = screen[i, j] * screen[j, k]
                if ss < eps1 or not (d[i, j] > 1e-8 and d[k, j] > 1e-8):
                    continue
                v1, v2, l1, l2 = diff[i, j], diff[k, j], d[i, j], d[k, j]
                bn, cs, s2 = swart_angle_rows(v1, v2, l1, l2)
                hb = 0.075 * ss ** 2 * (f + (1.0 - f) * np.sqrt(s2)) ** 2
                th1 = 1.0 - cs if cs > 1.0 - tolth else 1.0 + cs
                idx = np.r_[3 * i:3 * i + 3, 3 * j:3 * j + 3, 3 * k:3 * k + 3]
                if th1 >= tolth:
                    blk = hb * np.outer(bn, bn)
                else:
                    sl = (1.0 - (th1 / tolth) ** 2) ** 2
                    if cs > 1.0 - tolth:
                        bl = swart_linear_rows(v1, v2, l1, l2)
                        bc = sl * bl[0] + (1.0 - sl) * bn
                        blk = hb * np.outer(bl[1], bl[1]) + hb * np.outer(bc, bc)
                    else:
                        bs = (1.0 - sl) * bn
                        blk = hb * np.outer(bs, bs)
                H[np.ix_(idx, idx)] += blk
    return H


def swart_hessian(xyz, radii):
    """SwartApproxHessian.main, swart.py:317-355 (NaN fallback to bonds only, then projection)."""
    xyz = np.asarray(xyz, float)
    H = swart_raw_hessian(xyz, radii)
    if not np.all(np.isfinite(H)):
        H = swart_raw_hessian(xyz, radii, angles=False)
    return project_hessian_trrot(H, xyz.reshape(-1))


# ---------------------------------------------------------------------------------------------
# Redundant internal coordinates: Coordinate/redundant_coordinate.py
# ---------------------------------------------------------------------------------------------
def ric_bmatrix(xyz):
    """All-pairs distance B matrix, rows in itertools.combinations order (:15-43)."""
    xyz = np.asarray(xyz, float); N = len(xyz)
    rows = []
    for i in range(N):
        for j in range(i + 1, N):
            e = (xyz[i] - xyz[j]) / np.linalg.norm(xyz[i] - xyz[j])
            r = np.zeros(3 * N); r[3 * i:3 * i + 3] = e; r[3 * j:3 * j + 3] = -e
            rows.append(r)
    return np.array(rows)


def ric_partial_row(xyz, labels):
    """partial_stretch / bend / torsion B rows (:150-320); labels are 1-based, 2 / 3 / 4 of them."""
    xyz = np.asarray(xyz, float); N = len(xyz)
    out = np.zeros(3 * N)
    idx = [l - 1 for l in labels]
    if len(idx) == 2:
        i, j = idx
        e = (xyz[i] - xyz[j]) / np.linalg.norm(xyz[i] - xyz[j])
        parts = [e, -e]
    elif len(idx) == 3:
        i, j, k = idx
        u, w = xyz[i] - xyz[j], xyz[k] - xyz[j]
        lu, lw = np.linalg.norm(u), np.linalg.norm(w)
        c = min(max(u @ w / (lu * lw), -1.0), 1.0)
        th = np.arccos(c)
        if abs(th) > np.pi - 1e-6:
            parts = [(np.pi - th) / (2 * lu ** 2) * u, (1 / lu - 1 / lw) * (np.pi - th) / (2 * lu) * u,
                     (np.pi - th) / (2 * lw ** 2) * w]
        else:
            ct, st = 1 / np.tan(th), np.sin(th)
            parts = [ct * u / lu ** 2 - w / (lu * lw * st),
                     (u + w) / (lu * lw * st) - ct * (u / lu ** 2 + w / lw ** 2),
                     ct * w / lw ** 2 - u / (lu * lw * st)]
    else:
        i, j, k, l = idx
        vij, vlk, vkj = xyz[i] - xyz[j], xyz[l] - xyz[k], xyz[k] - xyz[j]
        nkj = np.linalg.norm(vkj); ukj = vkj / nkj
        a1 = vij - (vij @ ukj) * ukj
        a2 = vlk - (vlk @ ukj) * ukj
        n1, n2 = np.linalg.norm(a1), np.linalg.norm(a2)
        sg = np.sign(np.linalg.det(np.array([vlk, vij, vkj]))) or 1
        c = min(max(a1 @ a2 / (n1 * n2), -1.0), 1.0)
        phi = np.arccos(c) * sg
        A = (vij @ ukj) / nkj; Bc = (vlk @ ukj) / nkj
        if abs(phi) > np.pi - 1e-6 or abs(phi) < 1e-6:
            G = np.cross(vkj, a1); uG = G / np.linalg.norm(G)
            last = uG / n2 if abs(phi) > np.pi - 1e-6 else -uG / n2
            parts = [uG / n1, -((1 - A) / n1 - Bc / n2) * uG, -((1 + Bc) / n2 + A / n1) * uG, last]
        else:
            ct, st = 1 / np.tan(phi), np.sin(phi)
            parts = [ct * a1 / n1 ** 2 - a2 / (n1 * n2 * st),
                     ((1 - A) * a2 - Bc * a1) / (n1 * n2 * st) - ct * ((1 - A) * a1 / n1 ** 2 - Bc * a2 / n2 ** 2),
                     ((1 + Bc) * a1 + A * a2) / (n1 * n2 * st) - ct * ((1 + Bc) * a2 / n2 ** 2 + A * a1 / n1 ** 2),
                     ct * a2 / n2 ** 2 - a1 / (n1 * n2 * st)]
    for n_ in range(N):                 # the reference's chain picks the FIRST matching label
        for a, v in zip(idx, parts):
            if n_ == a:
                out[3 * n_:3 * n_ + 3] = v
                break
    return out


def _ric_coordinate_torch(c):
    """TorchDerivatives.distance / angle / dihedral_angle (:442-477) on a (m, 3) tensor."""
    import torch
    if c.shape[0] == 2:
        return torch.linalg.norm(c[0] - c[1])
    if c.shape[0] == 3:
        v1, v2 = c[0] - c[1], c[2] - c[1]
        return torch.arccos(torch.matmul(v1, v2) / (torch.linalg.norm(v1) * torch.linalg.norm(v2) + 1e-15))
    a1, a2, a3 = c[1] - c[0], c[2] - c[1], c[3] - c[2]
    v1 = torch.linalg.cross(a1, a2); v1 = v1 / torch.linalg.norm(v1, ord=2)
    v2 = torch.linalg.cross(a2, a3); v2 = v2 / torch.linalg.norm(v2, ord=2)
    ca = torch.sum(v1 * v2) / torch.sum((v1 ** 2) * torch.sum(v2 ** 2) + 1e-15) ** 0.5
    return torch.abs(torch.arccos(ca))


def ric_kmatrix(xyz, tables, ricgrad):
    """K of RIChess2carthess (:63-143): sum_t ricgrad[t] * d2 q_t / dx2 over bonds, angles, dihedrals,
    t counting through the three tables in order (ricgrad is indexed by this counter)."""
    import torch
    xyz = np.asarray(xyz, float); N = len(xyz)
    K = np.zeros((3 * N, 3 * N))
    t = 0
    for tab in tables:
        for atoms in tab:
            atoms = [int(a) for a in atoms]
            c = torch.tensor(xyz[atoms], dtype=torch.float64)
            h = torch.func.hessian(_ric_coordinate_torch)(c).reshape(3 * len(atoms), 3 * len(atoms)).numpy()
            idx = np.concatenate([np.arange(3 * a, 3 * a + 3) for a in atoms])
            K[np.ix_(idx, idx)] += h * ricgrad[t]
            t += 1
    return K


def ric_inv_G(G, threshold=1e-6):
    """calc_inv_G_mat (:381-394): SVD with s > threshold -> 1/s, otherwise s is KEPT (not zeroed)."""
    U, s, VT = np.linalg.svd(G)
    f = np.where(s > threshold, 1.0 / np.where(s > threshold, s, 1.0), s)
    return VT.T @ np.diag(f) @ U.T


def ric_int_grad(pB, cart_grad):
    """calc_int_grad_from_pBmat (:432-435): (G^+ pB^T)^T cart_grad with G = pB^T pB."""
    pB = np.asarray(pB, float)
    Binv = (ric_inv_G(pB.T @ pB) @ pB.T).T
    return Binv @ np.asarray(cart_grad, float).reshape(-1)


def ric_cart_grad(pB, int_grad):
    """calc_cart_grad_from_pBmat (:437-439)."""
    return np.asarray(pB, float).T @ np.asarray(int_grad, float).reshape(-1)


# ---------------------------------------------------------------------------------------------
# Step post-processing either side of the path
# ---------------------------------------------------------------------------------------------
def kabsch(P, Q):
    """Calculationtools.kabsch_algorithm (Utils/calc_tools.py:412-425): returns (P rotated onto Q and
    centred, Q centred); the reference mutates both arguments in place."""
    P = np.array(P, float); Q = np.array(Q, float)
    P -= P.mean(axis=0); Q -= Q.mean(axis=0)
    U, S, Vt = np.linalg.svd(P.T @ Q)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[-1, :] *= -1
        R = Vt.T @ U.T
    return (R @ P.T).T, Q


def rms_safely(v, threshold=1e-10):
    """calculate_rms_safely (optimization.py:1244-1250)."""
    v = np.asarray(v, float).ravel()
    f = v[np.abs(v) > threshold]
    return float(np.sqrt((f ** 2).mean())) if f.size else 0.0


def check_convergence(grad, disp, max_force_thr, rms_force_thr, max_disp_thr, rms_disp_thr):
    """ConvergenceChecker.check_convergence (optimization.py:1252-1289) without the optimizer-instance
    override: (converged, max_displacement_threshold, rms_displacement_threshold, the four measures)."""
    g = np.asarray(grad, float); d = np.asarray(disp, float)
    mf, rf = float(np.abs(g).max()), rms_safely(g)
    md, rd = float(np.abs(d).max()), rms_safely(d)
    mdt = max(max_disp_thr, max_disp_thr + max(0.0, max_force_thr - mf))
    rdt = max(rms_disp_thr, rms_disp_thr + max(0.0, rms_force_thr - rf))
    ok = mf < max_force_thr and rf < rms_force_thr and md < mdt and rd < rdt
    return bool(ok), mdt, rdt, (mf, rf, md, rd)


class FIRENEBOracle:
    """FIREOptimizer.optimize (Optimizer/fire_neb.py:38-92) up to the move vector: per-atom velocity /
    force blend, the global power test P = sum v_prev . F, the (dt, a, n_reset) schedule (note that `a` is
    multiplied by FIRE_f_inc, :65), velocity Verlet-like update, TR_calc.  Arrays are (nimg, natoms, 3)."""

    def __init__(self, dt=0.5, a=0.10, n_reset=0, N_accelerate=5, f_inc=1.10, f_decelerate=0.5, a_start=0.1, dt_max=3.0):
        self.dt, self.a, self.n_reset = dt, a, n_reset
        self.N_acc, self.f_inc, self.f_dec, self.a_start, self.dt_max = N_accelerate, f_inc, f_decelerate, a_start, dt_max

    def step(self, X, F, V, V_prev, optimize_num):
        X = np.asarray(X, float); F = np.asarray(F, float); V = np.asarray(V, float)
        have_prev = V_prev is not None and len(V_prev) > 1
        fn = np.linalg.norm(F, axis=2, keepdims=True); vn = np.linalg.norm(V, axis=2, keepdims=True)
        with np.errstate(all="ignore"):
            blend = (1.0 - self.a) * V + self.a * (vn / fn) * F
        vneb = np.where(fn > 1e-10, blend, V)
        P = float(np.sum(np.asarray(V_prev, float) * F)) if (optimize_num != 0 and have_prev) else 0.0
        if optimize_num > 0 and P > 0 and have_prev:
            if self.n_reset > self.N_acc:
                self.dt = min(self.dt * self.f_inc, self.dt_max)
                self.a *= self.f_inc
            self.n_reset += 1
        else:
            vneb = vneb * 0
            self.a = self.a_start
            self.dt *= self.f_dec
            self.n_reset = 0
        Vnew = vneb + self.dt * F
        delta = self.dt * (Vnew + np.asarray(V_prev, float)) if (optimize_num != 0 and have_prev) else self.dt * Vnew
        nimg = X.shape[0]
        move = neb_limit_tr(X.reshape(nimg, -1), F.reshape(nimg, -1), delta.reshape(nimg, -1), step_limit=False)
        return Vnew, delta, move.reshape(X.shape), P


def neb_optimize_step(orc, X, E, G, V, V_prev, optimize_num, ratio=0.5, fire_cfg=None):
    """RFOOptimizer.optimize (Optimizer/rfo_neb.py:104-208) from the pieces above: RFO move vectors (orc: a
    NEBRFOOracle carrying the Hessians), a FRESH FIREOptimizer per call (rfo_neb.py:185) driven with the projected
    NEB force, and the combine of :196-203 - ends take -rfo, the interior (1 - r) fire - r rfo.
    X (nimg, n), V / V_prev (nimg, natoms, 3).  Returns the new geometry in Bohr, (nimg, n)."""
    nimg, n = X.shape
    F, T, gam, delta, rfo = orc.step(X, E, G)
    fire = FIRENEBOracle(**(fire_cfg or {}))
    _, _, fmove, _ = fire.step(X.reshape(nimg, -1, 3), F.reshape(nimg, -1, 3), V, V_prev, optimize_num)
    fmove = fmove.reshape(nimg, n)
    move = (1.0 - ratio) * fmove - ratio * rfo
    move[0] = -rfo[0]; move[-1] = -rfo[-1]
    return X + move


# ---------------------------------------------------------------------------------------------
# Restraint bias potentials: Potential/keep_potential.py, keep_angle_potential.py
# ---------------------------------------------------------------------------------------------
def _keep_energy_torch(geom, kind, f1, f2, k, p):
    """calc_energy of StructKeepPotential (kind 1), StructKeepPotentialv2 (kind 2), StructKeepAnglePotential
    (kind 3), StructKeepDihedralAnglePotential (kind 4; keep_dihedral_angle_potential.py:62-154) restated on a
    torch tensor (atoms 0-based; p = distance in Angstrom, angle in degrees, or - kind 4 - phi0 in RADIANS as
    the reference's float32 / float64 deg2rad left it)."""
    import math
    import torch
    if kind == 4:
        i1, i2, i3, i4 = f1
        b1, b2, b3 = geom[i2] - geom[i1], geom[i3] - geom[i2], geom[i4] - geom[i3]
        n1, n2 = torch.linalg.cross(b1, b2), torch.linalg.cross(b2, b3)
        n1sq, n2sq = torch.sum(n1 ** 2), torch.sum(n2 ** 2)

        def switch(val):
            t = torch.clamp((val - 1e-10) / (1e-8 - 1e-10), 0.0, 1.0)
            return t * t * (3.0 - 2.0 * t)
        n1h = n1 / torch.clamp(torch.sqrt(n1sq), min=1e-12)
        n2h = n2 / torch.clamp(torch.sqrt(n2sq), min=1e-12)
        b2h = b2 / torch.clamp(torch.linalg.norm(b2), min=1e-12)
        phi = torch.atan2(torch.sum(torch.linalg.cross(n1h, n2h) * b2h), torch.sum(n1h * n2h))
        diff = phi - p
        diff = diff - 2.0 * math.pi * torch.round(diff / (2.0 * math.pi))
        return 0.5 * k * diff ** 2 * switch(n1sq) * switch(n2sq)
    if kind in (1, 2):
        v = geom[list(f1)].mean(dim=0) - geom[list(f2)].mean(dim=0)
        d = torch.clamp(torch.sqrt(torch.sum(v ** 2)), min=1e-12)
        return 0.5 * k * (d - p / BOHR2ANG) ** 2
    i, j, kk = f1
    theta0 = math.radians(p) if False else p * (math.pi / 180.0)
    v1, v2 = geom[i] - geom[j], geom[kk] - geom[j]
    u = torch.dot(v1, v2) / torch.clamp(torch.linalg.norm(v1) * torch.linalg.norm(v2), min=1e-12)
    u = torch.clamp(u, -1.0, 1.0)
    ucp, ucn = math.cos(1e-3), math.cos(math.pi - 1e-3)
    C = [128.0 / 1575.0, 4.0 / 35.0, 8.0 / 45.0, 1.0 / 3.0, 2.0]

    def taylor(delta):
        t = C[0]
        for c in C[1:]:
            t = c + delta * t
        return delta * t
    near0, nearpi = bool(u > ucp), bool(u < ucn)
    if abs(theta0) < 1e-8:
        if near0:
            return 0.5 * k * taylor(1.0 - u)
        if nearpi:
            return 0.5 * k * (math.pi - torch.sqrt(torch.clamp(taylor(1.0 + u), min=1e-30))) ** 2
        return 0.5 * k * torch.acos(torch.clamp(u, ucn, ucp)) ** 2
    if abs(theta0 - math.pi) < 1e-8:
        if nearpi:
            return 0.5 * k * taylor(1.0 + u)
        if near0:
            return 0.5 * k * (torch.sqrt(torch.clamp(taylor(1.0 - u), min=1e-30)) - math.pi) ** 2
        return 0.5 * k * (torch.acos(torch.clamp(u, ucn, ucp)) - math.pi) ** 2
    if near0:
        th = torch.sqrt(torch.clamp(taylor(1.0 - u), min=1e-30))
    elif nearpi:
        th = math.pi - torch.sqrt(torch.clamp(taylor(1.0 + u), min=1e-30))
    else:
        th = torch.acos(u)
    return 0.5 * k * (th - theta0) ** 2


BOHR2ANG = 0.52917721067


def _bias2_energy_torch(geom, kind, f1, f2, k, p, q):
    """calc_energy of one atom pair of LJRepulsivePotentialScale / Value (kind 5: k = eps, p = sigma, both already in
    atomic units; LJ_repulsive_potential.py:42-62,97-114), StructAnharmonicKeepPotential (kind 6;
    anharmonic_keep_potential.py:14-27), WellPotential (kind 7: k = wall energy in Hartree, q = limits in Bohr;
    switching_potential.py:14-67), StructKeepOutofPlainAnglePotential (kind 8: p = phi0 in radians;
    keep_outofplain_angle_potential.py:33-146) and the fragment-centroid restraints StructKeepAnglePotentialv2 /
    StructKeepDihedralAnglePotentialv2 / StructKeepOutofPlainAnglePotentialv2 (kinds 9 - 11), restated on a torch
    tensor (atoms 0-based)."""
    import math
    import torch
    if kind == 5:
        r = torch.linalg.norm(geom[f1[0]] - geom[f2[0]])
        return k * (-2 * (p / r) ** 6 + (p / r) ** 12)
    if kind == 6:
        r = torch.linalg.norm(geom[f1[0]] - geom[f2[0]])
        return q[0] * (1.0 - torch.exp(-math.sqrt(k / (2 * q[0])) * (r - p / BOHR2ANG))) ** 2
    if kind in (7, 12, 13):
        if kind == 7:      # WellPotential, and WellPotentialAround per target atom (switching_potential.py:14-67,172-224)
            r = torch.linalg.norm(geom[list(f1)].sum(dim=0) / len(f1) - geom[list(f2)].sum(dim=0) / len(f2))
        elif kind == 12:   # WellPotentialVP (:121-170): p3 = the point, float32-rounded by the reference
            r = torch.linalg.norm(geom[f1[0]] - torch.tensor(list(q[4:7]), dtype=torch.float64))
        else:              # WellPotentialWall (:69-119): f2[0] = axis
            r = abs(torch.linalg.norm(geom[f1[0]][f2[0]]))
        q = q[:4]
        a, b, c, d = q
        xs = 0.5 / (b - a) * r + (1.0 - 0.5 * b / (b - a))
        xl = 0.5 / (c - d) * r + (1.0 - 0.5 * c / (c - d))
        if r <= a:
            return k * (-3.75 * xs + 2.875)
        if r <= b:
            return k * (2.0 - 20.0 * xs ** 3 + 30.0 * xs ** 4 - 12.0 * xs ** 5)
        if r < c:
            return 0.0 * r
        if r < d:
            return k * (2.0 - 20.0 * xl ** 3 + 30.0 * xl ** 4 - 12.0 * xl ** 5)
        return k * (-3.75 * xl + 2.875)
    if kind in (9, 10, 11):   # fragment-centroid angle / dihedral / out-of-plane angle: q = fragment sizes, f1 = all atoms
        off = np.concatenate([[0], np.cumsum([int(v) for v in q])])
        cen = [geom[list(f1[off[g]:off[g + 1]])].mean(dim=0) for g in range(len(q)) if int(q[g]) > 0]
    if kind == 9:             # StructKeepAnglePotentialv2 (keep_angle_potential.py:293-478): p = theta0 in DEGREES
        th0 = p * (math.pi / 180.0)
        v1, v2 = cen[0] - cen[1], cen[2] - cen[1]
        u = torch.clamp(torch.dot(v1, v2) / torch.clamp(torch.linalg.norm(v1) * torch.linalg.norm(v2), min=1e-12), -1.0, 1.0)
        cut = 1e-3
        ucp, ucn = math.cos(cut), math.cos(math.pi - cut)
        co = [128.0 / 1575.0, 4.0 / 35.0, 8.0 / 45.0, 1.0 / 3.0, 2.0]

        def taylor(delta):
            term = co[0]
            for c in co[1:]:
                term = c + delta * term
            return delta * term

        def quad(th_cut, ucut):
            dth = -1.0 / math.sin(th_cut)
            return 0.5 * k * (th_cut - th0) ** 2 + k * (th_cut - th0) * dth * (u - ucut) + 0.5 * k * dth ** 2 * (u - ucut) ** 2
        if abs(th0) < 1e-8:
            if u > ucp:
                return 0.5 * k * taylor(1.0 - u)
            return quad(math.pi - cut, ucn) if u < ucn else 0.5 * k * torch.acos(torch.clamp(u, -1.0, ucp)) ** 2
        if abs(th0 - math.pi) < 1e-8:
            if u < ucn:
                return 0.5 * k * taylor(1.0 + u)
            return quad(cut, ucp) if u > ucp else 0.5 * k * (torch.acos(torch.clamp(u, ucn, 1.0)) - th0) ** 2
        if u > ucp:
            return quad(cut, ucp)
        return quad(math.pi - cut, ucn) if u < ucn else 0.5 * k * (torch.acos(u) - th0) ** 2
    if kind == 10:            # StructKeepDihedralAnglePotentialv2 (keep_dihedral_angle_potential.py:186-257): p in radians
        b1, b2, b3 = cen[1] - cen[0], cen[2] - cen[1], cen[3] - cen[2]
        n1, n2 = torch.linalg.cross(b1, b2), torch.linalg.cross(b2, b3)
        n1sq, n2sq = torch.sum(n1 ** 2), torch.sum(n2 ** 2)

        def sw(val):
            t = torch.clamp((val - 1e-10) / (1e-8 - 1e-10), 0.0, 1.0)
            return t * t * (3.0 - 2.0 * t)
        n1h = n1 / torch.clamp(torch.sqrt(n1sq), min=1e-12)
        n2h = n2 / torch.clamp(torch.sqrt(n2sq), min=1e-12)
        b2h = b2 / torch.clamp(torch.linalg.norm(b2), min=1e-12)
        x = torch.sum(n1h * n2h)
        y = torch.sum(torch.linalg.cross(n1h, n2h) * b2h)
        diff = torch.atan2(y, x) - p
        diff = diff - 2.0 * math.pi * torch.round(diff / (2.0 * math.pi))
        return 0.5 * k * diff ** 2 * sw(n1sq) * sw(n2sq)
    if kind == 11:
        a1, a2, a3 = cen[1] - cen[0], cen[2] - cen[0], cen[3] - cen[0]
    else:
        ci, i1, i2, i3 = f1
        a1, a2, a3 = geom[i1] - geom[ci], geom[i2] - geom[ci], geom[i3] - geom[ci]
    n = torch.linalg.cross(a2, a3)
    nsq = torch.sum(n ** 2)
    if nsq < 1e-8:
        return 0.0 * nsq
    nh = n / torch.clamp(torch.sqrt(nsq), min=1e-12)
    h = torch.sum(a1 * nh)
    rp = torch.sqrt(torch.clamp(torch.linalg.norm(a1) ** 2 - h ** 2, min=0.0))
    return 0.5 * k * (torch.atan2(h, rp) - p) ** 2


def bias2_egh(coord, terms):
    """(E, grad (N,3), hess (3N,3N)) of a list of (kind, f1, f2, k, p, q) terms by torch.func."""
    import torch
    geom = torch.tensor(np.asarray(coord, float), dtype=torch.float64)

    def f(x):
        e = 0.0
        for kind, f1, f2, k, p, q in terms:
            e = e + _bias2_energy_torch(x, kind, f1, f2, k, p, q)
        return e
    E = f(geom)
    g = torch.func.jacrev(f)(geom)
    H = torch.func.hessian(f)(geom).reshape(geom.numel(), geom.numel())
    return float(E), g.numpy(), H.numpy()


def keep_egh(coord, kind, f1, f2, k, p):
    """(E, grad (N,3), hess (3N,3N)) by torch.func, as Potential/potential.py:127-137 does."""
    import torch
    geom = torch.tensor(np.asarray(coord, float), dtype=torch.float64)
    f = lambda x: _keep_energy_torch(x, kind, f1, f2, k, p)
    E = f(geom)
    g = torch.func.jacrev(f)(geom)
    H = torch.func.hessian(f)(geom).reshape(geom.numel(), geom.numel())
    return float(E), g.numpy(), H.numpy()


def path_length_list(X):
    """Running path length of a chain X (M, N, 3), centroid-free (Utils/calc_tools.py:853-862)."""
    pl = [0.0]
    for i in range(len(X) - 1):
        a = X[i + 1] - np.mean(X[i + 1], axis=0)
        b = X[i] - np.mean(X[i], axis=0)
        pl.append(pl[-1] + np.linalg.norm(a - b))
    return np.array(pl)


def distribute_geometry(X):
    """Images at equal arc length on the piecewise-linear path (Interpolation/linear_interpolation.py:308-336)."""
    X = np.asarray(X, dtype=np.float64)
    M = len(X)
    pl = path_length_list(X)
    total = pl[-1]
    if total < 1e-8:
        return X.copy()
    node = total / (M - 1)
    out = [X[0]]
    for i in range(1, M - 1):
        dist = i * node
        hit = [j for j in range(M - 1) if pl[j] <= dist <= pl[j + 1]]
        if hit:
            j = hit[0]
            with np.errstate(divide="ignore", invalid="ignore"):
                dt = (dist - pl[j]) / (pl[j + 1] - pl[j])
            out.append(X[j] + (X[j + 1] - X[j]) * dt)
        else:
            out.append(X[-1])
    out.append(X[-1])
    return np.array(out)


def constraint_null_space(rows, svd_threshold=1e-5):
    """CRSIRFO._get_null_space_basis (Optimizer/crsirfo.py:16-45) from the raw constraint rows (k, n)."""
    import scipy.linalg
    rows = np.asarray(rows, float)
    n = rows.shape[1]
    if len(rows) == 0:
        return np.eye(n)
    nr = np.linalg.norm(rows, axis=1)
    nr[nr < 1e-12] = 1.0
    Bn = rows / nr[:, None]
    U, S, _ = scipy.linalg.svd(Bn.T, full_matrices=True)
    max_s = S[0] if len(S) > 0 else 1.0
    rank = int(np.sum(S > max(svd_threshold, max_s * 1e-6)))
    return U[:, rank:]


class CRSIRFOOracle(RSIRFOOracle):
    """CRSIRFO.run (Optimizer/crsirfo.py:47-170) for one structure: RS-I-RFO in the null space of the constraint
    rows.  ``rows`` (k, n) = constraints_obj._get_all_constraint_vectors(x), ``shake`` (n,) = the displacement of
    constraints_obj.adjust_init_coord (x is the CORRECTED geometry)."""

    def __init__(self, *a, gradient_norm_threshold=1e-4, svd_threshold=1e-5, **kw):
        super().__init__(*a, **kw)
        self.gradient_norm_threshold = gradient_norm_threshold
        self.svd_threshold = svd_threshold
        self.converged_sub = False

    def run(self, x, Bg, g, x_prev=None, g_prev=None, Be=0.0, rows=None, shake=None):
        x = np.asarray(x, float).ravel()
        gfull = np.array(Bg, float).ravel()
        g = np.asarray(g, float).ravel()
        info = {"updated": False, "alpha_search": False}
        if shake is not None and np.linalg.norm(shake) > 1e-6:            # (:70-80) H_eff aliases self.hessian
            if self.bias_hessian is not None:
                self.hessian += self.bias_hessian
            gfull = gfull + self.hessian @ np.asarray(shake, float).ravel()
        if self.have_prev and x_prev is not None and g_prev is not None and len(x_prev) > 0 and len(g_prev) > 0:
            self.hessian, info["updated"] = rsirfo_update_hessian(
                self.hessian, x, g, np.asarray(x_prev, float).ravel(), np.asarray(g_prev, float).ravel(), self.method_id)
        if self.bias_hessian is not None:                                   # (:88-90) in place as well
            self.hessian += self.bias_hessian
        U = constraint_null_space(rows if rows is not None else np.zeros((0, x.size)), self.svd_threshold)
        gs = U.T @ gfull
        Hs = U.T @ (self.hessian @ U)
        gnorm = np.linalg.norm(gs)
        self.converged_sub = False
        if gnorm < self.gradient_norm_threshold:                            # (:108-118)
            self.converged_sub = True
            self.have_prev = True
            self.prev_energy = Be
            self.last = dict(info, eigvals=None, pred=None, trust=self.trust_radius)
            return np.zeros_like(gfull)
        Hs = 0.5 * (Hs + Hs.T)
        lam, V, _ = eigh_with_shift(Hs)
        if self.prev_energy is not None:                                    # (:126-141)
            actual = Be - self.prev_energy
            if len(self.act) >= 3:
                self.act.pop(0)
            self.act.append(actual)
            if self.pred:
                self.trust_radius = adjust_trust_radius(self.trust_radius, actual, self.pred[-1], lam[0], gnorm,
                                                        self.saddle_order, self.trust_radius_min, self.trust_radius_max)
        m = lam.size
        P = np.eye(m)
        found = i = 0
        while found < self.saddle_order and i < m:
            if abs(lam[i]) > 1e-10:
                P = P - (1.0 if self.NEB_mode else 2.0) * np.outer(V[:, i], V[:, i])
                found += 1
            i += 1
        Hstar = P @ Hs
        Hstar = 0.5 * (Hstar + Hstar.T)
        gstar = P @ gs
        lam_s, V_s, _ = eigh_with_shift(Hstar)
        keep = ~(np.abs(lam_s) < 1e-6)
        step_sub, info["alpha_search"] = rs_step(lam_s[keep], V_s[:, keep], gstar, self.trust_radius)
        step = U @ step_sub
        pred = gs @ step_sub + 0.5 * (step_sub @ Hs @ step_sub)
        if len(self.pred) >= 3:
            self.pred.pop(0)
        self.pred.append(pred)
        self.have_prev = True
        self.prev_energy = Be
        self.iteration += 1
        self.last = dict(info, eigvals=lam, pred=pred, trust=self.trust_radius)
        return -step


PAULING_EN = {'H': 2.20, 'He': 0.00, 'Li': 0.98, 'Be': 1.57, 'B': 2.04, 'C': 2.55, 'N': 3.04, 'O': 3.44, 'F': 3.98, 'Ne': 0.00,
              'Na': 0.93, 'Mg': 1.31, 'Al': 1.61, 'Si': 1.90, 'P': 2.19, 'S': 2.58, 'Cl': 3.16, 'Ar': 0.00, 'K': 0.82,
              'Ca': 1.00, 'Sc': 1.36, 'Ti': 1.54, 'V': 1.63, 'Cr': 1.66, 'Mn': 1.55, 'Fe': 1.83, 'Co': 1.88, 'Ni': 1.91,
              'Cu': 1.90, 'Zn': 1.65, 'Ga': 1.81, 'Ge': 2.01, 'As': 2.18, 'Se': 2.55, 'Br': 2.96, 'Kr': 0.00}


def sr_charges(elems):
    """estimate_atomic_charges (ModelHessian/shortrange.py:148-184): 0.2 (mean electronegativity - electronegativity)."""
    en = [PAULING_EN.get(e, 2.0) for e in elems]
    avg = sum(en) / len(en)
    return np.array([0.2 * (avg - v) for v in en])


def sr_hessian(H, xyz, elems, radii, omega=0.2, cx_sr=0.78, scaling=0.5, cutoff=15.0):
    """ShortRangeCorrectionHessian.main (ModelHessian/shortrange.py:186-346): short-range Coulomb second derivatives of
    the non-bonded pairs, TR/ROT-projected, added to H, symmetrised.  radii = covalent radii (Bohr)."""
    from scipy.special import erf
    xyz = np.asarray(xyz, float)
    N = len(xyz)
    q = sr_charges(elems)
    C = np.zeros((3 * N, 3 * N))
    for i in range(N):
        for j in range(i + 1, N):
            if np.linalg.norm(xyz[j] - xyz[i]) <= (radii[j] + radii[i]) * 1.1:
                continue
            rv = xyz[j] - xyz[i]
            r = np.linalg.norm(rv)
            if r > cutoff:
                continue
            ef, ex = erf(omega * r), np.exp(-(omega * r) ** 2)
            d1 = 2 * omega * ex / (np.sqrt(np.pi) * r) + (ef - 1.0) / r ** 2
            d2 = 2 * (2 * ef - 1) / r ** 3 + 4 * omega * (ex / np.sqrt(np.pi)) / r ** 2 + 2 * omega ** 3 * (ex / np.sqrt(np.pi))
            u = rv / r
            blk = q[i] * q[j] * cx_sr * scaling * (d2 * np.outer(u, u) + d1 / r * (np.eye(3) - np.outer(u, u)))
            C[3 * i:3 * i + 3, 3 * i:3 * i + 3] += blk
            C[3 * j:3 * j + 3, 3 * j:3 * j + 3] += blk
            C[3 * i:3 * i + 3, 3 * j:3 * j + 3] -= blk
            C[3 * j:3 * j + 3, 3 * i:3 * i + 3] -= blk
    out = H + project_hessian_trrot(C, xyz.reshape(-1))
    return 0.5 * (out + out.T)
