"""NumPy prototype of the fast symmetric eigensolver (design study for
csrc/eigh_tridiag.cu): Householder tridiagonalisation, block splitting, Sturm
bisection with the division-free p-recurrence, twisted-factorisation
eigenvectors, cluster re-orthogonalisation, fallback detection.
Not used by the product or the tests; run directly to see accuracy statistics."""
import sys
import numpy as np

EPS = 2.0 ** -52


def householder_tridiag(A):
    """dsytd2-like (lower).  Returns d, e, list of (v, tau) reflectors (v[0]=1 at row k+1)."""
    A = A.copy()
    n = A.shape[0]
    refl = []
    for k in range(n - 2):
        x = A[k + 1:, k].copy()
        alpha = x[0]
        xnorm = np.linalg.norm(x[1:])
        if xnorm == 0.0:
            refl.append((None, 0.0))
            continue
        beta = -np.copysign(np.hypot(alpha, xnorm), alpha)
        tau = (beta - alpha) / beta
        v = x / (alpha - beta)
        v[0] = 1.0
        A22 = A[k + 1:, k + 1:]
        p = tau * (A22 @ v)
        w = p - 0.5 * tau * (p @ v) * v
        A22 -= np.outer(v, w) + np.outer(w, v)
        A[k + 1, k] = beta
        A[k + 2:, k] = 0
        A[k, k + 1] = beta
        A[k, k + 2:] = 0
        refl.append((v, tau))
    d = np.diag(A).copy()
    e = np.diag(A, -1).copy()
    return d, e, refl


def apply_Q(refl, Y, n):
    """Y <- Q Y with Q = H_0 H_1 ... ."""
    Y = Y.copy()
    for k in reversed(range(len(refl))):
        v, tau = refl[k]
        if v is None:
            continue
        sub = Y[k + 1:]
        sub -= tau * np.outer(v, v @ sub)
    return Y


def split_blocks(d, e):
    n = d.size
    e = e.copy()
    starts = [0]
    tnorm = max(np.abs(d).max(), np.abs(e).max() if e.size else 0.0)
    for k in range(n - 1):
        if abs(e[k]) <= EPS * (abs(d[k]) + abs(d[k + 1])) or abs(e[k]) <= EPS * tnorm * 1e-3 * 0 + 1e-300:
            e[k] = 0.0
            starts.append(k + 1)
    return e, starts + [n]


def sturm_count_p(d, e2, x):
    """# eigenvalues < x via sign changes of the (rescaled) p-recurrence."""
    pm1 = 1.0
    p = d[0] - x
    cnt = 1 if p < 0 else 0
    if p == 0:
        p = -1e-300; cnt = 1
    for k in range(1, d.size):
        pn = (d[k] - x) * p - e2[k - 1] * pm1
        pm1 = p
        p = pn
        if p == 0.0:
            p = -np.sign(pm1) * 1e-300 * abs(pm1) if pm1 != 0 else -1e-300
        if (p < 0) != (pm1 < 0):
            cnt += 1
        a = abs(p)
        if a > 1e100 or a < 1e-100:
            s = 1.0 / a
            p *= s; pm1 *= s
    return cnt


def bisect_block(d, e):
    n = d.size
    e2 = e * e
    ea = np.abs(np.concatenate([[0.0], e])) + np.abs(np.concatenate([e, [0.0]]))
    gl = (d - ea).min(); gu = (d + ea).max()
    tn = max(abs(gl), abs(gu))
    gl -= 2 * tn * EPS * n + 1e-300; gu += 2 * tn * EPS * n + 1e-300
    lam = np.empty(n)
    for i in range(n):
        lo, hi = gl, gu
        for it in range(200):
            mid = 0.5 * (lo + hi)
            if mid <= lo or mid >= hi:
                break
            if sturm_count_p(d, e2, mid) >= i + 1:
                hi = mid
            else:
                lo = mid
            if hi - lo <= 2 * EPS * max(abs(lo), abs(hi)) + 1e-300:
                break
        lam[i] = 0.5 * (lo + hi)
    return lam


def twisted_vector(d, e, lam):
    n = d.size
    if n == 1:
        return np.ones(1), 0.0
    a = d - lam
    piv = EPS * max(np.abs(d).max(), np.abs(e).max()) * 1e-3 + 1e-300
    Dp = np.empty(n); Dm = np.empty(n)
    Dp[0] = a[0]
    for k in range(1, n):
        q = Dp[k - 1]
        if abs(q) < piv:
            q = -piv if q <= 0 else piv
            Dp[k - 1] = q
        Dp[k] = a[k] - e[k - 1] ** 2 / q
    Dm[n - 1] = a[n - 1]
    for k in range(n - 2, -1, -1):
        q = Dm[k + 1]
        if abs(q) < piv:
            q = -piv if q <= 0 else piv
            Dm[k + 1] = q
        Dm[k] = a[k] - e[k] ** 2 / q
    gam = Dp + Dm - a
    r = int(np.argmin(np.abs(gam)))
    z = np.zeros(n)
    z[r] = 1.0
    for k in range(r - 1, -1, -1):
        q = Dp[k]
        if abs(q) < piv:
            q = -piv if q <= 0 else piv
        z[k] = -(e[k] / q) * z[k + 1]
    for k in range(r + 1, n):
        q = Dm[k]
        if abs(q) < piv:
            q = -piv if q <= 0 else piv
        z[k] = -(e[k - 1] / q) * z[k - 1]
    nz = np.linalg.norm(z)
    return z / nz, abs(gam[r]) / nz


def tridiag_eigh(d, e, gaptol_rel=1e-3, verbose=False):
    """Returns lam (sorted), Z (columns), info dict."""
    n = d.size
    e_s, bounds = split_blocks(d, e)
    tnorm = max(np.abs(d).max(), np.abs(e_s).max() if e_s.size else 0.0, 1e-300)
    lam_all = np.empty(n); Z = np.zeros((n, n))
    fallback = False
    worst_cancel = 1.0
    ncl = 0
    for b in range(len(bounds) - 1):
        s, t = bounds[b], bounds[b + 1]
        db, eb = d[s:t], e_s[s:t - 1]
        lam = bisect_block(db, eb)
        m = t - s
        Zb = np.zeros((m, m))
        for i in range(m):
            Zb[:, i], _ = twisted_vector(db, eb, lam[i])
        # clusters
        i = 0
        while i < m:
            j = i
            while j + 1 < m and lam[j + 1] - lam[j] < gaptol_rel * tnorm:
                j += 1
            if j > i:
                ncl += 1
                for c in range(i + 1, j + 1):
                    v = Zb[:, c]
                    for rep in range(2):
                        n0 = np.linalg.norm(v)
                        for p_ in range(i, c):
                            v = v - (Zb[:, p_] @ v) * Zb[:, p_]
                        n1 = np.linalg.norm(v)
                        if rep == 0:
                            worst_cancel = min(worst_cancel, n1 / n0)
                        if n1 > 0.7 * n0:
                            break
                    if n1 < 1e-3:
                        fallback = True
                    Zb[:, c] = v / n1 if n1 > 0 else v
            i = j + 1
        lam_all[s:t] = lam
        Z[s:t, s:t] = Zb
    order = np.argsort(lam_all, kind="stable")
    return lam_all[order], Z[:, order], dict(fallback=fallback, blocks=len(bounds) - 1, clusters=ncl,
                                             worst_cancel=worst_cancel)


def full_eigh(A, **kw):
    d, e, refl = householder_tridiag(A)
    lam, Z, info = tridiag_eigh(d, e, **kw)
    V = apply_Q(refl, Z, A.shape[0])
    return lam, V, info


def report(name, A):
    n = A.shape[0]
    lam, V, info = full_eigh(A)
    ref = np.linalg.eigvalsh(A)
    sc = max(np.abs(ref).max(), 1e-300)
    eerr = np.abs(lam - ref).max() / sc
    orth = np.abs(V.T @ V - np.eye(n)).max()
    res = np.abs(A @ V - V * lam).max() / sc
    print(f"{name:34s} n={n:4d} eval {eerr:.1e} orth {orth:.1e} resid {res:.1e} {info}")
    return eerr, orth, res, info


if __name__ == "__main__":
    sys.path.insert(0, ".")
    from multioptpy_b200 import synthetic
    from oracle import np_oracle as O
    rng = np.random.default_rng(0)
    for n in (24, 72, 150):
        A = rng.standard_normal((n, n)); A = 0.5 * (A + A.T)
        report("random symmetric", A)
        H = synthetic.spd_hessian(n, rng)
        report("spd hessian", H)
        x = synthetic.grid_geometry(n // 3, rng).reshape(-1)
        report("TR/ROT projected hessian", O.project_hessian_trrot(H, x))
        w, Vv = np.linalg.eigh(H); w[: n // 2] = 0.5
        D = (Vv * w) @ Vv.T
        report("half spectrum degenerate", 0.5 * (D + D.T))
        w2 = np.sort(rng.uniform(0, 1, n)); w2[5] = w2[4] + 1e-13; w2[9] = w2[8] + 1e-9; w2[12] = w2[11] + 1e-6
        D = (Vv * w2) @ Vv.T
        report("tight pairs 1e-13/1e-9/1e-6", 0.5 * (D + D.T))
    report("diagonal", np.diag(np.arange(30.0)))
    W = np.diag(np.abs(np.arange(-10, 11)).astype(float)) + np.diag(np.ones(20), 1) + np.diag(np.ones(20), -1)
    report("Wilkinson W21+", W)
    report("zeros", np.zeros((12, 12)))
    # symmetric molecule-like: ring of identical springs (circulant -> exact 2-fold degeneracies)
    m = 24
    C = np.zeros((m, m))
    for i in range(m):
        C[i, i] = 2.0; C[i, (i + 1) % m] = -1.0; C[i, (i - 1) % m] = -1.0
    report("circulant ring (2-fold degenerate)", C)
    K3 = np.kron(C, np.eye(3)) + 0.1 * np.kron(np.eye(m), np.ones((3, 3)))
    report("ring x 3 (6-fold structure)", K3)
