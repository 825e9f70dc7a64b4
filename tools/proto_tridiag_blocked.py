"""NumPy model of k_tridiag_blk (csrc/tridiag_blocked.cu): blocked Householder tridiagonalisation (LAPACK dlatrd
panels) in which the symv of every column runs on the PANEL-START matrix, the reflector scalars join afterwards by
linearity and all per-column scalar products travel through ONE reduction.  Development tool: checks the algebra
against numpy.linalg.eigvalsh and an unblocked dsytd2 reference (not part of the product, not a test oracle)."""
import numpy as np


def tridiag_blocked(A, g, NB=6):
    n = A.shape[0]
    L = np.tril(A).copy()          # lower triangle, panel-start values; eliminated columns hold the reflectors
    Wp = np.zeros((n, NB))
    d = np.zeros(n); e = np.zeros(n); tau = np.zeros(n)
    gq = g.copy()
    Vh = np.zeros((n, n))
    sym = lambda i, j: L[i, j] if i >= j else L[j, i]
    uu = np.zeros(n)
    uu[1:] = L[1:, 0]
    d[0] = L[0, 0]
    k0 = 0
    for k in range(n - 2):
        jj = k - k0
        V = L[:, k0:k0 + jj]       # V(i, l) (unit entries stored explicitly)
        W = Wp[:, :jj]
        u = uu.copy(); u[:k + 1] = 0.0
        # (a) c = updated column k+1, rows >= k+1
        c = np.zeros(n)
        for i in range(k + 1, n):
            c[i] = L[i, k + 1] - V[i] @ W[k + 1] - W[i] @ V[k + 1]
        VTu = V[k + 1:].T @ u[k + 1:]; WTu = W[k + 1:].T @ u[k + 1:]
        # (b) z = A_panelstart u over rows / cols > k
        z = np.zeros(n)
        for i in range(k + 1, n):
            z[i] = sum(sym(i, j) * u[j] for j in range(k + 1, n))
        # (c) one reduction
        S1p = z[k + 2:] @ u[k + 2:]; S2 = c[k + 2:] @ u[k + 2:]; S3 = u[k + 2:] @ gq[k + 2:]; xn2 = u[k + 2:] @ u[k + 2:]
        alpha = u[k + 1]
        beta, tk, s = alpha, 0.0, 0.0
        if xn2 > 0.0:
            beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
            tk = (beta - alpha) / beta
            s = 1.0 / (alpha - beta)
        ca = 1.0 - s * alpha
        # corrected q = A^(jj) u
        q = z - V @ WTu - W @ VTu
        S1 = S1p - sum(WTu[l] * (VTu[l] - alpha * V[k + 1, l]) + VTu[l] * (WTu[l] - alpha * W[k + 1, l]) for l in range(jj))
        assert abs(S1 - q[k + 2:] @ u[k + 2:]) <= 1e-12 * (abs(S1) + 1)
        p0 = tk * (s * q[k + 1] + ca * c[k + 1])
        pv = p0 + tk * s * (s * S1 + ca * S2)
        vg = gq[k + 1] + s * S3
        a2 = -0.5 * tk * pv
        w0 = p0 + a2
        v = s * u; v[k + 1] = 1.0; v[:k + 1] = 0.0
        p = tk * (s * q + ca * c); p[:k + 1] = 0.0
        w = p + a2 * v
        assert abs(w[k + 1] - w0) < 1e-13 * (abs(w0) + 1)
        un = c - v * w0 - w
        # (d) publish
        Wp[k + 1:, jj] = w[k + 1:]
        L[k + 1:, k] = v[k + 1:]
        Vh[k, k + 1:] = v[k + 1:]
        gq[k + 1:] -= tk * vg * v[k + 1:]
        e[k] = beta; tau[k] = tk; d[k + 1] = un[k + 1] + 0.0   # un[k+1] = c - 2 w0
        uu = un.copy(); uu[k + 1] = 0.0
        uu[:k + 2] = 0.0
        if jj == NB - 1 and k + 1 < n - 2 or False:
            # trailing update with the whole panel (DMMA tiles in the kernel)
            kn = k + 1
            V = L[:, k0:k0 + NB]; W = Wp[:, :NB]
            for i in range(kn, n):
                for j in range(kn, i + 1):
                    L[i, j] -= V[i] @ W[j] + W[i] @ V[j]
            k0 = kn
    # tail: e_{n-2} = raw column n-2, d_{n-1} = fully updated last diagonal element
    jj = (n - 2) - k0
    e[n - 2] = uu[n - 1]
    d[n - 1] = L[n - 1, n - 1] - 2.0 * (L[n - 1, k0:k0 + jj] @ Wp[n - 1, :jj])
    return d, e, tau, Vh, gq


def dsytd2(A, g):
    n = A.shape[0]
    A = A.copy(); gq = g.copy()
    d = np.zeros(n); e = np.zeros(n); tau = np.zeros(n); Vh = np.zeros((n, n))
    for k in range(n - 2):
        x = A[k + 1:, k].copy()
        alpha = x[0]; xn2 = x[1:] @ x[1:]
        beta, tk = alpha, 0.0
        v = np.zeros_like(x); v[0] = 1.0
        if xn2 > 0:
            beta = -np.copysign(np.sqrt(alpha * alpha + xn2), alpha)
            tk = (beta - alpha) / beta
            v[1:] = x[1:] / (alpha - beta)
        p = tk * (A[k + 1:, k + 1:] @ v)
        w = p - 0.5 * tk * (p @ v) * v
        A[k + 1:, k + 1:] -= np.outer(v, w) + np.outer(w, v)
        gq[k + 1:] -= tk * (v @ gq[k + 1:]) * v
        d[k] = A[k, k]; e[k] = beta; tau[k] = tk; Vh[k, k + 1:] = v
    d[n - 2] = A[n - 2, n - 2]; e[n - 2] = A[n - 1, n - 2]; d[n - 1] = A[n - 1, n - 1]
    return d, e, tau, Vh, gq


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in (3, 4, 7, 12, 13, 33, 60):
        for NB in (4, 6, 8):
            M = rng.standard_normal((n, n)); A = M + M.T
            g = rng.standard_normal(n)
            r1 = tridiag_blocked(A, g, NB); r2 = dsytd2(A, g)
            errs = [np.abs(a - b).max() for a, b in zip(r1, r2)]
            T = np.diag(r1[0]) + np.diag(r1[1][:-1], 1) + np.diag(r1[1][:-1], -1)
            ev = np.abs(np.linalg.eigvalsh(T) - np.linalg.eigvalsh(A)).max()
            print(n, NB, " ".join(f"{x:.1e}" for x in errs), f"eig {ev:.1e}")
            assert max(errs) < 1e-11 and ev < 1e-12
