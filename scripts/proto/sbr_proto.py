"""NumPy prototype of the two-stage eigensolve (stage 2a of DESIGN.md):
   (1) sbr_band: full symmetric -> band (bandwidth b) by bottom-up blocked Householder (compact WY,
       deferred rank-2b update),  (2) band_to_tridiag: Householder bulge chasing.
Mirrors the conventions of csrc/sbr_band.cuh / csrc/band_tridiag.cuh; used to pin the algebra before
the CUDA versions were written (development aid, not part of the product or the tests' oracle)."""
import numpy as np


def house(alpha, xnorm2):
    """reflector for (x[0:pc], alpha): returns beta, tau, vscale  (u = x*vscale, u[pc] = 1)"""
    if xnorm2 <= 0.0:
        return alpha, 0.0, 0.0
    nrm = np.sqrt(alpha * alpha + xnorm2)
    beta = -np.copysign(nrm, alpha)
    tau = (beta - alpha) / beta
    vscale = 1.0 / (alpha - beta)
    return beta, tau, vscale


def sbr_band(G, b):
    n = G.shape[0]
    A = np.tril(G).copy()          # lower triangle is the storage
    B = np.zeros_like(A)           # band output (lower)
    V = W = None                   # pending update, [k][r] layout: shape (b, n)
    m = n

    def updated_row(r):            # row r (cols <= r) with the pending update applied
        row = A[r, : r + 1].copy()
        if V is not None:
            row -= V[:, r] @ W[:, : r + 1] + W[:, r] @ V[:, : r + 1]
        return row

    while m >= b + 2:
        p0 = m - b
        # (1) mini-pass: panel rows
        P = np.zeros((b, p0))
        for t in range(b):
            row = updated_row(p0 + t)
            P[t] = row[:p0]
            B[p0 + t, p0 : p0 + t + 1] = row[p0:]          # diagonal block: final
        # (2) LQ of the panel, rows t = b-1 .. 0;  k = application order
        U = np.zeros((b, n))
        tau = np.zeros(b)
        for k in range(b):
            t = b - 1 - k
            pc = p0 - b + t
            if pc < 0:
                continue
            x = P[t, :pc]
            alpha = P[t, pc]
            beta, tk, vs = house(alpha, float(x @ x))
            tau[k] = tk
            if tk != 0.0:
                U[k, :pc] = x * vs
                U[k, pc] = 1.0
                for t2 in range(t):                        # apply to the rows above
                    s = P[t2, : pc + 1] @ U[k, : pc + 1]
                    P[t2, : pc + 1] -= tk * s * U[k, : pc + 1]
            P[t, :pc] = 0.0
            P[t, pc] = beta
        for t in range(b):                                 # R block -> band
            lo = max(p0 - b + t, 0)
            B[p0 + t, lo:p0] = P[t, lo:p0]
        # T (forward recurrence, application order)
        T = np.zeros((b, b))
        for k in range(b):
            T[k, k] = tau[k]
            if k > 0:
                z = U[:k, :] @ U[k, :]
                T[:k, k] = -tau[k] * (T[:k, :k] @ z)
        # (3) fused pass: apply pending update to the trailing block, Y = A U^T
        for r in range(p0):
            A[r, : r + 1] = updated_row(r)
        As = A[:p0, :p0]
        Af = As + np.tril(As, -1).T
        Y = (Af @ U[:, :p0].T).T                           # [k][r]
        # (4) X = Y^T T ; S = T^T U X ; W = X - 1/2 U S
        X = Y.T @ T                                        # p0 x b
        S = T.T @ (U[:, :p0] @ X)                          # b x b
        Wn = X - 0.5 * U[:, :p0].T @ S
        V = U.copy()
        W = np.zeros((b, n))
        W[:, :p0] = Wn.T
        m = p0
    for r in range(m):
        B[r, : r + 1] = updated_row(r)
    return B


def band_to_tridiag(B, b):
    """B: lower band (n x n, zeros outside).  Householder bulge chasing with zero padding."""
    n = B.shape[0]
    N = n + 3 * b
    M = np.zeros((N, N))
    M[:n, :n] = B + np.tril(B, -1).T
    for k in range(n - 2):
        j = 0
        while True:
            r0 = k + 1 + j * b
            if r0 > n - 2:
                break
            c0 = k if j == 0 else r0 - b
            x = M[r0 : r0 + b, c0].copy()
            beta, tau, vs = house(x[0], float(x[1:] @ x[1:]))
            if tau != 0.0:
                v = x * vs
                v[0] = 1.0
                H = np.eye(b) - tau * np.outer(v, v)
                M[r0 : r0 + b, :] = H @ M[r0 : r0 + b, :]
                M[:, r0 : r0 + b] = M[:, r0 : r0 + b] @ H
            j += 1
    d = np.diag(M)[:n].copy()
    e = np.diag(M, -1)[: n - 1].copy()
    off = M[:n, :n] - np.diag(d) - np.diag(e, 1) - np.diag(e, -1)
    return d, e, np.abs(off).max()


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, b in [(8, 4), (13, 4), (32, 4), (50, 4), (96, 4), (192, 4), (33, 2), (70, 8), (5, 4), (6, 4), (7, 4)]:
        Wm = rng.standard_normal((n, 3 * n)) * 0.02
        G = Wm @ Wm.T
        ref = np.linalg.eigvalsh(G)
        B = sbr_band(G, b)
        assert np.abs(np.tril(B, -b - 1)).max() == 0.0
        eb = np.linalg.eigvalsh(B + np.tril(B, -1).T)
        d, e, off = band_to_tridiag(B, b)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        et = np.linalg.eigvalsh(T)
        print(n, b, "band err", np.abs(eb - ref).max() / ref.max(), "tri err", np.abs(et - ref).max() / ref.max(), "off", off)
