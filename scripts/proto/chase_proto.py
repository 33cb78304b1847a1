"""Prototype of band_tridiag.cuh's schedule: lower band in L[r][jj] = B[r][r-jj], 8 groups per warp in
lock-step ticks, sweep k step j touches blocks (a),(b),(c) only.  Development aid."""
import numpy as np
from sbr_proto import sbr_band, house

B4 = 4


def chase_step(L, n, k, j):
    r0 = k + 1 + 4 * j
    xc = r0 - 1 if j == 0 else r0 - 4
    x = np.array([L[r0 + i, r0 + i - xc] for i in range(4)])
    beta, tau, vs = house(x[0], float(x[1:] @ x[1:]))
    if tau == 0.0:
        return
    v = x * vs
    v[0] = 1.0
    # (a)
    newL = {}
    for q in range(4):
        c = r0 - 4 + q
        if c < 0:
            continue
        m = np.array([L[r0 + i, 4 + i - q] for i in range(4)])
        s = m @ v
        m = m - tau * s * v
        if c == xc:
            m = np.array([beta, 0, 0, 0.0])
        for i in range(4):
            newL[(r0 + i, 4 + i - q)] = m[i]
    # (b)
    D = np.zeros((4, 4))
    for i in range(4):
        for c in range(i + 1):
            D[i, c] = D[c, i] = L[r0 + i, i - c]
    p = tau * D @ v
    K = 0.5 * tau * (v @ p)
    w = p - K * v
    for q in range(4):
        for c in range(q + 1):
            newL[(r0 + q, q - c)] = D[q, c] - v[q] * w[c] - w[q] * v[c]
    # (c)
    for q in range(4):
        r = r0 + 4 + q
        m = np.array([L[r, 4 + q - c] for c in range(4)])
        s = m @ v
        m = m - tau * s * v
        for c in range(4):
            newL[(r, 4 + q - c)] = m[c]
    for (r, jj), val in newL.items():
        L[r, jj] = val


def band_to_tridiag_pipelined(B, n):
    L = np.zeros((n + 12, 8))
    for r in range(n):
        for jj in range(5):
            if r - jj >= 0:
                L[r, jj] = B[r, r - jj]
    G = 8
    k = list(range(G))
    j = [0] * G
    active = [g <= n - 3 for g in range(G)]
    ticks = 0
    while any(active):
        ticks += 1
        snap = [(k[g], j[g], active[g]) for g in range(G)]
        for g in range(G):
            if not active[g]:
                continue
            kp, jp, ap = snap[(g + 7) % G]
            ok = k[g] == 0 or (not ap) or kp > k[g] - 1 or (kp == k[g] - 1 and jp >= j[g] + 4)
            if not ok:
                continue
            chase_step(L, n, k[g], j[g])
            j[g] += 1
            if k[g] + 1 + 4 * j[g] > n - 2:
                k[g] += G
                j[g] = 0
                if k[g] > n - 3:
                    active[g] = False
    d = L[:n, 0].copy()
    e = L[1:n, 1].copy()
    return d, e, ticks


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for n in [3, 4, 5, 6, 7, 8, 9, 13, 32, 50, 96, 192]:
        Wm = rng.standard_normal((n, 2 * n)) * 0.02
        G = Wm @ Wm.T
        ref = np.linalg.eigvalsh(G)
        B = sbr_band(G, 4)
        d, e, ticks = band_to_tridiag_pipelined(B, n)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        et = np.linalg.eigvalsh(T)
        print(n, "tri err", np.abs(et - ref).max() / ref.max(), "ticks", ticks)
