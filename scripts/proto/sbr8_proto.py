"""NumPy prototype of csrc/sbr8.cuh + csrc/chase8.cuh (development aid: pins the algebra and the index conventions
before the CUDA versions; not part of the product or the oracle).

Conventions mirrored from the kernels:
  * internal order N = round_up(n, 8); the matrix sits at the BOTTOM-RIGHT of the N x N frame (off = N - n zero
    rows/columns at the top-left, which stay zero and decouple);
  * elimination bottom-up, panel = tile row Ip = p0/8 (rows p0..p0+7), p0 = m - 8, while m >= 16;
  * the panel buffer holds the panel REVERSED: buffer row k = panel row t = 7 - k, so that reflector k (application
    order) is built from buffer row k and overwrites it (LQ in place: P becomes U);
  * pivot column of buffer row k: pc = p0 - 1 - k;
  * band output Bd[r][j] = A[r][r - j], j = 0..8.
"""
import numpy as np

B = 8


def house(alpha, xnorm2):
    if xnorm2 <= 1e-280:
        return alpha, 0.0, 0.0
    nrm = np.sqrt(alpha * alpha + xnorm2)
    beta = -np.copysign(nrm, alpha)
    tau = (beta - alpha) / beta
    vscale = 1.0 / (alpha - beta)
    return beta, tau, vscale


def sbr8(G):
    n = G.shape[0]
    N = (n + 7) // 8 * 8
    off = N - n
    A = np.zeros((N, N))
    A[off:, off:] = G  # full symmetric working copy (the kernel stores the lower tiles, diagonal tiles full)
    Bd = np.zeros((N, 9))
    V = W = None  # [8][N]
    m = N
    while m >= 16:
        p0 = m - 8
        # (1) mini-pass: panel rows with the pending update
        rows = A[p0:m, :m].copy()
        if V is not None:
            rows -= V[:, p0:m].T @ W[:, :m] + W[:, p0:m].T @ V[:, :m]
        for g in range(8):  # diagonal tile -> band
            for c in range(g + 1):
                Bd[p0 + g, g - c] = rows[g, p0 + c]
        P = rows[::-1, :p0].copy()  # buffer row k = panel row 7 - k
        # (2a) LQ in place
        tau = np.zeros(8)
        zz = np.zeros((8, 8))
        for k in range(8):
            t = 7 - k
            pc = p0 - 1 - k
            x = P[k, :pc].copy()
            alpha = P[k, pc]
            beta, tk, vs = house(alpha, float(x @ x))
            tau[k] = tk
            # band: beta at pc (j = 8), entries right of the pivot are final
            r = p0 + t
            Bd[r, r - pc] = beta if tk != 0.0 else alpha
            for c in range(pc + 1, p0):
                Bd[r, r - c] = P[k, c]
            dots = P[:, :pc] @ x  # Gram row k restricted to columns < pc
            for k2 in range(k + 1, 8):  # rows still to be reduced
                coef = tk * (vs * dots[k2] + P[k2, pc])
                P[k2, :pc] -= coef * x * vs
                P[k2, pc] -= coef
            for j in range(k):  # u_j . u_k for T
                zz[k, j] = (vs * dots[j] + P[j, pc]) if tk != 0.0 else 0.0
            P[k, :] = 0.0
            if tk != 0.0:
                P[k, :pc] = x * vs
                P[k, pc] = 1.0
        U = P  # [8][p0]
        T = np.zeros((8, 8))
        for i in range(8):  # "lane i computes row i"
            T[i, i] = tau[i]
            for k in range(i + 1, 8):
                s = 0.0
                for j in range(i, k):
                    s += T[i, j] * zz[k, j]
                T[i, k] = -tau[k] * s
        # (2b) pending update of the leading p0 x p0 block
        if V is not None:
            A[:p0, :p0] -= V[:, :p0].T @ W[:, :p0] + W[:, :p0].T @ V[:, :p0]
        # (3) Y = A U^T  ([8][p0], k-major)
        Y = (A[:p0, :p0] @ U.T).T
        # (4) X = Y^T T; Z = U X; S = T^T Z; W = X - U^T S / 2
        X = Y.T @ T  # p0 x 8
        Z = U @ X  # 8 x 8
        S = T.T @ Z
        Wn = X - 0.5 * U.T @ S
        V = np.zeros((8, N))
        V[:, :p0] = U
        W = np.zeros((8, N))
        W[:, :p0] = Wn.T
        m = p0
    # leading 8 x 8 tile
    blk = A[:8, :8].copy()
    if V is not None:
        blk -= V[:, :8].T @ W[:, :8] + W[:, :8].T @ V[:, :8]
    for g in range(8):
        for c in range(g + 1):
            Bd[g, g - c] = blk[g, c]
    return Bd, off


def band_dense(Bd, off):
    N = Bd.shape[0]
    M = np.zeros((N, N))
    for r in range(N):
        for j in range(9):
            if r - j >= 0:
                M[r, r - j] = M[r - j, r] = Bd[r, j]
    return M[off:, off:]


# ---------------------------------------------------------------- chase (b = 8), lock-step groups
def chase_step(L, k, j):
    """L[r][jj] = B[r][r - jj], jj = 0..15.  Step j of sweep k: rows R = r0..r0+7, r0 = k + 1 + 8 j."""
    r0 = k + 1 + 8 * j
    xj = 1 if j == 0 else 8  # jj of x_0 in row r0
    x = np.array([L[r0 + i, xj + i] for i in range(8)])
    xn2 = float(x[1:] @ x[1:])
    if xn2 <= 0.0:
        return
    beta, tau, vs = house(x[0], xn2)
    v = x * vs
    v[0] = 1.0
    new = {}
    # (b) diagonal block, two-sided
    D = np.zeros((8, 8))
    for i in range(8):
        for c in range(i + 1):
            D[i, c] = D[c, i] = L[r0 + i, i - c]
    p = tau * (D @ v)
    K = 0.5 * tau * float(v @ p)
    w = p - K * v
    for q in range(8):
        for c in range(q + 1):
            new[(r0 + q, q - c)] = D[q, c] - v[q] * w[c] - w[q] * v[c]
    # (a) block left of it: rows R, columns r0-8 .. r0-1, from the left; column q owned by lane q
    for q in range(8):
        c = r0 - 8 + q
        if c < 0:
            continue
        col = np.array([L[r0 + i, 8 + i - q] for i in range(8)])
        if 8 - q == xj:  # the column the reflector was built from
            col = np.array([beta] + [0.0] * 7)
        else:
            col = col - tau * float(col @ v) * v
        for i in range(8):
            new[(r0 + i, 8 + i - q)] = col[i]
    # (c) block below: rows r0+8 .. r0+15, columns R, from the right; row q owned by lane q
    for q in range(8):
        r = r0 + 8 + q
        row = np.array([L[r, 8 + q - c] for c in range(8)])
        row = row - tau * float(row @ v) * v
        for c in range(8):
            new[(r, 8 + q - c)] = row[c]
    for (r, jj), val in new.items():
        L[r, jj] = val


def chase8(Bd, off, groups=4, lag=4):
    N = Bd.shape[0]
    n = N - off
    L = np.zeros((n + 3 * 8, 16))
    for r in range(n):
        for jj in range(9):
            if r - jj >= 0:
                L[r, jj] = Bd[off + r, jj]
    k = list(range(groups))
    j = [0] * groups
    active = [g <= n - 3 for g in range(groups)]
    ticks = 0
    while any(active):
        ticks += 1
        snap = [(k[g], j[g], active[g]) for g in range(groups)]
        for g in range(groups):
            if not active[g]:
                continue
            kp, jp, ap = snap[(g + groups - 1) % groups]
            ok = k[g] == 0 or (not ap) or kp > k[g] - 1 or (kp == k[g] - 1 and jp >= j[g] + lag)
            if not ok:
                continue
            chase_step(L, k[g], j[g])
            j[g] += 1
            if k[g] + 1 + 8 * j[g] > n - 2:
                k[g] += groups
                j[g] = 0
                if k[g] > n - 3:
                    active[g] = False
    return L[:n, 0].copy(), L[1:n, 1].copy(), ticks


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n in [1, 2, 3, 7, 8, 9, 15, 16, 17, 23, 24, 25, 33, 50, 64, 96, 100, 192, 200]:
        Wm = rng.standard_normal((n, 2 * n + 3)) * 0.02
        G = Wm @ Wm.T
        ref = np.linalg.eigvalsh(G)
        Bd, off = sbr8(G)
        eb = np.linalg.eigvalsh(band_dense(Bd, off))
        d, e, ticks = chase8(Bd, off)
        Tm = np.diag(d) + (np.diag(e, 1) + np.diag(e, -1) if n > 1 else 0)
        et = np.linalg.eigvalsh(Tm)
        d3, e3, ticks3 = chase8(Bd, off, groups=8, lag=3)
        T3 = np.diag(d3) + (np.diag(e3, 1) + np.diag(e3, -1) if n > 1 else 0)
        e3v = np.linalg.eigvalsh(T3)
        print(n, "band err %.2e" % (np.abs(eb - ref).max() / ref.max()), "tri err %.2e" % (np.abs(et - ref).max() / ref.max()),
              "ticks", ticks, "| lag3 g8 err %.2e" % (np.abs(e3v - ref).max() / ref.max()), "ticks", ticks3)
