"""Lane-level Python model of the DMMA lift kernel (ls_spa_b200/csrc/lifts_mma.cu).

It mirrors the CUDA kernel's data layout and per-lane register contents (mma.sync m8n8k4
f64 fragments, 16-byte tile accesses, the C->A operand reuse), so that the index algebra can
be checked against the oracle on a machine without a GPU.  Test infrastructure only.

Layouts (lane = 4*c + q, c = lane>>2, q = lane&3):
  tile access   ld_tile(M, i0, j0): lane holds M[i0+2q+e][j0+c], e = 0,1   ("T layout")
  mma(C, A, B)  A[m=c][k=q], B[k=q][n=c], C[m=c][n=2q+e]
"""

import numpy as np

LANES = np.arange(32)
C_ = LANES >> 2
Q_ = LANES & 3


def mma(cfrag, afrag, bfrag):
    """cfrag (32,2), afrag (32,), bfrag (32,) -> D = A @ B + C in C layout."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    A[C_, Q_] = afrag
    B[Q_, C_] = bfrag
    Cm = np.zeros((8, 8))
    Cm[C_, 2 * Q_] = cfrag[:, 0]
    Cm[C_, 2 * Q_ + 1] = cfrag[:, 1]
    D = A @ B + Cm
    return np.stack([D[C_, 2 * Q_], D[C_, 2 * Q_ + 1]], 1)


def ld_tile(M, i0, j0):
    """M is a 2-D array indexed [row, col] (column-major in the kernel)."""
    return np.stack([M[i0 + 2 * Q_, j0 + C_], M[i0 + 2 * Q_ + 1, j0 + C_]], 1)


def st_tile(M, i0, j0, frag):
    M[i0 + 2 * Q_, j0 + C_] = frag[:, 0]
    M[i0 + 2 * Q_ + 1, j0 + C_] = frag[:, 1]


def quad_sum(v):
    out = v.copy()
    for l in range(32):
        out[l] = v[(l & ~3):(l & ~3) + 4].sum()
    return out


def lifts_one(R_tr, c_tr, R_te, c_te, ynsq, perm):
    p = len(perm)
    RT = -(-p // 8)
    PT = -(-(p + 1) // 8)
    NR, NC = 8 * RT, 8 * PT
    A = np.zeros((NR, NC))
    for k in range(p):
        col = perm[k]
        A[:col + 1, k] = R_tr[:col + 1, col]
    A[:p, p] = c_tr

    # ---------------- phase 1: blocked Householder, panel = 8 columns
    for s in range(RT):
        j0 = 8 * s
        nf = min(8, p - j0)
        # --- panel factorisation by warp 0; registers vr[t] = ld_tile(A, 8t, j0), t = s..RT-1
        vr = {t: ld_tile(A, 8 * t, j0) for t in range(s, RT)}
        Vb = np.zeros((NR, 8))            # column-major V (rows absolute), explicit unit diagonal
        tau = np.zeros(8)
        rtop = vr[s].copy()               # R entries of the top tile are read back from here at the end
        for cc in range(nf):
            qq, ee = cc >> 1, cc & 1
            # norm^2 of column cc strictly below the pivot (lanes of column cc)
            part = np.zeros(32)
            for t in range(s, RT):
                for e in range(2):
                    lr = 2 * Q_ + e
                    below = (t > s) | (lr > cc)
                    part += np.where(below, vr[t][:, e] ** 2, 0.0)
            sig = quad_sum(part)[4 * cc]                       # broadcast from the quad of column cc
            x0 = vr[s][4 * cc + qq, ee]                        # pivot, broadcast from lane 4cc+qq
            t_, scale, beta = 0.0, 0.0, x0
            if sig != 0.0:
                nrm = np.sqrt(x0 * x0 + sig)
                beta = -nrm if x0 >= 0 else nrm
                t_ = (beta - x0) / beta
                scale = 1.0 / (x0 - beta)
            tau[cc] = t_
            # lanes of column cc: scale sub-pivot entries, record beta, write v to Vb
            own = C_ == cc
            for t in range(s, RT):
                for e in range(2):
                    lr = 2 * Q_ + e
                    below = (t > s) | (lr > cc)
                    vr[t][:, e] = np.where(own & below, vr[t][:, e] * scale, vr[t][:, e])
            A[j0 + cc, j0 + cc] = beta
            for t in range(s, RT):
                for e in range(2):
                    lr = 2 * Q_ + e
                    rows = 8 * t + lr
                    val = np.where((t > s) | (lr > cc), vr[t][:, e], np.where(lr == cc, 1.0, 0.0))
                    Vb[rows[own], cc] = val[own]
            # apply H_cc to the later columns of the panel tile (lanes with c > cc)
            w = np.zeros(32)
            vv = {}
            for t in range(s, RT):
                vv[t] = np.stack([Vb[8 * t + 2 * Q_, cc], Vb[8 * t + 2 * Q_ + 1, cc]], 1)   # LDS.128 broadcast
                w += vv[t][:, 0] * vr[t][:, 0] + vv[t][:, 1] * vr[t][:, 1]
            w = quad_sum(w) * t_
            later = C_ > cc
            for t in range(s, RT):
                for e in range(2):
                    vr[t][:, e] = np.where(later, vr[t][:, e] - w * vv[t][:, e], vr[t][:, e])
        # write back: R part of the top tile (rows above / on the diagonal) and the updated
        # non-factored columns (c >= nf) of every tile of the strip
        for t in range(s, RT):
            fr = vr[t].copy()
            if t == s:
                for e in range(2):
                    lr = 2 * Q_ + e
                    diag = np.array([A[j0 + c, j0 + c] if c < nf else 0.0 for c in C_])
                    fr[:, e] = np.where(C_ >= nf, vr[t][:, e],
                                        np.where(lr < C_, vr[t][:, e], np.where(lr == C_, diag, 0.0)))
            else:
                for e in range(2):
                    fr[:, e] = np.where(C_ >= nf, vr[t][:, e], 0.0)
            st_tile(A, 8 * t, j0, fr)
        Vt = Vb.copy()                     # the kernel keeps a second, row-major copy for the U phase
        # G = V^T V from the registers (A and B operand are the same fragments), then T
        Vreg = {t: ld_tile(Vb, 8 * t, 0) for t in range(s, RT)}
        G = np.zeros((32, 2))
        for t in range(s, RT):
            for e in range(2):
                G = mma(G, Vreg[t][:, e], Vreg[t][:, e])
        Gm = np.zeros((8, 8))
        Gm[C_, 2 * Q_] = G[:, 0]
        Gm[C_, 2 * Q_ + 1] = G[:, 1]
        Tm = np.zeros((8, 8))
        for u in range(8):                 # lane u owns row u of T
            for c in range(u, 8):
                if c == u:
                    Tm[u, c] = tau[c]
                else:
                    Tm[u, c] = -tau[c] * sum(Tm[u, k] * Gm[k, c] for k in range(u, c))
        # --- trailing column tiles: one warp per column tile, W then U, no block sync in between
        for j in range(s + 1, PT):
            acc = np.zeros((32, 2))
            for t in range(s, RT):
                a2 = ld_tile(A, 8 * t, 8 * j)
                vb = ld_tile(Vb, 8 * t, 0)
                for e in range(2):
                    acc = mma(acc, a2[:, e], vb[:, e])          # C[m=jcol][n=v] = Wraw^T
            tt = ld_tile(Tm, 0, 0)
            wp = np.zeros((32, 2))
            for e in range(2):
                wp = mma(wp, acc[:, e], tt[:, e])               # W'^T = Wraw^T T  (C -> A reuse)
            for t in range(s, RT):
                cfr = ld_tile(A, 8 * t, 8 * j)                  # C layout of the transposed tile
                vt = np.stack([Vt[8 * t + C_, 2 * Q_], Vt[8 * t + C_, 2 * Q_ + 1]], 1)   # row-major V, LDS.128
                for e in range(2):
                    cfr = mma(cfr, -wp[:, e], vt[:, e])
                st_tile(A, 8 * t, 8 * j, cfr)

    # ---------------- phase 1.5: inverses of the diagonal blocks
    Dinv = np.zeros((RT, 8, 8))
    for J in range(RT):
        D = A[8 * J:8 * J + 8, 8 * J:8 * J + 8].copy()
        for k in range(8):
            if 8 * J + k >= p:
                D[k, :] = 0.0
                D[:, k] = 0.0
                D[k, k] = 1.0
        for jj in range(8):
            x = np.zeros(8)
            x[jj] = 1.0 / D[jj, jj]
            for u in range(jj - 1, -1, -1):
                x[u] = -sum(D[u, k] * x[k] for k in range(u + 1, jj + 1)) / D[u, u]
            if 8 * J + jj >= p:
                x[:] = 0.0
            Dinv[J][:, jj] = x
    cvec = np.zeros(NR)
    cvec[:p] = A[:p, p]

    # ---------------- phase 2: one warp per 8-row tile of X, registers only
    X = np.zeros((NR, NR))
    for l in range(p):
        col = perm[l]
        X[:min(col + 1, R_te.shape[0]), l] = R_te[:col + 1, col]
    cte = np.zeros(NR)
    cte[:len(c_te)] = c_te
    cost = np.zeros(p + 1)
    cost[0] = c_te @ c_te
    for it in range(RT):
        rows = 8 * it + C_                                      # C layout: lane (row = c, cols 2q+e)
        xr = {L: np.stack([X[rows, 8 * L + 2 * Q_], X[rows, 8 * L + 2 * Q_ + 1]], 1) for L in range(RT)}
        r_in = cte[rows].copy()
        for J in range(RT):
            dv = ld_tile(Dinv[J], 0, 0)
            M = np.zeros((32, 2))
            for e in range(2):
                M = mma(M, xr[J][:, e], dv[:, e])
            cv = np.stack([cvec[8 * J + 2 * Q_], cvec[8 * J + 2 * Q_ + 1]], 1)
            # residuals after each of the 8 columns as one more product: R = r_in 1^T - M Tc with
            # Tc[k][n] = c_k for k <= n (B fragment of k-step e: k = 2q + e, n = c)
            R = np.stack([r_in, r_in], 1)
            for e in range(2):
                bfrag = np.where(2 * Q_ + e <= C_, cv[:, e], 0.0)
                R = mma(R, -M[:, e], bfrag)
            d = R * R
            # row sums of both squares with three shuffles: round 1 (xor 4) hands column e to the
            # lanes with (c & 1) == e, rounds 2 and 3 (xor 8, 16) finish the sum over the 8 rows
            odd = (C_ & 1) != 0
            keep = np.where(odd, d[:, 1], d[:, 0])
            send = np.where(odd, d[:, 0], d[:, 1])
            tot = keep + send[LANES ^ 4]
            tot = tot + tot[LANES ^ 8]
            tot = tot + tot[LANES ^ 16]
            for l in range(32):
                if C_[l] < 2:                                    # lanes of rows 0 and 1 write
                    k = 8 * J + 2 * Q_[l] + C_[l]
                    if k < p:
                        cost[k + 1] += tot[l]
            r_in = R[(LANES & ~3) + 3, 1]                        # residual after the last column
            for L in range(J + 1, RT):
                rt = ld_tile(A, 8 * J, 8 * L)
                for e in range(2):
                    xr[L] = mma(xr[L], -M[:, e], rt[:, e])
    lifts = np.zeros(p)
    lifts[np.asarray(perm)] = (cost[:-1] - cost[1:]) / ynsq
    return lifts
