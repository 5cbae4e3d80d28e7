"""Prototype (CPU, NumPy) of the blocked warp LU used by rmx_tc.cuh: lane-level emulation of the data flow -- rows never
move (indirection through perm[] / rem[]), 8-column panels factored in registers with LAPACK's idamax tie rule, U12 by a
forward substitution per trailing column, trailing update by 8x8x4 FP64 tensor-core tiles (DMMA fragment layout of PTX
mma.m8n8k4.f64: A[g][t], B[t][g], C[g][2t..2t+1], g = lane>>2, t = lane&3).  Checked against scipy's LU / solve.
"""
import numpy as np
import scipy.linalg as sla


def dmma(a, b, c0, c1):
    """One mma.sync.m8n8k4.f64: a[32], b[32] fragments, c0/c1[32] accumulators -> d0, d1."""
    A = np.zeros((8, 4))
    B = np.zeros((4, 8))
    C = np.zeros((8, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, t] = a[lane]
        B[t, g] = b[lane]
        C[g, 2 * t] = c0[lane]
        C[g, 2 * t + 1] = c1[lane]
    D = A @ B + C
    d0 = np.array([D[l >> 2, 2 * (l & 3)] for l in range(32)])
    d1 = np.array([D[l >> 2, 2 * (l & 3) + 1] for l in range(32)])
    return d0, d1


def lu_solve_blocked(Hin, rhs, scale=1.0):
    nr = Hin.shape[0]
    H = Hin.copy()              # "shared memory", H[r, c]
    NP = (nr + 7) // 8
    lanes = np.arange(32)
    done = lanes >= nr          # padding rows never participate
    pos = lanes.copy()          # LAPACK position of each row
    mypos = -np.ones(32, dtype=int)
    b = np.where(lanes < nr, scale * np.pad(rhs, (0, 32 - nr)), 0.0)
    rdiag = np.ones(32)
    perm = -np.ones(32, dtype=int)
    for p in range(NP):
        c0 = 8 * p
        w = min(8, nr - c0)
        a = np.zeros((32, 8))
        for r in range(nr):
            a[r, :w] = H[r, c0:c0 + w]
        for i in range(w):
            k = c0 + i
            v = np.where(done, -1.0, np.abs(a[:, i]))
            m = v.max()
            cand = np.where((v == m) & ~done)[0]
            src = cand[np.argmin(pos[cand])]          # first maximum in current (LAPACK) row order
            kl = np.where(~done & (pos == k))[0][0]
            pos[kl] = pos[src]
            pos[src] = k
            done[src] = True
            mypos[src] = k
            perm[k] = src
            piv = a[src, i]
            rp = 1.0 / piv
            rdiag[src] = rp
            l = np.where(done, 0.0, a[:, i] * rp)
            a[:, i] = np.where(done, a[:, i], l)
            for j in range(i + 1, w):
                a[:, j] = a[:, j] - l * a[src, j]
            b = b - l * b[src]
        for r in range(nr):
            H[r, c0:c0 + w] = a[r, :w]
        ntrail = nr - (c0 + 8)
        if ntrail <= 0:
            continue
        # U12: lane = trailing column
        for c in range(c0 + 8, nr):
            x = np.zeros(8)
            for i in range(8):
                acc = H[perm[c0 + i], c]
                for j in range(i):
                    acc = acc - H[perm[c0 + i], c0 + j] * x[j]
                x[i] = acc
            for i in range(8):
                H[perm[c0 + i], c] = x[i]
        # compact list of the remaining rows (ballot + popc rank), padded with -1
        rem = [r for r in range(32) if not done[r]]
        nrem = len(rem)
        rem = rem + [-1] * (-nrem % 8)
        for I in range(len(rem) // 8):
            for J in range((ntrail + 7) // 8):
                cJ = c0 + 8 + 8 * J
                c0f = np.zeros(32)
                c1f = np.zeros(32)
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    r = rem[8 * I + g]
                    if r >= 0:
                        if cJ + 2 * t < nr:
                            c0f[lane] = H[r, cJ + 2 * t]
                        if cJ + 2 * t + 1 < nr:
                            c1f[lane] = H[r, cJ + 2 * t + 1]
                for ks in range(2):
                    af = np.zeros(32)
                    bf = np.zeros(32)
                    for lane in range(32):
                        g, t = lane >> 2, lane & 3
                        r = rem[8 * I + g]
                        if r >= 0:
                            af[lane] = -H[r, c0 + 4 * ks + t]
                        if cJ + g < nr:
                            bf[lane] = H[perm[c0 + 4 * ks + t], cJ + g]
                    c0f, c1f = dmma(af, bf, c0f, c1f)
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    r = rem[8 * I + g]
                    if r >= 0:
                        if cJ + 2 * t < nr:
                            H[r, cJ + 2 * t] = c0f[lane]
                        if cJ + 2 * t + 1 < nr:
                            H[r, cJ + 2 * t + 1] = c1f[lane]
    # back substitution: lane = row, U[k][c] = H[perm[k], c]
    x = np.zeros(nr)
    for k in range(nr - 1, -1, -1):
        src = perm[k]
        xk = b[src] * rdiag[src]
        x[k] = xk
        for r in range(nr):
            if 0 <= mypos[r] < k:
                b[r] = b[r] - H[r, k] * xk
    return x, perm[:nr], H


if __name__ == '__main__':
    rng = np.random.default_rng(0)
    worst = 0.0
    for nr in (1, 2, 5, 8, 9, 10, 16, 20, 24, 31, 32):
        for trial in range(20):
            A = rng.standard_normal((nr, nr))
            if trial % 3 == 0:
                A = A + 5 * np.eye(nr)
            if trial % 4 == 1 and nr > 2:
                A[:, 0] = np.sign(A[:, 0])          # exact ties in the first pivot column
            rhs = rng.standard_normal(nr)
            x, perm, Hf = lu_solve_blocked(A, rhs, -1.0)
            xr = np.linalg.solve(A, -rhs)
            err = np.abs(x - xr).max() / max(np.abs(xr).max(), 1e-300)
            worst = max(worst, err / np.linalg.cond(A))
            lu, piv = sla.lu_factor(A)
            # scipy piv (swap sequence) -> permutation vector
            pv = np.arange(nr)
            for i, pidx in enumerate(piv):
                pv[i], pv[pidx] = pv[pidx], pv[i]
            assert (pv == perm).all(), (nr, trial, pv, perm)
            Lr = np.array([Hf[perm[k], :] for k in range(nr)])
            assert np.abs(Lr - lu).max() <= 1e-9 * np.abs(lu).max(), (nr, trial)
    print('blocked indirection LU == LAPACK pivots and factors; worst err/cond = %.2e' % worst)
