"""Development prototype (NOT product, NOT oracle): the projected block-Jacobi preconditioner of c++/PCG
(ConstraintJoint.cpp:1236 preprocess_PCG_preconditioner, :1455 computeMinv_x; notes.pdf Alg. 10) restated in the world frame
used by the CUDA kernels: an exact O(n) solve of (J' blkdiag(M_j) J + Pr) y = x by an articulated-body recursion.
With world-frame screws s_k and world-frame body inertias no frame transforms are needed:
   leaves->root : IA_j = I_j + sum_c (IA_c - U_c U_c'/d_c),  U_c = IA_c s_c,  d_c = s_c'U_c + Pr_c   (fixed joints pass IA_c whole)
   leaves->root : u_j = x_j - s_j'p_j ;  p_parent += p_j + U_j u_j/d_j
   root->leaves : y_j = (u_j - U_j'a_parent)/d_j ;  a_j = a_parent + s_j y_j
Checked against a dense solve with the oracle's M_r.  Run: python tools/proto_precond.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import redmax_oracle as ro  # noqa: E402
from proto_worldframe import cross, flatten  # noqa: E402


def br(x):
    return ro.se3_brac(x)


def world_data(d, q):
    n = d['n']
    Rw, pw = [None] * n, [None] * n
    s = np.zeros((n, 6))
    Iw = np.zeros((n, 6, 6))
    for j in range(n):
        p = d['parent'][j]
        E = d['E0_pj'][j].copy()
        if d['ndof'][j]:
            Q = np.eye(4)
            Q[:3, :3] = ro.se3_aaToMat(d['axis'][j], q[d['idx'][j]])
            E = E @ Q
        Ew = E if p < 0 else np.block([[Rw[p], pw[p][:, None]], [np.zeros((1, 3)), np.ones((1, 1))]]) @ E
        Rw[j], pw[j] = Ew[:3, :3], Ew[:3, 3]
        if d['ndof'][j]:
            w = Rw[j] @ d['axis'][j]
            s[j] = np.concatenate([w, cross(pw[j], w)])
        Eb = Ew @ d['E0_ji'][j]
        R, pb = Eb[:3, :3], Eb[:3, 3]
        I = d['I'][j]
        m = I[3]
        Iw[j] = np.block([[R @ np.diag(I[:3]) @ R.T - m * br(pb) @ br(pb), m * br(pb)], [-m * br(pb), m * np.eye(3)]])
    return s, Iw


def aba_solve(d, s, Iw, Pr, x):
    n = d['n']
    par = d['parent']
    IA = Iw.copy()
    U = np.zeros((n, 6))
    dd = np.ones(n)
    for j in range(n - 1, -1, -1):
        if d['ndof'][j]:
            U[j] = IA[j] @ s[j]
            dd[j] = s[j] @ U[j] + Pr[d['idx'][j]]
            Ia = IA[j] - np.outer(U[j], U[j]) / dd[j]
        else:
            Ia = IA[j]
        if par[j] >= 0:
            IA[par[j]] += Ia
    p = np.zeros((n, 6))
    u = np.zeros(n)
    for j in range(n - 1, -1, -1):
        pj = p[j]
        if d['ndof'][j]:
            u[j] = x[d['idx'][j]] - s[j] @ pj
            pj = pj + U[j] * u[j] / dd[j]
        if par[j] >= 0:
            p[par[j]] += pj
    a = np.zeros((n, 6))
    y = np.zeros(len(x))
    for j in range(n):
        ap = a[par[j]] if par[j] >= 0 else np.zeros(6)
        if d['ndof'][j]:
            yj = (u[j] - U[j] @ ap) / dd[j]
            y[d['idx'][j]] = yj
            a[j] = ap + s[j] * yj
        else:
            a[j] = ap
    return y


if __name__ == '__main__':
    for name, sc in (('scene0', ro.scenes(0)), ('scene2', ro.scenes(2)), ('hand', ro.hand_scene()), ('chain32', ro.chain_scene(32))):
        sc.init()
        rng = np.random.default_rng(1)
        q = sc.qInit + 0.3 * rng.uniform(-1, 1, sc.nr)
        sc.setQ(q, np.zeros(sc.nr))
        sc.update()
        M, f = ro.compute_values(sc, False)
        d = flatten(sc)
        s, Iw = world_data(d, q)
        Pr = rng.uniform(0, 50, sc.nr)
        x = rng.uniform(-1, 1, sc.nr)
        y = aba_solve(d, s, Iw, Pr, x)
        yref = np.linalg.solve(M + np.diag(Pr), x)
        print('%-8s nr=%2d  rel err %.2e' % (name, sc.nr, np.linalg.norm(y - yref) / np.linalg.norm(yref)))
