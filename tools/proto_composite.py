"""Development prototype (NOT product, NOT oracle): the composite (world-frame, O(n) + n^2 dot products) formulation of the
Newton matrix that the fast CUDA path implements; checked here against the dense oracle before it is written in CUDA.

    T^i_j = B_j c2_i + A_j c1_i + sq C_j s_i     (tangent wrench of body j for column i; world frame)
    H[k][i] = (B^C_k s_k).c2_i + (A^C_k' s_k).c1_i + sq (C^C_k' s_k).s_i        k in sub(i)
            = s_k.(B^C_i c2_i + A^C_i c1_i + sq C^C_i s_i + sq ad*(s_i) F^C_i)  k a proper ancestor of i
with X^C_k = sum over the subtree of k.  Run: python tools/proto_composite.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import redmax_oracle as ro  # noqa: E402
from proto_worldframe import ad_mv, adstar_fv, body_ext, cross, flatten  # noqa: E402


def br(x):
    return ro.se3_brac(x)


def evaluate2(d, q, qdot, dq, c, beta, sq=1.0, sqd=None, sd=1.0):
    sqd = beta if sqd is None else sqd
    n, nr = d['n'], d['nr']
    par = d['parent']
    Rw = [None] * n
    pw = [None] * n
    s = np.zeros((n, 6))
    V = np.zeros((n, 6))
    U = np.zeros((n, 6))
    for j in range(n):
        p = par[j]
        E = d['E0_pj'][j].copy()
        if d['ndof'][j]:
            Q = np.eye(4)
            Q[:3, :3] = ro.se3_aaToMat(d['axis'][j], q[d['idx'][j]])
            E = E @ Q
        Ew = E if p < 0 else np.block([[Rw[p], pw[p][:, None]], [np.zeros((1, 3)), np.ones((1, 1))]]) @ E
        Rw[j], pw[j] = Ew[:3, :3], Ew[:3, 3]
        Vp = V[p] if p >= 0 else np.zeros(6)
        Up = U[p] if p >= 0 else np.zeros(6)
        if d['ndof'][j]:
            w = Rw[j] @ d['axis'][j]
            s[j] = np.concatenate([w, cross(pw[j], w)])
            qd = qdot[d['idx'][j]]
            V[j] = Vp + s[j] * qd
            U[j] = Up + s[j] * dq[d['idx'][j]] + c * ad_mv(Vp, s[j]) * qd
        else:
            V[j], U[j] = Vp, Up
    grav = d['grav']
    F = np.zeros((n, 6))
    Jb = np.zeros((n, 3, 3))   # sum (Ibar - m [p][p])
    mp = np.zeros((n, 3))
    M = np.zeros(n)
    Atl = np.zeros((n, 3, 3))
    mv = np.zeros((n, 3))
    Aext = np.zeros((n, 6, 6))
    Cext = np.zeros((n, 6, 6))
    for j in range(n):
        Eb = np.block([[Rw[j], pw[j][:, None]], [np.zeros((1, 3)), np.ones((1, 1))]]) @ d['E0_ji'][j]
        R, p = Eb[:3, :3], Eb[:3, 3]
        I = d['I'][j]
        m = I[3]
        ph = np.concatenate([R.T @ V[j][:3], R.T @ (V[j][3:] + cross(V[j][:3], p))])
        u = np.concatenate([R.T @ U[j][:3], R.T @ (U[j][3:] + cross(U[j][:3], p))])
        Iw = I[:3] * ph[:3]
        mvb = m * ph[3:]
        fcor = np.concatenate([cross(Iw, ph[:3]) + cross(mvb, ph[3:]), cross(mvb, ph[:3])])
        fgrav = np.concatenate([np.zeros(3), m * (R.T @ grav)])
        fext, Kext, Dext = body_ext(d['ground'][j], R, p, ph, d['sides'][j], True)
        Fb = I * u - c * (fcor + fgrav + fext)
        F[j] = np.concatenate([R @ Fb[:3] + cross(p, R @ Fb[3:]), R @ Fb[3:]])
        I3 = np.diag(I[:3])
        Jb[j] = R @ I3 @ R.T - m * br(p) @ br(p)
        mp[j] = m * p
        M[j] = m
        wb, vb = ph[:3], ph[3:]
        Ptl = I3 @ br(wb) - br(wb) @ I3 + br(I3 @ wb)
        vc = R @ vb
        Atl[j] = -c * (R @ Ptl @ R.T + 2 * m * br(p) @ br(vc))
        mv[j] = m * vc
        if d['ground'][j] is not None:
            X = np.block([[R.T, np.zeros((3, 3))], [-R.T @ br(p), R.T]])
            Aext[j] = -c * X.T @ Dext @ X
            Cext[j] = -c * X.T @ Kext @ X
    # composite sums, leaves -> root
    for j in range(n - 1, 0, -1):
        p = par[j]
        for arr in (F, Jb, mp, M, Atl, mv, Aext, Cext):
            arr[p] += arr[j]
    g = np.zeros(nr)
    H = np.zeros((nr, nr))
    size = d['size']
    L = np.zeros((n, 18))
    Rt = np.zeros((n, 18))
    Z = np.zeros((n, 6))
    dKD = np.zeros(n)
    for k in range(n):
        if not d['ndof'][k]:
            continue
        jt = d['joints'][k]
        r = d['idx'][k]
        qk, qdk = q[r], qdot[r]
        fr = jt.tau[0] + jt.stiffness * (jt.qRest[0] - qk) - jt.damping * qdk
        dK, dD = -jt.stiffness, -jt.damping
        if qk < jt.qLimL:
            fr += jt.qLimK * (jt.qLimL - qk) - jt.qLimD * qdk
            dK -= jt.qLimK
            dD -= jt.qLimD
        if qk > jt.qLimU:
            fr += jt.qLimK * (jt.qLimU - qk) - jt.qLimD * qdk
            dK -= jt.qLimK
            dD -= jt.qLimD
        g[r] = s[k] @ F[k] - c * fr
        dKD[k] = -c * (sq * dK + sqd * dD)
        sw, sv = s[k][:3], s[k][3:]
        # a = B^C s
        a = np.concatenate([Jb[k] @ sw + cross(mp[k], sv), -cross(mp[k], sw) + M[k] * sv])
        # b = A^C' s  (structured part only has the first three entries)
        b = np.concatenate([Atl[k].T @ sw + 2 * c * cross(mv[k], sv), np.zeros(3)]) + Aext[k].T @ s[k]
        # e = C^C' s
        e = np.concatenate([-c * cross(grav, cross(mp[k], sw) - M[k] * sv), np.zeros(3)]) + Cext[k].T @ s[k]
        L[k] = np.concatenate([a, b, e])
        p = par[k]
        Vp = V[p] if p >= 0 else np.zeros(6)
        Up = U[p] if p >= 0 else np.zeros(6)
        c1 = sqd * s[k] - sq * ad_mv(s[k], Vp)
        c2 = sd * s[k] - sq * ad_mv(s[k], Up) + c * (sqd * ad_mv(Vp, s[k]) - ad_mv(c1, Vp))
        Rt[k] = np.concatenate([c2, c1, sq * s[k]])
        c1w = c1[:3]
        Bc2 = np.concatenate([Jb[k] @ c2[:3] + cross(mp[k], c2[3:]), -cross(mp[k], c2[:3]) + M[k] * c2[3:]])
        Ac1 = np.concatenate([Atl[k] @ c1w, -2 * c * cross(mv[k], c1w)]) + Aext[k] @ c1
        Cs = sq * (np.concatenate([-c * cross(mp[k], cross(grav, sw)), -c * M[k] * cross(grav, sw)]) + Cext[k] @ s[k])
        Z[k] = Bc2 + Ac1 + Cs + sq * adstar_fv(s[k], F[k])
    for i in range(n):
        if not d['ndof'][i]:
            continue
        ci = d['idx'][i]
        for k in range(n):
            if not d['ndof'][k]:
                continue
            rk = d['idx'][k]
            if i <= k < i + size[i]:
                H[rk, ci] = L[k] @ Rt[i]
            elif k <= i < k + size[k]:
                H[rk, ci] = s[k] @ Z[i]
        H[ci, ci] += dKD[i]
    return g, H


def check(scene, seed, label):
    rng = np.random.default_rng(seed)
    nr = scene.nr
    q1 = scene.qInit + 0.3 * rng.uniform(-1, 1, nr)
    q0 = q1 - 0.01 * rng.uniform(-1, 1, nr)
    qdot0 = rng.uniform(-1, 1, nr)
    h = scene.h
    scene.setQ0(q0, qdot0)
    for j in scene.joints:
        j.tau = rng.uniform(-1, 1, j.ndof) * 100
    g_ref, H_ref, M_ref, f, K, D_ref, J = ro.eval_bdf1(q1, scene, True, True)
    d = flatten(scene)
    args = (d, q1, (q1 - q0) / h, q1 - q0 - h * qdot0, h * h, 1 / h)
    g, H = evaluate2(*args)
    _, Mm = evaluate2(*args, sq=0.0, sqd=0.0, sd=1.0)
    _, Dm = evaluate2(*args, sq=0.0, sqd=1.0, sd=0.0)
    Dm = -Dm / (h * h)
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    print('%-20s nr=%2d  rel err g %.2e  H %.2e  M %.2e  D %.2e' % (label, nr, rel(g, g_ref), rel(H, H_ref), rel(Mm, M_ref),
                                                                  rel(Dm, D_ref)))


if __name__ == '__main__':
    for sid in (0, 1, 2, 14):
        sc = ro.scenes(sid)
        sc.init()
        check(sc, sid, 'scene %d' % sid)
    sc = ro.chain_scene(8, ground=True, h=1e-3)
    sc.init()
    for f in sc.forces:
        f.E[2, 3] = -5.0
    check(sc, 7, 'chain8+ground')
    sc = ro.hand_scene()
    sc.init()
    check(sc, 8, 'hand')
    sc = ro.chain_scene(32)
    sc.init()
    check(sc, 9, 'chain32')
