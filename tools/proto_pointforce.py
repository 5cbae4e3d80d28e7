"""Development prototype (NOT product, NOT oracle): ForcePointPoint (matlab-diff/+redmax/ForcePointPoint.m) in the composite
world-frame formulation of tools/proto_composite.py, checked against the dense oracle before it is written in CUDA.

A point-point force couples two bodies a, b.  Its wrenches enter F_a, F_b; the diagonal blocks Km_aa, Dm_aa (body frame, as
the reference forms them) go through the same per-body external blocks as ground contact (Aext, Cext); the off-diagonal
blocks become a rank-structured update of the Newton matrix,
    H[k][i] += s_k . (Aext_ab c1_i + sq Cext_ab s_i)      for k in anc*(a), i in anc*(b)   (and a <-> b),
with Aext_ab = -c X_a' Dm_ab X_b, Cext_ab = -c X_a' Km_ab X_b.   Run: python tools/proto_pointforce.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import redmax_oracle as ro  # noqa: E402
from proto_composite import br  # noqa: E402
from proto_worldframe import ad_mv, adstar_fv, body_ext, cross, flatten  # noqa: E402


def flatten_pf(scene):
    d = flatten(scene)
    idx = {id(j.body): k for k, j in enumerate(scene.joints)}
    d['pointforces'] = [((-1 if f.body1 is None else idx[id(f.body1)]), f.x_1, (-1 if f.body2 is None else idx[id(f.body2)]),
                         f.x_2, f.stiffness, f.damping) for f in scene.forces if isinstance(f, ro.ForcePointPoint)]
    return d


def evaluate3(d, q, qdot, dq, c, beta, sq=1.0, sqd=None, sd=1.0):
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
    d['_Rb'], d['_pb'], d['_phib'] = [None] * n, [None] * n, [None] * n
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
        d['_Rb'][j], d['_pb'][j] = R, p
        I = d['I'][j]
        m = I[3]
        ph = np.concatenate([R.T @ V[j][:3], R.T @ (V[j][3:] + cross(V[j][:3], p))])
        u = np.concatenate([R.T @ U[j][:3], R.T @ (U[j][3:] + cross(U[j][:3], p))])
        d['_phib'][j] = ph
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
    # ---- point-point forces (ForcePointPoint.m:48-113): wrench on both bodies, own 6x6 blocks into Aext/Cext, cross blocks kept
    Rb, pb, phib = d['_Rb'], d['_pb'], d['_phib']
    cross_blocks = []  # (a, b, Aext_ab, Cext_ab): tangent wrench on body a per unit twist rate / displacement of body b (world)
    for (ba, x1, bb, x2, ks, kd) in d['pointforces']:
        def point(bj, xl):
            if bj < 0:
                return xl, np.zeros(3), np.zeros(3), np.eye(3), np.zeros(3), None
            R, p, ph = Rb[bj], pb[bj], phib[bj]
            G = ro.se3_Gamma(xl)
            vl = G @ ph
            return R @ xl + p, R @ vl, vl, R, p, G
        xw1, vw1, vl1, R1, p1, G1 = point(ba, x1)
        xw2, vw2, vl2, R2, p2, G2 = point(bb, x2)
        f = ks * (xw2 - xw1) + kd * (vw2 - vw1)
        I3 = np.eye(3)
        Z3 = np.zeros((3, 3))

        def Xof(R, p):
            return np.block([[R.T, np.zeros((3, 3))], [-R.T @ br(p), R.T]])
        if ba >= 0:
            fb1 = G1.T @ (R1.T @ f)                      # body-frame wrench on body 1
            X1 = Xof(R1, p1)
            F[ba] -= c * (X1.T @ fb1)
            K11 = ks * (G1.T @ np.hstack([br(R1.T @ (xw2 - p1)), -I3])) + kd * (G1.T @ np.hstack([br(R1.T @ vw2), Z3]))
            D11 = -kd * (G1.T @ G1)
            Aext[ba] += -c * X1.T @ D11 @ X1
            Cext[ba] += -c * X1.T @ K11 @ X1
        if bb >= 0:
            fb2 = -G2.T @ (R2.T @ f)
            X2 = Xof(R2, p2)
            F[bb] -= c * (X2.T @ fb2)
            K22 = ks * (G2.T @ np.hstack([br(R2.T @ (xw1 - p2)), -I3])) + kd * (G2.T @ np.hstack([br(R2.T @ vw1), Z3]))
            D22 = -kd * (G2.T @ G2)
            Aext[bb] += -c * X2.T @ D22 @ X2
            Cext[bb] += -c * X2.T @ K22 @ X2
        if ba >= 0 and bb >= 0:
            K12 = ks * (G1.T @ R1.T @ R2 @ np.hstack([-br(x2), I3])) - kd * (G1.T @ R1.T @ R2 @ np.hstack([br(vl2), Z3]))
            K21 = ks * (G2.T @ R2.T @ R1 @ np.hstack([-br(x1), I3])) - kd * (G2.T @ R2.T @ R1 @ np.hstack([br(vl1), Z3]))
            D12 = kd * (G1.T @ R1.T @ R2 @ G2)
            D21 = kd * (G2.T @ R2.T @ R1 @ G1)
            cross_blocks.append((ba, bb, -c * X1.T @ D12 @ X2, -c * X1.T @ K12 @ X2))
            cross_blocks.append((bb, ba, -c * X2.T @ D21 @ X1, -c * X2.T @ K21 @ X1))
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
    # cross terms: rows anc*(a), columns anc*(b):  H[k][i] += s_k . (Aext_ab c1_i + sq Cext_ab s_i)
    c1s = Rt[:, 6:12]
    for (ba, bb, Aab, Cab) in cross_blocks:
        i = bb
        while i >= 0:
            if d['ndof'][i]:
                y = Aab @ c1s[i] + sq * (Cab @ s[i])
                k = ba
                while k >= 0:
                    if d['ndof'][k]:
                        H[d['idx'][k], d['idx'][i]] += s[k] @ y
                    k = par[k]
            i = par[i]
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
    d = flatten_pf(scene)
    args = (d, q1, (q1 - q0) / h, q1 - q0 - h * qdot0, h * h, 1 / h)
    g, H = evaluate3(*args)
    _, Mm = evaluate3(*args, sq=0.0, sqd=0.0, sd=1.0)
    _, Dm = evaluate3(*args, sq=0.0, sqd=1.0, sd=0.0)
    Dm = -Dm / (h * h)
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    print('%-28s nr=%2d  rel err g %.2e  H %.2e  M %.2e  D %.2e' % (label, nr, rel(g, g_ref), rel(H, H_ref), rel(Mm, M_ref),
                                                                  rel(Dm, D_ref)))


def preorder(sc):
    """flatten() needs the joints listed in DFS preorder (the CUDA library reorders internally); scene 10 is not."""
    order = []

    def visit(j):
        order.append(j)
        for ch in j.children:
            visit(ch)
    for j in sc.joints:
        if j.parent is None:
            visit(j)
    sc.joints = order
    sc.bodies = [j.body for j in order]
    return sc


if __name__ == '__main__':
    import redmax_b200.scenes as scenes
    sc = preorder(scenes.scenesRedMax(10, api=ro))
    sc.init()
    check(sc, 10, 'scene 10 (loop)')
    sc = preorder(scenes.scenesRedMax(10, api=ro))
    sc.forces[0].setDamping(3e3)
    sc.forces.append(ro.ForcePointPoint(None, [3.0, 1.0, -12.0], sc.bodies[4], [0.5, 0.0, -4.0]))
    sc.forces[1].setStiffness(2e4)
    sc.forces[1].setDamping(5e2)
    sc.forces.append(ro.ForcePointPoint(sc.bodies[1], [0.2, 0.1, -3.0], sc.bodies[4], [0.0, 0.3, -1.0]))
    sc.forces[2].setStiffness(1e4)
    sc.forces[2].setDamping(1e3)
    sc.init()
    check(sc, 11, 'loop + damping + world + chain')
