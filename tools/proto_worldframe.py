"""Development prototype (NOT product, NOT oracle): the O(n^2) world-frame residual/Jacobian evaluation that
the CUDA kernel implements, written in NumPy so its maths can be checked against the dense oracle before it is
written in CUDA.  See DESIGN.md section "Algorithm".  Run:  python tools/proto_worldframe.py
"""
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'oracle'))
import redmax_oracle as ro  # noqa: E402


def cross(a, b):
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])


def ad_mv(a, b):
    """ad(a) b for motion vectors [w; v]."""
    return np.concatenate([cross(a[:3], b[:3]), cross(a[:3], b[3:]) + cross(a[3:], b[:3])])


def adstar_fv(s, F):
    """-ad(s)^T F for force vectors [tau; f]."""
    return np.concatenate([cross(s[:3], F[:3]) + cross(s[3:], F[3:]), cross(s[:3], F[3:])])


def flatten(scene):
    joints = scene.joints
    n = len(joints)
    idx = {id(j): k for k, j in enumerate(joints)}
    d = dict(n=n, nr=scene.nr)
    d['parent'] = [(-1 if j.parent is None else idx[id(j.parent)]) for j in joints]
    d['ndof'] = [j.ndof for j in joints]
    d['idx'] = [(int(j.idxR[0]) if j.ndof else -1) for j in joints]
    d['axis'] = [getattr(j, 'axis', np.zeros(3)) for j in joints]
    d['E0_pj'] = [np.eye(4) if j.E0_pj is None else j.E0_pj for j in joints]
    d['E0_ji'] = [j.body.E0_ji for j in joints]
    d['I'] = [j.body.I_i for j in joints]
    d['sides'] = [j.body.sides for j in joints]
    d['joints'] = joints
    d['grav'] = scene.grav
    ground = [None] * n
    for f in scene.forces:
        if isinstance(f, ro.ForceGroundCuboid):
            ground[idx[id(f.cuboid.joint)]] = f
    d['ground'] = ground
    # subtree sizes (requires DFS preorder listing)
    size = [1] * n
    for j in range(n - 1, 0, -1):
        size[d['parent'][j]] += size[j]
    d['size'] = size
    return d


def body_ext(f, R, p, phi, sides, deriv):
    """ForceGroundCuboid restated per body: returns wrench (6), Km (6x6), Dm (6x6) in body coordinates."""
    fm = np.zeros(6)
    Km = np.zeros((6, 6))
    Dm = np.zeros((6, 6))
    if f is None:
        return fm, Km, Dm
    fake_fm = np.zeros(6)

    class B:  # minimal stand-in to reuse the oracle's restatement
        pass
    b = B()
    b.idxM = np.arange(6)
    E = np.eye(4)
    E[:3, :3] = R
    E[:3, 3] = p
    b.E_wi = E
    b.phi = phi
    b.sides = sides
    old = f.cuboid
    f.cuboid = b
    if deriv:
        f.computeValues_(None, fake_fm, None, Km, None, Dm)
    else:
        f.computeValues_(None, fake_fm)
    f.cuboid = old
    return fake_fm, Km, Dm


def evaluate(d, q, qdot, dq, c, beta, deriv=True):
    """g = M(q) dq - c f(q, qdot); H = dg/dq with d(qdot)/dq = beta I, d(dq)/dq = I."""
    n, nr = d['n'], d['nr']
    Rw = [None] * n
    pw = [None] * n
    s = np.zeros((n, 6))
    V = np.zeros((n, 6))
    U = np.zeros((n, 6))
    Rb = [None] * n
    pb = [None] * n
    phi = np.zeros((n, 6))
    Fw = np.zeros((n, 6))
    Kb = [None] * n
    Db = [None] * n
    for j in range(n):
        p = d['parent'][j]
        E = d['E0_pj'][j].copy()
        if d['ndof'][j]:
            Q = np.eye(4)
            Q[:3, :3] = ro.se3_aaToMat(d['axis'][j], q[d['idx'][j]])
            E = E @ Q
        Ew = E if p < 0 else np.block([[Rw[p], pw[p][:, None]], [np.zeros((1, 3)), np.ones((1, 1))]]) @ E
        Rw[j] = Ew[:3, :3]
        pw[j] = Ew[:3, 3]
        Vp = V[p] if p >= 0 else np.zeros(6)
        Up = U[p] if p >= 0 else np.zeros(6)
        if d['ndof'][j]:
            w = Rw[j] @ d['axis'][j]
            s[j] = np.concatenate([w, cross(pw[j], w)])
            qd = qdot[d['idx'][j]]
            sdot = ad_mv(Vp, s[j])
            V[j] = Vp + s[j] * qd
            U[j] = Up + s[j] * dq[d['idx'][j]] + c * sdot * qd
        else:
            V[j] = Vp
            U[j] = Up
    grav = d['grav']
    for j in range(n):
        Eb = np.block([[Rw[j], pw[j][:, None]], [np.zeros((1, 3)), np.ones((1, 1))]]) @ d['E0_ji'][j]
        R = Eb[:3, :3]
        p = Eb[:3, 3]
        Rb[j], pb[j] = R, p
        I = d['I'][j]
        ph = np.concatenate([R.T @ V[j][:3], R.T @ (V[j][3:] + cross(V[j][:3], p))])
        u = np.concatenate([R.T @ U[j][:3], R.T @ (U[j][3:] + cross(U[j][:3], p))])
        phi[j] = ph
        Iw = I[:3] * ph[:3]
        mv = I[3] * ph[3:]
        fcor = np.concatenate([cross(Iw, ph[:3]) + cross(mv, ph[3:]), cross(mv, ph[:3])])
        fgrav = np.concatenate([np.zeros(3), I[3] * (R.T @ grav)])
        fext, Kext, Dext = body_ext(d['ground'][j], R, p, ph, d['sides'][j], deriv)
        Fb = I * u - c * (fcor + fgrav + fext)
        Fw[j] = np.concatenate([R @ Fb[:3] + cross(p, R @ Fb[3:]), R @ Fb[3:]])
        if deriv:
            K = Kext.copy()
            K[3:6, 0:3] += ro.se3_brac(fgrav[3:])
            Kb[j] = K
            Db[j] = Dext
    Fsub = Fw.copy()
    for j in range(n - 1, 0, -1):
        Fsub[d['parent'][j]] += Fsub[j]
    g = np.zeros(nr)
    H = np.zeros((nr, nr))
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
        g[r] = s[k] @ Fsub[k] - c * fr
        H[r, r] += -c * (dK + beta * dD)
    if not deriv:
        return g
    size = d['size']
    for i in range(n):
        if not d['ndof'][i]:
            continue
        ci = d['idx'][i]
        p = d['parent'][i]
        Vp = V[p] if p >= 0 else np.zeros(6)
        Up = U[p] if p >= 0 else np.zeros(6)
        c1 = beta * s[i] - ad_mv(s[i], Vp)
        c2 = s[i] - ad_mv(s[i], Up) + c * beta * ad_mv(Vp, s[i]) - c * ad_mv(c1, Vp)
        T = np.zeros((n, 6))
        for j in range(i, i + size[i]):
            R, pp, I, ph = Rb[j], pb[j], d['I'][j], phi[j]

            def X(x):
                return np.concatenate([R.T @ x[:3], R.T @ (x[3:] + cross(x[:3], pp))])
            xi = X(s[i])
            dphi = X(c1)
            du = X(c2) + c * ad_mv(dphi, ph)
            Iw, mv = I[:3] * ph[:3], I[3] * ph[3:]
            dIw, dmv = I[:3] * dphi[:3], I[3] * dphi[3:]
            dfcor = np.concatenate([cross(dIw, ph[:3]) + cross(Iw, dphi[:3]) + cross(dmv, ph[3:]) + cross(mv, dphi[3:]),
                                    cross(dmv, ph[:3]) + cross(mv, dphi[:3])])
            dFb = I * du - c * (dfcor + Kb[j] @ xi + Db[j] @ dphi)
            T[j] = np.concatenate([R @ dFb[:3] + cross(pp, R @ dFb[3:]), R @ dFb[3:]])
        # subtree sums within sub(i)
        Tsub = T.copy()
        for j in range(i + size[i] - 1, i, -1):
            Tsub[d['parent'][j]] += Tsub[j]
        for k in range(i, i + size[i]):
            if d['ndof'][k]:
                H[d['idx'][k], ci] += s[k] @ Tsub[k]
        Z = Tsub[i] + adstar_fv(s[i], Fsub[i])
        k = d['parent'][i]
        while k >= 0:
            if d['ndof'][k]:
                H[d['idx'][k], ci] += s[k] @ Z
            k = d['parent'][k]
    return g, H


def check(scene, c, beta, seed, label):
    rng = np.random.default_rng(seed)
    nr = scene.nr
    q1 = scene.qInit + 0.3 * rng.uniform(-1, 1, nr)
    q0 = q1 - 0.01 * rng.uniform(-1, 1, nr)
    qdot0 = rng.uniform(-1, 1, nr)
    h = scene.h
    scene.setQ0(q0, qdot0)
    for j in scene.joints:
        j.tau = rng.uniform(-1, 1, j.ndof) * 100
    g_ref, H_ref = ro.eval_bdf1(q1, scene, True)
    d = flatten(scene)
    g, H = evaluate(d, q1, (q1 - q0) / h, q1 - q0 - h * qdot0, h * h, 1 / h)
    eg = np.linalg.norm(g - g_ref) / np.linalg.norm(g_ref)
    eH = np.linalg.norm(H - H_ref) / np.linalg.norm(H_ref)
    print('%-28s nr=%2d  rel err g %.2e  H %.2e' % (label, nr, eg, eH))
    return eg, eH


if __name__ == '__main__':
    for sid in (0, 1, 2, 14):
        sc = ro.scenes(sid)
        sc.init()
        check(sc, 0, 0, sid, 'scene %d' % sid)
    sc = ro.chain_scene(8, ground=True, h=1e-3)
    sc.init()
    # push the chain down so corners are in contact
    for f in sc.forces:
        f.E[2, 3] = -5.0
    check(sc, 0, 0, 7, 'chain8+ground')
    sc = ro.hand_scene()
    sc.init()
    check(sc, 0, 0, 8, 'hand')
    sc = ro.chain_scene(32)
    sc.init()
    check(sc, 0, 0, 9, 'chain32')
