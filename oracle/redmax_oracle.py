"""CPU oracle: NumPy float64 restatement of sueda/redmax `matlab-diff` (TEST INFRASTRUCTURE ONLY).

This file is the parity checker for the CUDA path.  It is NOT part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import it.
The product (`redmax_b200/`) never imports anything under `oracle/`.

It restates, function by function and in the same operation order (dense O(n^4) algorithm, linked-list
traversal replaced by a loop over the same list order), these reference files (paths relative to
/root/reference/matlab-diff/):

    se3.m                     -> inv, Ad, ad, brac, Gamma, aaToMat, inertiaCuboid
    +redmax/Joint.m           -> Joint (countDofs:149, gather/scatter:173-369, update:382,
                                 computeForce:437, computeJacobian:490, computeEnergies:616)
    +redmax/JointRevolute.m   -> JointRevolute.update_:29
    +redmax/JointFixed.m      -> JointFixed
    +redmax/JointFree2D.m     -> JointFree2D.update_:20 (only to reach the scene-11 ground-contact pin)
    +redmax/JointSpherical.m  -> JointSpherical (update_:106, reparam_:63), euler_chart (getEuler:154 + the twelve generated
                                 chart bodies), euler_chart_inv (getEulerInv:186 + :1809-1964); JointFree3D.m -> JointFree3D
    +redmax/Body.m            -> Body (update:70, computeMassGrav:83, computeEnergies:167)
    +redmax/BodyCuboid.m      -> BodyCuboid.computeInertia_:16
    +redmax/Force.m, ForceNull.m, ForceGroundCuboid.m:54-183
    +redmax/Scene.m           -> Scene (init:59, reset:122, saveHistory:134, plotEnergies:164)
    driverRedMaxBDF1.m        -> sim_loop_bdf1, newton, eval_bdf1, compute_values
    driverRedMaxBDF2.m        -> sim_loop_bdf2, eval_sdirk2a/b, eval_bdf2
    driverRedMaxAdjointBDF1.m, driverRedMaxAdjointBDF2.m -> newton_adjoint, sim_loop_adjoint_bdf1/2
    +redmax/TaskBDF1.m, TaskBDF2.m, TaskBDF1PointPos.m, TaskBDF2PointPos.m
    scenesRedMax.m            -> scenes(sceneID) for IDs -2,-1,0,1,2,11,14,100,101 (3-10, 12, 13 through redmax_b200/scenes.py)

Pinning: `tests/test_oracle_pins.py` checks this oracle against every golden end-of-run energy
`Hexpected(BDF1/BDF2)` the reference holds for the in-scope joint/force types (scenes 0,1,2,14 and the
Free2D+ground scene 11, scenesRedMax.m:54,82,108,292,373; tolerance |dH|<=1e-2 as Scene.m:172; likewise scenes 3-10, 12, 13,
among them the Euler-chart joints of scenes 7 and 9 -- scene 7 under BDF2 switches charts twice, XYZ -> XYX -> YXZ).  The adjoint
scenes 100/101 carry no expected value in the reference ("parity unpinned" for P/dPdp by golden numbers);
they are pinned by the reference's own finite-difference recipe (driverRedMaxAdjointBDF1.m:47-61).

MATLAB built-ins on the path: `H\\g` (LAPACK dgesv, partial pivoting) -> numpy.linalg.solve;
`lu(H,'vector')` -> scipy.linalg.lu_factor (dgetrf).  Indices are 0-based here (1-based in MATLAB).
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla

THRESH = 1e-9  # se3.m:5


# ----------------------------------------------------------------------------------------------
# se3.m
# ----------------------------------------------------------------------------------------------
def se3_inv(E):
    """se3.m:11"""
    R = E[0:3, 0:3]
    p = E[0:3, 3]
    Ei = np.eye(4)
    Ei[0:3, 0:3] = R.T
    Ei[0:3, 3] = -R.T @ p
    return Ei


def se3_brac(x):
    """se3.m:89 (3-vector branch only)"""
    return np.array([[0.0, -x[2], x[1]], [x[2], 0.0, -x[0]], [-x[1], x[0], 0.0]])


def se3_Gamma(r):
    """se3.m:38"""
    return np.hstack([se3_brac(r[0:3]).T, np.eye(3)])


def se3_Ad(E):
    """se3.m:44"""
    A = np.zeros((6, 6))
    R = E[0:3, 0:3]
    p = E[0:3, 3]
    A[0:3, 0:3] = R
    A[3:6, 3:6] = R
    A[3:6, 0:3] = se3_brac(p) @ R
    return A


def se3_ad(phi):
    """se3.m:55 (6-vector branch)"""
    a = np.zeros((6, 6))
    w = phi[0:3]
    v = phi[3:6]
    W = se3_brac(w)
    a[0:3, 0:3] = W
    a[3:6, 0:3] = se3_brac(v)
    a[3:6, 3:6] = W
    return a


def se3_aaToMat(axis, angle):
    """se3.m:111 -- keeps the axis-aligned special cases (SURVEY note N3)."""
    R = np.eye(3)
    ax, ay, az = float(axis[0]), float(axis[1]), float(axis[2])
    mag = math.sqrt(ax * ax + ay * ay + az * az)
    if mag > THRESH:
        mag = 1.0 / mag
        ax = ax * mag
        ay = ay * mag
        az = az * mag
        if abs(ax) < THRESH and abs(ay) < THRESH:
            if az < 0:
                angle = -angle
            s = math.sin(angle)
            c = math.cos(angle)
            R[0, 0] = c
            R[0, 1] = -s
            R[1, 0] = s
            R[1, 1] = c
        elif abs(ay) < THRESH and abs(az) < THRESH:
            if ax < 0:
                angle = -angle
            s = math.sin(angle)
            c = math.cos(angle)
            R[1, 1] = c
            R[1, 2] = -s
            R[2, 1] = s
            R[2, 2] = c
        elif abs(az) < THRESH and abs(ax) < THRESH:
            if ay < 0:
                angle = -angle
            s = math.sin(angle)
            c = math.cos(angle)
            R[0, 0] = c
            R[0, 2] = s
            R[2, 0] = -s
            R[2, 2] = c
        else:
            s = math.sin(angle)
            c = math.cos(angle)
            t = 1.0 - c
            xz = ax * az
            xy = ax * ay
            yz = ay * az
            R[0, 0] = t * ax * ax + c
            R[0, 1] = t * xy - s * az
            R[0, 2] = t * xz + s * ay
            R[1, 0] = t * xy + s * az
            R[1, 1] = t * ay * ay + c
            R[1, 2] = t * yz - s * ax
            R[2, 0] = t * xz - s * ay
            R[2, 1] = t * yz + s * ax
            R[2, 2] = t * az * az + c
    return R


def se3_inertiaCuboid(whd, density):
    """se3.m:366"""
    whd = np.asarray(whd, dtype=float).reshape(3)
    m = np.zeros(6)
    mass = density * np.prod(whd)
    m[0] = (1.0 / 12.0) * mass * (whd[[1, 2]] @ whd[[1, 2]])
    m[1] = (1.0 / 12.0) * mass * (whd[[2, 0]] @ whd[[2, 0]])
    m[2] = (1.0 / 12.0) * mass * (whd[[0, 1]] @ whd[[0, 1]])
    m[3] = mass
    m[4] = mass
    m[5] = mass
    return m


def trans(p):
    E = np.eye(4)
    E[0:3, 3] = p
    return E


# ----------------------------------------------------------------------------------------------
# +redmax/Body.m, BodyCuboid.m
# ----------------------------------------------------------------------------------------------
class Body:
    def __init__(self, density):
        """Body.m:26"""
        self.density = density
        self.damping = 0.0
        self.I_i = np.ones(6)
        self.E0_ji = np.eye(4)
        self.E0_ij = np.eye(4)
        self.E_wi = np.eye(4)
        self.E_iw = np.eye(4)
        self.E_ip = np.eye(4)
        self.A0_ij = np.eye(6)
        self.phi = np.zeros(6)
        self.joint = None
        self.idxM = None

    def setBodyTransform(self, E):
        """Body.m:46"""
        self.E0_ji = np.array(E, dtype=float)
        self.E0_ij = se3_inv(self.E0_ji)
        self.A0_ij = se3_Ad(self.E0_ij)

    def countDofs(self, scene):
        """Body.m:54"""
        nm = scene.nm
        self.idxM = nm + np.arange(6)
        scene.nm = nm + 6

    def update(self):
        """Body.m:70"""
        self.E_wi = self.joint.E_wj @ self.E0_ji
        self.E_iw = se3_inv(self.E_wi)
        self.E_ip = np.eye(4)
        if self.joint.parent is not None:
            self.E_ip = self.E_iw @ self.joint.parent.body.E_wi
        self.phi = self.A0_ij @ self.joint.V

    def computeMassGrav(self, grav, Mm, fm, Km=None, Dm=None):
        """Body.m:83 (one body; the `next` recursion is the caller's loop)"""
        rows = self.idxM
        M_i = np.diag(self.I_i)
        Mm[np.ix_(rows, rows)] = M_i
        adt = se3_ad(self.phi).T
        fcor = adt @ M_i @ self.phi
        R_wi = self.E_wi[0:3, 0:3]
        R_iw = R_wi.T
        fgrav = np.zeros(6)
        mass = M_i[3, 3]
        grav_i = R_iw @ grav
        fgrav[3:6] = mass * grav_i
        fm[rows] = fm[rows] + fcor + fgrav
        if Km is not None:
            Km[np.ix_(rows[3:6], rows[0:3])] += se3_brac(fgrav[3:6])
            e1 = se3_brac([1, 0, 0])
            e2 = se3_brac([0, 1, 0])
            e3 = se3_brac([0, 0, 1])
            z3 = np.zeros(3)
            Iw = self.I_i[0:3] * self.phi[0:3]
            mv = mass * self.phi[3:6]
            blk = np.vstack([
                np.column_stack([e1 @ Iw, e2 @ Iw, e3 @ Iw, e1 @ mv, e2 @ mv, e3 @ mv]),
                np.column_stack([e1 @ mv, e2 @ mv, e3 @ mv, z3, z3, z3]),
            ])
            Dm[np.ix_(rows, rows)] += adt @ M_i - blk

    def computeEnergies(self, grav, T, V):
        """Body.m:167"""
        T = T + 0.5 * self.phi @ (np.diag(self.I_i) @ self.phi)
        V = V - self.I_i[5] * (grav @ self.E_wi[0:3, 3])
        return T, V


class BodyCuboid(Body):
    def __init__(self, density, sides):
        """BodyCuboid.m:10"""
        super().__init__(density)
        self.sides = np.asarray(sides, dtype=float).reshape(3)

    def computeInertia_(self):
        """BodyCuboid.m:16"""
        self.I_i = se3_inertiaCuboid(self.sides, self.density)


# ----------------------------------------------------------------------------------------------
# +redmax/Joint.m and subclasses
# ----------------------------------------------------------------------------------------------
class Joint:
    def __init__(self, parent, body, ndof):
        """Joint.m:56"""
        self.parent = parent
        self.body = body
        self.children = []
        self.ndof = ndof
        self.q = np.zeros(ndof)
        self.qdot = np.zeros(ndof)
        self.q0 = np.zeros(ndof)
        self.qdot0 = np.zeros(ndof)
        self.q1 = np.zeros(ndof)
        self.qdot1 = np.zeros(ndof)
        self.qRest = np.zeros(ndof)
        self.qLimL = -1e8
        self.qLimU = 1e8
        self.qLimK = 1e8
        self.qLimD = 0.0
        self.tau = np.zeros(ndof)
        self.stiffness = 0.0
        self.damping = 0.0
        self.S = np.zeros((6, ndof))
        self.Sdot = np.zeros((6, ndof))
        self.V = np.zeros(6)
        self.E0_pj = None
        self.E0_jp = None
        self.idxR = None
        body.joint = self
        if parent is not None:
            parent.children.append(self)

    def setJointTransform(self, E):
        """Joint.m:95"""
        self.E0_pj = np.array(E, dtype=float)
        self.E0_jp = se3_inv(self.E0_pj)

    def setStiffness(self, s):
        self.stiffness = s

    def setDamping(self, d):
        self.damping = d

    def setLimitLower(self, v):
        self.qLimL = v

    def setLimitUpper(self, v):
        self.qLimU = v

    def setLimitStiffness(self, K):
        self.qLimK = K

    def setLimitDamping(self, D):
        self.qLimD = D

    def countDofs(self, scene):
        """Joint.m:149"""
        nr = scene.nr
        self.idxR = nr + np.arange(self.ndof)
        scene.nr = nr + self.ndof
        self.body.countDofs(scene)
        self.qRest = self.q.copy()

    def update(self, deriv=True):
        """Joint.m:382 (one joint; `this.next.update()` is the caller's loop, always deriv=true)"""
        n = self.ndof
        self.Q = np.eye(4)
        self.A = np.eye(6)
        self.Adot = np.zeros((6, 6))
        self.S = np.zeros((6, n))
        self.Sdot = np.zeros((6, n))
        if deriv:
            self.dAdq = np.zeros((6, 6, n))
            self.dAdotdq = np.zeros((6, 6, n))
            self.dSdq = np.zeros((6, n, n))
            self.dSdotdq = np.zeros((6, n, n))
        self.update_(deriv)
        self.invQ = se3_inv(self.Q)
        self.invA = se3_Ad(self.invQ)
        if self.E0_pj is None:
            self.E_pj = self.Q
        else:
            self.E_pj = self.E0_pj @ self.Q
        self.E_jp = se3_inv(self.E_pj)
        self.A_jp = se3_Ad(self.E_jp)
        if self.parent is None:
            E_wp = np.eye(4)
        else:
            E_wp = self.parent.E_wj
        self.E_wj = E_wp @ self.E_pj
        if self.ndof == 0:
            self.V = np.zeros(6)
        else:
            self.V = self.S @ self.qdot
        if self.parent is not None:
            self.V = self.V + self.A_jp @ self.parent.V
        if self.body is not None:
            self.body.update()

    def update_(self, deriv):
        pass

    def reparam_(self):
        """Joint.m:790 -- subclasses may re-parameterise q, qdot (Euler-chart joints); True if they did"""
        return False

    def resetChart_(self):
        """test aid, see Scene.reset"""

    def setAux0_(self):
        """Joint.m:795"""

    def setAux1_(self):
        """Joint.m:800"""

    def computeForce(self, fr, Kr=None, Dr=None):
        """Joint.m:437 (one joint)"""
        rows = self.idxR
        q = self.q[: self.ndof]
        qdot = self.qdot[: self.ndof]
        fr[rows] = fr[rows] + self.tau + self.stiffness * (self.qRest[: self.ndof] - q) - self.damping * qdot
        hitL = (q < self.qLimL).astype(float)
        hitU = (q > self.qLimU).astype(float)
        fr[rows] = fr[rows] + hitL * (self.qLimK * (self.qLimL - q) - self.qLimD * qdot)
        fr[rows] = fr[rows] + hitU * (self.qLimK * (self.qLimU - q) - self.qLimD * qdot)
        if Kr is not None:
            I = np.eye(self.ndof)
            ix = np.ix_(rows, rows)
            Kr[ix] = Kr[ix] - self.stiffness * I
            Dr[ix] = Dr[ix] - self.damping * I
            Kr[ix] = Kr[ix] - np.outer(hitL, hitL) * (self.qLimK * I)
            Kr[ix] = Kr[ix] - np.outer(hitU, hitU) * (self.qLimK * I)
            Dr[ix] = Dr[ix] - np.outer(hitL, hitL) * (self.qLimD * I)
            Dr[ix] = Dr[ix] - np.outer(hitU, hitU) * (self.qLimD * I)

    def computeJacobian2(self, J, Jdot):
        """Joint.m:494-533 (J, Jdot only; O(n^2))"""
        invQ = self.invQ
        Adot = self.Adot
        S = self.S
        Sdot = self.Sdot
        idxmI = self.body.idxM
        idxrI = self.idxR
        E0_BiJi = self.body.E0_ij
        A0_BiJi = self.body.A0_ij
        J[np.ix_(idxmI, idxrI)] = A0_BiJi @ S
        Jdot[np.ix_(idxmI, idxrI)] = A0_BiJi @ Sdot
        if self.parent is not None:
            idxmP = self.parent.body.idxM
            E0_JpBp = self.parent.body.E0_ji
            E0_JiJp = self.E0_jp
            E0_JiBp = E0_JiJp @ E0_JpBp
            E_BiBp = E0_BiJi @ invQ @ E0_JiBp
            A_BiBp = se3_Ad(E_BiBp)
            Aleft = -se3_Ad(E0_BiJi @ invQ)
            Aright = se3_Ad(invQ @ E0_JiBp)
            Adot_BiBp = Aleft @ Adot @ Aright
            jointA = self.parent
            while jointA is not None:
                idxrA = jointA.idxR
                JPA = J[np.ix_(idxmP, idxrA)]
                JdotPA = Jdot[np.ix_(idxmP, idxrA)]
                J[np.ix_(idxmI, idxrA)] = A_BiBp @ JPA
                Jdot[np.ix_(idxmI, idxrA)] = A_BiBp @ JdotPA + Adot_BiBp @ JPA
                jointA = jointA.parent

    def computeJacobian4(self, J, Jdot, dJdq, dJdotdq):
        """Joint.m:535-612 (J, Jdot, dJdq, dJdotdq; O(n^3))"""
        invQ = self.invQ
        invA = self.invA
        dAdq = self.dAdq
        Adot = self.Adot
        dAdotdq = self.dAdotdq
        S = self.S
        dSdq = self.dSdq
        Sdot = self.Sdot
        dSdotdq = self.dSdotdq
        idxmI = self.body.idxM
        idxrI = self.idxR
        E0_BiJi = self.body.E0_ij
        A0_BiJi = self.body.A0_ij
        J[np.ix_(idxmI, idxrI)] = A0_BiJi @ S
        Jdot[np.ix_(idxmI, idxrI)] = A0_BiJi @ Sdot
        for ii in range(self.ndof):
            dJdq[np.ix_(idxmI, idxrI, [idxrI[ii]])] = (A0_BiJi @ dSdq[:, :, ii])[:, :, None]
            dJdotdq[np.ix_(idxmI, idxrI, [idxrI[ii]])] = (A0_BiJi @ dSdotdq[:, :, ii])[:, :, None]
        if self.parent is not None:
            idxmP = self.parent.body.idxM
            E0_JpBp = self.parent.body.E0_ji
            E0_JiJp = self.E0_jp
            E0_JiBp = E0_JiJp @ E0_JpBp
            E_BiBp = E0_BiJi @ invQ @ E0_JiBp
            A_BiBp = se3_Ad(E_BiBp)
            dAdq_BiBp = np.zeros((6, 6, self.ndof))
            dAdotdq_BiBp = np.zeros((6, 6, self.ndof))
            Aleft = -se3_Ad(E0_BiJi @ invQ)
            Aright = se3_Ad(invQ @ E0_JiBp)
            Adot_BiBp = Aleft @ Adot @ Aright
            for ii in range(self.ndof):
                dAdq_ii = dAdq[:, :, ii]
                dAdotdq_ii = dAdotdq[:, :, ii]
                tmp1 = dAdq_ii @ invA @ Adot
                tmp2 = Adot @ invA @ dAdq_ii
                dAdq_BiBp[:, :, ii] = Aleft @ dAdq_ii @ Aright
                dAdotdq_BiBp[:, :, ii] = Aleft @ (dAdotdq_ii - tmp1 - tmp2) @ Aright
            i0, p0 = idxmI[0], idxmP[0]
            jointA = self.parent
            while jointA is not None:
                idxrA = jointA.idxR
                if len(idxrA) > 0:
                    a0, a1 = idxrA[0], idxrA[-1] + 1
                    JPA = J[p0:p0 + 6, a0:a1].copy()
                    JdotPA = Jdot[p0:p0 + 6, a0:a1].copy()
                    J[i0:i0 + 6, a0:a1] = A_BiBp @ JPA
                    Jdot[i0:i0 + 6, a0:a1] = A_BiBp @ JdotPA + Adot_BiBp @ JPA
                    for ii in range(len(idxrI)):
                        dAdq_BiBp_ii = dAdq_BiBp[:, :, ii]
                        dAdotdq_BiBp_ii = dAdotdq_BiBp[:, :, ii]
                        dJdq[i0:i0 + 6, a0:a1, idxrI[ii]] = dAdq_BiBp_ii @ JPA
                        dJdotdq[i0:i0 + 6, a0:a1, idxrI[ii]] = dAdq_BiBp_ii @ JdotPA + dAdotdq_BiBp_ii @ JPA
                    jointK = self.parent
                    while jointK is not None:
                        idxrK = jointK.idxR
                        for kk in range(len(idxrK)):
                            dJdqPAK = dJdq[p0:p0 + 6, a0:a1, idxrK[kk]]
                            dJdotdqPAK = dJdotdq[p0:p0 + 6, a0:a1, idxrK[kk]]
                            dJdq[i0:i0 + 6, a0:a1, idxrK[kk]] = A_BiBp @ dJdqPAK
                            dJdotdq[i0:i0 + 6, a0:a1, idxrK[kk]] = A_BiBp @ dJdotdqPAK + Adot_BiBp @ dJdqPAK
                        jointK = jointK.parent
                jointA = jointA.parent

    def computeEnergies(self, grav, T, V):
        """Joint.m:616 (one joint)"""
        T, V = self.body.computeEnergies(grav, T, V)
        q = self.q[: self.ndof]
        dq = q - self.qRest[: self.ndof]
        V = V + 0.5 * self.stiffness * (dq @ dq)
        hitL = (q < self.qLimL).astype(float)
        hitU = (q > self.qLimU).astype(float)
        dqL = hitL * (self.qLimL - q)
        dqU = hitU * (self.qLimU - q)
        V = V + 0.5 * self.qLimK * (dqL @ dqL + dqU @ dqU)
        return T, V


class JointRevolute(Joint):
    def __init__(self, parent, body, axis):
        """JointRevolute.m:12"""
        super().__init__(parent, body, 1)
        axis = np.asarray(axis, dtype=float).reshape(3)
        self.axis = axis / np.linalg.norm(axis)

    def update_(self, deriv):
        """JointRevolute.m:29"""
        q = self.q[0]
        qdot = self.qdot[0]
        a = self.axis
        R = se3_aaToMat(a, q)
        self.Q[0:3, 0:3] = R
        self.A = se3_Ad(self.Q)
        self.S = np.concatenate([a, np.zeros(3)]).reshape(6, 1)
        abrac = se3_brac(a)
        dRdq = R @ abrac
        Rdot = dRdq * qdot
        self.Adot[0:3, 0:3] = Rdot
        self.Adot[3:6, 3:6] = Rdot
        if deriv:
            self.dAdq[0:3, 0:3, 0] = dRdq
            self.dAdq[3:6, 3:6, 0] = dRdq
            d2Rdq2 = dRdq @ abrac
            tmp = d2Rdq2 * qdot
            self.dAdotdq[0:3, 0:3, 0] = tmp
            self.dAdotdq[3:6, 3:6, 0] = tmp


class JointFixed(Joint):
    def __init__(self, parent, body):
        """JointFixed.m:6"""
        super().__init__(parent, body, 0)


class JointPrismatic(Joint):
    def __init__(self, parent, body, axis):
        """JointPrismatic.m:12"""
        super().__init__(parent, body, 1)
        axis = np.asarray(axis, dtype=float).reshape(3)
        self.axis = axis / np.linalg.norm(axis)

    def update_(self, deriv):
        """JointPrismatic.m:28"""
        a = self.axis
        self.Q[0:3, 3] = a * self.q[0]
        self.A = se3_Ad(self.Q)
        self.S = np.concatenate([np.zeros(3), a]).reshape(6, 1)
        abrac = se3_brac(a)
        self.Adot[3:6, 0:3] = abrac * self.qdot[0]
        if deriv:
            self.dAdq[3:6, 0:3, 0] = abrac


class JointPlanar(Joint):
    def __init__(self, parent, body, plane=None):
        """JointPlanar.m:11 -- `plane` is 3 x 2 (columns = the two in-plane directions)."""
        super().__init__(parent, body, 2)
        if plane is None:
            plane = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]).T
        plane = np.array(plane, dtype=float).reshape(3, 2)
        self.plane = np.stack([plane[:, 0] / np.linalg.norm(plane[:, 0]), plane[:, 1] / np.linalg.norm(plane[:, 1])], axis=1)

    def update_(self, deriv):
        """JointPlanar.m:24"""
        B = self.plane
        self.Q[0:3, 3] = B @ self.q
        self.A = se3_Ad(self.Q)
        self.S = np.vstack([np.zeros((3, 2)), B])
        self.Adot[3:6, 0:3] = se3_brac(B @ self.qdot)
        if deriv:
            for k in range(2):
                self.dAdq[3:6, 0:3, k] = se3_brac(B[:, k])


class JointTranslational(Joint):
    def __init__(self, parent, body):
        """JointTranslational.m:12"""
        super().__init__(parent, body, 3)

    def update_(self, deriv):
        """JointTranslational.m:20"""
        self.Q[0:3, 3] = self.q
        self.A = se3_Ad(self.Q)
        self.S = np.vstack([np.zeros((3, 3)), np.eye(3)])
        self.Adot[3:6, 0:3] = se3_brac(self.qdot)
        if deriv:
            for k in range(3):
                ek = np.zeros(3)
                ek[k] = 1.0
                self.dAdq[3:6, 0:3, k] = se3_brac(ek)


class JointUniversal(Joint):
    """JointUniversal.m -- rotation about X then Y, R = X(q1) Y(q2); the entries below are the reference's generated
    closed forms (JointUniversal.m:122-182, `XY`), written out per matrix instead of through the t-temporaries."""

    def __init__(self, parent, body):
        super().__init__(parent, body, 2)

    def update_(self, deriv):
        """JointUniversal.m:19"""
        q1, q2 = self.q
        qd1, qd2 = self.qdot
        c1, s1, c2, s2 = math.cos(q1), math.sin(q1), math.cos(q2), math.sin(q2)
        R = np.array([[c2, 0.0, s2], [s1 * s2, c1, -c2 * s1], [-c1 * s2, s1, c1 * c2]])
        dR1 = np.array([[0.0, 0.0, 0.0], [c1 * s2, -s1, -c1 * c2], [s1 * s2, c1, -c2 * s1]])
        dR2 = np.array([[-s2, 0.0, c2], [c2 * s1, 0.0, s1 * s2], [-c1 * c2, 0.0, -c1 * s2]])
        Rdot = dR1 * qd1 + dR2 * qd2
        self.Q[0:3, 0:3] = R
        self.A = se3_Ad(self.Q)
        self.Adot[0:3, 0:3] = Rdot
        self.Adot[3:6, 3:6] = Rdot
        self.S[0:3, 0:2] = [[c2, 0.0], [0.0, 1.0], [s2, 0.0]]
        self.Sdot[0:3, 0:2] = [[-s2 * qd2, 0.0], [0.0, 0.0], [c2 * qd2, 0.0]]
        if deriv:
            # d(Rdot)/dq_i = d2R/dq_i dq_1 qd1 + d2R/dq_i dq_2 qd2
            d11 = np.array([[0.0, 0.0, 0.0], [-s1 * s2, -c1, c2 * s1], [c1 * s2, -s1, -c1 * c2]])
            d12 = np.array([[0.0, 0.0, 0.0], [c1 * c2, 0.0, c1 * s2], [c2 * s1, 0.0, s1 * s2]])
            d22 = np.array([[-c2, 0.0, -s2], [-s1 * s2, 0.0, c2 * s1], [c1 * s2, 0.0, -c1 * c2]])
            dRdot = [d11 * qd1 + d12 * qd2, d12 * qd1 + d22 * qd2]
            for i, (dR, dRd) in enumerate(((dR1, dRdot[0]), (dR2, dRdot[1]))):
                self.dAdq[0:3, 0:3, i] = dR
                self.dAdq[3:6, 3:6, i] = dR
                self.dAdotdq[0:3, 0:3, i] = dRd
                self.dAdotdq[3:6, 3:6, i] = dRd
            self.dSdq[0:3, 0:2, 1] = [[-s2, 0.0], [0.0, 0.0], [c2, 0.0]]
            self.dSdotdq[0:3, 0:2, 1] = [[-c2 * qd2, 0.0], [0.0, 0.0], [-s2 * qd2, 0.0]]


class JointFree2D(Joint):
    """JointFree2D.m -- 2D free joint in XY (only used to reach the scene-11 pin for ForceGroundCuboid)."""

    def __init__(self, parent, body):
        super().__init__(parent, body, 3)

    def update_(self, deriv):
        """JointFree2D.m:20"""
        n = self.ndof
        p = np.array([self.q[0], self.q[1], 0.0])
        r = self.q[2]
        pdot = self.qdot[0:2]
        rdot = self.qdot[2]
        R = np.eye(3)
        s = math.sin(r)
        c = math.cos(r)
        R[0:2, 0:2] = [[c, -s], [s, c]]
        self.Q = np.eye(4)
        self.Q[0:3, 0:3] = R
        self.Q[0:3, 3] = p
        self.A = se3_Ad(self.Q)
        self.S[2, 2] = 1
        self.S[3:5, 0:2] = [[c, s], [-s, c]]
        pbrac = se3_brac(p)
        dRdq = np.zeros((3, 3))
        dRdq[0:2, 0:2] = [[-s, -c], [c, -s]]
        pdotbrac = se3_brac([pdot[0], pdot[1], 0.0])
        Rdot = dRdq * rdot
        self.Adot[0:3, 0:3] = Rdot
        self.Adot[3:6, 3:6] = Rdot
        self.Adot[3:6, 0:3] = pdotbrac @ R + pbrac @ Rdot
        self.Sdot[3:5, 0:2] = np.array([[-s, c], [-c, -s]]) * rdot
        if deriv:
            self.dSdotdq[3:5, 0:2, 2] = np.array([[-c, -s], [s, -c]]) * rdot
            e1brac = se3_brac([1, 0, 0])
            e2brac = se3_brac([0, 1, 0])
            self.dAdq[3:6, 0:3, 0] = e1brac @ R
            self.dAdq[3:6, 0:3, 1] = e2brac @ R
            self.dAdq[0:3, 0:3, 2] = dRdq
            self.dAdq[3:6, 0:3, 2] = pbrac @ dRdq
            self.dAdq[3:6, 3:6, 2] = dRdq
            self.dSdq = np.zeros((6, 3, n))
            self.dSdq[3:5, 0:2, 2] = [[-s, c], [-c, -s]]
            dRdotdq = np.zeros((3, 3))
            dRdotdq[0:2, 0:2] = np.array([[-c, s], [-s, -c]]) * rdot
            self.dAdotdq[3:6, 0:3, 0] = e1brac @ Rdot
            self.dAdotdq[3:6, 0:3, 1] = e2brac @ Rdot
            self.dAdotdq[0:3, 0:3, 2] = dRdotdq
            self.dAdotdq[3:6, 3:6, 2] = dRdotdq
            self.dAdotdq[3:6, 0:3, 2] = pdotbrac @ dRdq + pbrac @ dRdotdq


# Euler charts of JointSpherical.m:5-16, in the reference's numbering 1..12: the three rotation axes (0=X, 1=Y, 2=Z) of
# R = R_a(q1) R_b(q2) R_c(q3).
EULER_CHARTS = {1: (0, 1, 0), 2: (0, 2, 0), 3: (1, 2, 1), 4: (1, 0, 1), 5: (2, 0, 2), 6: (2, 1, 2),
                7: (0, 1, 2), 8: (0, 2, 1), 9: (1, 2, 0), 10: (1, 0, 2), 11: (2, 0, 1), 12: (2, 1, 0)}
EULER_CHART_NAMES = {k: ''.join('XYZ'[i] for i in v) for k, v in EULER_CHARTS.items()}
CHART_XYZ = 7


def _axis_rot(axis, angle):
    e = np.zeros(3)
    e[axis] = 1.0
    return se3_aaToMat(e, angle)


def euler_chart(chart, q, qdot):
    """JointSpherical.getEuler (JointSpherical.m:154-183 and the generated per-chart bodies :342-1806), restated from the
    definition the generated closed forms implement instead of term by term: R = R_a(q1) R_b(q2) R_c(q3),
    body angular velocity omega = T qdot with T = [(R_b R_c)' e_a, R_c' e_b, e_c], and their analytic first / second
    derivatives.  Checked against the reference's own entries for chart XYZ (R at JointSpherical.m:1162) and, for all
    twelve charts, by finite differences (tests/test_oracle_pins.py).
    Returns R, dRdq[3,3,3], Rdot, dRdotdq, T, detT, dTdq, Tdot, dTdotdq (third index = derivative direction)."""
    a, b, c = EULER_CHARTS[chart]
    E = [np.zeros(3), np.zeros(3), np.zeros(3)]
    E[0][a] = E[1][b] = E[2][c] = 1.0
    br = [se3_brac(E[0]), se3_brac(E[1]), se3_brac(E[2])]
    Rs = [_axis_rot(a, q[0]), _axis_rot(b, q[1]), _axis_rot(c, q[2])]

    def prod(ins):
        # R_a [a]^i0 R_b [b]^i1 R_c [c]^i2
        M = np.eye(3)
        for k in range(3):
            M = M @ Rs[k]
            for _ in range(ins[k]):
                M = M @ br[k]
        return M

    R = prod((0, 0, 0))
    dRdq = np.zeros((3, 3, 3))
    d2R = np.zeros((3, 3, 3, 3))
    for i in range(3):
        ins = [0, 0, 0]
        ins[i] = 1
        dRdq[:, :, i] = prod(ins)
        for j in range(3):
            ins2 = list(ins)
            ins2[j] += 1
            d2R[:, :, i, j] = prod(ins2)
    Rdot = dRdq @ qdot
    dRdotdq = d2R @ qdot  # d(Rdot)/dq_i = sum_j d2R/dq_i dq_j qdot_j
    # T and its derivatives: column 0 = R_c' R_b' e_a, column 1 = R_c' e_b, column 2 = e_c;  d(R')/dq = -[e] R'
    RbT, RcT = Rs[1].T, Rs[2].T
    T = np.column_stack([RcT @ RbT @ E[0], RcT @ E[1], E[2]])
    dTdq = np.zeros((3, 3, 3))
    d2T = np.zeros((3, 3, 3, 3))
    dTdq[:, 0, 1] = -RcT @ br[1] @ RbT @ E[0]
    dTdq[:, 0, 2] = -br[2] @ RcT @ RbT @ E[0]
    dTdq[:, 1, 2] = -br[2] @ RcT @ E[1]
    d2T[:, 0, 1, 1] = RcT @ br[1] @ br[1] @ RbT @ E[0]
    d2T[:, 0, 1, 2] = d2T[:, 0, 2, 1] = br[2] @ RcT @ br[1] @ RbT @ E[0]
    d2T[:, 0, 2, 2] = br[2] @ br[2] @ RcT @ RbT @ E[0]
    d2T[:, 1, 2, 2] = br[2] @ br[2] @ RcT @ E[1]
    Tdot = dTdq @ qdot
    dTdotdq = d2T @ qdot
    # det T in the reference's closed forms (detS = -sin q2 for the proper Euler charts, JointSpherical.m:452...1072;
    # +cos q2 for XYZ, YZX, ZXY and -cos q2 for XZY, YXZ, ZYX, :1190...1795): charts that share their middle angle (XYX / XZX,
    # YZY / YXY, ZXZ / ZYZ) tie exactly in reparam_'s max(min(abs(detTs))), and MATLAB's max takes the first
    if chart <= 6:
        detT = -math.sin(q[1])
    else:
        detT = math.cos(q[1]) if chart in (7, 9, 11) else -math.cos(q[1])
    return R, dRdq, Rdot, dRdotdq, T, detT, dTdq, Tdot, dTdotdq


def euler_chart_inv(chart, R):
    """JointSpherical.getEulerInv (JointSpherical.m:186-214, per-chart bodies :1809-1964): the twelve closed forms share
    one pattern in terms of the axes (a, b, c) and the parity eps of (a, b, third axis); gimbal lock returns NaNs."""
    a, b, c3 = EULER_CHARTS[chart]
    if a == c3:  # proper Euler a-b-a: q2 = acos(R_aa) in (0, pi)
        c = 3 - a - b
        eps = 1.0 if (b - a) % 3 == 1 else -1.0
        raa = R[a, a]
        if not (-1.0 < raa < 1.0):
            return np.full(3, np.nan)
        return np.array([math.atan2(R[b, a], -eps * R[c, a]), math.acos(raa), math.atan2(R[a, b], eps * R[a, c])])
    c = c3  # Tait-Bryan a-b-c: q2 = asin(eps R_ac) in (-pi/2, pi/2)
    eps = 1.0 if (b - a) % 3 == 1 else -1.0
    rac = R[a, c]
    if not (-1.0 < rac < 1.0):
        return np.full(3, np.nan)
    return np.array([math.atan2(-eps * R[b, c], R[c, c]), math.asin(eps * rac), math.atan2(-eps * R[a, b], R[a, a])])


class JointSpherical(Joint):
    """JointSpherical.m -- ball joint in Euler angles with twelve coordinate charts and re-parameterisation."""

    def __init__(self, parent, body):
        """JointSpherical.m:30"""
        super().__init__(parent, body, 3)
        self.chart = CHART_XYZ
        self.chart0 = None
        self.chart1 = None
        self.switches = []  # (old chart, new chart) of every re-parameterisation, for the tests

    def setAux0_(self):
        """JointSpherical.m:53"""
        self.chart0 = self.chart

    def setAux1_(self):
        """JointSpherical.m:58"""
        self.chart1 = self.chart

    def reparam_(self):
        """JointSpherical.m:63-103"""
        R, _, _, _, Told, detTold, _, _, _ = euler_chart(self.chart, self.q, self.qdot)
        if abs(detTold) > 0.5:
            return False
        if self.chart1 is None:
            # the reference reaches getEuler with an empty chart1 here (BDF1 never calls setQ1; JointFree3D never forwards
            # setAux1_ to its inner joint) and stops with an unassigned-output error
            raise RuntimeError('JointSpherical.reparam_: chart1 unset (the reference errors at JointSpherical.m:73)')
        z = np.zeros(3)
        R1 = euler_chart(self.chart1, self.q1, z)[0]
        detTs = np.zeros((2, 12))
        for k in range(1, 13):
            qk = euler_chart_inv(k, R)
            detTs[0, k - 1] = euler_chart(k, qk, z)[5] if np.isfinite(qk).all() else np.nan
            q1k = euler_chart_inv(k, R1)
            detTs[1, k - 1] = euler_chart(k, q1k, z)[5] if np.isfinite(q1k).all() else np.nan
        detTs[np.isnan(detTs)] = 0.0
        old = self.chart
        self.chart = int(np.argmax(np.min(np.abs(detTs), axis=0))) + 1  # first maximum, as MATLAB's max
        self.switches.append((old, self.chart))
        self.q = euler_chart_inv(self.chart, R)
        Tnew = euler_chart(self.chart, self.q, z)[4]
        self.qdot = np.linalg.solve(Tnew, Told @ self.qdot)
        Told1 = euler_chart(self.chart1, self.q1, z)[4]
        self.chart1 = self.chart
        self.q1 = euler_chart_inv(self.chart1, R1)
        Tnew1 = euler_chart(self.chart1, self.q1, z)[4]
        self.qdot1 = np.linalg.solve(Tnew1, Told1 @ self.qdot1)
        return True

    def resetChart_(self):
        self.chart = CHART_XYZ
        self.chart0 = self.chart1 = None
        self.switches = []

    def update_(self, deriv):
        """JointSpherical.m:106-130"""
        R, dRdq, Rdot, dRdotdq, T, _, dTdq, Tdot, dTdotdq = euler_chart(self.chart, self.q, self.qdot)
        self.Q[0:3, 0:3] = R
        self.A[0:3, 0:3] = R
        self.A[3:6, 3:6] = R
        self.Adot[0:3, 0:3] = Rdot
        self.Adot[3:6, 3:6] = Rdot
        self.S[0:3, 0:3] = T
        self.Sdot[0:3, 0:3] = Tdot
        if deriv:
            for k in range(3):
                self.dAdq[0:3, 0:3, k] = dRdq[:, :, k]
                self.dAdq[3:6, 3:6, k] = dRdq[:, :, k]
                self.dAdotdq[0:3, 0:3, k] = dRdotdq[:, :, k]
                self.dAdotdq[3:6, 3:6, k] = dRdotdq[:, :, k]
                self.dSdq[0:3, :, k] = dTdq[:, :, k]
                self.dSdotdq[0:3, :, k] = dTdotdq[:, :, k]


class _InnerBody:
    """stand-in for the body handle JointFree3D passes to its two inner joints (they never touch it)"""
    joint = None


class JointFree3D(Joint):
    """JointFree3D.m -- free joint composed of a translational joint (q(1:3) = p) and a spherical joint (q(4:6))."""

    def __init__(self, parent, body):
        """JointFree3D.m:16-21"""
        super().__init__(parent, body, 6)
        self.joint1 = JointTranslational(None, _InnerBody())
        self.joint2 = JointSpherical(None, _InnerBody())
        body.joint = self

    def reparam_(self):
        """JointFree3D.m:27-31"""
        did = self.joint2.reparam_()
        self.q[3:6] = self.joint2.q
        self.qdot[3:6] = self.joint2.qdot
        return did

    def resetChart_(self):
        self.joint2.resetChart_()

    def update_(self, deriv):
        """JointFree3D.m:34-122"""
        j1, j2 = self.joint1, self.joint2
        j1.q = self.q[0:3].copy()
        j2.q = self.q[3:6].copy()
        j1.qdot = self.qdot[0:3].copy()
        j2.qdot = self.qdot[3:6].copy()
        p = j1.q
        pdot = j1.qdot
        rdot = j2.qdot
        R, dRdr, Rdot, dRdotdr, T, _, dTdr, Tdot, dTdotdr = euler_chart(j2.chart, j2.q, j2.qdot)
        self.Q[0:3, 0:3] = R
        self.Q[0:3, 3] = p
        self.A[0:3, 0:3] = R
        self.A[3:6, 3:6] = R
        pbrac = se3_brac(p)
        self.A[3:6, 0:3] = pbrac @ R
        pdotbrac = se3_brac(pdot)
        self.Adot[0:3, 0:3] = Rdot
        self.Adot[3:6, 3:6] = Rdot
        self.Adot[3:6, 0:3] = pdotbrac @ R + pbrac @ Rdot
        self.S[3:6, 0:3] = R.T
        self.S[0:3, 3:6] = T
        self.Sdot[3:6, 0:3] = Rdot.T
        self.Sdot[0:3, 3:6] = Tdot
        if deriv:
            tmp = se3_brac(T @ rdot)
            for k in range(3):
                ek = np.zeros(3)
                ek[k] = 1.0
                ekbrac = se3_brac(ek)
                dRk = dRdr[:, :, k]
                dRdk = dRdotdr[:, :, k]
                self.dAdq[0:3, 0:3, 3 + k] = dRk
                self.dAdq[3:6, 3:6, 3 + k] = dRk
                self.dAdq[3:6, 0:3, k] = ekbrac @ R
                self.dAdq[3:6, 0:3, 3 + k] = pbrac @ dRk
                self.dAdotdq[3:6, 0:3, k] = ekbrac @ Rdot
                self.dAdotdq[0:3, 0:3, 3 + k] = dRdk
                self.dAdotdq[3:6, 3:6, 3 + k] = dRdk
                self.dAdotdq[3:6, 0:3, 3 + k] = pdotbrac @ dRk + pbrac @ dRdk
                self.dSdq[3:6, 0:3, 3 + k] = dRk.T
                self.dSdq[0:3, 3:6, 3 + k] = dTdr[:, :, k]
                self.dSdotdq[3:6, 0:3, 3 + k] = -se3_brac(dTdr[:, :, k] @ rdot) @ R.T - tmp @ dRk.T
                self.dSdotdq[0:3, 3:6, 3 + k] = dTdotdr[:, :, k]


# ----------------------------------------------------------------------------------------------
# +redmax/Force.m, ForceNull.m, ForceGroundCuboid.m
# ----------------------------------------------------------------------------------------------
class Force:
    def init_(self):
        pass

    def computeValues_(self, fr, fm, Kr=None, Km=None, Dr=None, Dm=None):
        pass

    def computeEnergy_(self, V):
        return V


class ForceNull(Force):
    pass


class ForceSpringGeneric(Force):
    """ForceSpringGeneric.m -- generic spring along the line between two body points (a body may be None = world)."""

    def __init__(self, body1, x_1, body2, x_2):
        """ForceSpringGeneric.m:15"""
        self.body1 = body1
        self.body2 = body2
        self.x_1 = np.asarray(x_1, dtype=float).reshape(3)
        self.x_2 = np.asarray(x_2, dtype=float).reshape(3)

    def _state(self):
        """ForceSpringGeneric.m:37-66"""
        E1 = np.eye(4) if self.body1 is None else self.body1.E_wi
        E2 = np.eye(4) if self.body2 is None else self.body2.E_wi
        phi1 = np.zeros(6) if self.body1 is None else self.body1.phi
        phi2 = np.zeros(6) if self.body2 is None else self.body2.phi
        G1 = se3_Gamma(self.x_1)
        G2 = se3_Gamma(self.x_2)
        R1, R2, p1, p2 = E1[0:3, 0:3], E2[0:3, 0:3], E1[0:3, 3], E2[0:3, 3]
        xw1 = R1 @ self.x_1 + p1
        xw2 = R2 @ self.x_2 + p2
        vw1 = R1 @ (G1 @ phi1)
        vw2 = R2 @ (G2 @ phi2)
        dx = xw2 - xw1
        l = np.linalg.norm(dx)
        dv = vw2 - vw1
        ldot = (dx @ dv) / l
        return R1, R2, p1, p2, G1, G2, phi1, phi2, xw1, xw2, dx, l, dv, ldot

    def computeValues_(self, fr, fm, Kr=None, Km=None, Dr=None, Dm=None):
        """ForceSpringGeneric.m:35-143"""
        R1, R2, p1, p2, G1, G2, phi1, phi2, xw1, xw2, dx, l, dv, ldot = self._state()
        xl1, xl2 = self.x_1, self.x_2
        _, fs, dfsdl, dfsdldot = self.computeSpringForce(l, ldot)
        fx_1 = G1.T @ (R1.T @ dx)
        fx_2 = -G2.T @ (R2.T @ dx)
        fx = np.concatenate([fx_1, fx_2])
        f = (fs / l) * fx
        b1, b2 = self.body1 is not None, self.body2 is not None
        if b1:
            fm[self.body1.idxM] += f[0:6]
        if b2:
            fm[self.body2.idxM] += f[6:12]
        if Km is None:
            return
        I = np.eye(3)
        A = np.hstack([-R1 @ G1, R2 @ G2])  # 3 x 12
        dldq = (dx / l) @ A
        dldotdq = (((dx @ dx) * I - np.outer(dx, dx)) / l ** 3 @ dv) @ A
        for ax in range(3):  # rotation columns of body 1 and body 2 (ForceSpringGeneric.m:88-90)
            e = np.zeros(3)
            e[ax] = 1.0
            eb = se3_brac(e)
            dldotdq[ax] += (dx / l) @ (-R1 @ eb @ (G1 @ phi1))
            dldotdq[6 + ax] += (dx / l) @ (R2 @ eb @ (G2 @ phi2))
        dfsdq = dfsdl * dldq + dfsdldot * dldotdq
        K1 = np.outer(fx, dfsdq / l - fs / l ** 2 * dldq)
        K2 = np.zeros((12, 12))
        x1b = se3_brac(xl1)
        x2b = se3_brac(xl2)
        R2R1 = R2.T @ R1
        R1R2 = R2R1.T
        K2[3:6, 0:3] = se3_brac(R1.T @ (p1 - xw2))
        K2[0:3, 0:3] = x1b @ K2[3:6, 0:3]
        K2[9:12, 0:3] = R2R1 @ x1b
        K2[6:9, 0:3] = x2b @ K2[9:12, 0:3]
        K2[3:6, 3:6] = I
        K2[0:3, 3:6] = x1b
        K2[9:12, 3:6] = -R2R1
        K2[6:9, 3:6] = x2b @ K2[9:12, 3:6]
        K2[3:6, 6:9] = R1R2 @ x2b
        K2[0:3, 6:9] = x1b @ K2[3:6, 6:9]
        K2[9:12, 6:9] = se3_brac(R2.T @ (p2 - xw1))
        K2[6:9, 6:9] = x2b @ K2[9:12, 6:9]
        K2[3:6, 9:12] = -R1R2
        K2[0:3, 9:12] = x1b @ K2[3:6, 9:12]
        K2[9:12, 9:12] = I
        K2[6:9, 9:12] = x2b
        K2 = -(fs / l) * K2
        K = K1 + K2
        d_w = dfsdldot * dx / l ** 2
        D = -np.outer(fx, np.concatenate([d_w @ R1 @ G1, -(d_w @ R2 @ G2)]))
        if b1:
            i1 = self.body1.idxM
            Km[np.ix_(i1, i1)] += K[0:6, 0:6]
            Dm[np.ix_(i1, i1)] += D[0:6, 0:6]
        if b2:
            i2 = self.body2.idxM
            Km[np.ix_(i2, i2)] += K[6:12, 6:12]
            Dm[np.ix_(i2, i2)] += D[6:12, 6:12]
        if b1 and b2:
            Km[np.ix_(i1, i2)] += K[0:6, 6:12]
            Km[np.ix_(i2, i1)] += K[6:12, 0:6]
            Dm[np.ix_(i1, i2)] += D[0:6, 6:12]
            Dm[np.ix_(i2, i1)] += D[6:12, 0:6]

    def computeEnergy_(self, V):
        """ForceSpringGeneric.m:146-177"""
        st = self._state()
        return V + self.computeSpringForce(st[11], st[13])[0]


class ForceSpringDamper(ForceSpringGeneric):
    """ForceSpringDamper.m -- damped spring with rest length L (taken from the initial configuration unless set)."""

    def __init__(self, body1, x_1, body2, x_2):
        super().__init__(body1, x_1, body2, x_2)
        self.stiffness = 1.0
        self.damping = 1.0
        self.L = 0.0

    def setStiffness(self, stiffness):
        self.stiffness = stiffness

    def setDamping(self, damping):
        self.damping = damping

    def setRetLength(self, L):
        """ForceSpringDamper.m:31 (sic)"""
        self.L = L

    def init_(self):
        """ForceSpringDamper.m:38-62"""
        if self.L > 0:
            return
        self.L = self._state()[11]

    def computeSpringForce(self, l, ldot):
        """ForceSpringDamper.m:65-72: positive force contracts the spring"""
        strain = (l - self.L) / self.L
        dstrain = ldot / self.L
        V = (self.stiffness / 2) * strain ** 2 * self.L
        f = self.stiffness * strain + self.damping * dstrain
        return V, f, self.stiffness / self.L, self.damping / self.L


class ForceSpringMultiPointGeneric(Force):
    """ForceSpringMultiPointGeneric.m -- generic spring routed along a sequence of body points (a body may be None = world)."""

    def __init__(self):
        self.bodies = []
        self.xls = []

    def addBodyPoint(self, body, xl):
        """ForceSpringMultiPointGeneric.m:20"""
        self.bodies.append(body)
        self.xls.append(np.asarray(xl, dtype=float).reshape(3))

    def _points(self):
        """ForceSpringMultiPointGeneric.m:31-53"""
        out = []
        for body, xl in zip(self.bodies, self.xls):
            E = np.eye(4) if body is None else body.E_wi
            phi = np.zeros(6) if body is None else body.phi
            G = se3_Gamma(xl)
            R, p = E[0:3, 0:3], E[0:3, 3]
            out.append(dict(R=R, p=p, G=G, phi=phi, xl=xl, xw=R @ xl + p, vw=R @ (G @ phi)))
        return out

    def _length(self, pts):
        l = 0.0
        ldot = 0.0
        for k in range(len(pts) - 1):
            dx = pts[k + 1]['xw'] - pts[k]['xw']
            dv = pts[k + 1]['vw'] - pts[k]['vw']
            dxlen = np.linalg.norm(dx)
            l += dxlen
            ldot += (dx @ dv) / dxlen
        return l, ldot

    def computeValues_(self, fr, fm, Kr=None, Km=None, Dr=None, Dm=None):
        """ForceSpringMultiPointGeneric.m:29-190"""
        pts = self._points()
        npts = len(pts)
        fn = np.zeros(6 * npts)
        for k in range(npts - 1):
            a, b = pts[k], pts[k + 1]
            dx = b['xw'] - a['xw']
            dxlen = np.linalg.norm(dx)
            fx = np.concatenate([a['G'].T @ (a['R'].T @ dx), -(b['G'].T @ (b['R'].T @ dx))])
            fn[6 * k:6 * k + 12] += fx / dxlen
        l, ldot = self._length(pts)
        _, fs, dfsdl, dfsdldot = self.computeSpringForce(l, ldot)
        f = fs * fn
        for k, body in enumerate(self.bodies):
            if body is not None:
                fm[body.idxM] += f[6 * k:6 * k + 6]
        if Km is None:
            return
        I = np.eye(3)
        Kn = np.zeros((6 * npts, 6 * npts))
        dfsdq = np.zeros(6 * npts)
        dfsdqdot = np.zeros(6 * npts)
        for k in range(npts - 1):
            a, b = pts[k], pts[k + 1]
            sl = slice(6 * k, 6 * k + 12)
            R1, R2, G1, G2 = a['R'], b['R'], a['G'], b['G']
            dx = b['xw'] - a['xw']
            dv = b['vw'] - a['vw']
            dxlen = np.linalg.norm(dx)
            dxnor = dx / dxlen
            A = np.hstack([-R1 @ G1, R2 @ G2])
            dldq = dxnor @ A
            dldotdq = ((I - np.outer(dxnor, dxnor)) / dxlen @ dv) @ A
            for ax in range(3):
                e = np.zeros(3)
                e[ax] = 1.0
                eb = se3_brac(e)
                dldotdq[ax] += dxnor @ (-R1 @ eb @ (G1 @ a['phi']))
                dldotdq[6 + ax] += dxnor @ (R2 @ eb @ (G2 @ b['phi']))
            dfsdq[sl] += dfsdl * dldq + dfsdldot * dldotdq
            fx = np.concatenate([G1.T @ (R1.T @ dx), -(G2.T @ (R2.T @ dx))])
            d = -dx / dxlen ** 3
            K1 = np.outer(fx, np.concatenate([d @ R1 @ G1, -(d @ R2 @ G2)]))
            K2 = np.zeros((12, 12))
            x1b = se3_brac(a['xl'])
            x2b = se3_brac(b['xl'])
            R2R1 = R2.T @ R1
            R1R2 = R2R1.T
            K2[3:6, 0:3] = se3_brac(R1.T @ (a['p'] - b['xw']))
            K2[0:3, 0:3] = x1b @ K2[3:6, 0:3]
            K2[9:12, 0:3] = R2R1 @ x1b
            K2[6:9, 0:3] = x2b @ K2[9:12, 0:3]
            K2[3:6, 3:6] = I
            K2[0:3, 3:6] = x1b
            K2[9:12, 3:6] = -R2R1
            K2[6:9, 3:6] = x2b @ K2[9:12, 3:6]
            K2[3:6, 6:9] = R1R2 @ x2b
            K2[0:3, 6:9] = x1b @ K2[3:6, 6:9]
            K2[9:12, 6:9] = se3_brac(R2.T @ (b['p'] - a['xw']))
            K2[6:9, 6:9] = x2b @ K2[9:12, 6:9]
            K2[3:6, 9:12] = -R1R2
            K2[0:3, 9:12] = x1b @ K2[3:6, 9:12]
            K2[9:12, 9:12] = I
            K2[6:9, 9:12] = x2b
            Kn[sl, sl] += K1 + K2 / dxlen
            d = dfsdldot * dxnor
            dfsdqdot[6 * k:6 * k + 6] -= d @ R1 @ G1
            dfsdqdot[6 * k + 6:6 * k + 12] += d @ R2 @ G2
        K = np.outer(fn, dfsdq) - fs * Kn
        D = np.outer(fn, dfsdqdot)
        for k1, b1 in enumerate(self.bodies):
            if b1 is None:
                continue
            s1 = slice(6 * k1, 6 * k1 + 6)
            Km[np.ix_(b1.idxM, b1.idxM)] += K[s1, s1]
            Dm[np.ix_(b1.idxM, b1.idxM)] += D[s1, s1]
            for k2 in range(k1 + 1, npts):
                b2 = self.bodies[k2]
                if b2 is None:
                    continue
                s2 = slice(6 * k2, 6 * k2 + 6)
                Km[np.ix_(b1.idxM, b2.idxM)] += K[s1, s2]
                Km[np.ix_(b2.idxM, b1.idxM)] += K[s2, s1]
                Dm[np.ix_(b1.idxM, b2.idxM)] += D[s1, s2]
                Dm[np.ix_(b2.idxM, b1.idxM)] += D[s2, s1]

    def computeEnergy_(self, V):
        """ForceSpringMultiPointGeneric.m:193-232"""
        l, ldot = self._length(self._points())
        return V + self.computeSpringForce(l, ldot)[0]


class ForceCable(ForceSpringMultiPointGeneric):
    """ForceCable.m -- a cable routed through several points: pulls only when stretched beyond its rest length."""

    def __init__(self):
        super().__init__()
        self.stiffness = 1.0
        self.damping = 1.0
        self.L = 0.0

    def setStiffness(self, stiffness):
        self.stiffness = stiffness

    def setDamping(self, damping):
        self.damping = damping

    def setRetLength(self, L):
        self.L = L

    def init_(self):
        """ForceCable.m:36-63"""
        if self.L > 0:
            return
        self.L = self._length(self._points())[0]

    def computeSpringForce(self, l, ldot):
        """ForceCable.m:66-81"""
        strain = (l - self.L) / self.L
        dstrain = ldot / self.L
        if strain > 0:
            return ((self.stiffness / 2) * strain ** 2 * self.L, self.stiffness * strain + self.damping * dstrain,
                    self.stiffness / self.L, self.damping / self.L)
        return 0.0, 0.0, 0.0, 0.0


class ForcePointPoint(Force):
    """ForcePointPoint.m -- linear zero-rest-length spring/damper between two body points (a body may be None = world)."""

    def __init__(self, body1, x_1, body2, x_2):
        """ForcePointPoint.m:15"""
        self.body1 = body1
        self.body2 = body2
        self.x_1 = np.asarray(x_1, dtype=float).reshape(3)
        self.x_2 = np.asarray(x_2, dtype=float).reshape(3)
        self.stiffness = 1.0
        self.damping = 0.0

    def setStiffness(self, stiffness):
        self.stiffness = stiffness

    def setDamping(self, damping):
        self.damping = damping

    def _point(self, body, xl):
        """world position / velocity of a body point, local velocity, R, p, Gamma (ForcePointPoint.m:50-73)"""
        if body is None:
            return xl, np.zeros(3), None, None, None, None
        E = body.E_wi
        R = E[0:3, 0:3]
        p = E[0:3, 3]
        G = se3_Gamma(xl)
        vl = G @ body.phi
        return R @ xl + p, R @ vl, vl, R, p, G

    def computeValues_(self, fr, fm, Kr=None, Km=None, Dr=None, Dm=None):
        """ForcePointPoint.m:48-113"""
        xw1, vw1, vl1, R1, p1, G1 = self._point(self.body1, self.x_1)
        xw2, vw2, vl2, R2, p2, G2 = self._point(self.body2, self.x_2)
        dx = xw2 - xw1
        dv = vw2 - vw1
        I = np.eye(3)
        Z = np.zeros((3, 3))
        ks = self.stiffness
        kd = self.damping
        f = ks * dx + kd * dv
        b1, b2 = self.body1 is not None, self.body2 is not None
        if b1:
            idx1 = self.body1.idxM
            fm[idx1] += G1.T @ (R1.T @ f)
        if b2:
            idx2 = self.body2.idxM
            fm[idx2] -= G2.T @ (R2.T @ f)
        if Km is None:
            return
        if b1:
            i11 = np.ix_(idx1, idx1)
            Km[i11] += ks * (G1.T @ np.hstack([se3_brac(R1.T @ (xw2 - p1)), -I]))
            Km[i11] += kd * (G1.T @ np.hstack([se3_brac(R1.T @ vw2), Z]))
            Dm[i11] -= kd * (G1.T @ G1)
        if b2:
            i22 = np.ix_(idx2, idx2)
            Km[i22] += ks * (G2.T @ np.hstack([se3_brac(R2.T @ (xw1 - p2)), -I]))
            Km[i22] += kd * (G2.T @ np.hstack([se3_brac(R2.T @ vw1), Z]))
            Dm[i22] -= kd * (G2.T @ G2)
        if b1 and b2:
            i12 = np.ix_(idx1, idx2)
            i21 = np.ix_(idx2, idx1)
            Km[i12] += ks * (G1.T @ R1.T @ R2 @ np.hstack([-se3_brac(self.x_2), I]))
            Km[i21] += ks * (G2.T @ R2.T @ R1 @ np.hstack([-se3_brac(self.x_1), I]))
            Km[i12] -= kd * (G1.T @ R1.T @ R2 @ np.hstack([se3_brac(vl2), Z]))
            Km[i21] -= kd * (G2.T @ R2.T @ R1 @ np.hstack([se3_brac(vl1), Z]))
            Dm[i12] += kd * (G1.T @ R1.T @ R2 @ G2)
            Dm[i21] += kd * (G2.T @ R2.T @ R1 @ G1)

    def computeEnergy_(self, V):
        """ForcePointPoint.m:116-132"""
        E1 = np.eye(4) if self.body1 is None else self.body1.E_wi
        E2 = np.eye(4) if self.body2 is None else self.body2.E_wi
        x1w = E1[0:3, :] @ np.append(self.x_1, 1.0)
        x2w = E2[0:3, :] @ np.append(self.x_2, 1.0)
        d = x2w - x1w
        return V + 0.5 * self.stiffness * (d @ d)


_CORNERS = np.array([
    [-1, -1, -1, 1],
    [-1, -1, 1, 1],
    [-1, 1, -1, 1],
    [-1, 1, 1, 1],
    [1, -1, -1, 1],
    [1, -1, 1, 1],
    [1, 1, -1, 1],
    [1, 1, 1, 1],
], dtype=float).T  # ForceGroundCuboid.m:73-82


class ForceGroundCuboid(Force):
    def __init__(self, cuboid):
        """ForceGroundCuboid.m:18"""
        self.cuboid = cuboid
        self.E = np.eye(4)
        self.kn = 1.0
        self.kt = 0.0
        self.mu = 0.0
        self.kd = 0.0

    def setTransform(self, E):
        self.E = np.array(E, dtype=float)

    def setStiffness(self, kn, kt):
        self.kn = kn
        self.kt = kt

    def setDamping(self, kd):
        self.kd = kd

    def setFriction(self, mu):
        self.mu = mu

    def computeValues_(self, fr, fm, Kr=None, Km=None, Dr=None, Dm=None):
        """ForceGroundCuboid.m:54-153"""
        deriv = Km is not None
        idxM = self.cuboid.idxM
        ixM = np.ix_(idxM, idxM)
        xg = self.E[0:3, 3]
        ng = self.E[0:3, 2]
        N = np.outer(ng, ng)
        I = np.eye(3)
        Z = np.zeros((3, 3))
        T = I - N
        R = self.cuboid.E_wi[0:3, 0:3]
        p = self.cuboid.E_wi[0:3, 3]
        phi = self.cuboid.phi
        e1b = se3_brac([1, 0, 0])
        e2b = se3_brac([0, 1, 0])
        e3b = se3_brac([0, 0, 1])
        RNR = R.T @ N @ R
        pxgtmp = se3_brac(R.T @ N @ (p - xg))
        S = np.eye(4)
        S[0:3, 0:3] = np.diag(0.5 * self.cuboid.sides)
        xl = S @ _CORNERS
        xw = self.cuboid.E_wi @ xl
        for i in range(8):
            xli = xl[0:3, i]
            xwi = xw[0:3, i]
            xlbrac = se3_brac(xli)
            d = ng @ (xwi - xg)
            if d > 0:
                continue
            G = se3_Gamma(xli)
            Gphi = G @ phi
            vwi = R @ Gphi
            fc = -self.kn * ng * d - self.kd * (N @ vwi)
            fm[idxM] = fm[idxM] + G.T @ R.T @ fc
            if deriv:
                RNRxl = RNR @ xli
                RNRGphi = RNR @ Gphi
                Gphibrac = se3_brac(Gphi)
                tmp1 = -np.column_stack([e1b @ RNRxl, e2b @ RNRxl, e3b @ RNRxl]) - RNR @ xlbrac + pxgtmp
                tmp2 = -np.column_stack([e1b @ RNRGphi, e2b @ RNRGphi, e3b @ RNRGphi]) - RNR @ Gphibrac
                Km[ixM] = Km[ixM] - self.kn * G.T @ np.hstack([tmp1, RNR]) - self.kd * G.T @ np.hstack([tmp2, np.zeros((3, 3))])
                Dm[ixM] = Dm[ixM] - self.kd * G.T @ RNR @ G
            if self.mu == 0:
                continue
            xwdot = R @ G @ phi
            a = T @ xwdot
            anorm = np.linalg.norm(a)
            if self.mu * abs(self.kn * d) > self.kt * anorm:
                fs = -self.kt * a
                fm[idxM] = fm[idxM] + G.T @ R.T @ fs
                if deriv:
                    D = -self.kt * G.T @ R.T @ T @ R @ G
                    Gphi = G @ phi
                    B = R.T @ T @ R
                    K = -self.kt * G.T @ np.column_stack([
                        (B @ e1b - e1b @ B) @ Gphi, (B @ e2b - e2b @ B) @ Gphi, (B @ e3b - e3b @ B) @ Gphi, Z])
                    Dm[ixM] = Dm[ixM] + D
                    Km[ixM] = Km[ixM] + K
            else:
                mukn = self.mu * self.kn
                t = a / anorm
                fd = -mukn * d * t
                fm[idxM] = fm[idxM] + G.T @ R.T @ fd
                if deriv:
                    A = ((a @ a) * I - np.outer(a, a)) / np.linalg.norm(a) ** 3
                    D = -mukn * G.T @ R.T @ (d * A) @ T @ R @ G
                    Rt = R.T @ t
                    K1 = -d * np.column_stack([e1b @ Rt, e2b @ Rt, e3b @ Rt, Z])
                    K2 = np.outer(Rt, ng) @ R @ G
                    K3 = -d * R.T @ A @ T @ R @ np.hstack([se3_brac(G @ phi), Z])
                    K = -mukn * G.T @ (K1 + K2 + K3)
                    Dm[ixM] = Dm[ixM] + D
                    Km[ixM] = Km[ixM] + K

    def computeEnergy_(self, V):
        """ForceGroundCuboid.m:156"""
        xg = self.E[0:3, 3]
        ng = self.E[0:3, 2]
        S = np.eye(4)
        S[0:3, 0:3] = np.diag(0.5 * self.cuboid.sides)
        xl = S @ _CORNERS
        xw = self.cuboid.E_wi @ xl
        for i in range(8):
            x = xw[0:3, i]
            d = ng @ (x - xg)
            if d > 0:
                continue
            V = V + 0.5 * self.kn * (d * d)
        return V


# ----------------------------------------------------------------------------------------------
# +redmax/Scene.m
# ----------------------------------------------------------------------------------------------
class Scene:
    def __init__(self):
        """Scene.m:31"""
        self.nm = 0  # redmax.Scene.countM (global in MATLAB)
        self.nr = 0  # redmax.Scene.countR
        self.name = ''
        self.bodies = []
        self.joints = []
        self.forces = []
        self.tEnd = 1.0
        self.qInit = None
        self.qdotInit = None
        self.h = 1e-2
        self.t = 0.0
        self.k = 0
        self.T0 = 0.0
        self.V0 = 0.0
        self.history = []
        self.nsteps = 0
        self.grav = np.array([0.0, 0.0, -980.0])
        self.computeH = True
        self.Hexpected = np.zeros(2)
        self.task = None
        self.messages = []  # stands in for the reference's fprintf diagnostics (N6)

    # --- gather/scatter, Joint.m:173-369 (linked-list recursion -> loop over the list) ---
    def getQ(self):
        q = np.zeros(self.nr)
        qdot = np.zeros(self.nr)
        for j in self.joints:
            q[j.idxR] = j.q[: j.ndof]
            qdot[j.idxR] = j.qdot[: j.ndof]
        return q, qdot

    def getQdot(self):
        return self.getQ()[1]

    def setQ(self, q, qdot=None):
        for j in self.joints:
            j.q[: j.ndof] = q[j.idxR]
            if qdot is not None:
                j.qdot[: j.ndof] = qdot[j.idxR]

    def setQdot(self, qdot):
        for j in self.joints:
            j.qdot[: j.ndof] = qdot[j.idxR]

    def getQ0(self):
        q = np.zeros(self.nr)
        qdot = np.zeros(self.nr)
        for j in self.joints:
            q[j.idxR] = j.q0[: j.ndof]
            qdot[j.idxR] = j.qdot0[: j.ndof]
        return q, qdot

    def setQ0(self, q, qdot):
        for j in self.joints:
            j.q0[: j.ndof] = q[j.idxR]
            j.qdot0[: j.ndof] = qdot[j.idxR]
            j.setAux0_()  # Joint.m:283

    def getQ1(self):
        q = np.zeros(self.nr)
        qdot = np.zeros(self.nr)
        for j in self.joints:
            q[j.idxR] = j.q1[: j.ndof]
            qdot[j.idxR] = j.qdot1[: j.ndof]
        return q, qdot

    def getQdot1(self):
        return self.getQ1()[1]

    def setQ1(self, q, qdot):
        for j in self.joints:
            j.q1[: j.ndof] = q[j.idxR]
            j.qdot1[: j.ndof] = qdot[j.idxR]
            j.setAux1_()  # Joint.m:350

    def reparam(self):
        """jroot.reparam(): Joint.m:372-379.  chart_switch_steps (test aid): 0-based index of every step whose result was
        re-parameterised."""
        for j in self.joints:
            if j.reparam_():
                self.chart_switch_steps.append(self.k)

    def update(self, deriv=True):
        """jroot.update(deriv): Joint.m:382-434; only the root sees `deriv` (Joint.m:432)."""
        for i, j in enumerate(self.joints):
            j.update(deriv if i == 0 else True)

    def init(self):
        """Scene.m:59"""
        njoints = len(self.joints)
        # getTraversalOrder returns 1:n (Joint.m:134-146): parents must be listed before children
        for i in range(njoints - 1, -1, -1):
            self.joints[i].countDofs(self)
        if len(self.forces) == 0:
            self.forces = [ForceNull()]
        for j in self.joints:
            j.q0 = j.q.copy()
            j.qdot0 = j.qdot.copy()
            j.q1 = j.q.copy()
            j.qdot1 = j.qdot.copy()
        self.update()
        self.qInit, self.qdotInit = self.getQ()
        for b in self.bodies:
            b.computeInertia_()
        for f in self.forces:
            f.init_()
        self.nsteps = int(math.ceil(self.tEnd / self.h))
        self.reset()

    def computeEnergies(self):
        T = 0.0
        V = 0.0
        for j in self.joints:
            T, V = j.computeEnergies(self.grav, T, V)
        for f in self.forces:
            V = f.computeEnergy_(V)
        return T, V

    def reset(self):
        """Scene.m:122 (plus, so that one scene object can be rolled out repeatedly by the tests: Euler-chart joints go back
        to the chart they were built in -- the reference builds a fresh scene per run)"""
        for j in self.joints:
            j.resetChart_()
        self.chart_switch_steps = []
        self.setQ(self.qInit, self.qdotInit)
        self.t = 0.0
        self.k = 0
        self.T0, self.V0 = self.computeEnergies()
        self.history = []
        if self.task is not None:
            self.task.P = 0.0

    def saveHistory(self, tape=None):
        """Scene.m:134"""
        q, qdot = self.getQ()
        rec = {'q': q, 'qdot': qdot}
        while len(self.history) < self.k:
            self.history.append(None)
        self.history[self.k - 1] = rec
        if self.task is not None:
            rec.update(tape)
            self.task.calcStep()
        if self.computeH:
            T, V = self.computeEnergies()
            rec['T'] = T
            rec['V'] = V
            rec['t'] = self.t
        return rec

    def finalEnergy(self):
        """Scene.m:164-170: H(end) with V shifted by V(1)=V0."""
        return self.history[-1]['T'] + (self.history[-1]['V'] - self.V0)

    def checkEnergy(self, itype):
        """Scene.m:171-177.  itype: 1=BDF1, 2=BDF2.  Returns (pass, H_end)."""
        Hend = self.finalEnergy()
        return abs(Hend - self.Hexpected[itype - 1]) <= 1e-2, Hend


# ----------------------------------------------------------------------------------------------
# driverRedMaxBDF1.m / driverRedMaxBDF2.m
# ----------------------------------------------------------------------------------------------
def compute_values(scene, deriv):
    """driverRedMaxBDF1.m:190-243 (identical copies in the other three drivers).

    deriv=False: returns (M, f); deriv=True: returns (M, f, dMdq, K, D, J)."""
    nr = scene.nr
    nm = scene.nm
    qdot = scene.getQdot()
    J = np.zeros((nm, nr))
    Jdot = np.zeros((nm, nr))
    Mm = np.zeros((nm, nm))
    fm = np.zeros(nm)
    fr = np.zeros(nr)
    if not deriv:
        for j in scene.joints:
            j.computeJacobian2(J, Jdot)
        for b in scene.bodies:
            b.computeMassGrav(scene.grav, Mm, fm)
        for j in scene.joints:
            j.computeForce(fr)
        for f in scene.forces:
            f.computeValues_(fr, fm)
    else:
        dJdq = np.zeros((nm, nr, nr))
        dJdotdq = np.zeros((nm, nr, nr))
        Km = np.zeros((nm, nm))
        Dm = np.zeros((nm, nm))
        Kr = np.zeros((nr, nr))
        Dr = np.zeros((nr, nr))
        for j in scene.joints:
            j.computeJacobian4(J, Jdot, dJdq, dJdotdq)
        for b in scene.bodies:
            b.computeMassGrav(scene.grav, Mm, fm, Km, Dm)
        for j in scene.joints:
            j.computeForce(fr, Kr, Dr)
        for f in scene.forces:
            f.computeValues_(fr, fm, Kr, Km, Dr, Dm)

    JtMm = J.T @ Mm
    M = JtMm @ J
    fqvv = -JtMm @ Jdot @ qdot
    f = fr + J.T @ fm + fqvv
    if not deriv:
        return M, f

    dMdq = np.zeros((nr, nr, nr))
    for i in range(nr):
        tmp = JtMm @ dJdq[:, :, i]
        dMdq[:, :, i] = tmp.T + tmp
    Kqvv = np.zeros((nr, nr))
    Dqvv = -JtMm @ Jdot
    MmJdotqdot = Mm @ Jdot @ qdot
    for i in range(nr):
        dJdqi = dJdq[:, :, i]
        dJdotdqi = dJdotdq[:, :, i]
        Kqvv[:, i] = -dJdqi.T @ MmJdotqdot - JtMm @ dJdotdqi @ qdot
        Dqvv[:, i] = Dqvv[:, i] - JtMm @ dJdqi @ qdot
    K = Kr + J.T @ Km @ J + Kqvv
    D = Dr + J.T @ Dm @ J + Dqvv
    JtDm = J.T @ Dm
    for i in range(nr):
        dJdqi = dJdq[:, :, i]
        K[:, i] = K[:, i] + dJdqi.T @ fm + JtDm @ dJdqi @ qdot
    return M, f, dMdq, K, D, J


SDIRK_A = (2 - math.sqrt(2)) / 2  # driverRedMaxBDF2.m:75


def _eval_common(scene, dqtmp, cD, cK, deriv, want_tape):
    """Shared tail of evalBDF1/evalSDIRK2a/evalSDIRK2b/evalBDF2: g = M*dqtmp - cK*f,
    H = M - cD*D - cK*K + sum_i dMdq(:,:,i)*dqtmp (driverRedMaxBDF1.m:173-185)."""
    nr = scene.nr
    if not deriv:
        scene.update(False)
        M, f = compute_values(scene, False)
        return M @ dqtmp - cK * f
    scene.update()
    M, f, dMdq, K, D, J = compute_values(scene, True)
    g = M @ dqtmp - cK * f
    H = M - cD * D - cK * K
    for i in range(nr):
        H[:, i] = H[:, i] + dMdq[:, :, i] @ dqtmp
    if want_tape:
        return g, H, M, f, K, D, J
    return g, H


def eval_bdf1(q1, scene, deriv=True, want_tape=False):
    """driverRedMaxBDF1.m:160"""
    h = scene.h
    h2 = h * h
    q0, qdot0 = scene.getQ0()
    dqtmp = q1 - q0 - h * qdot0
    qdot1 = (q1 - q0) / h
    scene.setQ(q1, qdot1)
    return _eval_common(scene, dqtmp, h, h2, deriv, want_tape)


def eval_sdirk2a(qa, scene, deriv=True, want_tape=False):
    """driverRedMaxBDF2.m:194"""
    h = scene.h
    a = SDIRK_A
    ah = a * h
    ah2 = ah * ah
    q0, qdot0 = scene.getQ0()
    dqtmp = qa - q0 - ah * qdot0
    qdota = (qa - q0) / ah
    scene.setQ(qa, qdota)
    return _eval_common(scene, dqtmp, ah, ah2, deriv, want_tape)


def eval_sdirk2b(q1, scene, deriv=True, want_tape=False):
    """driverRedMaxBDF2.m:228"""
    h = scene.h
    a = SDIRK_A
    ah = a * h
    ah2 = ah * ah
    q0, qdot0 = scene.getQ0()
    qdota = scene.getQdot1()
    dqtmp = q1 - q0 - (2 * a - 1) * h * qdot0 - 2 * (1 - a) * h * qdota
    qdot1 = (q1 - q0 - (1 - a) * h * qdota) / ah
    scene.setQ(q1, qdot1)
    return _eval_common(scene, dqtmp, ah, ah2, deriv, want_tape)


def eval_bdf2(q2, scene, deriv=True, want_tape=False):
    """driverRedMaxBDF2.m:263"""
    h = scene.h
    h2 = h * h
    q0, qdot0 = scene.getQ0()
    q1, qdot1 = scene.getQ1()
    dqtmp = q2 - (4 / 3) * q1 + (1 / 3) * q0 - (8 / 9) * h * qdot1 + (2 / 9) * h * qdot0
    qdot2 = (3 / (2 * h)) * (q2 - (4 / 3) * q1 + (1 / 3) * q0)
    scene.setQ(q2, qdot2)
    return _eval_common(scene, dqtmp, (2 / 3) * h, (4 / 9) * h2, deriv, want_tape)


def newton(evalFcn, xInit, scene=None, stats=None):
    """driverRedMaxBDF1.m:94-157: damped Newton with backtracking line search."""
    tol = 1e-9
    dxMax = 1e3
    iterMax = 10 * len(xInit)
    iterLsMax = 20
    x = xInit.copy()
    it = 1
    nls = 0
    status = 0
    while True:
        g, H = evalFcn(x, True)
        dx = -np.linalg.solve(H, g)
        if np.linalg.norm(dx) > dxMax:
            status |= 1  # 'Newton diverged'
            break
        alpha = 1.0
        g0 = g
        x0 = x
        f0 = 0.5 * (g0 @ g0)
        iterLs = 1
        while True:
            x = x0 + alpha * dx
            g = evalFcn(x, False)
            nls += 1
            f = 0.5 * (g @ g)
            if f < f0:
                break
            if iterLs >= iterLsMax:
                status |= 4
                break
            alpha = 0.5 * alpha
            iterLs = iterLs + 1
        if np.linalg.norm(g) < tol:
            break
        if it >= iterMax:
            status |= 2  # 'Newton did not converge'
            break
        it = it + 1
    if stats is not None:
        stats.append((it, nls, status))
    return x


def sim_loop_bdf1(scene, nsteps=None, stats=None):
    """driverRedMaxBDF1.m:57-91"""
    h = scene.h
    nsteps = scene.nsteps if nsteps is None else nsteps
    for k in range(nsteps):
        q0, qdot0 = scene.getQ()
        scene.setQ0(q0, qdot0)
        q1 = q0 + h * qdot0
        q1 = newton(lambda x, d: eval_bdf1(x, scene, d), q1, scene, stats)
        qdot1 = (q1 - q0) / h
        scene.setQ(q1, qdot1)
        scene.reparam()  # :78
        scene.update()
        scene.t = scene.t + h
        scene.k = k + 1
        scene.saveHistory()


def sim_loop_bdf2(scene, nsteps=None, stats=None):
    """driverRedMaxBDF2.m:57-125"""
    h = scene.h
    nsteps = scene.nsteps if nsteps is None else nsteps
    for k in range(nsteps):
        if k == 0:
            q0, qdot0 = scene.getQ()
            scene.setQ0(q0, qdot0)
            a = SDIRK_A
            qa = q0 + a * h * qdot0
            qa = newton(lambda x, d: eval_sdirk2a(x, scene, d), qa, scene, stats)
            qdota = (qa - q0) / (a * h)
            scene.setQ1(qa, qdota)
            q1 = qa + (1 - a) * h * qdota
            q1 = newton(lambda x, d: eval_sdirk2b(x, scene, d), q1, scene, stats)
            qdot1 = (q1 - q0 - (1 - a) * h * qdota) / (a * h)
            scene.setQ(q1, qdot1)
            scene.setQ1(q0, qdot0)
        else:
            q0, qdot0 = scene.getQ1()
            scene.setQ0(q0, qdot0)
            q1, qdot1 = scene.getQ()
            scene.setQ1(q1, qdot1)
            q2 = q1 + h * qdot1
            q2 = newton(lambda x, d: eval_bdf2(x, scene, d), q2, scene, stats)
            qdot2 = (3 / (2 * h)) * (q2 - (4 / 3) * q1 + (1 / 3) * q0)
            scene.setQ(q2, qdot2)
        scene.reparam()  # :112
        scene.update()
        scene.t = scene.t + h
        scene.k = k + 1
        scene.saveHistory()


# ----------------------------------------------------------------------------------------------
# driverRedMaxAdjointBDF1.m / driverRedMaxAdjointBDF2.m
# ----------------------------------------------------------------------------------------------
def _lu_vector(H):
    """[Hl,Hu,Hp] = lu(H,'vector'): H(Hp,:) = Hl*Hu (LAPACK dgetrf)."""
    lu, piv = sla.lu_factor(H)
    n = H.shape[0]
    perm = np.arange(n)
    for i in range(n):
        perm[i], perm[piv[i]] = perm[piv[i]], perm[i]
    Hl = np.tril(lu, -1) + np.eye(n)
    Hu = np.triu(lu)
    return Hl, Hu, perm


def newton_adjoint(evalFcn, xInit, stats=None):
    """driverRedMaxAdjointBDF1.m:105-146: no line search, convergence test on the PRE-update g (N4)."""
    tol = 1e-9
    dxMax = 1e3
    iterMax = 5 * len(xInit)
    x = xInit.copy()
    it = 1
    status = 0
    while True:
        g, H, M, f, K, D, J = evalFcn(x)
        Hl, Hu, Hp = _lu_vector(H)
        dx = -sla.solve_triangular(Hu, sla.solve_triangular(Hl, g[Hp], lower=True, unit_diagonal=True))
        if np.linalg.norm(dx) > dxMax:
            status |= 1
            break
        x = x + dx
        if np.linalg.norm(g) < tol:
            break
        if it >= iterMax:
            status |= 2
            break
        it = it + 1
    if stats is not None:
        stats.append((it, 0, status))
    tape = {'Hl': Hl, 'Hu': Hu, 'Hp': Hp, 'M': M, 'f': f, 'K': K, 'D': D, 'J': J, 'H': H}
    return x, tape


def sim_loop_adjoint_bdf1(scene, stats=None):
    """driverRedMaxAdjointBDF1.m:65-102"""
    h = scene.h
    for k in range(scene.nsteps):
        scene.task.applyStep()
        q0, qdot0 = scene.getQ()
        scene.setQ0(q0, qdot0)
        q1 = q0 + h * qdot0
        q1, tape = newton_adjoint(lambda x: eval_bdf1(x, scene, True, True), q1, stats)
        qdot1 = (q1 - q0) / h
        scene.setQ(q1, qdot1)
        scene.reparam()  # driverRedMaxAdjointBDF1.m:89 / driverRedMaxAdjointBDF2.m:123
        scene.update()
        scene.t = scene.t + h
        scene.k = k + 1
        scene.saveHistory(tape)


def sim_loop_adjoint_bdf2(scene, stats=None):
    """driverRedMaxAdjointBDF2.m:65-136 (only the second SDIRK sub-solve's tape is saved, :88,:96)"""
    h = scene.h
    for k in range(scene.nsteps):
        scene.task.applyStep()
        if k == 0:
            q0, qdot0 = scene.getQ()
            scene.setQ0(q0, qdot0)
            a = SDIRK_A
            qa = q0 + a * h * qdot0
            qa, _ = newton_adjoint(lambda x: eval_sdirk2a(x, scene, True, True), qa, stats)
            qdota = (qa - q0) / (a * h)
            scene.setQ1(qa, qdota)
            q1 = qa + (1 - a) * h * qdota
            q1, tape = newton_adjoint(lambda x: eval_sdirk2b(x, scene, True, True), q1, stats)
            qdot1 = (q1 - q0 - (1 - a) * h * qdota) / (a * h)
            scene.setQ(q1, qdot1)
            scene.setQ1(q0, qdot0)
        else:
            q0, qdot0 = scene.getQ1()
            scene.setQ0(q0, qdot0)
            q1, qdot1 = scene.getQ()
            scene.setQ1(q1, qdot1)
            q2 = q1 + h * qdot1
            q2, tape = newton_adjoint(lambda x: eval_bdf2(x, scene, True, True), q2, stats)
            qdot2 = (3 / (2 * h)) * (q2 - (4 / 3) * q1 + (1 / 3) * q0)
            scene.setQ(q2, qdot2)
        scene.reparam()  # driverRedMaxAdjointBDF1.m:89 / driverRedMaxAdjointBDF2.m:123
        scene.update()
        scene.t = scene.t + h
        scene.k = k + 1
        scene.saveHistory(tape)


def task_objective(p, scene, scheme):
    """driverRedMaxAdjointBDF1.m:39-45 (taskObjective without the FD self-test)."""
    scene.reset()
    scene.task.p = np.asarray(p, dtype=float).copy()
    scene.task.init()
    if scheme == 1:
        sim_loop_adjoint_bdf1(scene)
    else:
        sim_loop_adjoint_bdf2(scene)
    return scene.task.calcFinal()


# ----------------------------------------------------------------------------------------------
# +redmax/TaskBDF1.m, TaskBDF2.m, TaskBDF1PointPos.m, TaskBDF2PointPos.m
# ----------------------------------------------------------------------------------------------
class _TaskBase:
    def __init__(self, scene, nparams):
        """TaskBDF1.m:17"""
        self.scene = scene
        self.p = np.zeros(nparams)
        self.P = 0.0
        self.wreg = 1.0
        self.dgdp = None
        self.dPdq = None

    def init(self):
        """TaskBDF1.m:27"""
        nr = self.scene.nr
        nsteps = self.scene.nsteps
        self.P = 0.0
        self.dPdq = [None] * nsteps
        self.dgdp = np.zeros((nsteps * nr, len(self.p)))

    def _solve_diag(self, k, yk):
        """zkk0(Hp) = Hl'\\(Hu'\\yk)  (TaskBDF1.m:74-77)"""
        rec = self.scene.history[k]
        w = sla.solve_triangular(rec['Hu'].T, yk, lower=True)
        w = sla.solve_triangular(rec['Hl'].T, w, lower=False, unit_diagonal=True)
        z = np.zeros_like(w)
        z[rec['Hp']] = w
        return z


class TaskBDF1(_TaskBase):
    def calcFinal(self):
        """TaskBDF1.m:45-81"""
        nr = self.scene.nr
        P = self.P + self.wreg * 0.5 * (self.p @ self.p)
        nsteps = self.scene.nsteps
        z = np.zeros(nsteps * nr)
        h = self.scene.h
        hist = self.scene.history
        for k in range(nsteps, 0, -1):  # 1-based k as in the reference
            yk = self.dPdq[k - 1].copy()
            kk0 = (k - 1) * nr + np.arange(nr)
            kk1 = kk0 + nr
            kk2 = kk1 + nr
            if k < nsteps:
                M = hist[k]['M']
                D = hist[k]['D']
                block = -2 * M + h * D
                yk = yk - block.T @ z[kk1]
            if k < nsteps - 1:
                M = hist[k + 1]['M']
                block = M
                yk = yk - block.T @ z[kk2]
            z[kk0] = self._solve_diag(k - 1, yk)
        dPdp = self.wreg * self.p - z @ self.dgdp
        self.z = z
        return P, dPdp


class TaskBDF2(_TaskBase):
    def calcFinal(self):
        """TaskBDF2.m:45-108"""
        nr = self.scene.nr
        P = self.P + self.wreg * 0.5 * (self.p @ self.p)
        nsteps = self.scene.nsteps
        z = np.zeros(nsteps * nr)
        h = self.scene.h
        a = SDIRK_A
        hist = self.scene.history
        for k in range(nsteps, 0, -1):
            yk = self.dPdq[k - 1].copy()
            kk0 = (k - 1) * nr + np.arange(nr)
            kk1 = kk0 + nr
            kk2 = kk1 + nr
            kk3 = kk2 + nr
            kk4 = kk3 + nr
            if k < nsteps:
                M = hist[k]['M']
                D = hist[k]['D']
                if k == 1:
                    block = -((8 / (9 * a)) + (4 / 3)) * M + (8 / 9) * h * D
                else:
                    block = -(8 / 3) * M + (8 / 9) * h * D
                yk = yk - block.T @ z[kk1]
            if k < nsteps - 1:
                M = hist[k + 1]['M']
                D = hist[k + 1]['D']
                if k == 1:
                    block = ((2 / (9 * a)) + (19 / 9)) * M - (2 / 9) * h * D
                else:
                    block = (22 / 9) * M - (2 / 9) * h * D
                yk = yk - block.T @ z[kk2]
            if k < nsteps - 2:
                M = hist[k + 2]['M']
                block = -(8 / 9) * M
                yk = yk - block.T @ z[kk3]
            if k < nsteps - 3:
                M = hist[k + 3]['M']
                block = (1 / 9) * M
                yk = yk - block.T @ z[kk4]
            z[kk0] = self._solve_diag(k - 1, yk)
        dPdp = self.wreg * self.p - z @ self.dgdp
        self.z = z
        return P, dPdp


class _PointPosMixin:
    """TaskBDF1PointPos.m / TaskBDF2PointPos.m (they differ only in the dgdp coefficient, :106)."""

    def _init_pointpos(self, scene):
        nparams = 0
        for j in scene.joints:
            nparams += j.ndof
        return nparams

    def setTime(self, t):
        self.t = t

    def setBody(self, body):
        self.body = body

    def setPoint(self, xlocal):
        self.xlocal = np.asarray(xlocal, dtype=float).reshape(3)

    def setTarget(self, xtarget):
        self.xtarget = np.asarray(xtarget, dtype=float).reshape(3)

    def setScale(self, pscale):
        self.pscale = pscale

    def setWeights(self, wreg, wpos):
        self.wreg = wreg
        self.wpos = wpos

    def applyStep(self):
        """TaskBDF1PointPos.m:58"""
        # first_step_free: NOT the reference -- diagnostic mode of tests/test_oracle_pins.py in which the first time step takes
        # no control (its dgdp block is zero), so that the parameters act through the plain BDF steps only
        off = getattr(self, 'first_step_free', False) and self.scene.k == 0
        for j in self.scene.joints:
            j.tau = (0.0 if off else self.pscale) * self.p[j.idxR]

    def calcStep(self):
        """TaskBDF1PointPos.m:67-107"""
        scene = self.scene
        nm = scene.nm
        nr = scene.nr
        wp = self.wpos
        k = scene.k
        dt = self.t - scene.t
        if abs(dt) < 1e-6:
            rec = scene.history[k - 1]
            scene.setQ(rec['q'], rec['qdot'])
            scene.update()
            E = self.body.E_wi
            xworld = E[0:3, :] @ np.append(self.xlocal, 1.0)
            dx = xworld - self.xtarget
            self.P = self.P + wp * 0.5 * (dx @ dx)
            dxdqm = np.zeros((3, nm))
            R = E[0:3, 0:3]
            dxdqm[:, self.body.idxM] = R @ se3_Gamma(self.xlocal)
            J = rec['J']
            self.dPdq[k - 1] = J.T @ dxdqm.T @ dx * wp
        else:
            self.dPdq[k - 1] = np.zeros(nr)
        h = scene.h
        kk = (k - 1) * nr + np.arange(nr)
        coeff = self._dgdp_coeff
        if k == 1 and getattr(self, 'first_step_free', False):
            coeff = 0.0
        self.dgdp[kk, :] = coeff * h ** 2 * self.pscale * np.eye(nr)


class TaskBDF1PointPos(_PointPosMixin, TaskBDF1):
    _dgdp_coeff = -1.0  # TaskBDF1PointPos.m:106

    def __init__(self, scene):
        TaskBDF1.__init__(self, scene, self._init_pointpos(scene))


class TaskBDF2PointPos(_PointPosMixin, TaskBDF2):
    _dgdp_coeff = -(4 / 9)  # TaskBDF2PointPos.m:106 (used for every step, incl. the SDIRK first step: N7)

    def __init__(self, scene):
        TaskBDF2.__init__(self, scene, self._init_pointpos(scene))


# ----------------------------------------------------------------------------------------------
# scenesRedMax.m
# ----------------------------------------------------------------------------------------------
def scenes(sceneID):
    """scenesRedMax.m -- IDs -2,-1,0,1,2,11,14,100,101 (revolute/fixed/Free2D + ground only)."""
    scene = Scene()
    density = 1.0
    if sceneID == -2:  # :13
        scene.name = 'Single revolute'
        b = BodyCuboid(density, [2, 0.2, 0.2])
        scene.bodies.append(b)
        j = JointRevolute(None, b, [0, 1, 0])
        j.setJointTransform(np.eye(4))
        j.q[0] = 0
        j.qdot[0] = 1
        scene.joints.append(j)
        b.setBodyTransform(trans([1, 0, 0]))
    elif sceneID == -1:  # :27
        scene.name = 'Simpler serial chain'
        sides = [10, 1, 1]
        nbodies = 1
        for i in range(1, nbodies + 1):
            b = BodyCuboid(density, sides)
            scene.bodies.append(b)
            if i == 1:
                j = JointRevolute(None, b, [0, 1, 0])
                j.setJointTransform(np.eye(4))
                j.q[0] = 0
                j.qdot[0] = 1
            else:
                j = JointRevolute(scene.joints[i - 2], b, [0, 1, 0])
                j.setJointTransform(trans([10, 0, 0]))
                j.q[0] = math.pi / 4
                j.qdot[0] = 1
            scene.joints.append(j)
            b.setBodyTransform(trans([5, 0, 0]))
            j.setStiffness(1e6)
            j.setDamping(1e4)
    elif sceneID == 0:  # :52
        scene.name = 'Simple serial chain'
        scene.Hexpected[:] = [-1.2705398823489915e+05, 2.6058008179021417e+03]
        sides = [10, 1, 1]
        nbodies = 5
        for i in range(1, nbodies + 1):
            b = BodyCuboid(density, sides)
            scene.bodies.append(b)
            if i == 1:
                j = JointRevolute(None, b, [0, 1, 0])
                j.setJointTransform(np.eye(4))
            else:
                if i % 2 == 1:
                    j = JointRevolute(scene.joints[i - 2], b, [0, 1, 0])
                else:
                    j = JointFixed(scene.joints[i - 2], b)
                j.setJointTransform(trans([10, 0, 0]))
            scene.joints.append(j)
            b.setBodyTransform(trans([5, 0, 0]))
            if j.ndof > 0:  # `q(1)=...` on a 0-dof joint has no effect on the dynamics
                j.q[0] = math.pi / 4 if i % 2 == 1 else 0.0
    elif sceneID == 1:  # :80
        scene.name = 'Different revolute axes'
        scene.Hexpected[:] = [-3.8359074258588909e+04, -9.7138545812971279e+02]
        sides = [10, 1, 1]
        bs = [BodyCuboid(density, sides) for _ in range(3)]
        scene.bodies = bs
        j1 = JointRevolute(None, bs[0], [0, 0, 1])
        j2 = JointRevolute(j1, bs[1], [0, 1, 0])
        j3 = JointRevolute(j2, bs[2], [0, 0, 1])
        scene.joints = [j1, j2, j3]
        for b in bs:
            b.setBodyTransform(trans([5, 0, 0]))
        j1.setJointTransform(np.eye(4))
        j2.setJointTransform(trans([10, 0, 0]))
        j3.setJointTransform(trans([10, 0, 0]))
        j1.q[0] = 0
        j2.q[0] = math.pi / 2
        j3.q[0] = math.pi / 2
    elif sceneID == 2:  # :101
        scene.name = 'Branching'
        scene.Hexpected[:] = [-2.2826101928480086e+04, -2.4159349151742754e+02]
        bs = [BodyCuboid(density, [1, 1, 10]), BodyCuboid(density, [1, 20, 1]),
              BodyCuboid(density, [1, 1, 10]), BodyCuboid(density, [1, 1, 10])]
        scene.bodies = bs
        j1 = JointRevolute(None, bs[0], [1, 0, 0])
        j2 = JointRevolute(j1, bs[1], [0, 0, 1])
        j3 = JointRevolute(j2, bs[2], [1, 0, 0])
        j4 = JointRevolute(j2, bs[3], [0, 1, 0])
        scene.joints = [j1, j2, j3, j4]
        bs[0].setBodyTransform(trans([0, 0, -5]))
        bs[1].setBodyTransform(trans([0, 0, 0]))
        bs[2].setBodyTransform(trans([0, 0, -5]))
        bs[3].setBodyTransform(trans([0, 0, -5]))
        j1.setJointTransform(trans([0, 0, 15]))
        j2.setJointTransform(trans([0, 0, -10]))
        j3.setJointTransform(trans([0, -10, 0]))
        j4.setJointTransform(trans([0, 10, 0]))
        j1.q[0] = 0
        j2.q[0] = 0
        j3.q[0] = math.pi / 4
        j4.q[0] = math.pi / 4
    elif sceneID == 11:  # :290
        scene.name = 'Free2D with ground'
        scene.Hexpected[:] = [-4.4208045000000002e+03, -2.7811251900394832e+03]
        scene.h = 5e-4
        scene.tEnd = 0.6
        scene.grav = np.array([0.0, -980.0, 0.0])
        b = BodyCuboid(density, [3, 1, 1])
        scene.bodies = [b]
        j = JointFree2D(None, b)
        j.q = np.array([-1.0, 2.0, 0.0])
        j.qdot = np.array([5.0, 70.0, 2.0])
        j.setJointTransform(np.eye(4))
        scene.joints = [j]
        b.setBodyTransform(np.eye(4))
        f = ForceGroundCuboid(b)
        E = np.eye(4)
        E[0:3, 0:3] = se3_aaToMat([1, 0, 0], -math.pi / 2)
        f.setTransform(E)
        f.setStiffness(1e5, 1e2)
        f.setDamping(3e1)
        f.setFriction(0.5)
        scene.forces = [f]
    elif sceneID == 14:  # :371
        scene.name = 'Joint limits'
        scene.Hexpected[:] = [-2.5928305306546572e+04, -1.8476279319765570e+04]
        scene.h = 5e-3
        sides = [10, 1, 1]
        nbodies = 3
        for i in range(1, nbodies + 1):
            b = BodyCuboid(density, sides)
            scene.bodies.append(b)
            if i == 1:
                j = JointRevolute(None, b, [0, 1, 0])
                E = np.eye(4)
                E[0:3, 0:3] = se3_aaToMat([0, 1, 0], math.pi / 2)
                j.setJointTransform(E)
                j.q[0] = 0
                j.qdot[0] = 0
            else:
                j = JointRevolute(scene.joints[i - 2], b, [0, 1, 0])
                j.setJointTransform(trans([10, 0, 0]))
                j.q[0] = -math.pi / 6
                j.qdot[0] = 0
            scene.joints.append(j)
            b.setBodyTransform(trans([5, 0, 0]))
            j.setLimitLower(-math.pi / 2)
            j.setLimitUpper(0)
            j.setLimitStiffness(1e5)
            j.setLimitDamping(1e2)
            j.setDamping(1e2)
    elif sceneID in (100, 101):  # :402, :437
        scene.name = 'Adjoint BDF1' if sceneID == 100 else 'Adjoint BDF2'
        sides = [10, 1, 1]
        nbodies = 2
        for i in range(1, nbodies + 1):
            b = BodyCuboid(density, sides)
            scene.bodies.append(b)
            if i == 1:
                j = JointRevolute(None, b, [0, 1, 0])
                j.setJointTransform(np.eye(4))
                j.q[0] = math.pi / 2
                j.qdot[0] = 1
            else:
                j = JointRevolute(scene.joints[i - 2], b, [0, 1, 0])
                j.setJointTransform(trans([10, 0, 0]))
                j.q[0] = math.pi / 4
                j.qdot[0] = 1
            scene.joints.append(j)
            b.setBodyTransform(trans([5, 0, 0]))
            j.setStiffness(1e4)
            j.setDamping(1e4)
        scene.task = TaskBDF1PointPos(scene) if sceneID == 100 else TaskBDF2PointPos(scene)
        scene.task.setTime(scene.tEnd)
        scene.task.setBody(scene.bodies[-1])
        scene.task.setPoint([5, 0, 0])
        scene.task.setTarget([10, 0, -10] if sceneID == 100 else [-10, 0, -10])
        scene.task.setScale(1e5)
        scene.task.setWeights(1e-2, 1e2)
    else:
        raise ValueError('scene %r not restated (out of scope, SURVEY.md section 8f)' % (sceneID,))
    return scene


# ----------------------------------------------------------------------------------------------
# Synthetic benchmark scenes: same object API, our constants
# ----------------------------------------------------------------------------------------------
def chain_scene(n, ground=False, h=1e-2, nsteps=100, axis=(0, 1, 0)):
    """C2/C3/C5: n-link serial chain after the scene 0/-1 pattern (scenesRedMax.m:27-79)."""
    scene = Scene()
    scene.name = '%d-link chain' % n
    scene.h = h
    scene.tEnd = nsteps * h
    for i in range(1, n + 1):
        b = BodyCuboid(1.0, [10, 1, 1])
        scene.bodies.append(b)
        if i == 1:
            j = JointRevolute(None, b, axis)
            j.setJointTransform(np.eye(4))
        else:
            j = JointRevolute(scene.joints[i - 2], b, axis)
            j.setJointTransform(trans([10, 0, 0]))
        scene.joints.append(j)
        b.setBodyTransform(trans([5, 0, 0]))
        j.q[0] = math.pi / 4 if i % 2 == 1 else 0.0
    if ground:
        for b in scene.bodies:
            f = ForceGroundCuboid(b)
            E = np.eye(4)
            E[0:3, 0:3] = se3_aaToMat([1, 0, 0], 0.0)
            E[0:3, 3] = [0, 0, -40]
            f.setTransform(E)
            f.setStiffness(1e5, 1e2)
            f.setDamping(3e1)
            f.setFriction(0.5)
            scene.forces.append(f)
    return scene


def hand_scene(h=1e-2, nsteps=100, scheme=1):
    """C4: fixed palm + 5 fingers x 4 revolute phalanges, TaskBDF*PointPos on the index fingertip."""
    scene = Scene()
    scene.name = 'hand'
    scene.h = h
    scene.tEnd = nsteps * h
    palm = BodyCuboid(1.0, [8, 8, 1])
    scene.bodies.append(palm)
    jp = JointFixed(None, palm)
    jp.setJointTransform(np.eye(4))
    scene.joints.append(jp)
    palm.setBodyTransform(np.eye(4))
    tip_index = None
    for fi, y in enumerate([-3.0, -1.5, 0.0, 1.5, 3.0]):
        parent = jp
        for k in range(4):
            b = BodyCuboid(1.0, [3, 0.8, 0.8])
            scene.bodies.append(b)
            ax = [0, 0, 1] if (fi == 0 and k == 0) else [0, 1, 0]
            j = JointRevolute(parent, b, ax)
            j.setJointTransform(trans([4, y, 0]) if k == 0 else trans([3, 0, 0]))
            b.setBodyTransform(trans([1.5, 0, 0]))
            j.setStiffness(1e4)
            j.setDamping(1e4)
            scene.joints.append(j)
            parent = j
        if fi == 1:
            tip_index = scene.bodies[-1]
    scene.task = TaskBDF1PointPos(scene) if scheme == 1 else TaskBDF2PointPos(scene)
    scene.task.setTime(scene.tEnd)
    scene.task.setBody(tip_index)
    scene.task.setPoint([1.5, 0, 0])
    scene.task.setTarget([10, 0, -5])
    scene.task.setScale(1e5)
    scene.task.setWeights(1e-2, 1e2)
    return scene


def run_forward(scene, scheme, q0=None, qdot0=None, tau=None, nsteps=None, stats=None):
    """Convenience used by tests/bench: one forward rollout from (q0, qdot0) with constant joint torques
    `tau` (indexed like q).  Returns q(t), qdot(t) as [nsteps, nr] arrays."""
    if q0 is not None:
        scene.qInit = np.asarray(q0, dtype=float).copy()
    if qdot0 is not None:
        scene.qdotInit = np.asarray(qdot0, dtype=float).copy()
    scene.reset()
    scene.update()
    scene.T0, scene.V0 = scene.computeEnergies()
    if tau is not None:
        for j in scene.joints:
            j.tau = np.asarray(tau, dtype=float)[j.idxR].copy()
    nsteps = scene.nsteps if nsteps is None else nsteps
    task, scene.task = scene.task, None  # forward drivers know no task (driverRedMaxBDF1.m)
    try:
        if scheme == 1:
            sim_loop_bdf1(scene, nsteps, stats)
        else:
            sim_loop_bdf2(scene, nsteps, stats)
    finally:
        scene.task = task
    qs = np.array([r['q'] for r in scene.history[:nsteps]])
    qds = np.array([r['qdot'] for r in scene.history[:nsteps]])
    return qs, qds
