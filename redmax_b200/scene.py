"""Host-side mirror of the reference's `+redmax` object API (matlab-diff/+redmax/*.m), description only.

The class and method names, argument meaning and defaults are the reference's, so a scene script written for
`scenesRedMax.m` reads the same here:

    scene = Scene()
    scene.bodies.append(BodyCuboid(density, sides))
    scene.joints.append(JointRevolute(parent, body, axis)); joint.setJointTransform(E); body.setBodyTransform(E)
    scene.forces.append(ForceGroundCuboid(body)); ...
    scene.init()

`Scene.init()` (Scene.m:59-119) numbers the degrees of freedom leaf-to-root exactly as the reference does,
flattens the object graph into `rmx_scene_desc` and hands it to the CUDA library through the C ABI.  All
dynamics (Joint.update, computeJacobian, computeValues, newton, simLoop, Task.calcFinal ...) then run on the
GPU for a whole batch of rollouts: `Scene.rollout(...)`, `Scene.rollout_adjoint(...)`.  Nothing in this file
computes dynamics on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _ffi, euler
from ._ffi import RmxError, f64, i32, ptr


def inertiaCuboid(whd, density):
    """se3.inertiaCuboid (se3.m:366): diagonal inertia [Ixx Iyy Izz m m m]."""
    whd = np.asarray(whd, dtype=float).reshape(3)
    m = np.zeros(6)
    mass = density * np.prod(whd)
    m[0] = (1.0 / 12.0) * mass * (whd[[1, 2]] @ whd[[1, 2]])
    m[1] = (1.0 / 12.0) * mass * (whd[[2, 0]] @ whd[[2, 0]])
    m[2] = (1.0 / 12.0) * mass * (whd[[0, 1]] @ whd[[0, 1]])
    m[3:6] = mass
    return m


class Body:
    """+redmax/Body.m (description fields only)."""

    def __init__(self, density):
        self.name = ''
        self.density = density
        self.E0_ji = np.eye(4)
        self.I_i = np.ones(6)
        self.joint = None
        self.idxM = None

    def setBodyTransform(self, E):
        """Body.m:46 -- transform of this body wrt its joint."""
        self.E0_ji = np.array(E, dtype=float).reshape(4, 4)


class BodyCuboid(Body):
    """+redmax/BodyCuboid.m"""

    def __init__(self, density, sides):
        super().__init__(density)
        self.sides = np.asarray(sides, dtype=float).reshape(3)

    def computeInertia_(self):
        self.I_i = inertiaCuboid(self.sides, self.density)


class Joint:
    """+redmax/Joint.m (description fields only)."""

    def __init__(self, parent, body, ndof):
        self.parent = parent
        self.body = body
        self.children = []
        self.ndof = ndof
        self.q = np.zeros(ndof)
        self.qdot = np.zeros(ndof)
        self.qRest = np.zeros(ndof)
        self.qLimL = -1e8   # Joint.m:77-80
        self.qLimU = 1e8
        self.qLimK = 1e8
        self.qLimD = 0.0
        self.tau = np.zeros(ndof)
        self.stiffness = 0.0
        self.damping = 0.0
        self.E0_pj = np.eye(4)
        self.idxR = None
        body.joint = self
        if parent is not None:
            parent.children.append(self)

    def setJointTransform(self, E):
        """Joint.m:95 -- transform of this joint wrt the parent joint at q = 0."""
        self.E0_pj = np.array(E, dtype=float).reshape(4, 4)

    def setStiffness(self, stiffness):
        self.stiffness = float(stiffness)

    def setDamping(self, damping):
        self.damping = float(damping)

    def setLimitLower(self, limit):
        self.qLimL = float(limit)

    def setLimitUpper(self, limit):
        self.qLimU = float(limit)

    def setLimitStiffness(self, K):
        self.qLimK = float(K)

    def setLimitDamping(self, D):
        self.qLimD = float(D)


class JointRevolute(Joint):
    """+redmax/JointRevolute.m"""
    jtype = _ffi.RMX_JOINT_REVOLUTE

    def __init__(self, parent, body, axis):
        super().__init__(parent, body, 1)
        axis = np.asarray(axis, dtype=float).reshape(3)
        self.axis = axis / np.linalg.norm(axis)  # JointRevolute.m:14


class JointFixed(Joint):
    """+redmax/JointFixed.m"""
    jtype = _ffi.RMX_JOINT_FIXED

    def __init__(self, parent, body):
        super().__init__(parent, body, 0)
        self.axis = np.zeros(3)


class JointPrismatic(Joint):
    """+redmax/JointPrismatic.m"""
    jtype = _ffi.RMX_JOINT_PRISMATIC

    def __init__(self, parent, body, axis):
        super().__init__(parent, body, 1)
        axis = np.asarray(axis, dtype=float).reshape(3)
        self.axis = axis / np.linalg.norm(axis)  # JointPrismatic.m:14


class JointPlanar(Joint):
    """+redmax/JointPlanar.m -- `plane` is 3 x 2, columns = the two in-plane directions (default x, y)."""
    jtype = _ffi.RMX_JOINT_PLANAR

    def __init__(self, parent, body, plane=None):
        super().__init__(parent, body, 2)
        if plane is None:
            plane = np.array([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]]).T  # JointPlanar.m:14
        plane = np.array(plane, dtype=float).reshape(3, 2)
        self.plane = np.stack([plane[:, 0] / np.linalg.norm(plane[:, 0]), plane[:, 1] / np.linalg.norm(plane[:, 1])], axis=1)
        self.axis = self.plane[:, 0].copy()
        self.axis2 = self.plane[:, 1].copy()


class JointTranslational(Joint):
    """+redmax/JointTranslational.m"""
    jtype = _ffi.RMX_JOINT_TRANSLATIONAL

    def __init__(self, parent, body):
        super().__init__(parent, body, 3)
        self.axis = np.zeros(3)


class JointFree2D(Joint):
    """+redmax/JointFree2D.m -- q = [x y theta]"""
    jtype = _ffi.RMX_JOINT_FREE2D

    def __init__(self, parent, body):
        super().__init__(parent, body, 3)
        self.axis = np.zeros(3)


class JointUniversal(Joint):
    """+redmax/JointUniversal.m -- rotation about X then Y"""
    jtype = _ffi.RMX_JOINT_UNIVERSAL

    def __init__(self, parent, body):
        super().__init__(parent, body, 2)
        self.axis = np.zeros(3)


class JointSpherical(Joint):
    """+redmax/JointSpherical.m -- Euler angles in one of the reference's twelve charts (`chart`, default XYZ as
    JointSpherical.m:33).  Under BDF2 a rollout that leaves the chart's well-conditioned range is re-parameterised as the
    reference does (Scene._reparam_rollouts); under BDF1 the reference itself stops there and the rollout stays flagged
    RMX_ST_CHART."""
    jtype = _ffi.RMX_JOINT_SPHERICAL
    CHART_XYZ = 7

    def __init__(self, parent, body):
        super().__init__(parent, body, 3)
        self.axis = np.zeros(3)
        self.chart = self.CHART_XYZ


class JointFree3D(Joint):
    """+redmax/JointFree3D.m -- q = [x y z, Euler XYZ]"""
    jtype = _ffi.RMX_JOINT_FREE3D

    def __init__(self, parent, body):
        super().__init__(parent, body, 6)
        self.axis = np.zeros(3)


class ForceGroundCuboid:
    """+redmax/ForceGroundCuboid.m:18-48"""

    def __init__(self, cuboid):
        self.cuboid = cuboid
        self.E = np.eye(4)
        self.kn = 1.0
        self.kt = 0.0
        self.mu = 0.0
        self.kd = 0.0

    def setTransform(self, E):
        self.E = np.array(E, dtype=float).reshape(4, 4)

    def setStiffness(self, kn, kt):
        self.kn = float(kn)
        self.kt = float(kt)

    def setDamping(self, kd):
        self.kd = float(kd)

    def setFriction(self, mu):
        self.mu = float(mu)


class ForcePointPoint:
    """+redmax/ForcePointPoint.m:15-45 -- linear zero-rest-length spring / damper between two body points; a body may be
    None (the point is then fixed in the world)."""

    kind = _ffi.RMX_FORCE_POINTPOINT

    def __init__(self, body1, x_1, body2, x_2):
        self.body1 = body1
        self.body2 = body2
        self.x_1 = np.asarray(x_1, dtype=float).reshape(3)
        self.x_2 = np.asarray(x_2, dtype=float).reshape(3)
        self.stiffness = 1.0
        self.damping = 0.0

    def setStiffness(self, stiffness):
        self.stiffness = float(stiffness)

    def setDamping(self, damping):
        self.damping = float(damping)


class ForceSpringDamper(ForcePointPoint):
    """+redmax/ForceSpringDamper.m -- damped spring with rest length L along the line between two body points; L defaults to
    the distance in the initial configuration (init_, ForceSpringDamper.m:38-62; computed by the library)."""
    kind = _ffi.RMX_FORCE_SPRINGDAMPER

    def __init__(self, body1, x_1, body2, x_2):
        super().__init__(body1, x_1, body2, x_2)
        self.stiffness = 1.0
        self.damping = 1.0  # ForceSpringDamper.m:15
        self.L = 0.0

    def setRetLength(self, L):
        """ForceSpringDamper.m:31 (sic)"""
        self.L = float(L)


class ForceCable:
    """+redmax/ForceCable.m (ForceSpringMultiPointGeneric.m) -- a cable routed through up to four body points (a body may be
    None = world); pulls only when stretched beyond its rest length L (default: routed length in the initial configuration)."""

    def __init__(self):
        self.bodies = []
        self.xls = []
        self.stiffness = 1.0
        self.damping = 1.0
        self.L = 0.0

    def addBodyPoint(self, body, xl):
        """ForceSpringMultiPointGeneric.m:20"""
        self.bodies.append(body)
        self.xls.append(np.asarray(xl, dtype=float).reshape(3))

    def setStiffness(self, stiffness):
        self.stiffness = float(stiffness)

    def setDamping(self, damping):
        self.damping = float(damping)

    def setRetLength(self, L):
        self.L = float(L)


class _TaskPointPos:
    """+redmax/TaskBDF1PointPos.m / TaskBDF2PointPos.m (parameters = constant joint torques, objective = a body
    point reaching a target at time t)."""
    scheme = 1

    def __init__(self, scene):
        self.scene = scene
        self.t = scene.tEnd
        self.body = None
        self.xlocal = np.zeros(3)
        self.xtarget = np.zeros(3)
        self.pscale = 1.0
        self.wreg = 1.0
        self.wpos = 1.0

    def setTime(self, t):
        self.t = float(t)

    def setBody(self, body):
        self.body = body

    def setPoint(self, xlocal):
        self.xlocal = np.asarray(xlocal, dtype=float).reshape(3)

    def setTarget(self, xtarget):
        self.xtarget = np.asarray(xtarget, dtype=float).reshape(3)

    def setScale(self, pscale):
        self.pscale = float(pscale)

    def setWeights(self, wreg, wpos):
        self.wreg = float(wreg)
        self.wpos = float(wpos)


class TaskBDF1PointPos(_TaskPointPos):
    scheme = 1


class TaskBDF2PointPos(_TaskPointPos):
    scheme = 2


class Scene:
    """+redmax/Scene.m.  `init()` replaces Scene.init; `rollout*` replace the drivers' simLoop."""

    def __init__(self):
        self.name = ''
        self.bodies = []
        self.joints = []
        self.forces = []
        self.tEnd = 1.0        # Scene.m:38
        self.h = 1e-2          # Scene.m:41
        self.grav = np.array([0.0, 0.0, -980.0])  # Scene.m:48
        self.Hexpected = np.zeros(2)
        self.task = None
        self.qInit = None
        self.qdotInit = None
        self.nsteps = 0
        self.nr = 0
        self.nm = 0
        self._handle = None

    # -- Scene.init, Scene.m:59-119 ---------------------------------------------------------------------
    def init(self):
        n = len(self.joints)
        if n == 0 or len(self.bodies) != n:
            raise ValueError('scene needs one body per joint')
        index = {id(j): i for i, j in enumerate(self.joints)}
        for i, j in enumerate(self.joints):
            if j.parent is not None and index[id(j.parent)] >= i:
                raise ValueError('joints must be listed parents-before-children (Joint.getTraversalOrder)')
            if j.body is not self.bodies[i]:
                raise ValueError('bodies must be listed in the same order as their joints (Scene.m:104)')
        nr = 0
        nm = 0
        for i in range(n - 1, -1, -1):  # Scene.m:69-71: leaf-to-root numbering
            j = self.joints[i]
            j.idxR = nr + np.arange(j.ndof)
            nr += j.ndof
            j.body.idxM = nm + np.arange(6)
            nm += 6
            j.qRest = np.array(j.q[: j.ndof], dtype=float)  # Joint.m:157
        self.nr, self.nm = nr, nm
        for b in self.bodies:
            b.computeInertia_()
        self.qInit = np.zeros(nr)
        self.qdotInit = np.zeros(nr)
        for j in self.joints:
            self.qInit[j.idxR] = j.q[: j.ndof]
            self.qdotInit[j.idxR] = j.qdot[: j.ndof]
        self.nsteps = int(math.ceil(self.tEnd / self.h))  # Scene.m:117
        self._create_handle(index)
        return self

    def _create_handle(self, index):
        h = self._build_handle(index)
        self.close()
        self._handle = h
        self._index = index
        self._variants = {}

    def _spherical(self):
        return [j for j in self.joints if isinstance(j, JointSpherical)]

    def _variant(self, charts):
        """Library handle of this scene with its JointSpherical joints in the given Euler charts (the scene's own charts: the
        main handle).  Built on first use, destroyed with the scene."""
        sph = self._spherical()
        charts = tuple(int(c) for c in charts)
        if charts == tuple(j.chart for j in sph):
            return self._handle
        if charts not in self._variants:
            own = [j.chart for j in sph]
            try:
                for j, c in zip(sph, charts):
                    j.chart = c
                self._variants[charts] = self._build_handle(self._index)
            finally:
                for j, c in zip(sph, own):
                    j.chart = c
        return self._variants[charts]

    def _build_handle(self, index):
        n = len(self.joints)
        L = _ffi.lib()
        d = _ffi.rmx_scene_desc()
        keep = []

        def arr(a, conv):
            a = conv(a)
            keep.append(a)
            return a.ctypes.data_as(_ffi._pd if a.dtype == np.float64 else _ffi._pi)
        d.n = n
        d.parent = arr([(-1 if j.parent is None else index[id(j.parent)]) for j in self.joints], i32)
        for j in self.joints:
            if getattr(j, 'jtype', None) is None:
                raise RmxError('joint type %s is not on the GPU hot path (SURVEY.md section 8)' % type(j).__name__)
        d.jtype = arr([j.jtype for j in self.joints], i32)
        # MATLAB stores 4x4 column-major: E.T.ravel()
        d.E0_pj = arr(np.concatenate([j.E0_pj.T.ravel() for j in self.joints]), f64)
        d.E0_ji = arr(np.concatenate([j.body.E0_ji.T.ravel() for j in self.joints]), f64)
        d.axis = arr(np.concatenate([j.axis for j in self.joints]), f64)
        d.axis2 = arr(np.concatenate([getattr(j, 'axis2', np.array([0.0, 1.0, 0.0])) for j in self.joints]), f64)
        d.I_i = arr(np.concatenate([j.body.I_i for j in self.joints]), f64)
        d.sides = arr(np.concatenate([j.body.sides for j in self.joints]), f64)
        d.stiffness = arr([j.stiffness for j in self.joints], f64)
        d.damping = arr([j.damping for j in self.joints], f64)
        qrest = np.zeros((n, _ffi.RMX_MAX_JOINT_DOF))
        for i, j in enumerate(self.joints):
            qrest[i, : j.ndof] = j.qRest
        d.qRest = arr(qrest, f64)
        d.qLimL = arr([j.qLimL for j in self.joints], f64)
        d.qLimU = arr([j.qLimU for j in self.joints], f64)
        d.qLimK = arr([j.qLimK for j in self.joints], f64)
        d.qLimD = arr([j.qLimD for j in self.joints], f64)
        d.grav = (C.c_double * 3)(*[float(x) for x in self.grav])
        grounds = [f for f in self.forces if isinstance(f, ForceGroundCuboid)]
        points = [f for f in self.forces if isinstance(f, ForcePointPoint)]
        cables = [f for f in self.forces if isinstance(f, ForceCable)]
        bidx = {id(j.body): i for i, j in enumerate(self.joints)}
        d.ncable = len(cables)
        if cables:
            mp = _ffi.RMX_MAX_CABLE_POINTS
            if any(not (2 <= len(f.bodies) <= mp) for f in cables):
                raise RmxError('a cable has 2 .. %d points' % mp)
            cb = -np.ones((len(cables), mp), dtype=np.int32)
            cx = np.zeros((len(cables), mp, 3))
            for i, f in enumerate(cables):
                for k, (b, xl) in enumerate(zip(f.bodies, f.xls)):
                    cb[i, k] = -1 if b is None else bidx[id(b)]
                    cx[i, k] = xl
            d.cable_npts = arr([len(f.bodies) for f in cables], i32)
            d.cable_body = arr(cb, i32)
            d.cable_x = arr(cx, f64)
            d.cable_ks = arr([f.stiffness for f in cables], f64)
            d.cable_kd = arr([f.damping for f in cables], f64)
            d.cable_L = arr([f.L for f in cables], f64)
        if len(grounds) + len(points) + len(cables) != len(self.forces):
            raise RmxError('only ForceGroundCuboid, ForcePointPoint, ForceSpringDamper and ForceCable are on the GPU hot path '
                           '(SURVEY.md section 8)')
        d.npointforce = len(points)
        if points:
            d.pf_body1 = arr([(-1 if f.body1 is None else bidx[id(f.body1)]) for f in points], i32)
            d.pf_body2 = arr([(-1 if f.body2 is None else bidx[id(f.body2)]) for f in points], i32)
            d.pf_x1 = arr(np.concatenate([f.x_1 for f in points]), f64)
            d.pf_x2 = arr(np.concatenate([f.x_2 for f in points]), f64)
            d.pf_ks = arr([f.stiffness for f in points], f64)
            d.pf_kd = arr([f.damping for f in points], f64)
            d.pf_kind = arr([f.kind for f in points], i32)
            d.pf_L = arr([getattr(f, 'L', 0.0) for f in points], f64)
        d.nground = len(grounds)
        if grounds:
            d.ground_body = arr([index[id(f.cuboid.joint)] for f in grounds], i32)
            d.ground_E = arr(np.concatenate([f.E.T.ravel() for f in grounds]), f64)
            d.ground_kn = arr([f.kn for f in grounds], f64)
            d.ground_kt = arr([f.kt for f in grounds], f64)
            d.ground_kd = arr([f.kd for f in grounds], f64)
            d.ground_mu = arr([f.mu for f in grounds], f64)
        d.chart = arr([int(getattr(j, 'chart', 0)) for j in self.joints], i32)
        h = C.c_void_p()
        _ffi.check(L.rmx_scene_create(C.byref(d), C.byref(h)), 'rmx_scene_create')
        assert L.rmx_scene_nr(h) == self.nr and L.rmx_scene_nm(h) == self.nm
        return h

    def close(self):
        for h in getattr(self, '_variants', {}).values():
            _ffi.lib().rmx_scene_destroy(h)
        self._variants = {}
        if self._handle is not None:
            _ffi.lib().rmx_scene_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def countR(scene):
        return scene.nr

    # -- options ----------------------------------------------------------------------------------------
    def opts(self, scheme=1, adjoint=False, nsteps=None, h=None, ngpus=1, tau_mode=_ffi.RMX_TAU_NONE,
             linsolve=_ffi.RMX_LINSOLVE_LU, **kw):
        o = _ffi.rmx_opts()
        _ffi.lib().rmx_opts_default(C.byref(o), int(scheme), int(bool(adjoint)))
        o.nsteps = int(self.nsteps if nsteps is None else nsteps)
        o.h = float(self.h if h is None else h)
        o.ngpus = int(ngpus)
        o.tau_mode = int(tau_mode)
        o.linsolve = int(linsolve)
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def _require(self):
        if self._handle is None:
            raise RmxError('call scene.init() first')
        return _ffi.lib()

    # -- forward rollouts: simLoop of driverRedMaxBDF1.m:57 / driverRedMaxBDF2.m:57, batched ----------------
    def rollout(self, q0=None, qdot0=None, tau=None, scheme=1, nsteps=None, ngpus=1, want_qdot=True, reparam=True, **kw):
        """Host-buffer call.  q0, qdot0: [B, nr] (row b = rollout b; memory layout nr x B column-major as the
        C ABI wants).  tau: None, [B, nr] (constant) or [B, nsteps, nr].  Returns dict(q=[B,nsteps,nr],
        qdot=..., status=[B], iters=[B,2])."""
        L = self._require()
        q0 = f64(self.qInit[None, :] if q0 is None else q0)
        qdot0 = f64(self.qdotInit[None, :] if qdot0 is None else qdot0)
        if q0.ndim == 1:
            q0 = q0[None, :]
        if qdot0.ndim == 1:
            qdot0 = qdot0[None, :]
        B = q0.shape[0]
        assert q0.shape == (B, self.nr) and qdot0.shape == (B, self.nr)
        tau_mode = _ffi.RMX_TAU_NONE
        if tau is not None:
            tau = f64(tau)
            tau_mode = _ffi.RMX_TAU_CONST if tau.ndim == 2 else _ffi.RMX_TAU_PER_STEP
        o = self.opts(scheme=scheme, nsteps=nsteps, ngpus=ngpus, tau_mode=tau_mode, **kw)
        ns = o.nsteps
        if tau is not None:
            assert tau.shape == ((B, self.nr) if tau_mode == _ffi.RMX_TAU_CONST else (B, ns, self.nr))
        q = np.empty((B, ns, self.nr))
        qd = np.empty((B, ns, self.nr)) if want_qdot else None
        status = np.empty(B, dtype=np.int32)
        iters = np.empty((B, 2), dtype=np.int32)
        _ffi.check(L.rmx_rollout(self._handle, C.byref(o), B, ptr(q0), ptr(qdot0), ptr(tau), ptr(q), ptr(qd),
                                 ptr(status), ptr(iters)), 'rmx_rollout')
        out = dict(q=q, qdot=qd, status=status, iters=iters)
        if reparam and scheme == 2 and want_qdot and self._spherical() and (status & _ffi.RMX_ST_CHART).any():
            self._reparam_rollouts(out, q0, qdot0, tau, tau_mode, o)
        return out

    def _reparam_rollouts(self, out, q0, qdot0, tau, tau_mode, o):
        """jroot.reparam() of driverRedMaxBDF2.m:112 for the rollouts the library flagged with RMX_ST_CHART, all of them
        together, in rounds: (1) every active rollout is cut at the first step whose result leaves the well-conditioned range
        of a JointSpherical chart (|det T| <= 0.5); the piece up to that step is run once more on its own (one
        rmx_rollout_resume per group of rollouts with equal charts, per-rollout [k_begin, k_end)), which gives its status and
        iteration counts without the discarded tail; (2) that step and the BDF2 history joint.q1 / qdot1 are re-expressed in
        the chart JointSpherical.reparam_ picks (euler.reparam); (3) the rollouts resume from the next step to the end under
        the scene variant of their new charts, and the next round cuts them again if they switch again.
        Adds out['chart'] [B, nspherical] (the charts q(t_end) is expressed in) and out['chart_switches'] (per rollout:
        (step, spherical joint, old chart, new chart)); q, qdot of a switch step are the re-parameterised ones, as in the
        reference's history.  JointFree3D cannot switch in the reference (its inner joint never gets chart1:
        JointFree3D.m:27-31, JointSpherical.m:73 stops with an error) and stays flagged."""
        L = _ffi.lib()
        sph = self._spherical()
        nsteps, B = o.nsteps, out['q'].shape[0]
        Q, QD = out['q'], out['qdot']
        own = tuple(j.chart for j in sph)
        out['chart'] = np.tile(np.array(own, dtype=np.int32), (B, 1))
        out['chart_switches'] = [[] for _ in range(B)]
        active = [int(b) for b in np.nonzero(out['status'] & _ffi.RMX_ST_CHART)[0]]
        charts = {b: list(own) for b in active}
        kb = {b: 0 for b in active}
        hist = {b: None for b in active}  # (step or -1 = initial state, q, qdot): joint.q1 / qdot1 in the charts in force
        status = {b: 0 for b in active}
        iters = {b: np.zeros(2, dtype=np.int64) for b in active}

        def run(rows, k0, k1, verify):
            """rmx_rollout_resume of the steps [k0[b], k1[b]) for the rollouts `rows`, grouped by chart tuple, each with its
            re-expressed BDF2 history in place of the stored (old-chart) step; results go back into Q, QD."""
            groups = {}
            for b in rows:
                groups.setdefault(tuple(charts[b]), []).append(b)
            for ct, rs in groups.items():
                idx = np.array(rs)
                qs, qds = np.ascontiguousarray(Q[idx]), np.ascontiguousarray(QD[idx])
                q0s, qd0s = np.ascontiguousarray(q0[idx]), np.ascontiguousarray(qdot0[idx])
                ts = None if tau is None else np.ascontiguousarray(tau[idx])
                saved = {}
                for i, b in enumerate(rs):
                    if hist[b] is None:
                        continue
                    k, hq, hqd = hist[b]
                    if k < 0:
                        q0s[i], qd0s[i] = hq, hqd
                    else:
                        saved[i] = (k, qs[i, k].copy(), qds[i, k].copy())
                        qs[i, k], qds[i, k] = hq, hqd
                st = np.zeros(len(rs), dtype=np.int32)
                it = np.zeros((len(rs), 2), dtype=np.int32)
                ka = np.array([k0[b] for b in rs], dtype=np.int32)
                ke = np.array([k1[b] for b in rs], dtype=np.int32)
                _ffi.check(L.rmx_rollout_resume(self._variant(ct), C.byref(o), len(rs), ptr(ka), ptr(ke), ptr(q0s), ptr(qd0s),
                                                ptr(ts), ptr(qs), ptr(qds), ptr(st), ptr(it)), 'rmx_rollout_resume')
                for i, (k, sq, sqd) in saved.items():
                    qs[i, k], qds[i, k] = sq, sqd
                for i, b in enumerate(rs):
                    if verify:
                        a, e = k0[b], k1[b]
                        if np.isfinite(qs[i]).all() and not np.array_equal(qs[i, a:e], Q[b, a:e]):
                            raise RmxError('re-parameterised rollout %d: piece [%d, %d) does not reproduce the pass it was '
                                           'cut from' % (b, a, e))
                        status[b] |= int(st[i]) & ~_ffi.RMX_ST_CHART
                        iters[b] += it[i]
                    Q[b], QD[b] = qs[i], qds[i]

        while active:
            # (1) cut: first step at or after kb whose result needs a new chart; the piece [kb, k1] on its own
            k1, sw = {}, {}
            for b in active:
                need = np.zeros(nsteps - kb[b], dtype=bool)
                for j, c in zip(sph, charts[b]):
                    need |= euler.chart_det(c, Q[b, kb[b]:, j.idxR[1]]) <= 0.5
                sw[b] = bool(need.any())
                k1[b] = kb[b] + int(np.argmax(need)) if sw[b] else nsteps - 1
            run(active, kb, {b: k1[b] + 1 for b in active}, verify=True)
            # (2) JointSpherical.reparam_ at step k1 for every spherical joint that asks for it
            nxt = []
            for b in active:
                if not sw[b]:
                    continue
                k = k1[b]
                hq = (q0[b] if k == 0 else Q[b, k - 1]).copy()
                hqd = (qdot0[b] if k == 0 else QD[b, k - 1]).copy()
                if hist[b] is not None and hist[b][0] == k - 1:  # consecutive switch steps: history already re-expressed
                    hq, hqd = hist[b][1].copy(), hist[b][2].copy()
                for i, (j, c) in enumerate(zip(sph, charts[b])):
                    r = j.idxR
                    if euler.chart_det(c, Q[b, k, r[1]]) > 0.5:
                        continue
                    new, qn, qdn, q1n, qd1n = euler.reparam(c, Q[b, k, r], QD[b, k, r], c, hq[r], hqd[r])
                    out['chart_switches'][b].append((int(k), i, int(c), int(new)))
                    charts[b][i] = new
                    Q[b, k, r], QD[b, k, r] = qn, qdn
                    hq[r], hqd[r] = q1n, qd1n
                hist[b] = (k - 1, hq, hqd)
                kb[b] = k + 1
                if kb[b] < nsteps:
                    nxt.append(b)
            # (3) the rest of those rollouts in their new charts
            if nxt:
                run(nxt, kb, {b: nsteps for b in nxt}, verify=False)
            active = nxt
        free3d_mid = [j.idxR[4] for j in self.joints if isinstance(j, JointFree3D)]
        for b in charts:
            for m in free3d_mid:  # a JointFree3D that left its chart stays flagged
                if (np.abs(np.cos(Q[b, :, m])) <= 0.5).any():
                    status[b] |= _ffi.RMX_ST_CHART
            out['status'][b] = status[b]
            out['iters'][b] = iters[b]
            out['chart'][b] = charts[b]

    def rollout_into(self, q0, qdot0, q_out, qdot_out=None, tau=None, scheme=1, nsteps=None, ngpus=1, **kw):
        """rmx_rollout with caller-owned host buffers (e.g. pinned memory): q0, qdot0 [B, nr] float64 C-contiguous,
        q_out / qdot_out [B, nsteps, nr].  Returns dict(status, iters)."""
        L = self._require()
        B = q0.shape[0]
        tau_mode = _ffi.RMX_TAU_NONE
        if tau is not None:
            tau_mode = _ffi.RMX_TAU_CONST if tau.ndim == 2 else _ffi.RMX_TAU_PER_STEP
        o = self.opts(scheme=scheme, nsteps=nsteps, ngpus=ngpus, tau_mode=tau_mode, **kw)
        for a in (q0, qdot0, q_out, qdot_out, tau):
            if a is not None and not (a.dtype == np.float64 and a.flags['C_CONTIGUOUS']):
                raise ValueError('rollout_into needs C-contiguous float64 arrays')
        if q0.shape != (B, self.nr) or qdot0.shape != (B, self.nr) or q_out.shape != (B, o.nsteps, self.nr):
            raise ValueError('rollout_into: shape mismatch')
        status = np.empty(B, dtype=np.int32)
        iters = np.empty((B, 2), dtype=np.int32)
        _ffi.check(L.rmx_rollout(self._handle, C.byref(o), B, ptr(q0), ptr(qdot0), ptr(tau), ptr(q_out), ptr(qdot_out),
                                 ptr(status), ptr(iters)), 'rmx_rollout')
        return dict(status=status, iters=iters)

    def rollout_dev(self, q0, qdot0, q_out, qdot_out, status, iters=None, tau=None, scheme=1, nsteps=None,
                    stream=None, **kw):
        """Device-buffer call (torch CUDA tensors or raw device addresses), enqueued on `stream` (a
        torch.cuda.Stream, a raw cudaStream_t address, or None for the legacy default stream)."""
        L = self._require()
        B = int(q0.shape[0])
        tau_mode = _ffi.RMX_TAU_NONE
        if tau is not None:
            tau_mode = _ffi.RMX_TAU_CONST if tau.dim() == 2 else _ffi.RMX_TAU_PER_STEP
        o = self.opts(scheme=scheme, nsteps=nsteps, tau_mode=tau_mode, **kw)
        st = None
        if stream is not None:
            st = C.c_void_p(int(getattr(stream, 'cuda_stream', stream)))
        _ffi.check(L.rmx_rollout_dev(self._handle, C.byref(o), B, ptr(q0), ptr(qdot0), ptr(tau), ptr(q_out),
                                     ptr(qdot_out), ptr(status), ptr(iters), st), 'rmx_rollout_dev')

    def rollout_multi_dev(self, q0, qdot0, q_out, qdot_out, status, iters=None, tau=None, scheme=1, nsteps=None, gather=True, **kw):
        """rmx_rollout_multi_dev: one process drives G GPUs.  Every argument is a list with one torch CUDA tensor per device:
        q0[g], qdot0[g] (and tau[g], status[g], iters[g]) hold device g's shard, q_out[g] / qdot_out[g] are full-size
        [B, nsteps, nr]; with gather=True one NCCL all-gather per array leaves all trajectories on every device."""
        L = self._require()
        G = len(q0)
        devs = (C.c_int32 * G)(*[int(t.device.index) for t in q0])
        B = int(q_out[0].shape[0])
        tau_mode = _ffi.RMX_TAU_NONE
        if tau is not None:
            tau_mode = _ffi.RMX_TAU_CONST if tau[0].dim() == 2 else _ffi.RMX_TAU_PER_STEP
        o = self.opts(scheme=scheme, nsteps=nsteps, tau_mode=tau_mode, **kw)

        def plist(ts):
            if ts is None:
                return None
            return (C.c_void_p * G)(*[None if t is None else t.data_ptr() for t in ts])
        keep = [plist(x) for x in (q0, qdot0, tau, q_out, qdot_out, status, iters)]
        _ffi.check(L.rmx_rollout_multi_dev(self._handle, C.byref(o), G, devs, B, *keep, int(bool(gather))), 'rmx_rollout_multi_dev')

    @staticmethod
    def check_status(status):
        """Raise if a device-buffer launch flagged an internal scheduling failure (RMX_ST_SCHED): those trajectories are not
        valid.  The host-pointer calls (rollout, rollout_into) re-run such rollouts inside the library; rollout_dev is
        asynchronous, so its caller checks the status array once the stream has been synchronised."""
        st = status.cpu().numpy() if hasattr(status, 'cpu') else np.asarray(status)
        bad = np.nonzero(st & _ffi.RMX_ST_SCHED)[0]
        if len(bad):
            raise _ffi.RmxError('rmx_rollout_dev: %d rollout(s) hit RMX_ST_SCHED (first: %d); re-run them with RMX_SCHED=0 or '
                                'through the host-pointer call' % (len(bad), int(bad[0])))

    def linsolve_stats(self):
        """Total Krylov iterations of the last rollout with linsolve=RMX_LINSOLVE_PCG (rmx_linsolve_stats)."""
        L = self._require()
        v = C.c_int64(0)
        _ffi.check(L.rmx_linsolve_stats(self._handle, C.byref(v)), 'rmx_linsolve_stats')
        return int(v.value)

    # -- adjoint: taskObjective of driverRedMaxAdjointBDF1.m:39 / ...BDF2.m:39, batched ----------------------
    def _task_struct(self):
        if self.task is None or self.task.body is None:
            raise RmxError('scene.task with a body is required for the adjoint path')
        index = {id(b): i for i, b in enumerate(self.bodies)}
        t = _ffi.rmx_task_pointpos()
        t.body = index[id(self.task.body)]
        t.xlocal = (C.c_double * 3)(*self.task.xlocal)
        t.t_target = self.task.t
        t.pscale = self.task.pscale
        t.wreg = self.task.wreg
        t.wpos = self.task.wpos
        return t

    def rollout_adjoint(self, p, xtarget=None, q0=None, qdot0=None, nsteps=None, ngpus=1, want_q=False, **kw):
        """(P, dPdp) for B parameter vectors p: [B, nr]; xtarget: [B, 3] (default: the task's target)."""
        L = self._require()
        p = f64(p)
        if p.ndim == 1:
            p = p[None, :]
        B = p.shape[0]
        q0 = f64(np.broadcast_to(self.qInit, (B, self.nr)) if q0 is None else q0)
        qdot0 = f64(np.broadcast_to(self.qdotInit, (B, self.nr)) if qdot0 is None else qdot0)
        xt = f64(np.broadcast_to(self.task.xtarget, (B, 3)) if xtarget is None else xtarget)
        o = self.opts(scheme=self.task.scheme, adjoint=True, nsteps=nsteps, ngpus=ngpus, **kw)
        t = self._task_struct()
        P = np.empty(B)
        dPdp = np.empty((B, self.nr))
        q = np.empty((B, o.nsteps, self.nr)) if want_q else None
        status = np.empty(B, dtype=np.int32)
        _ffi.check(L.rmx_rollout_adjoint(self._handle, C.byref(o), C.byref(t), B, ptr(q0), ptr(qdot0), ptr(p), ptr(xt),
                                         ptr(P), ptr(dPdp), ptr(q), ptr(status)), 'rmx_rollout_adjoint')
        return dict(P=P, dPdp=dPdp, q=q, status=status)

    def rollout_adjoint_dev(self, q0, qdot0, p, xtarget, P, dPdp, status, q_out=None, nsteps=None, stream=None, **kw):
        """Device-buffer form of rollout_adjoint (rmx_rollout_adjoint_dev): tape-writing forward rollout + backward sweep
        enqueued on `stream`; every argument is a torch CUDA tensor (or a raw device address), q_out optional."""
        L = self._require()
        B = int(p.shape[0])
        o = self.opts(scheme=self.task.scheme, adjoint=True, nsteps=nsteps, **kw)
        t = self._task_struct()
        st = None
        if stream is not None:
            st = C.c_void_p(int(getattr(stream, 'cuda_stream', stream)))
        _ffi.check(L.rmx_rollout_adjoint_dev(self._handle, C.byref(o), C.byref(t), B, ptr(q0), ptr(qdot0), ptr(p), ptr(xtarget),
                                             ptr(P), ptr(dPdp), ptr(q_out), ptr(status), st), 'rmx_rollout_adjoint_dev')

    def rollout_adjoint_into(self, q0, qdot0, p, xtarget, P, dPdp, status, nsteps=None, ngpus=1, **kw):
        """rmx_rollout_adjoint with caller-owned host buffers (float64 / int32, C-contiguous), nothing allocated here."""
        L = self._require()
        B = int(p.shape[0])
        o = self.opts(scheme=self.task.scheme, adjoint=True, nsteps=nsteps, ngpus=ngpus, **kw)
        t = self._task_struct()
        _ffi.check(L.rmx_rollout_adjoint(self._handle, C.byref(o), C.byref(t), B, ptr(q0), ptr(qdot0), ptr(p), ptr(xtarget),
                                         ptr(P), ptr(dPdp), None, ptr(status)), 'rmx_rollout_adjoint')

    def adjoint_tape_bytes(self, B, nsteps=None):
        """Bytes of the adjoint tape (LU(H) + perm + dP/dq, M, D per rollout-step) a batch of B rollouts writes and reads."""
        o = self.opts(scheme=self.task.scheme if self.task is not None else 1, adjoint=True, nsteps=nsteps)
        return int(self._require().rmx_adjoint_tape_bytes(self._handle, C.byref(o), B))

    # -- test hooks ---------------------------------------------------------------------------------------
    def eval(self, q, qdot, dqtmp, cK, beta, tau=None):
        """One residual/Jacobian evaluation on the GPU (rmx_eval)."""
        L = self._require()
        nr = self.nr
        q, qdot, dqtmp = f64(q), f64(qdot), f64(dqtmp)
        tau = None if tau is None else f64(tau)
        g = np.empty(nr)
        f = np.empty(nr)
        H = np.empty((nr, nr))
        M = np.empty((nr, nr))
        D = np.empty((nr, nr))
        _ffi.check(L.rmx_eval(self._handle, ptr(q), ptr(qdot), ptr(dqtmp), ptr(tau), float(cK), float(beta),
                              ptr(g), ptr(H), ptr(M), ptr(D), ptr(f)), 'rmx_eval')
        # column-major nr x nr -> numpy [row, col]
        return dict(g=g, f=f, H=H.T.copy(), M=M.T.copy(), D=D.T.copy())

    def eval_newton(self, q, qdot, dqtmp, cK, beta, tau=None):
        """H and dx = -H\\g through the forward kernel's own assembly + LU path (rmx_eval_newton)."""
        L = self._require()
        nr = self.nr
        q, qdot, dqtmp = f64(q), f64(qdot), f64(dqtmp)
        tau = None if tau is None else f64(tau)
        H = np.empty((nr, nr))
        dx = np.empty(nr)
        _ffi.check(L.rmx_eval_newton(self._handle, ptr(q), ptr(qdot), ptr(dqtmp), ptr(tau), float(cK), float(beta),
                                     ptr(H), ptr(dx)), 'rmx_eval_newton')
        return dict(H=H.T.copy(), dx=dx)

    def eval_krylov(self, q, qdot, dqtmp, cK, beta, x, tau=None):
        """The operators of the Krylov linear solve at one evaluation point (rmx_eval_krylov): Hx = H x through the
        matrix-free tree sweeps, Pinv_x = (J' blkdiag(M_j) J + Pr)^-1 x through the projected block-Jacobi preconditioner."""
        L = self._require()
        nr = self.nr
        q, qdot, dqtmp, x = f64(q), f64(qdot), f64(dqtmp), f64(x)
        tau = None if tau is None else f64(tau)
        hx, px = np.empty(nr), np.empty(nr)
        _ffi.check(L.rmx_eval_krylov(self._handle, ptr(q), ptr(qdot), ptr(dqtmp), ptr(tau), float(cK), float(beta), ptr(x),
                                     ptr(hx), ptr(px)), 'rmx_eval_krylov')
        return dict(Hx=hx, Pinv_x=px)

    def body_frames(self, q, chart=None):
        """World frames E_wi of all bodies (Body.update, Body.m:70-80) for B configurations q [B, nr] -> [B, nbodies, 4, 4].
        chart [B, nspherical]: the Euler charts the configurations are expressed in (see chart_history); default: the
        scene's own."""
        L = self._require()
        q = f64(q)
        if q.ndim == 1:
            q = q[None, :]
        B = q.shape[0]
        E = np.empty((B, len(self.bodies), 4, 4))
        if chart is None or not self._spherical():
            _ffi.check(L.rmx_body_frames(self._handle, B, ptr(q), ptr(E)), 'rmx_body_frames')
        else:
            chart = np.asarray(chart, dtype=np.int32).reshape(B, -1)
            for ct in {tuple(r) for r in chart.tolist()}:
                sel = np.nonzero((chart == np.array(ct, dtype=np.int32)).all(axis=1))[0]
                qs = np.ascontiguousarray(q[sel])
                Es = np.empty((len(sel), len(self.bodies), 4, 4))
                _ffi.check(L.rmx_body_frames(self._variant(ct), len(sel), ptr(qs), ptr(Es)), 'rmx_body_frames')
                E[sel] = Es
        return np.ascontiguousarray(np.swapaxes(E, 2, 3))  # column-major 4x4 blocks -> [row, col]

    def chart_history(self, out, b=0):
        """[nsteps, nspherical] Euler charts in which q(t_k) of rollout b of a rollout() result is expressed (the scene's own
        up to the first re-parameterised step, JointSpherical.m:84-87)."""
        sph = self._spherical()
        ch = np.tile(np.array([j.chart for j in sph], dtype=np.int32), (out['q'].shape[1], 1))
        for k, i, _, new in (out.get('chart_switches') or [[]] * (b + 1))[b]:
            ch[k:, i] = new
        return ch

    def energies(self, q, qdot, chart=None):
        """T, V of Scene.saveHistory (Scene.m:155-160) for B states.  chart [B, nspherical]: the Euler charts the states are
        expressed in (out['chart'] of a rollout that re-parameterised); default: the scene's own."""
        L = self._require()
        q, qdot = f64(q), f64(qdot)
        if q.ndim == 1:
            q, qdot = q[None, :], qdot[None, :]
        B = q.shape[0]
        T = np.empty(B)
        V = np.empty(B)
        if chart is None or not self._spherical():
            _ffi.check(L.rmx_energies(self._handle, B, ptr(q), ptr(qdot), ptr(T), ptr(V)), 'rmx_energies')
            return T, V
        chart = np.asarray(chart, dtype=np.int32).reshape(B, -1)
        for ct in {tuple(r) for r in chart.tolist()}:
            sel = np.nonzero((chart == np.array(ct, dtype=np.int32)).all(axis=1))[0]
            qs, qds = np.ascontiguousarray(q[sel]), np.ascontiguousarray(qdot[sel])
            Ts, Vs = np.empty(len(sel)), np.empty(len(sel))
            _ffi.check(L.rmx_energies(self._variant(ct), len(sel), ptr(qs), ptr(qds), ptr(Ts), ptr(Vs)), 'rmx_energies')
            T[sel], V[sel] = Ts, Vs
        return T, V
