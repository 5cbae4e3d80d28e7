"""ctypes wrapper of oracle/_ref/liboracle_c.so, the compiled twin of redmax_oracle.py (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Scenes are oracle (redmax_oracle) Scene objects after init(); they are flattened here into the C struct `oc_desc`."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_ref', 'liboracle_c.so')
_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int32)


class oc_desc(C.Structure):
    _fields_ = [('n', C.c_int32), ('parent', _pi), ('jtype', _pi), ('E0_pj', _pd), ('E0_ji', _pd), ('axis', _pd),
                ('I_i', _pd), ('sides', _pd), ('stiffness', _pd), ('damping', _pd), ('qRest', _pd), ('qLimL', _pd),
                ('qLimU', _pd), ('qLimK', _pd), ('qLimD', _pd), ('grav', C.c_double * 3), ('has_ground', _pi),
                ('ground_E', _pd), ('ground_kn', _pd), ('ground_kt', _pd), ('ground_kd', _pd), ('ground_mu', _pd)]


_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.oc_rollout.argtypes = [C.POINTER(oc_desc), C.c_int, C.c_double, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int]
        L.oc_eval.argtypes = [C.POINTER(oc_desc), vp, vp, vp, vp, C.c_double, C.c_double, vp, vp, vp, vp, vp, vp]
        L.oc_nr.argtypes = [C.POINTER(oc_desc)]
        _lib = L
    return _lib


def max_threads():
    return lib().oc_max_threads()


def flatten(scene):
    """oracle Scene (after init) -> (oc_desc, keepalive list).  Joints in list order; 4x4 transforms row-major."""
    joints = scene.joints
    n = len(joints)
    index = {id(j): i for i, j in enumerate(joints)}
    keep = []

    def arr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data_as(_pd if dt == np.float64 else _pi)
    d = oc_desc()
    d.n = n
    d.parent = arr([(-1 if j.parent is None else index[id(j.parent)]) for j in joints], np.int32)
    for j in joints:
        if j.ndof not in (0, 1) or (j.ndof == 1 and not hasattr(j, 'axis')):
            raise ValueError('the C oracle covers JointRevolute / JointFixed only')
    d.jtype = arr([j.ndof for j in joints], np.int32)
    d.E0_pj = arr(np.concatenate([(np.eye(4) if j.E0_pj is None else j.E0_pj).ravel() for j in joints]), np.float64)
    d.E0_ji = arr(np.concatenate([j.body.E0_ji.ravel() for j in joints]), np.float64)
    d.axis = arr(np.concatenate([(j.axis if j.ndof else np.zeros(3)) for j in joints]), np.float64)
    d.I_i = arr(np.concatenate([j.body.I_i for j in joints]), np.float64)
    d.sides = arr(np.concatenate([j.body.sides for j in joints]), np.float64)
    d.stiffness = arr([j.stiffness for j in joints], np.float64)
    d.damping = arr([j.damping for j in joints], np.float64)
    d.qRest = arr([(j.qRest[0] if j.ndof else 0.0) for j in joints], np.float64)
    d.qLimL = arr([j.qLimL for j in joints], np.float64)
    d.qLimU = arr([j.qLimU for j in joints], np.float64)
    d.qLimK = arr([j.qLimK for j in joints], np.float64)
    d.qLimD = arr([j.qLimD for j in joints], np.float64)
    d.grav = (C.c_double * 3)(*[float(x) for x in scene.grav])
    hg = np.zeros(n, dtype=np.int32)
    gE = np.tile(np.eye(4).ravel(), n).reshape(n, 16)
    gp = np.zeros((4, n))
    for f in scene.forces:
        if type(f).__name__ == 'ForceNull':
            continue
        if type(f).__name__ != 'ForceGroundCuboid':
            raise ValueError('the C oracle covers ForceGroundCuboid only')
        i = index[id(f.cuboid.joint)]
        hg[i] = 1
        gE[i] = np.asarray(f.E, dtype=float).ravel()
        gp[:, i] = [f.kn, f.kt, f.kd, f.mu]
    d.has_ground = arr(hg, np.int32)
    d.ground_E = arr(gE.ravel(), np.float64)
    d.ground_kn = arr(gp[0], np.float64)
    d.ground_kt = arr(gp[1], np.float64)
    d.ground_kd = arr(gp[2], np.float64)
    d.ground_mu = arr(gp[3], np.float64)
    return d, keep


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def run_forward_batch(scene, scheme, q0, qdot0, tau=None, nsteps=None, threads=1, want_qdot=True):
    """B forward rollouts (simLoop of driverRedMaxBDF1/2.m).  q0, qdot0, tau: [B, nr].  Returns q, qdot [B, nsteps, nr],
    stats [B, 3] = (Newton iterations, line-search evaluations, status bits)."""
    d, keep = flatten(scene)
    q0 = np.ascontiguousarray(q0, dtype=np.float64)
    qdot0 = np.ascontiguousarray(qdot0, dtype=np.float64)
    if q0.ndim == 1:
        q0, qdot0 = q0[None, :], qdot0[None, :]
    B, nr = q0.shape
    assert nr == lib().oc_nr(C.byref(d)) == scene.nr
    nsteps = scene.nsteps if nsteps is None else nsteps
    tau = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
    q = np.empty((B, nsteps, nr))
    qd = np.empty((B, nsteps, nr)) if want_qdot else None
    stats = np.zeros((B, 3), dtype=np.int32)
    rc = lib().oc_rollout(C.byref(d), int(scheme), float(scene.h), int(nsteps), int(B), _p(q0), _p(qdot0), _p(tau), _p(q),
                          _p(qd), _p(stats), int(threads))
    assert rc == 0
    return q, qd, stats


def eval_direct(scene, q, qdot, dqtmp, cD, cK, tau=None):
    """One evaluation of g, H, M, D, K, f at (q, qdot) with dqtmp given (cD, cK the stage coefficients)."""
    d, keep = flatten(scene)
    nr = scene.nr
    q, qdot, dqtmp = (np.ascontiguousarray(a, dtype=np.float64) for a in (q, qdot, dqtmp))
    tau = None if tau is None else np.ascontiguousarray(tau, dtype=np.float64)
    g, f = np.empty(nr), np.empty(nr)
    H, M, D, K = (np.empty((nr, nr)) for _ in range(4))
    rc = lib().oc_eval(C.byref(d), _p(q), _p(qdot), _p(dqtmp), _p(tau), float(cD), float(cK), _p(g), _p(H), _p(M), _p(D),
                       _p(K), _p(f))
    assert rc == 0
    return dict(g=g, H=H, M=M, D=D, K=K, f=f)
