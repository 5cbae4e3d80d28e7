"""Reference-named drivers over the batched GPU path (host mirror of matlab-diff/driverRedMax*.m).

    driverRedMaxBDF1(sceneID, batch)          driverRedMaxBDF1.m:1-55
    driverRedMaxBDF2(sceneID, batch)          driverRedMaxBDF2.m:1-55
    driverRedMaxAdjointBDF1() / ...BDF2()     driverRedMaxAdjointBDF1.m:1-37, driverRedMaxAdjointBDF2.m
    taskObjective(p, scene)                   driverRedMaxAdjointBDF1.m:39-63

Same names, argument meaning and console lines as the reference; what differs is only that `simLoop` is one call into the
CUDA library for all rollouts at once (`q0`, `qdot0`, `tau` add the batch the reference does not have), and that the
adjoint drivers minimise with SciPy's BFGS (the reference calls MATLAB's `fminunc` quasi-Newton with the analytic
gradient, which is outside the parity boundary: SURVEY.md section 8(c)).  Failure behaviour follows the reference: Newton
divergence / non-convergence is reported per rollout and the run continues (driverRedMaxBDF1.m:118-121, 150-153); an
energy mismatch prints '### FAIL' (Scene.m:172-177) and is returned, not raised.  Nothing here computes dynamics.
"""
from __future__ import annotations

import sys

import numpy as np

from . import _ffi, euler
from .scenes import scenesRedMax

BDF1, BDF2 = 1, 2


def _report_status(status, out=sys.stdout):
    """The reference prints one line per failing Newton solve; here one line per failing rollout."""
    for bit, msg in ((_ffi.RMX_ST_DIVERGED, 'Newton diverged'), (_ffi.RMX_ST_MAXITER, 'Newton did not converge')):
        bad = np.nonzero(status & bit)[0]
        if len(bad):
            print('%s (rollouts %s%s)' % (msg, bad[:8].tolist(), ' ...' if len(bad) > 8 else ''), file=out)
    bad = np.nonzero(status & _ffi.RMX_ST_CHART)[0]
    if len(bad):
        # the states at which the reference itself stops: reparam_ without chart1 (driverRedMaxBDF1, JointFree3D)
        print('Euler chart left its well-conditioned range and cannot be re-parameterised here (rollouts %s%s)'
              % (bad[:8].tolist(), ' ...' if len(bad) > 8 else ''), file=out)


def simLoop(scene, itype, q0=None, qdot0=None, tau=None, ngpus=1, out=sys.stdout):
    """simLoop(scene) of driverRedMaxBDF1.m:57-91 (itype 1) / driverRedMaxBDF2.m:57-125 (itype 2) for a batch of initial
    states (default: the scene's own, batch of one).  Returns dict(q, qdot, status, iters, T, V, H): q[b, k] = history(k).q,
    H[b] = T_end + V_end - V_0 as Scene.plotEnergies forms it."""
    res = scene.rollout(q0, qdot0, tau=tau, scheme=itype, ngpus=ngpus)
    for b, sw in enumerate(res.get('chart_switches', ())):
        for _, _, old, new in sw:  # JointSpherical.m:84
            print('%s->%s' % (euler.CHART_NAMES[old], euler.CHART_NAMES[new]), file=out)
    _report_status(res['status'], out)
    q_start = np.atleast_2d(scene.qInit if q0 is None else q0)
    qd_start = np.atleast_2d(scene.qdotInit if qdot0 is None else qdot0)
    _, V0 = scene.energies(q_start, qd_start)
    T1, V1 = scene.energies(res['q'][:, -1], res['qdot'][:, -1], chart=res.get('chart'))
    res['T'], res['V'], res['H'] = T1, V1 - V0, T1 + (V1 - V0)
    return res


def plotEnergies(scene, itype, H_end, out=sys.stdout):
    """The pass/fail line of Scene.plotEnergies (Scene.m:171-177); returns True / False / None (no expected value)."""
    if scene.Hexpected[itype - 1] == 0:
        return None
    ok = abs(H_end - scene.Hexpected[itype - 1]) <= 1e-2
    print('### PASS ###' if ok else '### FAIL: %.16e ###' % H_end, file=(out if ok else sys.stderr))
    return bool(ok)


def _driver(itype, sceneID, batch, q0, qdot0, tau, ngpus, out):
    scene = scenesRedMax(sceneID)
    scene.init()
    print("(%d) '%s': tEnd=%.1f, nsteps=%d, nr=%d, nm=%d" % (sceneID, scene.name, scene.tEnd, scene.nsteps, scene.nr, scene.nm),
          file=out)
    res = simLoop(scene, itype, q0, qdot0, tau, ngpus, out)
    res['scene'] = scene
    # the energy pin belongs to the scene's own initial state; with a user batch it is checked for rollout 0 only if that
    # rollout starts there
    own = q0 is None or (np.array_equal(np.atleast_2d(q0)[0], scene.qInit) and
                         (qdot0 is None or np.array_equal(np.atleast_2d(qdot0)[0], scene.qdotInit)))
    res['pass'] = plotEnergies(scene, itype, float(res['H'][0]), out) if own else None
    return res


def driverRedMaxBDF1(sceneID=0, batch=False, q0=None, qdot0=None, tau=None, ngpus=1, out=sys.stdout):
    """driverRedMaxBDF1(sceneID, batch).  `batch` is accepted for signature parity (the reference uses it to switch drawing
    off; nothing is drawn here).  q0 / qdot0 [B, nr], tau [B, nr] or [B, nsteps, nr]: optional batch of rollouts."""
    return _driver(BDF1, sceneID, batch, q0, qdot0, tau, ngpus, out)


def driverRedMaxBDF2(sceneID=0, batch=False, q0=None, qdot0=None, tau=None, ngpus=1, out=sys.stdout):
    """driverRedMaxBDF2(sceneID, batch): one SDIRK2 step, then BDF2."""
    return _driver(BDF2, sceneID, batch, q0, qdot0, tau, ngpus, out)


def taskObjective(p, scene, xtarget=None, ngpus=1):
    """[P, dPdp] = taskObjective(p, scene) (driverRedMaxAdjointBDF1.m:39-45): scene.reset, task.p = p, task.init, adjoint
    simLoop, task.calcFinal -- for one parameter vector (1-D p: returns scalars/1-D) or a batch ([B, np])."""
    p = np.asarray(p, dtype=float)
    single = p.ndim == 1
    res = scene.rollout_adjoint(p, xtarget=xtarget, ngpus=ngpus)
    _report_status(res['status'])
    if single:
        return float(res['P'][0]), res['dPdp'][0]
    return res['P'], res['dPdp']


def _adjoint_driver(sceneID, optimize, maxiter, out):
    from scipy.optimize import minimize
    scene = scenesRedMax(sceneID)
    scene.init()
    print("(%d) '%s': tEnd=%.1f, nsteps=%d, nr=%d, nm=%d" % (sceneID, scene.name, scene.tEnd, scene.nsteps, scene.nr, scene.nm),
          file=out)
    pInit = np.zeros(scene.nr) if getattr(scene.task, 'p', None) is None else np.asarray(scene.task.p, dtype=float)
    if not optimize:
        P, dPdp = taskObjective(pInit, scene)
        return dict(scene=scene, p=pInit, P=P, dPdp=dPdp)
    hist = []

    def fun(p):
        P, g = taskObjective(p, scene)
        hist.append(P)
        return P, g
    r = minimize(fun, pInit, jac=True, method='BFGS', options=dict(maxiter=maxiter, gtol=1e-6))
    print('p = [', file=out)
    print(r.x, file=out)
    print('];', file=out)
    return dict(scene=scene, p=r.x, P=float(r.fun), dPdp=r.jac, history=hist, result=r)


def driverRedMaxAdjointBDF1(optimize=True, maxiter=100, out=sys.stdout):
    """driverRedMaxAdjointBDF1(): scene 100, optimise the constant joint torques so that the tip reaches the target."""
    return _adjoint_driver(100, optimize, maxiter, out)


def driverRedMaxAdjointBDF2(optimize=True, maxiter=100, out=sys.stdout):
    """driverRedMaxAdjointBDF2(): scene 101, the same through SDIRK2 + BDF2."""
    return _adjoint_driver(101, optimize, maxiter, out)
