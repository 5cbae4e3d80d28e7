"""CPU: pin the oracle against every golden vector the reference holds for the hot path.

Golden end-of-run energies Hexpected(BDF1/BDF2) from matlab-diff/scenesRedMax.m:54-55, 82-83, 108-109, 292-293,
373-374 and 132-133, 146-147, 166-167, 190-191, 238-239; pass criterion |H(end) - Hexpected| <= 1e-2 as Scene.plotEnergies (Scene.m:171-177)."""
import math

import numpy as np
import pytest

PINS = {
    0: (-1.2705398823489915e+05, 2.6058008179021417e+03),
    1: (-3.8359074258588909e+04, -9.7138545812971279e+02),
    2: (-2.2826101928480086e+04, -2.4159349151742754e+02),
    14: (-2.5928305306546572e+04, -1.8476279319765570e+04),
    11: (-4.4208045000000002e+03, -2.7811251900394832e+03),
    # prismatic / planar / translational / Free2D / universal joints (scenesRedMax.m:132-133,
    # 146-147, 166-167, 190-191, 238-239)
    3: (-3.7579402399569808e+04, -6.1132876082600706e+02),
    4: (-4.5738939646068720e+04, -4.7000178355609387e+02),
    5: (3.3661704151378050e+04, 3.3377464890219308e+04),
    6: (2.0322933333333378e+04, 2.1283333333333332e+04),
    8: (-2.5276246935781084e+04, -1.3781281283808785e+03),
    # rank 3, Euler-chart joints: JointSpherical (scene 7, scenesRedMax.m:206-207) and JointFree3D (scene 9, :250-251)
    7: (-8.7859815791305155e+03, 8.6544602745403390e+03),
    9: (4.3970920953724946e+00, 4.5466508559364156e+00),
    # rank 2, first force with off-diagonal blocks: ForcePointPoint (scene 10 'Loop', scenesRedMax.m:264-265)
    10: (1.2376477982839792e+03, 4.1146190850293169e+03),
    # ForceSpringDamper (scene 12 'Spring-damper', scenesRedMax.m:314-315)
    12: (-2.2145412057327565e+04, -8.9887693524038732e+03),
    # ForceCable through a prismatic, a fixed and two revolute joints (scene 13 'Cables', scenesRedMax.m:342-343)
    13: (-3.1874892332895153e+04, -2.7872894793863266e+04),
}


@pytest.mark.parametrize('sid', [0, 1, 2, 14])
@pytest.mark.parametrize('itype', [1, 2])
def test_hexpected(oracle, sid, itype):
    s = oracle.scenes(sid)
    s.init()
    assert s.Hexpected[itype - 1] == PINS[sid][itype - 1]
    (oracle.sim_loop_bdf1 if itype == 1 else oracle.sim_loop_bdf2)(s)
    ok, H = s.checkEnergy(itype)
    assert ok, (sid, itype, H)
    # far tighter than the reference's 1e-2: the restatement reproduces the printed 17 digits to ~1e-8
    assert abs(H - PINS[sid][itype - 1]) < 1e-6


@pytest.mark.parametrize('sid', [3, 4, 5, 6, 8, 10, 12, 13])
@pytest.mark.parametrize('itype', [1, 2])
def test_hexpected_more_joint_types(oracle, sid, itype):
    """JointPrismatic / JointPlanar / JointTranslational / JointFree2D / JointUniversal restatements against the
    reference's recorded end-of-run energies (scenes built by the api-agnostic factory the GPU path uses too)."""
    import redmax_b200.scenes as scenes
    s = scenes.scenesRedMax(sid, api=oracle)
    s.init()
    assert s.Hexpected[itype - 1] == PINS[sid][itype - 1]
    (oracle.sim_loop_bdf1 if itype == 1 else oracle.sim_loop_bdf2)(s)
    ok, H = s.checkEnergy(itype)
    assert ok, (sid, itype, H)
    assert abs(H - PINS[sid][itype - 1]) < 1e-6


@pytest.mark.parametrize('sid', [7, 9])
@pytest.mark.parametrize('itype', [1, 2])
def test_hexpected_euler_chart_joints(oracle, sid, itype):
    """JointSpherical / JointFree3D against the reference's recorded energies.  Scene 7 under BDF2 is the one run of the
    reference's scene list that re-parameterises (JointSpherical.reparam_, JointSpherical.m:63-103): joint 2 goes
    XYZ -> XYX -> YXZ; under BDF1 and for JointFree3D no switch happens (the reference could not perform one there:
    chart1 is never set)."""
    import redmax_b200.scenes as scenes
    s = scenes.scenesRedMax(sid, api=oracle)
    s.init()
    assert s.Hexpected[itype - 1] == PINS[sid][itype - 1]
    (oracle.sim_loop_bdf1 if itype == 1 else oracle.sim_loop_bdf2)(s)
    ok, H = s.checkEnergy(itype)
    assert ok, (sid, itype, H)
    assert abs(H - PINS[sid][itype - 1]) < 1e-6
    sw = [j.switches for j in s.joints if hasattr(j, 'switches')]
    if (sid, itype) == (7, 2):
        assert sw == [[], [(7, 1), (1, 10)]] and len(s.chart_switch_steps) == 2
    else:
        assert not any(sw) and not s.chart_switch_steps


@pytest.mark.parametrize('chart', range(1, 13))
def test_euler_charts_fd_and_inverse(oracle, chart):
    """euler_chart is restated from the definition (R = R_a R_b R_c, omega_body = T qdot), not from the generated closed
    forms: check every output against central differences, T against R' dR/dq, and the inverse map's round trip."""
    rng = np.random.default_rng(700 + chart)
    q = rng.uniform(-1.2, 1.2, 3)
    if chart <= 6:
        q[1] = abs(q[1]) + 0.2  # proper Euler charts: q2 in (0, pi)
    qd = rng.uniform(-1, 1, 3)
    R, dR, Rdot, dRdot, T, detT, dT, Tdot, dTdot = oracle.euler_chart(chart, q, qd)
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-14) and abs(np.linalg.det(R) - 1) < 1e-14
    np.testing.assert_allclose(oracle.euler_chart_inv(chart, R), q, atol=1e-12)
    e = 1e-6
    for k in range(3):
        dq = np.zeros(3)
        dq[k] = e
        p, m = oracle.euler_chart(chart, q + dq, qd), oracle.euler_chart(chart, q - dq, qd)
        for name, ana, i in (('dR', dR, 0), ('dRdot', dRdot, 2), ('dT', dT, 4), ('dTdot', dTdot, 7)):
            fd = (p[i] - m[i]) / (2 * e)
            assert np.max(np.abs(fd - ana[:, :, k])) < 1e-8, (chart, name, k)
        w = R.T @ dR[:, :, k]  # = [T(:,k)]
        np.testing.assert_allclose([w[2, 1], w[0, 2], w[1, 0]], T[:, k], atol=1e-14)
    np.testing.assert_allclose(Rdot, sum(dR[:, :, k] * qd[k] for k in range(3)), atol=1e-14)
    np.testing.assert_allclose(Tdot, sum(dT[:, :, k] * qd[k] for k in range(3)), atol=1e-14)
    assert abs(detT - np.linalg.det(T)) < 1e-14  # the reference's closed form (-sin q2 / +-cos q2) is det T, sign included


@pytest.mark.parametrize('itype', [2, 1])
def test_hexpected_scene11_ground_contact(oracle, itype):
    """The only end-to-end pin of ForceGroundCuboid (Free2D body bouncing on the ground, 1200 steps)."""
    s = oracle.scenes(11)
    s.init()
    assert s.nsteps == 1200
    (oracle.sim_loop_bdf1 if itype == 1 else oracle.sim_loop_bdf2)(s)
    ok, H = s.checkEnergy(itype)
    assert ok, H
    assert abs(H - PINS[11][itype - 1]) < 1e-4


def _fd_check(name, analytic, numeric, tol=1e-6):
    """redmax.Scene.printError (Scene.m:424-450): relative error below 1e-6."""
    e = np.linalg.norm(analytic - numeric)
    n0, n1 = np.linalg.norm(analytic), np.linalg.norm(numeric)
    if n0 > 1e-4 and n1 > 1e-4:
        e = e / min(n0, n1)
    assert e < tol, (name, e)


@pytest.mark.parametrize('sid', [0, 1, 2, 14])
def test_scene_test_fd_selfcheck(oracle, sid):
    """Scene.test (Scene.m:224-378) restated with a fixed seed: H is the derivative of g, K and D of f."""
    s = oracle.scenes(sid)
    s.init()
    rng = np.random.default_rng(100 + sid)
    nr = s.nr
    q1 = s.qInit + 0.3 * rng.uniform(-1, 1, nr)
    q0 = q1 - 0.01 * rng.uniform(-1, 1, nr)
    s.setQ0(q0, rng.uniform(-1, 1, nr))
    g, H = oracle.eval_bdf1(q1, s, True)
    eps = np.sqrt(np.finfo(float).eps)
    Hn = np.zeros_like(H)
    for i in range(nr):
        x = q1.copy()
        x[i] += eps
        Hn[:, i] = (oracle.eval_bdf1(x, s, False) - g) / eps
    _fd_check('H', H, Hn, 2e-6)


def test_ground_contact_fd(oracle):
    """ForceGroundCuboid.test/test3/test4 (ForceGroundCuboid.m:197-423) restated: Km, Dm are the derivatives of fm
    in a state with corners in contact (static and dynamic friction both occur along the chain)."""
    s = oracle.chain_scene(6, ground=True, h=1e-3)
    for f in s.forces:
        f.E[2, 3] = -12.0
    s.init()
    rng = np.random.default_rng(5)
    nr = s.nr
    q1 = s.qInit + 0.2 * rng.uniform(-1, 1, nr)
    q0 = q1 - 1e-3 * rng.uniform(-1, 1, nr) * 5
    s.setQ0(q0, rng.uniform(-1, 1, nr))
    g, H = oracle.eval_bdf1(q1, s, True)
    # some corner must be in contact for this to test anything
    fm = np.zeros(s.nm)
    for f in s.forces:
        f.computeValues_(None, fm)
    assert np.linalg.norm(fm) > 0
    eps = 1e-7
    Hn = np.zeros_like(H)
    for i in range(nr):
        x = q1.copy()
        x[i] += eps
        Hn[:, i] = (oracle.eval_bdf1(x, s, False) - g) / eps
    _fd_check('H(ground)', H, Hn, 1e-5)


def _central_richardson(P, p, eps):
    g1, g2 = np.zeros_like(p), np.zeros_like(p)
    for i in range(len(p)):
        for g, e in ((g1, eps), (g2, 0.5 * eps)):
            a, b = p.copy(), p.copy()
            a[i] += e
            b[i] -= e
            g[i] = (P(a) - P(b)) / (2 * e)
    return (4.0 * g2 - g1) / 3.0


@pytest.mark.parametrize('sid,scheme,first_step_free,tol', [(100, 1, False, 1e-8), (101, 2, True, 1e-8), (101, 2, False, 5e-2)])
def test_adjoint_gradient_fd(oracle, sid, scheme, first_step_free, tol):
    """The reference's own check of the adjoint (driverRedMaxAdjointBDF1.m:47-61: dP/dp against differences of P obtained by
    re-simulating), sharpened: central differences with one Richardson step.  Scenes 100/101 carry no golden number, so this
    is what pins the oracle's (P, dP/dp).
      * BDF1: the adjoint gradient is the derivative of P to 1e-10 (observed).
      * BDF2 as the reference has it: 4.7 % off.  The reference treats the SDIRK2 first step like a BDF2 step (dgdp coefficient
        -(4/9) h^2 where neither stage has it, and the first stage is ignored altogether, TaskBDF2.m:52-54; SURVEY note N7).
      * BDF2 with the first step taking no control (oracle-only diagnostic switch `first_step_free`, so that the parameters act
        through the plain BDF2 steps only): 3e-10 (observed).  So the treatment of the first step is the ONLY source of the
        BDF2 gap; the backward recursion itself, special first-step blocks of the stencil included, is exact."""
    s = oracle.scenes(sid)
    s.tEnd = 0.2
    s.init()
    s.task.setTime(s.tEnd)
    s.task.first_step_free = first_step_free
    p = np.array([0.02, -0.01])
    P, dPdp = oracle.task_objective(p, s, scheme)
    d = _central_richardson(lambda x: oracle.task_objective(x, s, scheme)[0], p, 1e-4)
    rel = np.linalg.norm(d - dPdp) / np.linalg.norm(d)
    assert rel < tol, (sid, first_step_free, rel, d, dPdp)
    if tol > 1e-3:
        assert rel > 1e-3  # the reference's gradient really is inexact there; if this ever passes tightly, the note above is stale


def _system_quantities(oracle, s):
    """The quantities Scene.test forms before its checks (Scene.m:232-269): J, Jdot, dJdq, dJdotdq, M, dMdq, fqvv, Kqvv, Dqvv,
    f, K, D at the scene's current (q, qdot)."""
    nr, nm = s.nr, s.nm
    qdot = s.getQdot()
    J, Jdot = np.zeros((nm, nr)), np.zeros((nm, nr))
    dJdq, dJdotdq = np.zeros((nm, nr, nr)), np.zeros((nm, nr, nr))
    Mm, Km, Dm = np.zeros((nm, nm)), np.zeros((nm, nm)), np.zeros((nm, nm))
    fm, fr = np.zeros(nm), np.zeros(nr)
    Kr, Dr = np.zeros((nr, nr)), np.zeros((nr, nr))
    for j in s.joints:
        j.computeJacobian4(J, Jdot, dJdq, dJdotdq)
    for b in s.bodies:
        b.computeMassGrav(s.grav, Mm, fm, Km, Dm)
    for j in s.joints:
        j.computeForce(fr, Kr, Dr)
    M = J.T @ Mm @ J
    dMdq = np.zeros((nr, nr, nr))
    for i in range(nr):
        tmp = J.T @ Mm @ dJdq[:, :, i]
        dMdq[:, :, i] = tmp.T + tmp
    fqvv = -J.T @ Mm @ Jdot @ qdot
    Kqvv = np.zeros((nr, nr))
    Dqvv = -J.T @ Mm @ Jdot
    for i in range(nr):
        Kqvv[:, i] = -dJdq[:, :, i].T @ Mm @ Jdot @ qdot - J.T @ Mm @ dJdotdq[:, :, i] @ qdot
        Dqvv[:, i] = Dqvv[:, i] - J.T @ Mm @ dJdq[:, :, i] @ qdot
    f = fr + J.T @ fm + fqvv
    K = Kr + J.T @ Km @ J + Kqvv
    D = Dr + J.T @ Dm @ J + Dqvv
    for i in range(nr):
        K[:, i] = K[:, i] + dJdq[:, :, i].T @ fm + J.T @ Dm @ dJdq[:, :, i] @ qdot
    return dict(J=J, Jdot=Jdot, dJdq=dJdq, dJdotdq=dJdotdq, Mm=Mm, M=M, dMdq=dMdq, fqvv=fqvv, Kqvv=Kqvv, Dqvv=Dqvv, f=f, K=K, D=D)


def _jac(s):
    J, Jdot = np.zeros((s.nm, s.nr)), np.zeros((s.nm, s.nr))
    for j in s.joints:
        j.computeJacobian2(J, Jdot)
    return J, Jdot


def _force(s, J, Jdot, Mm, qdot):
    """f as the K / D checks of Scene.test re-form it (Scene.m:347-353, 362-368)."""
    fm, fr = np.zeros(s.nm), np.zeros(s.nr)
    Mm_ = np.zeros((s.nm, s.nm))
    for b in s.bodies:
        b.computeMassGrav(s.grav, Mm_, fm)
    for j in s.joints:
        j.computeForce(fr)
    return fr + J.T @ fm - J.T @ Mm @ Jdot @ qdot


@pytest.mark.parametrize('sid', [0, 1, 2, 14])
def test_scene_test_all_checks(oracle, sid):
    """The whole list of Scene.test (Scene.m:271-377), every check with the reference's own perturbation (sqrt(eps)) and its
    pass criterion (printError: relative error < 1e-6), at a seeded state instead of the scene's initial one: Jdot, dJ/dq,
    dJdot/dq, dM/dq, Kqvv, Dqvv, K, D.  These are the derivative identities the kernels' world-frame formulation is checked
    against through g, H, M, D (tests/test_gpu_parity.py), here for the dense quantities the reference itself forms."""
    s = oracle.scenes(sid)
    s.init()
    rng = np.random.default_rng(300 + sid)
    nr, nm = s.nr, s.nm
    q = s.qInit + 0.3 * rng.uniform(-1, 1, nr)
    qdot = rng.uniform(-1, 1, nr)
    if sid == 14:
        q[0] = -2.0  # below the lower joint limit: the limit terms of Joint.computeForce take part in K, D
    eps = np.sqrt(np.finfo(float).eps)

    def at(q_, qdot_):
        s.setQ(q_, qdot_)
        s.update()

    at(q, qdot)
    v = _system_quantities(oracle, s)
    J, Jdot, Mm = v['J'], v['Jdot'], v['Mm']
    # Jdot = d/dt J along qdot (Scene.m:276-283)
    at(q + eps * qdot, qdot)
    J_, _ = _jac(s)
    _fd_check('Jdot', Jdot, (J_ - J) / eps)
    # dJ/dq, dJdot/dq (:286-299); dM/dq (:302-314); Kqvv, Dqvv (:317-341); K, D (:344-376)
    dJdq_, dJdotdq_ = np.zeros((nm, nr, nr)), np.zeros((nm, nr, nr))
    dMdq_ = np.zeros((nr, nr, nr))
    Kqvv_, Dqvv_, K_, D_ = (np.zeros((nr, nr)) for _ in range(4))
    for i in range(nr):
        q_ = q.copy()
        q_[i] += eps
        at(q_, qdot)
        J_, Jdot_ = _jac(s)
        dJdq_[:, :, i] = (J_ - J) / eps
        dJdotdq_[:, :, i] = (Jdot_ - Jdot) / eps
        dMdq_[:, :, i] = (J_.T @ Mm @ J_ - v['M']) / eps
        Kqvv_[:, i] = (-J_.T @ Mm @ Jdot_ @ qdot - v['fqvv']) / eps
        K_[:, i] = (_force(s, J_, Jdot_, Mm, qdot) - v['f']) / eps
        qd_ = qdot.copy()
        qd_[i] += eps
        at(q, qd_)
        J_, Jdot_ = _jac(s)
        Dqvv_[:, i] = (-J_.T @ Mm @ Jdot_ @ qd_ - v['fqvv']) / eps
        D_[:, i] = (_force(s, J_, Jdot_, Mm, qd_) - v['f']) / eps
    at(q, qdot)
    _fd_check('dJ/dq', v['dJdq'], dJdq_)
    _fd_check('dJdot/dq', v['dJdotdq'], dJdotdq_)
    _fd_check('dM/dq', v['dMdq'], dMdq_)
    _fd_check('Kqvv', v['Kqvv'], Kqvv_)
    _fd_check('Dqvv', v['Dqvv'], Dqvv_)
    # K, D: forward differences of f, which is quadratic in qdot -- their truncation error at this seeded state reaches 1e-6
    _fd_check('K', v['K'], K_, 2e-6)
    _fd_check('D', v['D'], D_, 2e-6)
