"""GPU parity for the Euler-angle joints of SURVEY.md 8(f) rank 3 (JointSpherical, JointFree3D; scenes 7 and 9 of
scenesRedMax.m).  The CUDA path integrates them in the reference's initial chart XYZ as three revolute (plus three prismatic)
virtual joints (rmx_api.cu expand_scene); the oracle restates the reference's classes, T(q), Tdot and their derivatives, and
the chart switching of JointSpherical.reparam_.  Same bars as test_gpu_parity.py."""
import numpy as np
import pytest

import redmax_b200 as rmx
from conftest import rel_err
from test_gpu_parity import TOL_EVAL, TOL_Q, both, oracle_eval

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('sid', [7, 9])
def test_eval_matches_oracle(rb, oracle, sid):
    sg, so = both(rb, oracle, rb.scenesRedMax, sid)
    rng = np.random.default_rng(1500 + sid)
    nr, h = sg.nr, sg.h
    for trial in range(3):
        q = sg.qInit + 0.4 * rng.uniform(-1, 1, nr)
        q0 = q - 0.02 * rng.uniform(-1, 1, nr)
        qdot0 = rng.uniform(-1, 1, nr)
        tau = 100 * rng.uniform(-1, 1, nr)
        g, H, M, D, f = oracle_eval(oracle, so, q, qdot0, q0, tau)
        out = sg.eval(q, (q - q0) / h, q - q0 - h * qdot0, h * h, 1.0 / h, tau=tau)
        for nm_, ref, tol in (('g', g, TOL_EVAL), ('H', H, TOL_EVAL), ('M', M, TOL_EVAL), ('f', f, 1e-9)):
            assert rel_err(out[nm_], ref) < tol, (sid, trial, nm_, rel_err(out[nm_], ref))
        dscale = max(np.max(np.abs(D)), np.max(np.abs(M)))
        assert np.max(np.abs(out['D'] - D)) < 1e-10 * dscale, (sid, trial, 'D', np.max(np.abs(out['D'] - D)), dscale)
        on = sg.eval_newton(q, (q - q0) / h, q - q0 - h * qdot0, h * h, 1.0 / h, tau=tau)
        assert rel_err(on['H'], H) < TOL_EVAL, rel_err(on['H'], H)
        assert rel_err(on['dx'], np.linalg.solve(H, -g)) < 1e-12 * max(10.0, np.linalg.cond(H))


@pytest.mark.timeout(600)
@pytest.mark.parametrize('sid,scheme', [(7, 1), (9, 1), (9, 2)])
def test_rollout_and_golden_energy(rb, oracle, sid, scheme):
    """The runs of the reference's scene list in which no chart switch happens: rollout 0 is the reference run itself -> its
    end-of-run energy must hit the recorded Hexpected (|dH| <= 1e-2, Scene.m:172) through the CUDA path; every rollout
    matches the oracle's q(t) with identical Newton / line-search counts and reports no chart flag."""
    sg, so = both(rb, oracle, rb.scenesRedMax, sid)
    B = 2
    rng = np.random.default_rng(20260100 + sid)
    q0 = sg.qInit[None, :] + 0.05 * rng.uniform(-1, 1, (B, sg.nr))
    qd0 = sg.qdotInit[None, :] + 0.05 * rng.uniform(-1, 1, (B, sg.nr))
    q0[0], qd0[0] = sg.qInit, sg.qdotInit
    out = sg.rollout(q0, qd0, scheme=scheme)
    assert out['status'].tolist() == [0] * B
    for b in range(B):
        stats = []
        qs, qds = oracle.run_forward(so, scheme, q0[b], qd0[b], stats=stats)
        assert not so.chart_switch_steps
        assert rel_err(out['q'][b], qs) < TOL_Q, (b, rel_err(out['q'][b], qs))
        it = np.array(stats)
        assert out['iters'][b, 0] == it[:, 0].sum()
        assert out['iters'][b, 1] == it[:, 1].sum()
    T0, V0 = sg.energies(q0[0], qd0[0])
    T1, V1 = sg.energies(out['q'][0, -1], out['qdot'][0, -1])
    Hend = T1[0] + V1[0] - V0[0]
    assert abs(Hend - sg.Hexpected[scheme - 1]) <= 1e-2, (Hend, sg.Hexpected[scheme - 1])


@pytest.mark.timeout(600)
def test_scene7_bdf2_reports_the_chart_switch(rb, oracle):
    """Scene 7 under BDF2 is the one reference run that re-parameterises (joint 2: XYZ -> XYX -> YXZ).  The CUDA path keeps
    chart XYZ: it must agree with the oracle up to the step whose result the reference re-parameterises, and flag the
    rollout with RMX_ST_CHART (the trajectory after that step is the same motion integrated in other coordinates)."""
    sg, so = both(rb, oracle, rb.scenesRedMax, 7)
    out = sg.rollout(scheme=2)
    qs, _ = oracle.run_forward(so, 2, sg.qInit, sg.qdotInit)
    assert len(so.chart_switch_steps) == 2
    k1 = so.chart_switch_steps[0]
    assert k1 > 50
    assert rel_err(out['q'][0, :k1], qs[:k1]) < TOL_Q, rel_err(out['q'][0, :k1], qs[:k1])
    assert out['status'][0] == rmx.RMX_ST_CHART
    # a rollout cut before the switch carries no flag
    short = sg.rollout(scheme=2, nsteps=k1)
    assert short['status'][0] == 0
    np.testing.assert_array_equal(short['q'][0], out['q'][0, :k1])


def test_free3d_under_a_revolute_parent_with_ground(rb, oracle):
    """JointFree3D below another joint, with ground contact on its body: the six virtual joints sit in the middle of a tree
    and carry an external-force block."""
    def build(api):
        s = api.Scene()
        b1 = api.BodyCuboid(1.0, [6, 1, 1])
        j1 = api.JointRevolute(None, b1, [0, 1, 0])
        j1.setJointTransform(np.eye(4))
        E = np.eye(4)
        E[0:3, 3] = [3, 0, 0]
        b1.setBodyTransform(E)
        j1.q[0] = 0.3
        b2 = api.BodyCuboid(1.0, [2, 1, 1])
        j2 = api.JointFree3D(j1, b2)
        E2 = np.eye(4)
        E2[0:3, 3] = [6, 0, 0]
        j2.setJointTransform(E2)
        b2.setBodyTransform(np.eye(4))
        j2.q[:] = [0.5, -0.2, 0.1, 0.3, -0.4, 0.2]
        j2.qdot[:] = [1.0, 0.5, -2.0, 0.4, -0.3, 0.8]
        j2.setStiffness(2e2)
        j2.setDamping(2e1)
        b3 = api.BodyCuboid(1.0, [1, 1, 4])
        j3 = api.JointSpherical(j2, b3)
        E3 = np.eye(4)
        E3[0:3, 3] = [1, 0, 0]
        j3.setJointTransform(E3)
        E4 = np.eye(4)
        E4[0:3, 3] = [0, 0, -2]
        b3.setBodyTransform(E4)
        j3.q[:] = [0.2, 0.3, -0.1]
        f = api.ForceGroundCuboid(b3)
        Eg = np.eye(4)
        Eg[0:3, 3] = [0, 0, -5.52]
        f.setTransform(Eg)
        f.setStiffness(1e5, 1e2)
        f.setDamping(3e1)
        f.setFriction(0.5)
        s.bodies = [b1, b2, b3]
        s.joints = [j1, j2, j3]
        s.forces = [f]
        s.grav = np.array([0.0, 0.0, -98.0])
        s.h = 1e-3
        s.tEnd = 0.05
        return s
    sg, so = build(rb), build(oracle)
    sg.init()
    so.init()
    assert sg.nr == so.nr == 10
    for scheme in (1, 2):
        out = sg.rollout(scheme=scheme)
        assert out['status'][0] == 0
        stats = []
        qs, _ = oracle.run_forward(so, scheme, sg.qInit, sg.qdotInit, stats=stats)
        assert rel_err(out['q'][0], qs) < TOL_Q, (scheme, rel_err(out['q'][0], qs))
        assert out['iters'][0, 0] == np.array(stats)[:, 0].sum()
    T, V = sg.energies(sg.qInit, sg.qdotInit)
    so.setQ(sg.qInit, sg.qdotInit)
    so.update()
    To, Vo = so.computeEnergies()
    assert abs(V[0] - Vo) <= 1e-10 * max(1.0, abs(Vo)) and abs(T[0] - To) <= 1e-10 * max(1.0, abs(To))
