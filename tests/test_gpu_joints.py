"""GPU parity for the joint types of SURVEY.md 8(f) rank 1 (JointPrismatic, JointPlanar, JointTranslational, JointFree2D,
JointUniversal): the CUDA path integrates them as chains of one-DOF virtual joints (rmx_api.cu expand_scene); the oracle
restates the reference's own classes (S(q), Sdot, dSdq, ...).  Same bars as test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import rel_err
from test_gpu_parity import TOL_EVAL, TOL_Q, both, oracle_eval

pytestmark = pytest.mark.gpu

SCENES = [3, 4, 5, 6, 8, 11]


@pytest.mark.parametrize('sid', SCENES)
def test_eval_matches_oracle(rb, oracle, sid):
    sg, so = both(rb, oracle, rb.scenesRedMax, sid)
    rng = np.random.default_rng(500 + sid)
    nr, h = sg.nr, sg.h
    for trial in range(3):
        q = sg.qInit + 0.4 * rng.uniform(-1, 1, nr)
        if sid == 11 and trial > 0:
            q[1] = 0.2 + 0.2 * trial  # cuboid centre 0.4 / 0.6 above the ground plane: corners in contact
        q0 = q - 0.02 * rng.uniform(-1, 1, nr)
        qdot0 = rng.uniform(-1, 1, nr)
        tau = 100 * rng.uniform(-1, 1, nr)
        g, H, M, D, f = oracle_eval(oracle, so, q, qdot0, q0, tau)
        out = sg.eval(q, (q - q0) / h, q - q0 - h * qdot0, h * h, 1.0 / h, tau=tau)
        for nm_, ref, tol in (('g', g, TOL_EVAL), ('H', H, TOL_EVAL), ('M', M, TOL_EVAL), ('f', f, 1e-9)):
            assert rel_err(out[nm_], ref) < tol, (sid, trial, nm_, rel_err(out[nm_], ref))
        # D vanishes identically for a single free body (scene 6): measure it on the scale it enters H = M - h D - h^2 K with
        dscale = max(np.max(np.abs(D)), np.max(np.abs(M)))
        assert np.max(np.abs(out['D'] - D)) < 1e-10 * dscale, (sid, trial, 'D', np.max(np.abs(out['D'] - D)), dscale)


@pytest.mark.parametrize('sid', [3, 4, 5, 6, 8])
@pytest.mark.parametrize('scheme', [1, 2])
def test_rollout_and_golden_energy(rb, oracle, sid, scheme):
    """Full-length rollouts: rollout 0 is the reference run itself -> its end-of-run energy must hit the reference's
    recorded Hexpected (|dH| <= 1e-2, Scene.m:172) through the CUDA path; all rollouts match the oracle's q(t)."""
    sg, so = both(rb, oracle, rb.scenesRedMax, sid)
    q0, qd0 = rb.synthetic_inputs(sg, 3, seed=20260100 + sid)
    q0[0], qd0[0] = sg.qInit, sg.qdotInit
    out = sg.rollout(q0, qd0, scheme=scheme)
    assert out['status'].tolist() == [0, 0, 0]
    for b in range(3):
        stats = []
        qs, qds = oracle.run_forward(so, scheme, q0[b], qd0[b], stats=stats)
        assert rel_err(out['q'][b], qs) < TOL_Q, (b, rel_err(out['q'][b], qs))
        it = np.array(stats)
        assert out['iters'][b, 0] == it[:, 0].sum()
        assert out['iters'][b, 1] == it[:, 1].sum()
    T0, V0 = sg.energies(q0[0], qd0[0])
    T1, V1 = sg.energies(out['q'][0, -1], out['qdot'][0, -1])
    Hend = T1[0] + V1[0] - V0[0]
    assert abs(Hend - sg.Hexpected[scheme - 1]) <= 1e-2, (Hend, sg.Hexpected[scheme - 1])


def test_scene11_free2d_ground_golden_energy(rb, oracle):
    """The reference's only end-to-end pin of ForceGroundCuboid (scene 11, Free2D body bouncing on the ground, BDF2,
    1200 steps) reproduced by the CUDA path; q(t) against the oracle over the first bounce."""
    sg, so = both(rb, oracle, rb.scenesRedMax, 11)
    out = sg.rollout(scheme=2)
    T1, V1 = sg.energies(out['q'][0, -1], out['qdot'][0, -1])
    _, V0 = sg.energies(sg.qInit, sg.qdotInit)
    Hend = T1[0] + V1[0] - V0[0]
    assert abs(Hend - sg.Hexpected[1]) <= 1e-2, (Hend, sg.Hexpected[1])
    ns = 400
    qs, _ = oracle.run_forward(so, 2, sg.qInit, sg.qdotInit, nsteps=ns)
    assert rel_err(out['q'][0, :ns], qs) < TOL_Q, rel_err(out['q'][0, :ns], qs)


def test_planar_with_oblique_plane_and_limits(rb, oracle):
    """A planar joint in a non-axis-aligned plane carrying a universal joint with stiffness, damping and active limits:
    exercises axis2, per-DOF qRest and the per-DOF limit masks of Joint.computeForce (Joint.m:448-454)."""
    def build(api):
        s = api.Scene()
        b1 = api.BodyCuboid(1.0, [4, 3, 1])
        j1 = api.JointPlanar(None, b1, np.array([[1.0, 0.0, 1.0], [0.0, 2.0, 0.0]]).T)
        j1.setJointTransform(np.eye(4))
        b1.setBodyTransform(np.eye(4))
        j1.q[:] = [0.3, -0.2]
        j1.setStiffness(5e3)
        j1.setDamping(1e2)
        b2 = api.BodyCuboid(1.0, [1, 1, 6])
        j2 = api.JointUniversal(j1, b2)
        E = np.eye(4)
        E[0:3, 3] = [1, 0.5, -0.5]
        j2.setJointTransform(E)
        E2 = np.eye(4)
        E2[0:3, 3] = [0, 0, -3]
        b2.setBodyTransform(E2)
        j2.q[:] = [0.2, -0.3]
        j2.setLimitLower(-0.25)
        j2.setLimitUpper(0.15)
        j2.setLimitStiffness(1e5)
        j2.setLimitDamping(1e2)
        j2.setStiffness(1e3)
        s.bodies = [b1, b2]
        s.joints = [j1, j2]
        s.h = 2e-3
        s.tEnd = 0.1
        return s
    sg, so = build(rb), build(oracle)
    sg.init()
    so.init()
    assert sg.nr == so.nr == 4
    out = sg.rollout(scheme=2)
    qs, _ = oracle.run_forward(so, 2, sg.qInit, sg.qdotInit)
    assert rel_err(out['q'][0], qs) < TOL_Q, rel_err(out['q'][0], qs)
    # energies at a state that violates both limits of the universal joint (Joint.m:616-637 with per-DOF masks)
    qe = np.array(sg.qInit)
    qe[so.joints[1].idxR] = [0.4, -0.5]
    T, V = sg.energies(qe, sg.qdotInit)
    so.setQ(qe, sg.qdotInit)
    so.update()
    To, Vo = so.computeEnergies()
    assert abs(V[0] - Vo) <= 1e-10 * max(1.0, abs(Vo)) and abs(T[0] - To) <= 1e-10 * max(1.0, abs(To))


@pytest.mark.parametrize('sid', SCENES)
def test_newton_system_through_the_rollout_path(rb, oracle, sid):
    sg, so = both(rb, oracle, rb.scenesRedMax, sid)
    rng = np.random.default_rng(900 + sid)
    nr, h = sg.nr, sg.h
    q = sg.qInit + 0.4 * rng.uniform(-1, 1, nr)
    if sid == 11:
        q[1] = 0.4
    q0 = q - 0.02 * rng.uniform(-1, 1, nr)
    qdot0 = rng.uniform(-1, 1, nr)
    g, H, M, D, f = oracle_eval(oracle, so, q, qdot0, q0, None)
    out = sg.eval_newton(q, (q - q0) / h, q - q0 - h * qdot0, h * h, 1.0 / h)
    assert rel_err(out['H'], H) < TOL_EVAL, rel_err(out['H'], H)
    dx = np.linalg.solve(H, -g)
    assert rel_err(out['dx'], dx) < 1e-12 * max(10.0, np.linalg.cond(H))
