"""GPU parity tests of the adjoint path (rmx_rollout_adjoint: tape-recording forward simLoop of
driverRedMaxAdjointBDF1/2.m + Task*.calcFinal) and of rmx_energies, through the C ABI, against the oracle.

The reference holds no golden (P, dP/dp) for scenes 100/101 ("parity unpinned" by golden numbers);
the oracle's gradient is pinned by the reference's own FD recipe in tests/test_oracle_pins.py, and the CUDA path is
compared with the oracle here.  The energy tests pin the CUDA path directly on the reference's golden Hexpected."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL_P = 1e-10      # objective value
TOL_G = 1e-8       # gradient (z accumulates nsteps transposed solves; observed ~1e-12)


def _pair(rb, oracle, factory, *a, **kw):
    sg = factory(*a, **kw)
    sg.init()
    so = factory(*a, api=oracle, **kw)
    so.init()
    return sg, so


@pytest.mark.parametrize('sid,scheme', [(100, 1), (101, 2)])
def test_adjoint_scene100_101(rb, oracle, sid, scheme):
    """scenesRedMax.m:402-471 as the reference runs them (100 steps, objective at tEnd)."""
    sg, so = _pair(rb, oracle, rb.scenesRedMax, sid)
    rng = np.random.default_rng(sid)
    p = np.vstack([np.zeros(sg.nr), 0.02 * rng.uniform(-1, 1, (2, sg.nr))])
    res = sg.rollout_adjoint(p, want_q=True)
    assert res['status'].tolist() == [0, 0, 0]
    for b in range(3):
        P, dPdp = oracle.task_objective(p[b], so, scheme)
        qs = np.array([r['q'] for r in so.history])
        assert rel_err(res['q'][b], qs) < 1e-10
        assert abs(res['P'][b] - P) <= TOL_P * abs(P), (res['P'][b], P)
        assert rel_err(res['dPdp'][b], dPdp) < TOL_G, (res['dPdp'][b], dPdp)


@pytest.mark.parametrize('scheme', [1, 2])
def test_adjoint_mid_rollout_objective(rb, oracle, scheme):
    """Objective sampled in the middle of the rollout (exercises the 2-/4-step backward stencil on both sides of the
    objective step and the |t_target - t| < 1e-6 test on the accumulated time)."""
    sid = 100 if scheme == 1 else 101
    sg, so = _pair(rb, oracle, rb.scenesRedMax, sid)
    for s in (sg, so):
        s.task.setTime(0.37)
    p = np.array([[0.03, -0.02], [-0.01, 0.015]])
    res = sg.rollout_adjoint(p, nsteps=60)
    so.nsteps = 60
    for b in range(2):
        P, dPdp = oracle.task_objective(p[b], so, scheme)
        assert abs(res['P'][b] - P) <= TOL_P * abs(P)
        assert rel_err(res['dPdp'][b], dPdp) < TOL_G, (res['dPdp'][b], dPdp)
        assert np.linalg.norm(dPdp - sg.task.wreg * p[b]) > 1e-3  # the objective term really contributes


@pytest.mark.parametrize('scheme', [1, 2])
def test_adjoint_hand_c4(rb, oracle, scheme):
    """C4 shape (fixed palm + 5 fingers x 4 revolute, 20 dof, branching), per-rollout targets and initial states."""
    ns = 12
    sg, so = _pair(rb, oracle, rb.hand_scene, nsteps=ns, scheme=scheme)
    B = 3
    rng = np.random.default_rng(20260004)
    p = 0.01 * rng.uniform(-1, 1, (B, sg.nr))
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=20260004)
    xt = np.array(sg.task.xtarget)[None, :] + rng.uniform(-2, 2, (B, 3))
    res = sg.rollout_adjoint(p, xtarget=xt, q0=q0, qdot0=qd0)
    for b in range(B):
        so.qInit, so.qdotInit = q0[b].copy(), qd0[b].copy()
        so.task.setTarget(xt[b])
        P, dPdp = oracle.task_objective(p[b], so, scheme)
        assert abs(res['P'][b] - P) <= TOL_P * abs(P), (res['P'][b], P)
        assert rel_err(res['dPdp'][b], dPdp) < TOL_G, (rel_err(res['dPdp'][b], dPdp))


@pytest.mark.parametrize('scheme', [1, 2])
def test_adjoint_with_a_spherical_joint(rb, oracle, scheme):
    """Objective and gradient through an Euler-angle joint (JointSpherical in its chart XYZ, a revolute link below it): the
    parameters are the joint torques in the Euler coordinates (TaskBDF1PointPos.applyStep, :58-64), dP/dq uses the joint's
    S = T(q) (Joint.computeJacobian) -- reproduced by the three virtual revolute joints.  No chart switch in this run."""
    def build(api):
        s = api.Scene()
        b1 = api.BodyCuboid(1.0, [1, 1, 8])
        j1 = api.JointSpherical(None, b1)
        j1.setJointTransform(np.eye(4))
        E = np.eye(4)
        E[0:3, 3] = [0, 0, -4]
        b1.setBodyTransform(E)
        j1.q[:] = [0.4, -0.3, 0.2]
        j1.qdot[:] = [0.5, -0.2, 0.3]
        j1.setStiffness(1e4)
        j1.setDamping(1e4)
        b2 = api.BodyCuboid(1.0, [6, 1, 1])
        j2 = api.JointRevolute(j1, b2, [0, 1, 0])
        E2 = np.eye(4)
        E2[0:3, 3] = [0, 0, -8]
        j2.setJointTransform(E2)
        E3 = np.eye(4)
        E3[0:3, 3] = [3, 0, 0]
        b2.setBodyTransform(E3)
        j2.q[0] = 0.5
        j2.setStiffness(1e4)
        j2.setDamping(1e4)
        s.bodies = [b1, b2]
        s.joints = [j1, j2]
        s.h = 1e-2
        s.tEnd = 0.3
        s.task = (api.TaskBDF1PointPos if scheme == 1 else api.TaskBDF2PointPos)(s)
        s.task.setTime(s.tEnd)
        s.task.setBody(b2)
        s.task.setPoint([3, 0, 0])
        s.task.setTarget([4, 2, -9])
        s.task.setScale(1e5)
        s.task.setWeights(1e-2, 1e2)
        return s
    sg, so = build(rb), build(oracle)
    sg.init()
    so.init()
    assert sg.nr == so.nr == 4
    rng = np.random.default_rng(77 + scheme)
    p = np.vstack([np.zeros(4), 0.02 * rng.uniform(-1, 1, (2, 4))])
    res = sg.rollout_adjoint(p, want_q=True)
    assert res['status'].tolist() == [0, 0, 0]
    for b in range(3):
        P, dPdp = oracle.task_objective(p[b], so, scheme)
        assert not so.chart_switch_steps
        qs = np.array([r['q'] for r in so.history])
        assert rel_err(res['q'][b], qs) < 1e-10
        assert abs(res['P'][b] - P) <= TOL_P * abs(P), (res['P'][b], P)
        assert rel_err(res['dPdp'][b], dPdp) < TOL_G, (res['dPdp'][b], dPdp)
        assert np.linalg.norm(dPdp - sg.task.wreg * p[b]) > 1e-3


def test_adjoint_sharded_over_gpus_is_bitwise_identical(rb):
    """rmx_rollout_adjoint with opts.ngpus = G (C4: batch 8192 on 4 GPUs, 2048 each; here a small uneven batch): objective and
    gradient of every rollout equal the one-GPU call bit for bit.  Needs at least two GPUs."""
    from redmax_b200 import _ffi
    G = min(int(_ffi.lib().rmx_device_count()), 4)
    if G < 2:
        pytest.skip('needs >= 2 GPUs (run under gpurun --gpus 2)')
    for scheme in (1, 2):
        sg = rb.hand_scene(nsteps=12, scheme=scheme)
        sg.init()
        B = 301
        rng = np.random.default_rng(20260004)
        p = 0.01 * rng.uniform(-1, 1, (B, sg.nr))
        q0, qd0 = rb.synthetic_inputs(sg, B, seed=20260004)
        xt = np.array(sg.task.xtarget)[None, :] + rng.uniform(-2, 2, (B, 3))
        ref = sg.rollout_adjoint(p, xtarget=xt, q0=q0, qdot0=qd0, ngpus=1)
        out = sg.rollout_adjoint(p, xtarget=xt, q0=q0, qdot0=qd0, ngpus=G)
        for k in ('P', 'dPdp', 'status'):
            np.testing.assert_array_equal(out[k], ref[k])


def test_adjoint_gradient_is_a_gradient(rb):
    """Size-independent property at a larger batch: dP/dp from the adjoint kernels agrees with central differences
    of P from the same kernels (the reference's FD recipe, driverRedMaxAdjointBDF1.m:47-61), for every rollout."""
    sg = rb.scenesRedMax(100)
    sg.tEnd = 0.3
    sg.init()
    sg.task.setTime(sg.tEnd)
    B = 64
    rng = np.random.default_rng(9)
    p = 0.02 * rng.uniform(-1, 1, (B, sg.nr))
    base = sg.rollout_adjoint(p)
    eps = 1e-6
    for i in range(sg.nr):
        dp = np.zeros(sg.nr)
        dp[i] = eps
        Pp = sg.rollout_adjoint(p + dp)['P']
        Pm = sg.rollout_adjoint(p - dp)['P']
        fd = (Pp - Pm) / (2 * eps)
        np.testing.assert_allclose(base['dPdp'][:, i], fd, rtol=2e-4, atol=1e-6 * np.abs(fd).max())


@pytest.mark.parametrize('sid', [0, 1, 2, 14])
@pytest.mark.parametrize('itype', [1, 2])
def test_cuda_path_hits_reference_golden_energy(rb, oracle, sid, itype):
    """The reference's own pass criterion (Scene.plotEnergies, Scene.m:164-177) applied to the CUDA path: H(end) =
    T_end + V_end - V_0 from a full GPU rollout + rmx_energies must match scenesRedMax.m's Hexpected within 1e-2."""
    sg = rb.scenesRedMax(sid)
    sg.init()
    out = sg.rollout(scheme=itype)
    assert out['status'][0] == 0
    T, V = sg.energies(np.vstack([sg.qInit, out['q'][0, -1]]), np.vstack([sg.qdotInit, out['qdot'][0, -1]]))
    H = T[1] + (V[1] - V[0])
    assert abs(H - sg.Hexpected[itype - 1]) <= 1e-2, (H, sg.Hexpected[itype - 1])


def test_energies_match_oracle_with_ground(rb, oracle):
    kw = dict(ground=True, h=5e-4, ground_z=-48.5)
    sg, so = _pair(rb, oracle, rb.chain_scene, 6, **kw)
    q, qd = rb.synthetic_inputs(sg, 5, seed=3)
    q[:, 0] -= 0.3  # push corners into the ground
    T, V = sg.energies(q, qd)
    for b in range(5):
        so.setQ(q[b], qd[b])
        so.update()
        To, Vo = so.computeEnergies()
        assert abs(T[b] - To) <= 1e-11 * abs(To) and abs(V[b] - Vo) <= 1e-11 * abs(Vo)
    so.setQ(q[0], qd[0])
    so.update()
    Vg = 0.0
    for f in so.forces:
        Vg = f.computeEnergy_(Vg)
    assert Vg > 0  # the ground penalty energy is exercised
