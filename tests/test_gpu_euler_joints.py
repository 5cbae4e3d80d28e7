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
def test_scene7_bdf2_switches_charts_like_the_reference(rb, oracle):
    """Scene 7 under BDF2 is the one reference run that re-parameterises (joint 2: XYZ -> XYX -> YXZ,
    JointSpherical.reparam_).  The library flags the step (RMX_ST_CHART), the host re-expresses it and the BDF2 history in
    the chart the reference picks, and the rollout resumes under that chart (rmx_rollout_resume): same switch steps and
    charts as the oracle, q(t) within 1e-10 over the whole run, same Newton counts, and the recorded Hexpected(BDF2)."""
    sg, so = both(rb, oracle, rb.scenesRedMax, 7)
    B = 3
    rng = np.random.default_rng(20260107)
    q0 = sg.qInit[None, :] + 0.05 * rng.uniform(-1, 1, (B, sg.nr))
    qd0 = sg.qdotInit[None, :] + 0.05 * rng.uniform(-1, 1, (B, sg.nr))
    q0[0], qd0[0] = sg.qInit, sg.qdotInit
    out = sg.rollout(q0, qd0, scheme=2)
    assert out['status'].tolist() == [0] * B
    nsw = 0
    for b in range(B):
        stats = []
        qs, qds = oracle.run_forward(so, 2, q0[b], qd0[b], stats=stats)
        sw_o = [s for j in so.joints for s in j.switches]
        assert [k for k, _, _, _ in out['chart_switches'][b]] == so.chart_switch_steps
        assert sorted((o_, n_) for _, _, o_, n_ in out['chart_switches'][b]) == sorted(sw_o)
        assert out['chart'][b].tolist() == [j.chart for j in so.joints]
        assert rel_err(out['q'][b], qs) < TOL_Q, (b, rel_err(out['q'][b], qs))
        assert rel_err(out['qdot'][b], qds) < 1e-8, (b, rel_err(out['qdot'][b], qds))
        it = np.array(stats)
        assert out['iters'][b, 0] == it[:, 0].sum() and out['iters'][b, 1] == it[:, 1].sum()
        nsw += len(so.chart_switch_steps)
    assert out['chart_switches'][0] == [(out['chart_switches'][0][0][0], 1, 7, 1), (out['chart_switches'][0][1][0], 1, 1, 10)]
    assert nsw >= 2
    T1, V1 = sg.energies(out['q'][:, -1], out['qdot'][:, -1], chart=out['chart'])
    _, V0 = sg.energies(q0, qd0)
    assert abs(T1[0] + V1[0] - V0[0] - sg.Hexpected[1]) <= 1e-2, (T1[0] + V1[0] - V0[0], sg.Hexpected[1])
    # without the host step the library only reports: flag set, trajectory identical up to the first switch step
    raw = sg.rollout(q0, qd0, scheme=2, reparam=False)
    k1 = out['chart_switches'][0][0][0]
    assert raw['status'][0] == rmx.RMX_ST_CHART
    np.testing.assert_array_equal(raw['q'][0, :k1], out['q'][0, :k1])
    short = sg.rollout(scheme=2, nsteps=k1)  # cut before the switch: no flag
    assert short['status'][0] == 0
    np.testing.assert_array_equal(short['q'][0], out['q'][0, :k1])


@pytest.mark.timeout(900)
def test_chart_switching_in_a_batch(rb, oracle):
    """96 differently started rollouts of the double spherical pendulum, BDF2: they switch at different steps into different
    charts, some several times, some never; the rounds of Scene._reparam_rollouts handle them together.  Spot checks against
    the oracle, every rollout ends unflagged, and the final energies (evaluated in each rollout's own charts) are those of a
    plausible run (no energy blow-up from a mis-expressed history)."""
    sg, so = both(rb, oracle, rb.scenesRedMax, 7)
    B = 96
    rng = np.random.default_rng(20260777)
    q0 = sg.qInit[None, :] + 0.3 * rng.uniform(-1, 1, (B, sg.nr))
    qd0 = sg.qdotInit[None, :] + 0.5 * rng.uniform(-1, 1, (B, sg.nr))
    ns = 250
    out = sg.rollout(q0, qd0, scheme=2, nsteps=ns)
    nsw = np.array([len(s) for s in out['chart_switches']])
    assert (out['status'] == 0).all(), out['status']
    assert (nsw > 0).sum() >= 10 and (nsw == 0).sum() >= 1, nsw
    assert len({tuple(c) for c in out['chart'].tolist()}) >= 3  # several different final chart combinations
    multi = int(np.argmax(nsw))
    for b in sorted({0, multi, int(np.nonzero(nsw > 0)[0][-1]), int(np.nonzero(nsw == 0)[0][0])}):
        stats = []
        qs, qds = oracle.run_forward(so, 2, q0[b], qd0[b], nsteps=ns, stats=stats)
        assert [k for k, _, _, _ in out['chart_switches'][b]] == so.chart_switch_steps, b
        assert out['chart'][b].tolist() == [j.chart for j in so.joints], b
        assert rel_err(out['q'][b], qs) < TOL_Q, (b, rel_err(out['q'][b], qs))
        it = np.array(stats)
        assert out['iters'][b, 0] == it[:, 0].sum() and out['iters'][b, 1] == it[:, 1].sum()
    T1, V1 = sg.energies(out['q'][:, -1], out['qdot'][:, -1], chart=out['chart'])
    T0, V0 = sg.energies(q0, qd0)
    dH = (T1 + V1) - (T0 + V0)
    assert np.isfinite(dH).all() and np.abs(dH).max() < 0.5 * np.abs(V0).max(), (np.abs(dH).max(), np.abs(V0).max())


def test_resume_reproduces_the_uncut_rollout(rb):
    """rmx_rollout_resume from the states of an earlier call is bitwise the uncut rollout, for BDF1 and BDF2, from step 0, 1,
    2 and mid-way, per rollout."""
    import ctypes as C
    from redmax_b200 import _ffi
    sg = rb.chain_scene(6, nsteps=12, h=1e-3)
    sg.init()
    B = 5
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=3)
    for scheme in (1, 2):
        ref = sg.rollout(q0, qd0, scheme=scheme)
        q, qd = ref['q'].copy(), ref['qdot'].copy()
        kb = np.array([0, 1, 2, 7, 12], dtype=np.int32)
        for b in range(B):
            q[b, kb[b]:] = np.nan
            qd[b, kb[b]:] = np.nan
        st = np.zeros(B, dtype=np.int32)
        it = np.zeros((B, 2), dtype=np.int32)
        o = sg.opts(scheme=scheme)
        _ffi.check(_ffi.lib().rmx_rollout_resume(sg._handle, C.byref(o), B, _ffi.ptr(kb), None, _ffi.ptr(q0), _ffi.ptr(qd0), None,
                                                 _ffi.ptr(q), _ffi.ptr(qd), _ffi.ptr(st), _ffi.ptr(it)), 'rmx_rollout_resume')
        np.testing.assert_array_equal(q, ref['q'])
        np.testing.assert_array_equal(qd, ref['qdot'])
        assert st.tolist() == [0] * B and it[0].tolist() == ref['iters'][0].tolist() and it[4].tolist() == [0, 0]
        # with an end step: only [k_begin, k_end) is touched, and the pieces' iteration counts add up
        ke = np.array([5, 1, 9, 12, 12], dtype=np.int32)
        q2, qd2 = ref['q'].copy(), ref['qdot'].copy()
        for b in range(B):
            q2[b, kb[b]:ke[b]] = np.nan
        it2 = np.zeros((B, 2), dtype=np.int32)
        _ffi.check(_ffi.lib().rmx_rollout_resume(sg._handle, C.byref(o), B, _ffi.ptr(kb), _ffi.ptr(ke), _ffi.ptr(q0), _ffi.ptr(qd0),
                                                 None, _ffi.ptr(q2), _ffi.ptr(qd2), _ffi.ptr(st), _ffi.ptr(it2)),
                   'rmx_rollout_resume')
        np.testing.assert_array_equal(q2, ref['q'])
        it3 = np.zeros((B, 2), dtype=np.int32)
        _ffi.check(_ffi.lib().rmx_rollout_resume(sg._handle, C.byref(o), B, _ffi.ptr(ke), None, _ffi.ptr(q0), _ffi.ptr(qd0), None,
                                                 _ffi.ptr(q2), _ffi.ptr(qd2), _ffi.ptr(st), _ffi.ptr(it3)), 'rmx_rollout_resume')
        np.testing.assert_array_equal(it2 + it3, it)


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
