"""GPU parity at the edges of the path: BASELINE configuration C1 (the single pendulum of scenesRedMax.m scene -2, one rollout),
the smallest and largest sizes of each kernel family, one-step and one-rollout calls, ragged batches, per-step controls."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def oc():
    import oracle_c
    if not oracle_c.available():
        import __graft_entry__ as ge
        ge.build_oracle()
    return oracle_c


@pytest.mark.parametrize('scheme', [1, 2])
def test_c1_single_pendulum_one_rollout(rb, oracle, scheme):
    """BASELINE config C1: scene -2 verbatim (scenesRedMax.m:13-26), B = 1, the reference's defaults h = 1e-2, 100 steps."""
    sg = rb.scenesRedMax(-2)
    sg.init()
    so = rb.scenesRedMax(-2, api=oracle)
    so.init()
    assert sg.nr == 1 and sg.nsteps == so.nsteps
    out = sg.rollout(scheme=scheme)
    stats = []
    q, qd = oracle.run_forward(so, scheme, stats=stats)
    assert out['q'].shape == (1, sg.nsteps, 1)
    assert rel_err(out['q'][0], q) < 1e-10 and rel_err(out['qdot'][0], qd) < 1e-8
    it = np.array(stats)
    assert out['status'][0] == 0 and out['iters'][0, 0] == it[:, 0].sum() and out['iters'][0, 1] == it[:, 1].sum()


@pytest.mark.timeout(900)
@pytest.mark.parametrize('n', [1, 2, 31, 32, 33, 63, 64, 65, 128])
def test_eval_at_kernel_family_boundaries(rb, oracle, oc, n):
    """n = 32 | 33: one warp | two warps; 64 | 65: composite kernels | sweep kernels; 128: the largest scene the library takes."""
    sg = rb.chain_scene(n, h=2e-4)
    sg.init()
    so = rb.chain_scene(n, h=2e-4, api=oracle)
    so.init()
    rng = np.random.default_rng(500 + n)
    h = sg.h
    q = sg.qInit + 0.3 * rng.uniform(-1, 1, n)
    q0 = q - 0.002 * rng.uniform(-1, 1, n)
    qdot0 = rng.uniform(-1, 1, n)
    tau = 100 * rng.uniform(-1, 1, n)
    qd, dq = (q - q0) / h, q - q0 - h * qdot0
    ref = oc.eval_direct(so, q, qd, dq, h, h * h, tau=tau)
    out = sg.eval(q, qd, dq, h * h, 1.0 / h, tau=tau)
    for key, tol in (('g', 1e-11), ('H', 1e-11), ('M', 1e-11)):
        assert rel_err(out[key], ref[key]) < tol, (key, rel_err(out[key], ref[key]))
    # D of a single pendulum is exactly zero in the reference's formulas; the kernels form it with cancellation: absolute scale
    scale = max(np.linalg.norm(ref['D']), 1e-3 * np.linalg.norm(ref['M']))
    assert np.linalg.norm(out['D'] - ref['D']) < 1e-10 * scale


@pytest.mark.parametrize('n', [1, 2, 33, 65])
@pytest.mark.parametrize('scheme', [1, 2])
def test_short_rollouts_at_kernel_family_boundaries(rb, oracle, oc, n, scheme):
    sg = rb.chain_scene(n, h=2e-4)
    sg.init()
    so = rb.chain_scene(n, h=2e-4, api=oracle)
    so.init()
    B, ns = 3, 8
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=600 + n)
    out = sg.rollout(q0, qd0, scheme=scheme, nsteps=ns)
    q, qd, st = oc.run_forward_batch(so, scheme, q0, qd0, nsteps=ns, threads=B)
    assert (st[:, 2] == 0).all() and (out['status'] == 0).all()
    assert rel_err(out['q'], q) < 1e-10, rel_err(out['q'], q)
    assert rel_err(out['qdot'], qd) < 1e-8
    assert np.abs(out['iters'] - st[:, :2]).max() <= 1, (out['iters'], st[:, :2])


def test_one_step_one_rollout_and_ragged_batches(rb):
    """nsteps = 1, B = 1, and batch sizes around the co-resident block count: every rollout equals the same rollout run alone."""
    sg = rb.chain_scene(12, h=1e-3)
    sg.init()
    q0, qd0 = rb.synthetic_inputs(sg, 1300, seed=77)
    full = sg.rollout(q0, qd0, nsteps=5)
    one = sg.rollout(q0[:1], qd0[:1], nsteps=1)
    np.testing.assert_array_equal(one['q'][0, 0], full['q'][0, 0])
    for B in (1, 7, 1183, 1184, 1185):
        part = sg.rollout(q0[:B], qd0[:B], nsteps=5)
        for key in ('q', 'qdot', 'status', 'iters'):
            np.testing.assert_array_equal(part[key], full[key][:B])


@pytest.mark.parametrize('n', [10, 40])
def test_per_step_controls_equal_step_by_step_calls(rb, n):
    """tau given per step (rmx_opts.tau_mode = per step): under BDF1 the state after a step is all the next step needs, so the
    rollout must equal, bit for bit, a chain of one-step calls each given that step's torques as a constant."""
    sg = rb.chain_scene(n, h=1e-3)
    sg.init()
    B, ns = 5, 6
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=88)
    tau = 1e3 * np.random.default_rng(n).uniform(-1, 1, (B, ns, sg.nr))
    out = sg.rollout(q0, qd0, tau=tau, scheme=1, nsteps=ns)
    q, qd = q0, qd0
    for k in range(ns):
        step = sg.rollout(q, qd, tau=np.ascontiguousarray(tau[:, k]), scheme=1, nsteps=1)
        np.testing.assert_array_equal(step['q'][:, 0], out['q'][:, k])
        np.testing.assert_array_equal(step['qdot'][:, 0], out['qdot'][:, k])
        q, qd = step['q'][:, 0].copy(), step['qdot'][:, 0].copy()
    assert np.abs(out['q'] - sg.rollout(q0, qd0, scheme=1, nsteps=ns)['q']).max() > 1e-6   # the torques do something


def test_bad_calls_fail_with_codes_not_crashes(rb):
    from redmax_b200 import _ffi
    sg = rb.chain_scene(4)
    sg.init()
    q0, qd0 = rb.synthetic_inputs(sg, 2, seed=1)
    with pytest.raises(_ffi.RmxError):
        sg.rollout(q0, qd0, nsteps=0)
    with pytest.raises(_ffi.RmxError):
        sg.rollout(q0, qd0, scheme=3)
    big = rb.chain_scene(129)
    with pytest.raises(_ffi.RmxError):
        big.init()
        big.rollout(nsteps=1)
