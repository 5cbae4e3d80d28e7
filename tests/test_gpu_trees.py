"""GPU parity on seeded random joint trees (redmax_b200.scenes.tree_scene): branching, every se3.aaToMat axis case, general
axes, fixed joints in the middle of the tree, joint damping -- at the sizes of each kernel family (one warp with tensor-core
tiles: nr <= 32; two warps: nr <= 64; sweep kernels beyond).  The reference's own trees stop at 21 joints (scenesRedMax.m
scene 5 and the BASELINE hand); the two-warp and sweep kernels are otherwise exercised on chains only.  Checker: the C twin of
the oracle (bitwise the reference's dense algorithm; the NumPy oracle needs minutes at these sizes)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TREES = [(12, 3), (21, 4), (36, 5), (40, 1), (64, 2), (73, 6), (90, 7)]   # (links, seed): nr = 11, 19, 31, 35, 55, 63, 78


@pytest.fixture(scope='module')
def oc():
    import oracle_c
    if not oracle_c.available():
        import __graft_entry__ as ge
        ge.build_oracle()
    return oracle_c


def both(rb, oracle, n, seed, **kw):
    sg = rb.tree_scene(n, seed=seed, **kw)
    sg.init()
    so = rb.tree_scene(n, seed=seed, api=oracle, **kw)
    so.init()
    assert sg.nr == so.nr
    return sg, so


@pytest.mark.parametrize('n,seed', TREES)
def test_tree_eval_vs_c_oracle(rb, oracle, oc, n, seed):
    sg, so = both(rb, oracle, n, seed)
    rng = np.random.default_rng(100 + seed)
    h, nr = sg.h, sg.nr
    for trial in range(2):
        q = sg.qInit + 0.4 * rng.uniform(-1, 1, nr)
        q0 = q - 0.003 * rng.uniform(-1, 1, nr)
        qdot0 = rng.uniform(-1, 1, nr)
        tau = 100 * rng.uniform(-1, 1, nr)
        qd, dq = (q - q0) / h, q - q0 - h * qdot0
        ref = oc.eval_direct(so, q, qd, dq, h, h * h, tau=tau)
        out = sg.eval(q, qd, dq, h * h, 1.0 / h, tau=tau)
        for key, tol in (('g', 1e-11), ('H', 1e-11), ('M', 1e-11), ('D', 1e-10)):
            assert rel_err(out[key], ref[key]) < tol, (key, rel_err(out[key], ref[key]))
        if n <= 64:     # rmx_eval_newton is the composite kernels' entry (<= 64 joints)
            nw = sg.eval_newton(q, qd, dq, h * h, 1.0 / h, tau=tau)
            assert rel_err(nw['H'], ref['H']) < 1e-11
            dx = -np.linalg.solve(ref['H'], ref['g'])
            assert rel_err(nw['dx'], dx) < 1e-9, rel_err(nw['dx'], dx)


@pytest.mark.timeout(900)
@pytest.mark.parametrize('scheme', [1, 2])
@pytest.mark.parametrize('n,seed', TREES)
def test_tree_rollout_vs_c_oracle(rb, oracle, oc, n, seed, scheme):
    sg, so = both(rb, oracle, n, seed)
    B, ns = 6, 25
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=900 + seed)
    tau = 20.0 * np.random.default_rng(seed).uniform(-1, 1, (B, sg.nr))
    out = sg.rollout(q0, qd0, tau=tau, scheme=scheme, nsteps=ns)
    q, qd, st = oc.run_forward_batch(so, scheme, q0, qd0, tau=tau, nsteps=ns, threads=min(B, oc.max_threads()))
    assert (st[:, 2] == 0).all(), st
    assert (out['status'] == 0).all(), out['status']
    assert rel_err(out['q'], q) < 1e-10, rel_err(out['q'], q)
    assert rel_err(out['qdot'], qd) < 1e-8
    # Iteration counts: equal except where a step's second residual crosses the reference's absolute tolerance (1e-9,
    # driverRedMaxBDF1.m:95) within round-off -- e.g. tree (36, 5), BDF2, rollout 2: the oracle's residuals after the first
    # Newton iteration grow 8.6e-10, 9.5e-10, 9.64e-10, 1.00e-09, 1.04e-09 ... over steps 3..11, so the step at which a third
    # iteration first becomes necessary is decided by the last digits of ||g|| (GPU 68, oracle 69 iterations in total).
    assert np.abs(out['iters'] - st[:, :2]).max() <= max(1, 0.03 * st[:, 0].max()), (out['iters'], st[:, :2])


@pytest.mark.timeout(900)
@pytest.mark.parametrize('scheme', [1, 2])
@pytest.mark.parametrize('n,seed,ns', [(21, 4, 12), (36, 5, 10), (64, 2, 6)])
def test_tree_adjoint_vs_oracle(rb, oracle, n, seed, ns, scheme):
    """TaskBDF*PointPos objective and gradient through the adjoint kernels (one-warp tensor-core tape for nr <= 32, sweep tape
    beyond) on a branching tree with fixed joints, against the oracle's dense adjoint (NumPy: short horizons)."""
    sg, so = both(rb, oracle, n, seed, nsteps=ns, scheme=scheme)
    B = 3
    rng = np.random.default_rng(seed)
    p = 0.02 * rng.uniform(-1, 1, (B, sg.nr))
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=300 + seed)
    xt = np.array(sg.task.xtarget)[None, :] + rng.uniform(-2, 2, (B, 3))
    res = sg.rollout_adjoint(p, xtarget=xt, q0=q0, qdot0=qd0)
    assert (res['status'] == 0).all()
    for b in range(B):
        so.qInit, so.qdotInit = q0[b].copy(), qd0[b].copy()
        so.task.setTarget(xt[b])
        P, dPdp = oracle.task_objective(p[b], so, scheme)
        assert abs(res['P'][b] - P) <= 1e-10 * abs(P), (res['P'][b], P)
        assert rel_err(res['dPdp'][b], dPdp) < 1e-8, rel_err(res['dPdp'][b], dPdp)
        assert np.linalg.norm(dPdp - sg.task.wreg * p[b]) > 1e-3
