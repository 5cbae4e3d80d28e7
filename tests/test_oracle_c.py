"""CPU: the compiled C twin of the oracle (oracle/redmax_oracle_c.c, the bench's CPU baseline) against the NumPy oracle,
which is itself pinned on the reference's golden energies (tests/test_oracle_pins.py)."""
import numpy as np
import pytest

from conftest import rel_err


@pytest.fixture(scope='module')
def oc():
    import __graft_entry__ as ge
    ge.build_oracle()
    import oracle_c
    assert oracle_c.available()
    return oracle_c


CASES = [('scene0', 'scenes', (0,), {}), ('scene1', 'scenes', (1,), {}), ('scene2', 'scenes', (2,), {}),
         ('scene14', 'scenes', (14,), {}), ('hand', 'hand_scene', (), {}),
         ('chain6ground', 'chain_scene', (6,), dict(ground=True, h=5e-4))]


@pytest.mark.parametrize('name,factory,a,kw', CASES, ids=[c[0] for c in CASES])
def test_c_eval_matches_numpy_oracle(oracle, oc, name, factory, a, kw):
    s = getattr(oracle, factory)(*a, **kw)
    if name == 'chain6ground':
        for f in s.forces:
            f.E[2, 3] = -48.5
    s.init()
    rng = np.random.default_rng(11)
    nr, h = s.nr, s.h
    q = s.qInit + 0.3 * rng.uniform(-1, 1, nr)
    q0 = q - 0.02 * rng.uniform(-1, 1, nr)
    qdot0 = rng.uniform(-1, 1, nr)
    tau = 100 * rng.uniform(-1, 1, nr)
    if name == 'scene14':
        q[0], q[1] = -2.0, 0.4
    s.setQ0(q0, qdot0)
    for j in s.joints:
        j.tau = tau[j.idxR].copy()
    g, H, M, f, K, D, J = oracle.eval_bdf1(q, s, True, True)
    out = oc.eval_direct(s, q, (q - q0) / h, q - q0 - h * qdot0, h, h * h, tau=tau)
    for key, ref in (('g', g), ('H', H), ('M', M), ('D', D), ('K', K), ('f', f)):
        assert rel_err(out[key], ref) < 1e-12, (key, rel_err(out[key], ref))


@pytest.mark.parametrize('sid,scheme', [(0, 1), (0, 2), (2, 1), (2, 2), (14, 2)])
def test_c_rollout_matches_numpy_oracle(oracle, oc, sid, scheme):
    s = oracle.scenes(sid)
    s.init()
    stats = []
    qs, qds = oracle.run_forward(s, scheme, s.qInit.copy(), s.qdotInit.copy(), nsteps=40, stats=stats)
    q, qd, st = oc.run_forward_batch(s, scheme, s.qInit, s.qdotInit, nsteps=40, threads=2)
    assert rel_err(q[0], qs) < 1e-11 and rel_err(qd[0], qds) < 1e-9
    it = np.array(stats)
    assert st[0, 0] == it[:, 0].sum() and st[0, 1] == it[:, 1].sum() and st[0, 2] == 0


def test_c_rollout_ground_bdf2(oracle, oc):
    s = oracle.chain_scene(6, ground=True, h=5e-4)
    for f in s.forces:
        f.E[2, 3] = -48.5
    s.init()
    qs, _ = oracle.run_forward(s, 2, s.qInit.copy(), s.qdotInit.copy(), nsteps=40)
    q, _, st = oc.run_forward_batch(s, 2, s.qInit, s.qdotInit, nsteps=40)
    assert rel_err(q[0], qs) < 1e-10


def test_c_oracle_hits_reference_golden_energy(oracle, oc):
    """Scene 0, BDF1, full run: H(end) from the C twin's trajectory equals scenesRedMax.m:54 within the reference's 1e-2."""
    s = oracle.scenes(0)
    s.init()
    q, qd, _ = oc.run_forward_batch(s, 1, s.qInit, s.qdotInit)
    s.reset()
    s.update()
    T0, V0 = s.computeEnergies()
    s.setQ(q[0, -1], qd[0, -1])
    s.update()
    T, V = s.computeEnergies()
    assert abs(T + V - V0 - s.Hexpected[0]) <= 1e-2


@pytest.mark.parametrize('n,seed', [(12, 3), (21, 4), (36, 5)])
def test_c_oracle_on_random_trees(rb, oracle, oc, n, seed):
    """The checker of tests/test_gpu_trees.py (the C twin) against the NumPy oracle on the seeded random trees themselves:
    branching, general and sign-flipped axes, fixed joints inside the tree, damping.  One evaluation and a short rollout."""
    s = rb.tree_scene(n, seed=seed, api=oracle)
    s.init()
    rng = np.random.default_rng(100 + seed)
    nr, h = s.nr, s.h
    q = s.qInit + 0.4 * rng.uniform(-1, 1, nr)
    q0 = q - 0.003 * rng.uniform(-1, 1, nr)
    qdot0 = rng.uniform(-1, 1, nr)
    tau = 100 * rng.uniform(-1, 1, nr)
    s.setQ0(q0, qdot0)
    for j in s.joints:
        j.tau = tau[j.idxR].copy()
    g, H, M, f, K, D, J = oracle.eval_bdf1(q, s, True, True)
    out = oc.eval_direct(s, q, (q - q0) / h, q - q0 - h * qdot0, h, h * h, tau=tau)
    for key, ref in (('g', g), ('H', H), ('M', M), ('D', D), ('K', K), ('f', f)):
        assert rel_err(out[key], ref) < 1e-12, (key, rel_err(out[key], ref))
    if n <= 21:
        s2 = rb.tree_scene(n, seed=seed, api=oracle)
        s2.init()
        stats = []
        qs, qds = oracle.run_forward(s2, 2, s2.qInit.copy(), s2.qdotInit.copy(), nsteps=6, stats=stats)
        qc, qdc, st = oc.run_forward_batch(s2, 2, s2.qInit, s2.qdotInit, nsteps=6, threads=1)
        assert rel_err(qc[0], qs) < 1e-11 and rel_err(qdc[0], qds) < 1e-9
        assert st[0, 0] == np.array(stats)[:, 0].sum() and st[0, 2] == 0


def test_tree_scene_host_mirror_matches_oracle_numbering(rb, oracle):
    """Same seeded tree through the host mirror (what the library is given) and through the oracle's classes: same reduced
    numbering (leaf-to-root, Scene.m:69-71), same initial state."""
    for n, seed in ((12, 3), (40, 1), (73, 6)):
        a = rb.tree_scene(n, seed=seed)
        a.init()
        b = rb.tree_scene(n, seed=seed, api=oracle)
        b.init()
        assert a.nr == b.nr
        np.testing.assert_array_equal(a.qInit, b.qInit)
        for ja, jb in zip(a.joints, b.joints):
            assert list(np.atleast_1d(ja.idxR)) == list(np.atleast_1d(jb.idxR))
