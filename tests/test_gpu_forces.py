"""GPU parity for ForcePointPoint (first force with off-diagonal Km / Dm blocks): scene 10 ("Loop", a
1e7-stiff point-point spring closing a kinematic loop) and a variant with damping, a world-anchored force and a force between
two bodies of the same branch.  Same bars as test_gpu_parity.py."""
import numpy as np
import pytest

from conftest import rel_err
from test_gpu_parity import TOL_EVAL, TOL_Q, both, oracle_eval

pytestmark = pytest.mark.gpu


def loop_variant(api=None):
    import redmax_b200 as rb
    api = api or rb
    sc = rb.scenesRedMax(10, api=api)
    sc.forces[0].setStiffness(1e5)
    sc.forces[0].setDamping(3e3)
    f1 = api.ForcePointPoint(None, [3.0, 1.0, -12.0], sc.bodies[4], [0.5, 0.0, -4.0])
    f1.setStiffness(2e4)
    f1.setDamping(5e2)
    f2 = api.ForcePointPoint(sc.bodies[1], [0.2, 0.1, -3.0], sc.bodies[4], [0.0, 0.3, -1.0])
    f2.setStiffness(1e4)
    f2.setDamping(1e3)
    sc.forces += [f1, f2]
    sc.h = 2e-3
    sc.tEnd = 0.2
    return sc


CASES = [('scene10', lambda rb: (rb.scenesRedMax, (10,), {})), ('loop_variant', lambda rb: (loop_variant, (), {}))]


@pytest.mark.parametrize('name,mk', CASES, ids=[c[0] for c in CASES])
def test_eval_matches_oracle(rb, oracle, name, mk):
    factory, a, kw = mk(rb)
    sg, so = both(rb, oracle, factory, *a, **kw)
    rng = np.random.default_rng(77)
    nr, h = sg.nr, sg.h
    for trial in range(3):
        q = sg.qInit + 0.3 * rng.uniform(-1, 1, nr)
        q0 = q - 0.02 * rng.uniform(-1, 1, nr)
        qdot0 = rng.uniform(-1, 1, nr)
        tau = 100 * rng.uniform(-1, 1, nr)
        g, H, M, D, f = oracle_eval(oracle, so, q, qdot0, q0, tau)
        args = (q, (q - q0) / h, q - q0 - h * qdot0, h * h, 1.0 / h)
        out = sg.eval(*args, tau=tau)
        for nm_, ref, tol in (('g', g, TOL_EVAL), ('H', H, TOL_EVAL), ('M', M, TOL_EVAL), ('f', f, 1e-9)):
            assert rel_err(out[nm_], ref) < tol, (name, trial, nm_, rel_err(out[nm_], ref))
        dscale = max(np.max(np.abs(D)), np.max(np.abs(M)))
        assert np.max(np.abs(out['D'] - D)) < 1e-10 * dscale
        # the same system through the rollout kernel's own (tensor-core) assembly + LU
        nw = sg.eval_newton(*args, tau=tau)
        assert rel_err(nw['H'], H) < TOL_EVAL, rel_err(nw['H'], H)
        assert rel_err(nw['dx'], np.linalg.solve(H, -g)) < 1e-12 * max(10.0, np.linalg.cond(H))


@pytest.mark.parametrize('itype', [1, 2])
def test_scene10_golden_energy_and_trajectory(rb, oracle, itype):
    """The reference's recorded Hexpected of scene 10 through the CUDA path (|dH| <= 1e-2, Scene.m:172), and q(t) vs the oracle."""
    sg, so = both(rb, oracle, rb.scenesRedMax, 10)
    out = sg.rollout(scheme=itype)
    assert out['status'].tolist() == [0]
    T1, V1 = sg.energies(out['q'][0, -1], out['qdot'][0, -1])
    _, V0 = sg.energies(sg.qInit, sg.qdotInit)
    Hend = T1[0] + V1[0] - V0[0]
    assert abs(Hend - sg.Hexpected[itype - 1]) <= 1e-2, (Hend, sg.Hexpected[itype - 1])
    stats = []
    qs, _ = oracle.run_forward(so, itype, sg.qInit, sg.qdotInit, stats=stats)
    assert rel_err(out['q'][0], qs) < TOL_Q, rel_err(out['q'][0], qs)
    assert out['iters'][0, 0] == np.array(stats)[:, 0].sum()


def test_variant_rollouts_match_oracle(rb, oracle):
    sg, so = both(rb, oracle, loop_variant)
    q0, qd0 = rb.synthetic_inputs(sg, 3, seed=31)
    q0[0], qd0[0] = sg.qInit, sg.qdotInit
    for scheme in (1, 2):
        out = sg.rollout(q0, qd0, scheme=scheme)
        for b in range(3):
            qs, _ = oracle.run_forward(so, scheme, q0[b], qd0[b])
            assert rel_err(out['q'][b], qs) < TOL_Q, (scheme, b, rel_err(out['q'][b], qs))
    # energies incl. the spring potentials (ForcePointPoint.m:116-132)
    T, V = sg.energies(q0[1], qd0[1])
    so.setQ(q0[1], qd0[1])
    so.update()
    To, Vo = so.computeEnergies()
    assert abs(V[0] - Vo) <= 1e-10 * max(1.0, abs(Vo)) and abs(T[0] - To) <= 1e-10 * max(1.0, abs(To))


def spring_variant(api=None):
    """scene 12 plus a rest-length-set spring between the two bodies' far ends and a world spring on body 1."""
    import redmax_b200 as rb
    api = api or rb
    sc = rb.scenesRedMax(12, api=api)
    f3 = api.ForceSpringDamper(sc.bodies[0], [-4, 0.3, 0.2], sc.bodies[1], [4, -0.2, 0.1])
    f3.setStiffness(3e5)
    f3.setDamping(2e3)
    f3.setRetLength(17.0)
    f4 = api.ForceSpringDamper(sc.bodies[0], [1, 0, -0.5], None, [2.0, 1.0, -9.0])
    f4.setStiffness(2e5)
    f4.setDamping(5e2)
    sc.forces += [f3, f4]
    sc.joints[0].q[0] = 0.2
    sc.joints[1].q[0] = -0.3
    return sc


SPRING_CASES = [('scene12', lambda rb: (rb.scenesRedMax, (12,), {})), ('spring_variant', lambda rb: (spring_variant, (), {}))]


@pytest.mark.parametrize('name,mk', SPRING_CASES, ids=[c[0] for c in SPRING_CASES])
def test_spring_damper_eval_matches_oracle(rb, oracle, name, mk):
    """ForceSpringDamper (ForceSpringGeneric.m:35-143): g, H, M, D and the Newton step through both assembly paths."""
    factory, a, kw = mk(rb)
    sg, so = both(rb, oracle, factory, *a, **kw)
    rng = np.random.default_rng(78)
    nr, h = sg.nr, sg.h
    for trial in range(3):
        q = sg.qInit + 0.3 * rng.uniform(-1, 1, nr)
        q0 = q - 0.02 * rng.uniform(-1, 1, nr)
        qdot0 = rng.uniform(-1, 1, nr)
        g, H, M, D, f = oracle_eval(oracle, so, q, qdot0, q0, None)
        args = (q, (q - q0) / h, q - q0 - h * qdot0, h * h, 1.0 / h)
        out = sg.eval(*args)
        for nm_, ref, tol in (('g', g, TOL_EVAL), ('H', H, TOL_EVAL), ('M', M, TOL_EVAL), ('f', f, 1e-9)):
            assert rel_err(out[nm_], ref) < tol, (name, trial, nm_, rel_err(out[nm_], ref))
        assert np.max(np.abs(out['D'] - D)) < 1e-10 * max(np.max(np.abs(D)), np.max(np.abs(M)))
        nw = sg.eval_newton(*args)
        assert rel_err(nw['H'], H) < TOL_EVAL
        assert rel_err(nw['dx'], np.linalg.solve(H, -g)) < 1e-12 * max(10.0, np.linalg.cond(H))


@pytest.mark.parametrize('itype', [1, 2])
def test_scene12_golden_energy_and_trajectory(rb, oracle, itype):
    sg, so = both(rb, oracle, rb.scenesRedMax, 12)
    out = sg.rollout(scheme=itype)
    assert out['status'].tolist() == [0]
    T1, V1 = sg.energies(out['q'][0, -1], out['qdot'][0, -1])
    _, V0 = sg.energies(sg.qInit, sg.qdotInit)
    Hend = T1[0] + V1[0] - V0[0]
    assert abs(Hend - sg.Hexpected[itype - 1]) <= 1e-2, (Hend, sg.Hexpected[itype - 1])
    qs, _ = oracle.run_forward(so, itype, sg.qInit, sg.qdotInit)
    assert rel_err(out['q'][0], qs) < TOL_Q, rel_err(out['q'][0], qs)


def test_spring_variant_rollout_and_energy(rb, oracle):
    sg, so = both(rb, oracle, spring_variant)
    assert so.forces[0].L > 0 and so.forces[2].L == 17.0  # rest length from the initial configuration unless set
    q0, qd0 = rb.synthetic_inputs(sg, 2, seed=32)
    out = sg.rollout(q0, qd0, scheme=2, nsteps=60)
    for b in range(2):
        qs, _ = oracle.run_forward(so, 2, q0[b], qd0[b], nsteps=60)
        assert rel_err(out['q'][b], qs) < TOL_Q, (b, rel_err(out['q'][b], qs))
    T, V = sg.energies(q0[1], qd0[1])
    so.setQ(q0[1], qd0[1])
    so.update()
    To, Vo = so.computeEnergies()
    assert abs(V[0] - Vo) <= 1e-10 * max(1.0, abs(Vo)) and abs(T[0] - To) <= 1e-10 * max(1.0, abs(To))


def cable_variant(api=None):
    """scene 13 with a second, world-anchored two-point cable and a four-point cable with a set rest length."""
    import redmax_b200 as rb
    api = api or rb
    sc = rb.scenesRedMax(13, api=api)
    f2 = api.ForceCable()
    f2.setStiffness(4e5)
    f2.setDamping(2e3)
    f2.addBodyPoint(None, [12.0, 0.5, 6.0])
    f2.addBodyPoint(sc.bodies[2], [3.0, 0.0, 0.5])
    f3 = api.ForceCable()
    f3.setStiffness(2e5)
    f3.setDamping(1e3)
    f3.setRetLength(20.0)
    f3.addBodyPoint(sc.bodies[3], [0.2, 0.0, 0.3])
    f3.addBodyPoint(sc.bodies[1], [2.0, 0.0, -0.5])
    f3.addBodyPoint(sc.bodies[2], [1.0, 0.2, 0.5])
    f3.addBodyPoint(None, [-3.0, 1.0, -14.0])
    sc.forces += [f2, f3]
    return sc


CABLE_CASES = [('scene13', lambda rb: (rb.scenesRedMax, (13,), {})), ('cable_variant', lambda rb: (cable_variant, (), {}))]


@pytest.mark.parametrize('name,mk', CABLE_CASES, ids=[c[0] for c in CABLE_CASES])
def test_cable_eval_matches_oracle(rb, oracle, name, mk):
    """ForceCable (ForceSpringMultiPointGeneric.m:29-190): taut and slack states, g, H, M, D and the Newton step."""
    factory, a, kw = mk(rb)
    sg, so = both(rb, oracle, factory, *a, **kw)
    rng = np.random.default_rng(79)
    nr, h = sg.nr, sg.h
    taut = 0
    for trial in range(6):
        q = sg.qInit + 0.4 * rng.uniform(-1, 1, nr)
        q0 = q - 0.02 * rng.uniform(-1, 1, nr)
        qdot0 = rng.uniform(-1, 1, nr)
        g, H, M, D, f = oracle_eval(oracle, so, q, qdot0, q0, None)
        so.setQ(q, (q - q0) / h)
        so.update()
        pts = so.forces[0]._points()
        taut += so.forces[0]._length(pts)[0] > so.forces[0].L
        args = (q, (q - q0) / h, q - q0 - h * qdot0, h * h, 1.0 / h)
        out = sg.eval(*args)
        for nm_, ref, tol in (('g', g, TOL_EVAL), ('H', H, TOL_EVAL), ('M', M, TOL_EVAL), ('f', f, 1e-9)):
            assert rel_err(out[nm_], ref) < tol, (name, trial, nm_, rel_err(out[nm_], ref))
        assert np.max(np.abs(out['D'] - D)) < 1e-10 * max(np.max(np.abs(D)), np.max(np.abs(M)))
        nw = sg.eval_newton(*args)
        assert rel_err(nw['H'], H) < TOL_EVAL
    assert 0 < taut < 6 or name != 'scene13', taut  # both branches of ForceCable.computeSpringForce were exercised


@pytest.mark.parametrize('itype', [1, 2])
def test_scene13_golden_energy_and_trajectory(rb, oracle, itype):
    sg, so = both(rb, oracle, rb.scenesRedMax, 13)
    out = sg.rollout(scheme=itype)
    assert out['status'].tolist() == [0]
    T1, V1 = sg.energies(out['q'][0, -1], out['qdot'][0, -1])
    _, V0 = sg.energies(sg.qInit, sg.qdotInit)
    Hend = T1[0] + V1[0] - V0[0]
    assert abs(Hend - sg.Hexpected[itype - 1]) <= 1e-2, (Hend, sg.Hexpected[itype - 1])
    qs, _ = oracle.run_forward(so, itype, sg.qInit, sg.qdotInit)
    assert rel_err(out['q'][0], qs) < TOL_Q, rel_err(out['q'][0], qs)


def test_cable_variant_rollout_and_energy(rb, oracle):
    sg, so = both(rb, oracle, cable_variant)
    q0, qd0 = rb.synthetic_inputs(sg, 2, seed=33)
    out = sg.rollout(q0, qd0, scheme=1, nsteps=50)
    for b in range(2):
        qs, _ = oracle.run_forward(so, 1, q0[b], qd0[b], nsteps=50)
        assert rel_err(out['q'][b], qs) < TOL_Q, (b, rel_err(out['q'][b], qs))
    T, V = sg.energies(q0[1], qd0[1])
    so.setQ(q0[1], qd0[1])
    so.update()
    To, Vo = so.computeEnergies()
    assert abs(V[0] - Vo) <= 1e-10 * max(1.0, abs(Vo)) and abs(T[0] - To) <= 1e-10 * max(1.0, abs(To))
