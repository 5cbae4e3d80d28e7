"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on identical scenes and seeds.

Bars: q(t) within 1e-10 relative (BASELINE.json north_star); single evaluations (g, H, M, D) within 1e-11."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL_Q = 1e-10    # north_star: q(t) within 1e-10 relative of the reference on identical scenes and seeds
TOL_EVAL = 1e-11


def both(rb, oracle, factory, *a, **kw):
    sg = factory(*a, **kw)
    sg.init()
    so = factory(*a, api=oracle, **kw)
    so.init()
    assert sg.nr == so.nr and sg.nm == so.nm
    np.testing.assert_array_equal(sg.qInit, so.qInit)
    return sg, so


def oracle_eval(oracle, so, q, qdot0, q0, tau=None):
    """BDF1 evaluation at q with history (q0, qdot0): returns g, H, M, D."""
    so.setQ0(q0, qdot0)
    for j in so.joints:
        j.tau = (np.zeros(j.ndof) if tau is None else np.asarray(tau)[j.idxR].copy())
    g, H, M, f, K, D, J = oracle.eval_bdf1(q, so, True, True)
    return g, H, M, D, f


CASES = [
    ('scene0', lambda rb: (rb.scenesRedMax, (0,), {})),
    ('scene1', lambda rb: (rb.scenesRedMax, (1,), {})),
    ('scene2', lambda rb: (rb.scenesRedMax, (2,), {})),
    ('scene14', lambda rb: (rb.scenesRedMax, (14,), {})),
    ('hand', lambda rb: (rb.hand_scene, (), {})),
    ('chain10', lambda rb: (rb.chain_scene, (10,), {})),
    ('chain32', lambda rb: (rb.chain_scene, (32,), {})),
    ('chain6ground', lambda rb: (rb.chain_scene, (6,), dict(ground=True, h=5e-4, ground_z=-48.5))),
    ('chain40', lambda rb: (rb.chain_scene, (40,), {})),
]


@pytest.mark.parametrize('name,mk', CASES, ids=[c[0] for c in CASES])
def test_eval_matches_oracle(rb, oracle, name, mk):
    factory, a, kw = mk(rb)
    sg, so = both(rb, oracle, factory, *a, **kw)
    rng = np.random.default_rng(42)
    nr = sg.nr
    h = sg.h
    for trial in range(2):
        q = sg.qInit + 0.3 * rng.uniform(-1, 1, nr)
        q0 = q - 0.02 * rng.uniform(-1, 1, nr)
        qdot0 = rng.uniform(-1, 1, nr)
        tau = 100 * rng.uniform(-1, 1, nr)
        if name == 'scene14':
            q[0] = -2.0   # below the lower limit: exercises Joint.computeForce's hitL branch
            q[1] = 0.4    # above the upper limit
        g, H, M, D, f = oracle_eval(oracle, so, q, qdot0, q0, tau)
        out = sg.eval(q, (q - q0) / h, q - q0 - h * qdot0, h * h, 1.0 / h, tau=tau)
        assert rel_err(out['g'], g) < TOL_EVAL, ('g', rel_err(out['g'], g))
        assert rel_err(out['H'], H) < TOL_EVAL, ('H', rel_err(out['H'], H))
        assert rel_err(out['M'], M) < TOL_EVAL, ('M', rel_err(out['M'], M))
        assert rel_err(out['D'], D) < 1e-10, ('D', rel_err(out['D'], D))
        assert rel_err(out['f'], f) < 1e-9, ('f', rel_err(out['f'], f))


def test_ground_contact_active(rb, oracle):
    """The ground case above really has corners in contact (otherwise it would test nothing)."""
    so = rb.chain_scene(6, ground=True, h=5e-4, ground_z=-48.5, api=oracle)
    so.init()
    so.update()
    fm = np.zeros(so.nm)
    for f in so.forces:
        f.computeValues_(None, fm)
    assert np.linalg.norm(fm) > 0


ROLL = [
    ('scene0', 1), ('scene0', 2), ('scene1', 1), ('scene1', 2), ('scene2', 1), ('scene2', 2),
    ('scene14', 1), ('scene14', 2), ('hand', 1), ('chain10', 1), ('chain10', 2),
]


@pytest.mark.parametrize('name,scheme', ROLL, ids=['%s-bdf%d' % r for r in ROLL])
def test_rollout_matches_oracle(rb, oracle, name, scheme):
    factory, a, kw = dict(CASES)[name](rb)
    sg, so = both(rb, oracle, factory, *a, **kw)
    ns = sg.nsteps if sg.nr <= 5 else 25
    # rollout 0: the scene's own initial state (the reference run); rollouts 1..: seeded perturbations
    q0, qd0 = rb.synthetic_inputs(sg, 3, seed=20260000 + len(name))
    q0[0], qd0[0] = sg.qInit, sg.qdotInit
    out = sg.rollout(q0, qd0, scheme=scheme, nsteps=ns)
    assert out['status'].tolist() == [0, 0, 0]
    for b in range(3 if sg.nr <= 5 else 2):
        stats = []
        qs, qds = oracle.run_forward(so, scheme, q0[b], qd0[b], nsteps=ns, stats=stats)
        assert rel_err(out['q'][b], qs) < TOL_Q, (b, rel_err(out['q'][b], qs))
        assert rel_err(out['qdot'][b], qds) < 1e-8, (b, rel_err(out['qdot'][b], qds))
        it = np.array(stats)
        assert out['iters'][b, 0] == it[:, 0].sum(), (out['iters'][b], it[:, 0].sum())
        assert out['iters'][b, 1] == it[:, 1].sum()


def test_rollout_chain32_bdf1_short(rb, oracle):
    """Headline shape (32-link chain, BDF1), a few steps against the dense oracle."""
    sg, so = both(rb, oracle, rb.chain_scene, 32)
    q0, qd0 = rb.synthetic_inputs(sg, 2, seed=20260003)
    ns = 6
    out = sg.rollout(q0, qd0, scheme=1, nsteps=ns)
    for b in range(2):
        qs, _ = oracle.run_forward(so, 1, q0[b], qd0[b], nsteps=ns)
        assert rel_err(out['q'][b], qs) < TOL_Q, rel_err(out['q'][b], qs)


def test_rollout_ground_bdf2(rb, oracle):
    """C3 pattern at a size the oracle finishes quickly: chain + ForceGroundCuboid per link (normal spring/damper,
    static and dynamic Coulomb friction), SDIRK2 start + BDF2.  The plane sits 1.4 mm inside the lowest corner
    of the rest pose, so the chain starts in light contact and keeps bouncing/sliding."""
    kw = dict(ground=True, h=5e-4, ground_z=-48.5)
    sg, so = both(rb, oracle, rb.chain_scene, 6, **kw)
    q0, qd0 = rb.synthetic_inputs(sg, 3, seed=20260002)
    q0[0], qd0[0] = sg.qInit, sg.qdotInit
    ns = 80
    out = sg.rollout(q0, qd0, scheme=2, nsteps=ns)
    for b in range(3):
        stats = []
        qs, _ = oracle.run_forward(so, 2, q0[b], qd0[b], nsteps=ns, stats=stats)
        it = np.array(stats)
        err = rel_err(out['q'][b], qs)
        print('ground rollout %d: rel err %.2e, newton %d vs %d, status %d vs %d'
              % (b, err, out['iters'][b, 0], it[:, 0].sum(), out['status'][b], np.bitwise_or.reduce(it[:, 2])))
        assert err < TOL_Q, err


def test_batch_independence_and_determinism(rb):
    """Size-independent property at full batch: every rollout depends only on its own inputs (bitwise), and a
    rerun is bitwise identical."""
    sg = rb.chain_scene(32, nsteps=10)
    sg.init()
    B = 1024
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=7)
    a = sg.rollout(q0, qd0, scheme=1)
    b = sg.rollout(q0, qd0, scheme=1)
    np.testing.assert_array_equal(a['q'], b['q'])
    perm = np.random.default_rng(0).permutation(B)
    c = sg.rollout(q0[perm], qd0[perm], scheme=1)
    np.testing.assert_array_equal(a['q'][perm], c['q'])
    assert (a['status'] == 0).all()


@pytest.fixture(scope='module')
def oc():
    """compiled twin of the oracle (oracle/redmax_oracle_c.c), prebuilt by __graft_entry__.build()"""
    import oracle_c
    if not oracle_c.available():
        import __graft_entry__ as ge
        ge.build_oracle()
    return oracle_c


@pytest.mark.parametrize('scheme', [1, 2])
def test_headline_workload_full_length_vs_c_oracle(rb, oracle, oc, scheme):
    """BASELINE.json's headline shape at full length: 32-link chain, 100 steps, h = 1e-3 (bench.py's workload), a
    sample of the seeded batch against the reference's dense algorithm (C twin of the oracle; the NumPy oracle needs
    minutes per rollout at this size)."""
    sg, so = both(rb, oracle, rb.chain_scene, 32, h=1e-3)
    B = 8
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=20260003)
    out = sg.rollout(q0, qd0, scheme=scheme)
    q, qd, st = oc.run_forward_batch(so, scheme, q0, qd0, threads=min(B, oc.max_threads()))
    assert (out['status'] == 0).all() and (st[:, 2] == 0).all()
    assert rel_err(out['q'], q) < TOL_Q, rel_err(out['q'], q)
    np.testing.assert_array_equal(out['iters'], st[:, :2])


def test_c3_ground_friction_bdf2_vs_c_oracle(rb, oracle, oc):
    """Config C3 pattern: 32-link chain with a ForceGroundCuboid on every link, SDIRK2 start + BDF2, contacts active."""
    kw = dict(ground=True, h=5e-4, ground_z=-40.0)
    sg, so = both(rb, oracle, rb.chain_scene, 32, **kw)
    B = 4
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=20260002)
    ns = 40
    out = sg.rollout(q0, qd0, scheme=2, nsteps=ns)
    q, qd, st = oc.run_forward_batch(so, 2, q0, qd0, nsteps=ns, threads=min(B, oc.max_threads()))
    free = rb.chain_scene(32, h=5e-4)
    free.init()
    assert rel_err(free.rollout(q0, qd0, scheme=2, nsteps=ns)['q'], out['q']) > 1e-6  # the ground really acts
    for b in range(B):
        err = rel_err(out['q'][b], q[b])
        print('C3 rollout %d: rel err %.2e newton %d vs %d status %d vs %d' % (b, err, out['iters'][b, 0], st[b, 0],
                                                                              out['status'][b], st[b, 2]))
    ok = st[:, 2] == 0  # compare where the reference's Newton converged in every step
    assert ok.any()
    assert rel_err(out['q'][ok], q[ok]) < TOL_Q


NEWTON_CASES = [c for c in CASES if c[0] != 'chain40'] + [
    ('chain40', dict(CASES)['chain40']),
    ('chain20ground', lambda rb: (rb.chain_scene, (20,), dict(ground=True, h=5e-4, ground_z=-48.5))),
    ('chain9', lambda rb: (rb.chain_scene, (9,), {})),
    ('chain17', lambda rb: (rb.chain_scene, (17,), {})),
    ('chain31', lambda rb: (rb.chain_scene, (31,), {})),
]


@pytest.mark.parametrize('name,mk', NEWTON_CASES, ids=[c[0] for c in NEWTON_CASES])
def test_newton_system_through_the_rollout_path(rb, oracle, name, mk):
    """rmx_eval_newton: H assembled and dx = -H\\g solved by the very code path the forward rollout kernel runs (FP64
    tensor-core tiles + blocked partial-pivot LU for one warp; the scalar path for two warps), against the oracle's dense
    H and LAPACK's solve.  Sizes cover partial 8x8 tiles and partial LU panels."""
    factory, a, kw = mk(rb)
    sg, so = both(rb, oracle, factory, *a, **kw)
    rng = np.random.default_rng(4242)
    nr, h = sg.nr, sg.h
    for trial in range(2):
        q = sg.qInit + 0.3 * rng.uniform(-1, 1, nr)
        q0 = q - 0.02 * rng.uniform(-1, 1, nr)
        qdot0 = rng.uniform(-1, 1, nr)
        tau = 100 * rng.uniform(-1, 1, nr)
        g, H, M, D, f = oracle_eval(oracle, so, q, qdot0, q0, tau)
        out = sg.eval_newton(q, (q - q0) / h, q - q0 - h * qdot0, h * h, 1.0 / h, tau=tau)
        assert rel_err(out['H'], H) < TOL_EVAL, ('H', rel_err(out['H'], H))
        dx = np.linalg.solve(H, -g)
        # backward error of the in-block LU, then the forward error on the scale the conditioning allows
        back = np.linalg.norm(H @ out['dx'] + g) / (np.linalg.norm(H, 2) * np.linalg.norm(out['dx']) + np.linalg.norm(g))
        assert back < 1e-14, back
        assert rel_err(out['dx'], dx) < 1e-12 * max(10.0, np.linalg.cond(H)), (rel_err(out['dx'], dx), np.linalg.cond(H))


@pytest.mark.timeout(300)
@pytest.mark.parametrize('scheme', [1, 2])
def test_page_locked_outputs_are_written_by_the_kernel_and_bitwise_identical(rb, scheme):
    """rmx_rollout with page-locked q_out / qdot_out: the kernel mirrors every step into the mapped host buffers (no
    device-to-host copy of the trajectories afterwards).  Must equal the staged copy into pageable buffers bit for bit, also
    for a batch the launch cuts across blocks, and with only q_out page-locked."""
    import torch
    sg = rb.chain_scene(8, nsteps=9, h=1e-3)
    sg.init()
    B = 3001
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=77)
    ref = sg.rollout(q0, qd0, scheme=scheme)  # pageable numpy buffers: staged copy
    hq = torch.full((B, sg.nsteps, sg.nr), float('nan'), dtype=torch.float64).pin_memory()
    hqd = torch.full((B, sg.nsteps, sg.nr), float('nan'), dtype=torch.float64).pin_memory()
    out = sg.rollout_into(q0, qd0, hq.numpy(), hqd.numpy(), scheme=scheme)
    np.testing.assert_array_equal(hq.numpy(), ref['q'])
    np.testing.assert_array_equal(hqd.numpy(), ref['qdot'])
    np.testing.assert_array_equal(out['status'], ref['status'])
    np.testing.assert_array_equal(out['iters'], ref['iters'])
    hq.fill_(float('nan'))
    qd_pageable = np.full((B, sg.nsteps, sg.nr), np.nan)
    sg.rollout_into(q0, qd0, hq.numpy(), qd_pageable, scheme=scheme)
    np.testing.assert_array_equal(hq.numpy(), ref['q'])
    np.testing.assert_array_equal(qd_pageable, ref['qdot'])


@pytest.mark.timeout(300)
def test_batch_sharded_over_gpus_is_bitwise_identical(rb):
    """rmx_rollout with opts.ngpus = G shards the batch contiguously over G devices from one process, no communication
    (SURVEY.md 8(e)): every trajectory, status and iteration count equals the one-GPU call bit for bit -- with pageable
    buffers (staged copies per device) and with page-locked ones (each device's kernel writes its slice of the caller's
    buffers).  Needs at least two GPUs; the one-process-per-GPU path is what bench.py --gpus N runs."""
    import torch
    G = _ffi_device_count(rb)
    if G < 2:
        pytest.skip('needs >= 2 GPUs (run under gpurun --gpus 2)')
    G = min(G, 4)
    sg = rb.chain_scene(8, nsteps=9, h=1e-3)
    sg.init()
    B = 2 * 1184 + 37  # uneven shards, each above the resident blocks of a device
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=11)
    tau = 30.0 * np.random.default_rng(6).uniform(-1, 1, (B, sg.nsteps, sg.nr))
    for scheme in (1, 2):
        ref = sg.rollout(q0, qd0, tau=tau, scheme=scheme, ngpus=1)
        out = sg.rollout(q0, qd0, tau=tau, scheme=scheme, ngpus=G)
        for k in ('q', 'qdot', 'status', 'iters'):
            np.testing.assert_array_equal(out[k], ref[k])
        hq = torch.full((B, sg.nsteps, sg.nr), float('nan'), dtype=torch.float64).pin_memory()
        hqd = torch.full((B, sg.nsteps, sg.nr), float('nan'), dtype=torch.float64).pin_memory()
        res = sg.rollout_into(q0, qd0, hq.numpy(), hqd.numpy(), tau=tau, scheme=scheme, ngpus=G)
        np.testing.assert_array_equal(hq.numpy(), ref['q'])
        np.testing.assert_array_equal(hqd.numpy(), ref['qdot'])
        np.testing.assert_array_equal(res['iters'], ref['iters'])


def _ffi_device_count(rb):
    from redmax_b200 import _ffi
    return int(_ffi.lib().rmx_device_count())


@pytest.mark.timeout(600)
@pytest.mark.parametrize('n,B,chunk', [(40, 1301, 500), (70, 1001, 250)])
def test_load_balanced_schedule_multi_warp_kernels(rb, n, B, chunk):
    """The same for the kernels with several warps per rollout (two-warp tensor-core kernel, four-warp sweep kernel), where the
    block's first warp claims the segments and broadcasts them (claim_segment)."""
    sg = rb.chain_scene(n, nsteps=5, h=2e-4)
    sg.init()
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=98)
    out = sg.rollout(q0, qd0, scheme=1)
    for lo in range(0, B, chunk):  # `chunk` rollouts fit the resident blocks: plain launches
        ref = sg.rollout(q0[lo:lo + chunk], qd0[lo:lo + chunk], scheme=1)
        for key in ('q', 'qdot', 'iters', 'status'):
            np.testing.assert_array_equal(out[key][lo:lo + chunk], ref[key])


@pytest.mark.timeout(300)
@pytest.mark.parametrize('scheme', [1, 2])
def test_load_balanced_schedule_is_bitwise_identical(rb, scheme):
    """More rollouts than co-resident blocks, and not a multiple of them: the launch cuts rollouts across blocks (McNaughton
    wrap-around schedule, second part resumes from the trajectory in global memory).  Every trajectory, status and iteration
    count must equal the one-block-per-rollout launches bit for bit."""
    sg = rb.chain_scene(8, nsteps=9, h=1e-3)
    sg.init()
    B = 3001
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=99)
    # per-step joint torques: the second part of a cut rollout must pick up its controls at the right step
    tau = 50.0 * np.random.default_rng(5).uniform(-1, 1, (B, sg.nsteps, sg.nr))
    out = sg.rollout(q0, qd0, tau=tau, scheme=scheme)
    assert (out['status'] == 0).all()
    for lo in range(0, B, 500):  # 500 rollouts fit the resident blocks: plain launches
        ref = sg.rollout(q0[lo:lo + 500], qd0[lo:lo + 500], tau=tau[lo:lo + 500], scheme=scheme)
        np.testing.assert_array_equal(out['q'][lo:lo + 500], ref['q'])
        np.testing.assert_array_equal(out['qdot'][lo:lo + 500], ref['qdot'])
        np.testing.assert_array_equal(out['iters'][lo:lo + 500], ref['iters'])
        np.testing.assert_array_equal(out['status'][lo:lo + 500], ref['status'])


def test_one_process_multi_gpu_device_rollout_with_nccl_gather(rb):
    """rmx_rollout_multi_dev: one process, G devices, device-resident shards, no communication during the
    rollouts and ONE in-place NCCL all-gather of the trajectory shards after them.  Every device must end up with the whole
    batch, bitwise equal to the one-GPU run.  Needs at least two GPUs (gpurun --gpus 2)."""
    import torch
    G = min(_ffi_device_count(rb), 4)
    if G < 2:
        pytest.skip('needs >= 2 GPUs (run under gpurun --gpus 2)')
    sg = rb.chain_scene(12, nsteps=15, h=1e-3)
    sg.init()
    B = 96 * G
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=4)
    ref = sg.rollout(q0, qd0, scheme=2)
    nr, ns = sg.nr, sg.nsteps
    n = B // G
    dq0 = [torch.from_numpy(q0[g * n:(g + 1) * n].copy()).to('cuda:%d' % g) for g in range(G)]
    dqd0 = [torch.from_numpy(qd0[g * n:(g + 1) * n].copy()).to('cuda:%d' % g) for g in range(G)]
    qo = [torch.full((B, ns, nr), float('nan'), dtype=torch.float64, device='cuda:%d' % g) for g in range(G)]
    qdo = [torch.full((B, ns, nr), float('nan'), dtype=torch.float64, device='cuda:%d' % g) for g in range(G)]
    st = [torch.empty(n, dtype=torch.int32, device='cuda:%d' % g) for g in range(G)]
    it = [torch.empty((n, 2), dtype=torch.int32, device='cuda:%d' % g) for g in range(G)]
    for g in range(G):
        torch.cuda.synchronize(g)
    sg.rollout_multi_dev(dq0, dqd0, qo, qdo, st, it, scheme=2, gather=True)
    for g in range(G):
        np.testing.assert_array_equal(qo[g].cpu().numpy(), ref['q'])
        np.testing.assert_array_equal(qdo[g].cpu().numpy(), ref['qdot'])
        np.testing.assert_array_equal(st[g].cpu().numpy(), ref['status'][g * n:(g + 1) * n])
        np.testing.assert_array_equal(it[g].cpu().numpy(), ref['iters'][g * n:(g + 1) * n])
    # gather = False: only the device's own slice is written
    qo2 = [torch.full((B, ns, nr), float('nan'), dtype=torch.float64, device='cuda:%d' % g) for g in range(G)]
    for g in range(G):
        torch.cuda.synchronize(g)
    sg.rollout_multi_dev(dq0, dqd0, qo2, None, st, None, scheme=2, gather=False)
    for g in range(G):
        a = qo2[g].cpu().numpy()
        np.testing.assert_array_equal(a[g * n:(g + 1) * n], ref['q'][g * n:(g + 1) * n])
        assert np.isnan(np.delete(a, np.s_[g * n:(g + 1) * n], axis=0)).all()
