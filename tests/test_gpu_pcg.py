"""GPU: the alternative Newton linear solve (RMX_LINSOLVE_PCG: BiCGStab + the projected block-Jacobi preconditioner of the
reference's c++/PCG solver) lands on the same root as the LU path (SURVEY.md hard part H2: "PCG path must be validated to
land on the same root, ||g|| < 1e-9"), through the C ABI."""
import numpy as np
import pytest

from conftest import rel_err
from redmax_b200 import _ffi

pytestmark = pytest.mark.gpu

CASES = [
    ('scene0', lambda rb: rb.scenesRedMax(0), 1, 100),
    ('scene2-branching', lambda rb: rb.scenesRedMax(2), 2, 100),
    ('scene14-limits', lambda rb: rb.scenesRedMax(14), 1, 200),
    ('hand', lambda rb: rb.hand_scene(nsteps=30), 1, 30),
    ('chain32', lambda rb: rb.chain_scene(32, h=1e-3, nsteps=40), 1, 40),
    ('chain6ground', lambda rb: rb.chain_scene(6, ground=True, h=5e-4, ground_z=-48.5, nsteps=80), 2, 80),
    ('chain40-two-warps', lambda rb: rb.chain_scene(40, h=1e-3, nsteps=10), 1, 10),
]


@pytest.mark.parametrize('name,mk,scheme,ns', CASES, ids=[c[0] for c in CASES])
def test_pcg_lands_on_the_lu_root(rb, name, mk, scheme, ns):
    sg = mk(rb)
    sg.init()
    B = 4
    q0, qd0 = rb.synthetic_inputs(sg, B, seed=11)
    q0[0], qd0[0] = sg.qInit, sg.qdotInit
    lu = sg.rollout(q0, qd0, scheme=scheme, nsteps=ns)
    # tight linear tolerance: Newton sees (numerically) the direct-solve step, so trajectories agree to solver accuracy
    pcg = sg.rollout(q0, qd0, scheme=scheme, nsteps=ns, linsolve=_ffi.RMX_LINSOLVE_PCG, pcg_tol=1e-12)
    kry = sg.linsolve_stats()
    assert (lu['status'] == 0).all() and (pcg['status'] == 0).all()
    assert kry > 0
    err = rel_err(pcg['q'], lu['q'])
    print('%s: PCG vs LU rel err %.2e, Krylov iterations per Newton iteration %.2f' % (name, err, kry / pcg['iters'][:, 0].sum()))
    assert err < 1e-9, err
    # the reference's own tolerance (1e-6): inexact Newton steps, same converged states (||g|| < 1e-9 every step)
    pcg6 = sg.rollout(q0, qd0, scheme=scheme, nsteps=ns, linsolve=_ffi.RMX_LINSOLVE_PCG)
    assert (pcg6['status'] == 0).all()
    assert rel_err(pcg6['q'], lu['q']) < 1e-7


def test_pcg_rejected_for_adjoint(rb):
    sg = rb.scenesRedMax(100)
    sg.init()
    with pytest.raises(rb.RmxError):
        sg.rollout_adjoint(np.zeros((1, sg.nr)), linsolve=_ffi.RMX_LINSOLVE_PCG)


OPS = [
    ('chain10', lambda rb: rb.chain_scene(10), 1e-3),
    ('chain32', lambda rb: rb.chain_scene(32), 1e-3),
    ('scene2-branching', lambda rb: rb.scenesRedMax(2), 1e-2),
    ('hand', lambda rb: rb.hand_scene(), 1e-2),
    ('scene14-limits', lambda rb: rb.scenesRedMax(14), 5e-3),
    ('chain6ground', lambda rb: rb.chain_scene(6, ground=True, h=5e-4, ground_z=-48.5), 5e-4),
    ('chain48-two-warps', lambda rb: rb.chain_scene(48), 2e-4),
]


@pytest.mark.parametrize('name,mk,h', OPS, ids=[c[0] for c in OPS])
def test_krylov_operators_against_dense(rb, name, mk, h):
    """The two operators of the Krylov solve, directly (rmx_eval_krylov), against dense algebra on the same evaluation point:
    H x applied matrix-free by the two tree sweeps (the analogue of the reference's computeJ_x / computeLHS_x / computeJT_x,
    c++/PCG/src/ConstraintJoint.cpp:1090-1234) equals the assembled Newton matrix times x, and the projected block-Jacobi
    preconditioner (ConstraintJoint.cpp:1236, 1455; notes.pdf Alg. 10) equals the dense solve with J' blkdiag(M_j) J + Pr,
    Pr = cK (k + k_lim) + cK beta (d + d_lim) the joint-space diagonal."""
    sg = mk(rb)
    sg.init()
    nr = sg.nr
    rng = np.random.default_rng(17)
    for trial in range(2):
        q = sg.qInit + 0.3 * rng.uniform(-1, 1, nr)
        q0 = q - 0.02 * rng.uniform(-1, 1, nr)
        qdot0 = rng.uniform(-1, 1, nr)
        if name == 'scene14-limits':
            q[0], q[1] = -2.0, 0.4  # outside the joint limits: the limit stiffness enters Pr
        x = rng.uniform(-1, 1, nr)
        qd, dq = (q - q0) / h, q - q0 - h * qdot0
        cK, beta = h * h, 1.0 / h
        dense = sg.eval(q, qd, dq, cK, beta)
        ops = sg.eval_krylov(q, qd, dq, cK, beta, x)
        assert rel_err(ops['Hx'], dense['H'] @ x) < 1e-12, (name, rel_err(ops['Hx'], dense['H'] @ x))
        # joint-space diagonal: -cK (dfr/dq + beta dfr/dqdot) of Joint.computeForce (Joint.m:448-481)
        Pr = np.zeros(nr)
        for j in sg.joints:
            for k, r in enumerate(np.atleast_1d(j.idxR)):
                kk, dd = j.stiffness, j.damping
                if q[r] < j.qLimL or q[r] > j.qLimU:
                    kk, dd = kk + j.qLimK, dd + j.qLimD
                Pr[r] = cK * (kk + beta * dd)
        z = np.linalg.solve(dense['M'] + np.diag(Pr), x)
        assert rel_err(ops['Pinv_x'], z) < 1e-10, (name, rel_err(ops['Pinv_x'], z))
