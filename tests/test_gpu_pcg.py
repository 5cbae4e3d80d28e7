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
