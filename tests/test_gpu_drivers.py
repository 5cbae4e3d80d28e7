"""GPU: the reference-named drivers (redmax_b200/drivers.py) reproduce the reference's console verdicts: every in-scope
scene of `for sceneID = 0:14, driverRedMaxBDF1(sceneID,true)` (driverRedMaxBDF1.m:22-27) prints '### PASS ###'."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# every scene of the reference's loop except 11 (the reference's own BDF1 driver skips it: "doesn't work",
# driverRedMaxBDF1.m:16; its BDF2 pin is covered in test_gpu_joints.py) ...
IN_SCOPE = [(sid, itype) for sid in (0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 12, 13, 14) for itype in (1, 2)]
# ... and scene 7 (two spherical joints): under BDF2 the reference switches Euler charts twice in that scene, and so does
# the host side of the rollout (Scene._reparam_rollouts)
IN_SCOPE += [(7, 1), (7, 2)]


@pytest.mark.parametrize('sid,itype', IN_SCOPE)
def test_batch_mode_drivers_pass_the_energy_pins(rb, sid, itype):
    out = io.StringIO()
    drv = rb.driverRedMaxBDF1 if itype == 1 else rb.driverRedMaxBDF2
    res = drv(sid, True, out=out)
    text = out.getvalue()
    assert text.startswith("(%d) '%s': tEnd=" % (sid, res['scene'].name))
    assert res['pass'] is True and '### PASS ###' in text, text
    assert res['status'].tolist() == [0]


def test_scene7_bdf2_prints_the_chart_switches(rb):
    """driverRedMaxBDF2(7): the reference prints 'XYZ->XYX' and 'XYX->YXZ' (JointSpherical.m:84) on its way to PASS."""
    out = io.StringIO()
    res = rb.driverRedMaxBDF2(7, True, out=out)
    text = out.getvalue()
    assert 'XYZ->XYX' in text and 'XYX->YXZ' in text and '### PASS ###' in text, text
    assert res['status'].tolist() == [0] and res['chart'].tolist() == [[7, 10]]


def test_driver_with_a_batch_keeps_rollout_zero_on_the_pin(rb):
    sc = rb.scenesRedMax(2)
    sc.init()
    q0, qd0 = rb.synthetic_inputs(sc, 64, seed=5)
    q0[0], qd0[0] = sc.qInit, sc.qdotInit
    res = rb.driverRedMaxBDF2(2, True, q0=q0, qdot0=qd0, out=io.StringIO())
    assert res['pass'] is True and res['q'].shape == (64, sc.nsteps, sc.nr)
    assert abs(res['H'][0] - sc.Hexpected[1]) <= 1e-2 and np.isfinite(res['H']).all()


@pytest.mark.parametrize('drv', ['driverRedMaxAdjointBDF1', 'driverRedMaxAdjointBDF2'])
def test_adjoint_driver_descends(rb, drv):
    """fminunc's role (outside the parity boundary): the analytic gradient from the GPU adjoint drives a quasi-Newton
    descent of the task objective."""
    out = io.StringIO()
    res = getattr(rb, drv)(optimize=True, maxiter=12, out=out)
    h = res['history']
    assert h[-1] < 0.2 * h[0], (h[0], h[-1])
    P, g = rb.taskObjective(res['p'], res['scene'])
    assert abs(P - res['P']) <= 1e-9 * abs(P)
    assert 'p = [' in out.getvalue()
