"""CPU: the C-ABI library loads and exports every symbol include/redmax_b200.h declares; host-side logic of the
object-API mirror (DOF numbering, flattening) without any compute call."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'redmax_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(rmx_[a-z_0-9]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol(rb):
    from redmax_b200 import _ffi
    L = _ffi.lib()
    syms = header_symbols()
    assert sorted(_ffi.SYMBOLS) == syms
    for s in syms:
        assert hasattr(L, s), s
    assert L.rmx_version() == 109


def test_opts_default_are_the_references(rb):
    from redmax_b200 import _ffi
    o = _ffi.rmx_opts()
    _ffi.lib().rmx_opts_default(ctypes.byref(o), 1, 0)
    # driverRedMaxBDF1.m:95-98
    assert (o.tol, o.dxMax, o.iterMaxFactor, o.iterLsMax) == (1e-9, 1e3, 10, 20)
    _ffi.lib().rmx_opts_default(ctypes.byref(o), 2, 1)
    assert o.iterMaxFactor == 5  # driverRedMaxAdjointBDF1.m:108


def test_dof_numbering_is_leaf_to_root(rb, oracle):
    """Scene.m:69-71: the last joint owns q(1)."""
    for factory, a in ((rb.scenesRedMax, (0,)), (rb.scenesRedMax, (2,)), (rb.hand_scene, ())):
        sg = factory(*a)
        sg.init()
        so = factory(*a, api=oracle)
        so.init()
        assert sg.nr == so.nr and sg.nm == so.nm and sg.nsteps == so.nsteps
        for jg, jo in zip(sg.joints, so.joints):
            np.testing.assert_array_equal(jg.idxR, jo.idxR)
            np.testing.assert_array_equal(jg.body.idxM, jo.body.idxM)
            np.testing.assert_allclose(jg.body.I_i, jo.body.I_i, rtol=0, atol=0)
        np.testing.assert_array_equal(sg.qInit, so.qInit)
    sg = rb.scenesRedMax(1)
    sg.init()
    assert sg.joints[-1].idxR[0] == 0 and sg.joints[0].idxR[0] == sg.nr - 1


def test_scene_create_rejects_bad_input(rb):
    import pytest
    s = rb.Scene()
    b1, b2 = rb.BodyCuboid(1, [1, 1, 1]), rb.BodyCuboid(1, [1, 1, 1])
    j2 = rb.JointRevolute(None, b2, [0, 1, 0])
    j1 = rb.JointRevolute(j2, b1, [0, 1, 0])
    s.bodies = [b1, b2]
    s.joints = [j1, j2]  # child listed before parent
    with pytest.raises(ValueError):
        s.init()


def test_compute_fails_loudly_without_gpu(rb):
    """No CPU fallback: on a box without a CUDA device a compute call raises RmxError."""
    import pytest
    from redmax_b200 import _ffi
    if _ffi.lib().rmx_device_count() > 0:
        pytest.skip('GPU present')
    s = rb.scenesRedMax(-2)
    s.init()
    with pytest.raises(rb.RmxError):
        s.rollout()


def _two_link(rb):
    s = rb.Scene()
    b1, b2 = rb.BodyCuboid(1, [4, 1, 1]), rb.BodyCuboid(1, [4, 1, 1])
    j1 = rb.JointRevolute(None, b1, [0, 1, 0])
    j2 = rb.JointRevolute(j1, b2, [0, 1, 0])
    E = np.eye(4)
    E[0, 3] = 4.0
    j2.setJointTransform(E)
    s.bodies = [b1, b2]
    s.joints = [j1, j2]
    return s, b1, b2


def test_scene_create_validates_forces(rb):
    """rmx_scene_create (host only): error behaviour for the two-point forces and cables."""
    import pytest
    s, b1, b2 = _two_link(rb)
    s.forces = [rb.ForcePointPoint(b1, [0, 0, 0], b1, [1, 0, 0])]          # both points on one body
    with pytest.raises(rb.RmxError):
        s.init()
    s, b1, b2 = _two_link(rb)
    s.forces = [rb.ForcePointPoint(b1, [0, 0, 0], b2, [1, 0, 0]) for _ in range(9)]   # more than RMX_MAX_POINTFORCE
    with pytest.raises(rb.RmxError):
        s.init()
    s, b1, b2 = _two_link(rb)
    f = rb.ForceSpringDamper(b1, [2, 0, 0], b2, [-2, 0, 0])               # coincident points: zero rest length
    s.forces = [f]
    with pytest.raises(rb.RmxError):
        s.init()
    s, b1, b2 = _two_link(rb)
    c = rb.ForceCable()
    c.addBodyPoint(b1, [0, 0, 0])                                          # a cable needs at least two points
    s.forces = [c]
    with pytest.raises(rb.RmxError):
        s.init()
    s, b1, b2 = _two_link(rb)
    c = rb.ForceCable()
    c.addBodyPoint(None, [0, 0, 5])
    c.addBodyPoint(b1, [1, 0, 0])
    c.addBodyPoint(b2, [1, 0, 0])
    f = rb.ForceSpringDamper(None, [0, 0, -3], b2, [2, 0, 0])
    s.forces = [c, f, rb.ForcePointPoint(b1, [0, 0, 1], b2, [0, 0, 1])]
    s.init()                                                               # a valid mix is accepted
    assert s.nr == 2 and s.nm == 12


def test_scene_create_accepts_every_hot_path_joint_type(rb):
    s = rb.Scene()
    bs = [rb.BodyCuboid(1, [1, 1, 1]) for _ in range(9)]
    js = [rb.JointFixed(None, bs[0])]
    js.append(rb.JointRevolute(js[0], bs[1], [0, 0, 1]))
    js.append(rb.JointPrismatic(js[1], bs[2], [1, 0, 0]))
    js.append(rb.JointPlanar(js[2], bs[3]))
    js.append(rb.JointTranslational(js[3], bs[4]))
    js.append(rb.JointFree2D(js[4], bs[5]))
    js.append(rb.JointUniversal(js[5], bs[6]))
    js.append(rb.JointSpherical(js[6], bs[7]))
    js.append(rb.JointFree3D(js[7], bs[8]))
    s.bodies, s.joints = bs, js
    s.init()
    assert s.nr == 0 + 1 + 1 + 2 + 3 + 3 + 2 + 3 + 6 and s.nm == 54
    # leaf-to-root numbering with consecutive DOFs per joint (Scene.m:69-71, Joint.m:152)
    assert js[-1].idxR.tolist() == [0, 1, 2, 3, 4, 5] and js[-2].idxR.tolist() == [6, 7, 8] and js[1].idxR.tolist() == [20]
