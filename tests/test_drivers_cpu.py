"""CPU: host logic of the reference-named drivers (no compute call)."""
import io

import numpy as np


def test_driver_names_mirror_the_reference(rb):
    for name in ('driverRedMaxBDF1', 'driverRedMaxBDF2', 'driverRedMaxAdjointBDF1', 'driverRedMaxAdjointBDF2', 'taskObjective'):
        assert callable(getattr(rb, name))
    import inspect
    assert list(inspect.signature(rb.driverRedMaxBDF1).parameters)[:2] == ['sceneID', 'batch']


def test_plot_energies_verdict(rb):
    """Scene.m:171-177: |dH| > 1e-2 fails, no expected value -> no verdict."""
    sc = rb.scenesRedMax(0)
    out = io.StringIO()
    assert rb.plotEnergies(sc, 1, sc.Hexpected[0] + 5e-3, out) is True and '### PASS ###' in out.getvalue()
    assert rb.plotEnergies(sc, 1, sc.Hexpected[0] + 2e-2, out) is False
    sc2 = rb.scenesRedMax(-2)
    assert rb.plotEnergies(sc2, 1, 1.0, out) is None
    assert np.all(sc2.Hexpected == 0)


def test_quaternion_of_rotation_roundtrip():
    from redmax_b200 import export
    rng = np.random.default_rng(3)
    for _ in range(50):
        A = rng.standard_normal((3, 3))
        Q, _ = np.linalg.qr(A)
        if np.linalg.det(Q) < 0:
            Q[:, 0] = -Q[:, 0]
        w, x, y, z = export.rotation_to_quaternion(Q)
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        assert np.abs(R - Q).max() < 1e-12 and abs(w * w + x * x + y * y + z * z - 1) < 1e-12
