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
