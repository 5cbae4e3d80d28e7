"""CPU: the host side of the Euler-chart switching (redmax_b200/euler.py, what the reference does in jroot.reparam() between
steps) against the oracle's restatement of JointSpherical.reparam_ / getEuler / getEulerInv (JointSpherical.m:63-103,154-214)."""
import numpy as np
import pytest

from redmax_b200 import euler


class _B:
    joint = None


@pytest.mark.parametrize('chart', range(1, 13))
def test_chart_maps_match_oracle(oracle, chart):
    rng = np.random.default_rng(40 + chart)
    for _ in range(5):
        q = rng.uniform(-1.3, 1.3, 3)
        if chart <= 6:
            q[1] = abs(q[1]) + 0.1
        R, T = euler.chart_R_T(chart, q)
        o = oracle.euler_chart(chart, q, np.zeros(3))
        np.testing.assert_allclose(R, o[0], atol=1e-15)
        np.testing.assert_allclose(T, o[4], atol=1e-15)
        assert abs(euler.chart_det(chart, q[1]) - abs(o[5])) < 1e-15
        np.testing.assert_allclose(euler.chart_inv(chart, R), oracle.euler_chart_inv(chart, o[0]), atol=1e-13)


@pytest.mark.parametrize('chart', range(1, 13))
def test_reparam_matches_oracle(oracle, chart):
    """Same chart choice (including the exact ties between charts that share their middle angle, where the first wins) and
    the same re-expressed q, qdot, q1, qdot1 as JointSpherical.reparam_, from states just inside the |det T| <= 0.5 zone."""
    rng = np.random.default_rng(90 + chart)
    for trial in range(6):
        q = rng.uniform(-1.0, 1.0, 3)
        mid = rng.uniform(0.05, 0.5) * rng.choice([-1.0, 1.0])  # |det| = |sin| or |cos| of the middle angle <= 0.5
        q[1] = (np.arcsin(mid) % np.pi) if chart <= 6 else np.sign(mid) * np.arccos(abs(mid))
        qd = rng.uniform(-2, 2, 3)
        q1 = q + 0.01 * rng.uniform(-1, 1, 3)
        qd1 = qd + 0.1 * rng.uniform(-1, 1, 3)
        assert euler.chart_det(chart, q[1]) <= 0.5
        j = oracle.JointSpherical(None, _B())
        j.chart = j.chart1 = chart
        j.q, j.qdot, j.q1, j.qdot1 = q.copy(), qd.copy(), q1.copy(), qd1.copy()
        assert j.reparam_() is True
        new, qn, qdn, q1n, qd1n = euler.reparam(chart, q, qd, chart, q1, qd1)
        assert new == j.chart and new != chart
        np.testing.assert_allclose(qn, j.q, atol=1e-12)
        np.testing.assert_allclose(qdn, j.qdot, atol=1e-10)
        np.testing.assert_allclose(q1n, j.q1, atol=1e-12)
        np.testing.assert_allclose(qd1n, j.qdot1, atol=1e-10)
        # the motion is unchanged: same rotation, same body angular velocity
        R0, T0 = euler.chart_R_T(chart, q)
        R1, T1 = euler.chart_R_T(new, qn)
        np.testing.assert_allclose(R1, R0, atol=1e-12)
        np.testing.assert_allclose(T1 @ qdn, T0 @ qd, atol=1e-10)


def test_well_conditioned_state_is_left_alone(oracle):
    j = oracle.JointSpherical(None, _B())
    j.chart1 = j.chart
    j.q[:] = [0.3, 0.4, -0.2]
    assert j.reparam_() is False and j.chart == oracle.CHART_XYZ
    assert euler.chart_det(euler.CHART_XYZ, 0.4) > 0.5
    assert euler.CHART_NAMES[7] == 'XYZ' and euler.CHART_NAMES[1] == 'XYX' and euler.CHART_NAMES[10] == 'YXZ'


def test_chart_history_and_scene_variants(rb):
    """Host bookkeeping of a re-parameterised rollout (no GPU needed: scene handles are host objects until a rollout runs)."""
    sg = rb.scenesRedMax(7)
    sg.init()
    out = dict(q=np.zeros((2, 10, sg.nr)), chart_switches=[[(3, 1, 7, 1), (6, 1, 1, 10)], []])
    ch = sg.chart_history(out, 0)
    assert ch.shape == (10, 2) and ch[:, 0].tolist() == [7] * 10
    assert ch[:, 1].tolist() == [7, 7, 7, 1, 1, 1, 10, 10, 10, 10]
    assert (sg.chart_history(out, 1) == 7).all()
    # one library handle per chart combination, the scene's own charts map to the main handle, joints keep their charts
    h1 = sg._variant((7, 1))
    assert h1 is not None and sg._variant((7, 1)) is h1 and sg._variant((7, 7)) is sg._handle
    assert [j.chart for j in sg.joints] == [7, 7] and len(sg._variants) == 1
    with pytest.raises(rb.RmxError):
        sg._variant((7, 13))  # rmx_scene_create rejects charts outside 0..12
    assert [j.chart for j in sg.joints] == [7, 7]
    sg.close()
    assert sg._variants == {} and sg._handle is None
