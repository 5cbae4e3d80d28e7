"""GPU: world frames of the bodies (rmx_body_frames) against the oracle's Joint.update / Body.update, and the trajectory export
built on them."""
import json

import numpy as np
import pytest

from test_gpu_parity import both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('sid', [0, 2, 4, 8, 13])
def test_body_frames_match_oracle(rb, oracle, sid):
    sg, so = both(rb, oracle, rb.scenesRedMax, sid)
    rng = np.random.default_rng(sid)
    q = sg.qInit[None, :] + 0.5 * rng.uniform(-1, 1, (5, sg.nr))
    E = sg.body_frames(q)
    assert E.shape == (5, len(sg.bodies), 4, 4)
    for b in range(5):
        so.setQ(q[b], np.zeros(sg.nr))
        so.update()
        for i, body in enumerate(so.bodies):
            np.testing.assert_allclose(E[b, i], body.E_wi, rtol=0, atol=1e-12 * max(1.0, np.abs(body.E_wi).max()))


def test_body_frames_across_a_chart_switch(rb, oracle, tmp_path):
    """Scene 7 under BDF2: q(t) changes coordinates at the two re-parameterised steps, the bodies do not -- frames evaluated
    in each step's own charts are continuous across the switches and equal the oracle's at the switch steps and at the end;
    the exporter takes the same chart history."""
    from redmax_b200 import export
    sg, so = both(rb, oracle, rb.scenesRedMax, 7)
    out = sg.rollout(scheme=2)
    ch = sg.chart_history(out, 0)
    ks = [k for k, _, _, _ in out['chart_switches'][0]]
    assert len(ks) == 2 and ch[0].tolist() == [7, 7] and ch[-1].tolist() == out['chart'][0].tolist() == [7, 10]
    E = sg.body_frames(out['q'][0], chart=ch)
    step = np.abs(np.diff(E, axis=0)).max(axis=(1, 2, 3))
    for k in ks:
        assert np.abs(out['q'][0, k] - out['q'][0, k - 1]).max() > 0.5       # the coordinates jump ...
        assert step[k - 1] < 2.0 * max(step[k - 2], step[k])                # ... the frames do not
    qs, _ = oracle.run_forward(so, 2, sg.qInit, sg.qdotInit)                # leaves `so` at the end of the run
    for i, body in enumerate(so.bodies):
        np.testing.assert_allclose(E[-1, i], body.E_wi, rtol=0, atol=1e-8)
    d = export.brender_scene(sg, out['q'][0], every=50, chart=ch)
    np.testing.assert_allclose(d['body'][-1]['body1']['location'], E[450, 1, :3, 3], atol=1e-12)


def test_export_layout_and_content(rb, oracle, tmp_path):
    from redmax_b200 import export
    sg, so = both(rb, oracle, rb.scenesRedMax, 1)
    out = sg.rollout(scheme=1, nsteps=6)
    p = export.export_brender(sg, out['q'][0], str(tmp_path / 'traj.json'), every=2)
    d = json.load(open(p))
    assert [s['name'] for s in d['header']['states']] == ['body0', 'body1', 'body2'] and len(d['body']) == 3
    rec = d['body'][-1]
    assert rec['frame'] == 4
    so.setQ(out['q'][0, 4], np.zeros(sg.nr))
    so.update()
    for i, body in enumerate(so.bodies):
        r = rec['body%d' % i]
        np.testing.assert_allclose(r['location'], body.E_wi[:3, 3], atol=1e-10)
        w, x, y, z = r['quat']
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        np.testing.assert_allclose(R, body.E_wi[:3, :3], atol=1e-10)
        assert r['scale'] == [10.0, 1.0, 1.0]
