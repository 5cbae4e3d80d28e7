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
