"""Generates tests/golden/*.npz from the NumPy oracle (oracle/redmax_oracle.py).

The reference is MATLAB and cannot run in this image (no MATLAB/Octave), so these fixtures are outputs of the oracle,
which is pinned on the reference's golden energies Hexpected (tests/test_oracle_pins.py).  Each file holds the inputs and
the oracle's outputs; tests/test_golden.py checks the oracle still reproduces them (CPU) and the CUDA path matches them
(GPU).  Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import redmax_oracle as oracle  # noqa: E402
import redmax_b200.scenes as scenes  # noqa: E402  (scene factories only; no compute)

FORWARD = [
    # name, factory, args, kwargs, nsteps (None = the scene's own)
    ('scene0', 'scenesRedMax', (0,), {}, None),
    ('scene1', 'scenesRedMax', (1,), {}, None),
    ('scene2', 'scenesRedMax', (2,), {}, None),
    ('scene14', 'scenesRedMax', (14,), {}, None),
    # Euler-chart joints; scene 7 under BDF2 re-parameterises twice (XYZ -> XYX -> YXZ)
    ('scene7', 'scenesRedMax', (7,), {}, None),
    ('scene9', 'scenesRedMax', (9,), {}, None),
    ('chain10', 'chain_scene', (10,), dict(h=1e-3), 30),
    ('chain6ground', 'chain_scene', (6,), dict(ground=True, h=5e-4, ground_z=-48.5), 80),
]


def forward(only=None):
    for name, fac, a, kw, ns in FORWARD:
        if only and name not in only:
            continue
        for scheme in (1, 2):
            so = getattr(scenes, fac)(*a, api=oracle, **kw)
            so.init()
            ns_ = so.nsteps if ns is None else ns
            stats = []
            qs, qds = oracle.run_forward(so, scheme, so.qInit.copy(), so.qdotInit.copy(), nsteps=ns_, stats=stats)
            T0, V0 = so.T0, so.V0
            so.setQ(qs[-1], qds[-1])
            so.update()
            T, V = so.computeEnergies()
            np.savez_compressed(os.path.join(HERE, 'fwd_%s_bdf%d.npz' % (name, scheme)), q=qs, qdot=qds,
                                iters=np.array(stats), H_end=T + V - V0, Hexpected=so.Hexpected[scheme - 1], nsteps=ns_,
                                switch_steps=np.array(so.chart_switch_steps, dtype=np.int32),
                                charts=np.array([j.chart for j in so.joints if hasattr(j, 'switches')], dtype=np.int32))
            print(name, scheme, 'H_end', T + V - V0, 'Hexpected', so.Hexpected[scheme - 1])


def adjoint():
    for sid, scheme in ((100, 1), (101, 2)):
        so = scenes.scenesRedMax(sid, api=oracle)
        so.init()
        rng = np.random.default_rng(sid)
        p = np.vstack([np.zeros(so.nr), 0.02 * rng.uniform(-1, 1, (2, so.nr))])
        P, G = [], []
        for b in range(3):
            Pb, Gb = oracle.task_objective(p[b], so, scheme)
            P.append(Pb)
            G.append(Gb)
        np.savez_compressed(os.path.join(HERE, 'adj_scene%d.npz' % sid), p=p, P=np.array(P), dPdp=np.array(G))
        print(sid, P)
    for scheme in (1, 2):
        ns = 12
        so = scenes.hand_scene(nsteps=ns, scheme=scheme, api=oracle)
        so.init()
        B = 3
        rng = np.random.default_rng(20260004)
        p = 0.01 * rng.uniform(-1, 1, (B, so.nr))
        q0, qd0 = scenes.synthetic_inputs(so, B, seed=20260004)
        xt = np.array(so.task.xtarget)[None, :] + rng.uniform(-2, 2, (B, 3))
        P, G = [], []
        for b in range(B):
            so.qInit, so.qdotInit = q0[b].copy(), qd0[b].copy()
            so.task.setTarget(xt[b])
            Pb, Gb = oracle.task_objective(p[b], so, scheme)
            P.append(Pb)
            G.append(Gb)
        np.savez_compressed(os.path.join(HERE, 'adj_hand_bdf%d.npz' % scheme), p=p, q0=q0, qdot0=qd0, xtarget=xt,
                            P=np.array(P), dPdp=np.array(G), nsteps=ns)
        print('hand', scheme, P)


if __name__ == '__main__':
    if len(sys.argv) > 1:  # python make_golden.py scene7 scene9: only these forward fixtures
        forward(set(sys.argv[1:]))
    else:
        forward()
        adjoint()
