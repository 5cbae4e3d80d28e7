"""Developer probe (round 2): what the bit-exact shortcuts of stalled Newton solves buy on the straggler-bound workloads
(C3 at h = 5e-4, C5), timed with and without them (RMX_NO_SHORTCUTS=1), results compared bit for bit."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402
from explore_r2 import run, cost_report  # noqa: E402


def ab(label, sc, B, scheme, nsteps, slots, **kw):
    res = {}
    for mode in ('shortcuts', 'long way'):
        if mode == 'long way':
            os.environ['RMX_NO_SHORTCUTS'] = '1'
        else:
            os.environ.pop('RMX_NO_SHORTCUTS', None)
        q0, qd0 = rb.synthetic_inputs(sc, B, seed=20260003)
        dq0, dqd0 = torch.from_numpy(q0).cuda(), torch.from_numpy(qd0).cuda()
        qo = torch.empty((B, nsteps, sc.nr), dtype=torch.float64, device='cuda')
        qdo = torch.empty_like(qo)
        st = torch.empty(B, dtype=torch.int32, device='cuda')
        it = torch.empty((B, 2), dtype=torch.int32, device='cuda')
        stream = torch.cuda.current_stream()
        best = 1e30
        for r in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=scheme, nsteps=nsteps, stream=stream, **kw)
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[mode] = (best, qo.cpu().numpy(), qdo.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy())
        cost_report('%s [%s]' % (label, mode), best, res[mode][4], res[mode][3], nsteps, slots)
    os.environ.pop('RMX_NO_SHORTCUTS', None)
    a, b = res['shortcuts'], res['long way']
    same = all(np.array_equal(a[i], b[i]) for i in range(1, 5))
    print('   bitwise identical q, qdot, status, iters: %s ; speed-up %.2fx' % (same, b[0] / a[0]), flush=True)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0), flush=True)
    sc = rb.chain_scene(32, ground=True, h=5e-4, nsteps=100)
    sc.init()
    ab('C3 chain32+ground BDF2 h=5e-4', sc, 4096, 2, 100, 148 * 5)
    for h in (5e-4, 2e-4, 1e-4):
        sc = rb.chain_scene(64, h=h, nsteps=100)
        sc.init()
        ab('C5 chain64 BDF1 h=%g' % h, sc, 2048, 1, 100, 148 * 3, iterMaxFactor=2)
    sc = rb.chain_scene(64, h=2e-4, nsteps=100)
    sc.init()
    ms, it, st = run(sc, 8192, 1, 100)
    cost_report('C5 chain64 BDF1 h=2e-4 full iterMax, B=8192', ms, it, st, 100, 148 * 3)
