"""Developer probe: is the straggler-heavy workload (C3 at h = 5e-4) bound by the serial time of its costliest rollout or by
throughput?  Time against batch size (the first B rollouts of the same seeded batch), and the costliest rollout run alone."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402


def timed(sc, q0, qd0, scheme, nsteps):
    B = q0.shape[0]
    dq0, dqd0 = torch.from_numpy(np.ascontiguousarray(q0)).cuda(), torch.from_numpy(np.ascontiguousarray(qd0)).cuda()
    qo = torch.empty((B, nsteps, sc.nr), dtype=torch.float64, device='cuda')
    qdo = torch.empty_like(qo)
    st = torch.empty(B, dtype=torch.int32, device='cuda')
    it = torch.empty((B, 2), dtype=torch.int32, device='cuda')
    stream = torch.cuda.current_stream()
    best = 1e30
    for r in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=scheme, nsteps=nsteps, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, it.cpu().numpy(), st.cpu().numpy()


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0), flush=True)
    sc = rb.chain_scene(32, ground=True, h=5e-4, nsteps=100)
    sc.init()
    q0, qd0 = rb.synthetic_inputs(sc, 8192, seed=20260003)
    ms, it, st = timed(sc, q0[:4096], qd0[:4096], 2, 100)
    worst = np.argsort(-it[:, 1])[:8]
    print('B=4096: %.1f ms; costliest rollouts by reported line-search count: %s' % (ms, [(int(b), int(it[b, 1])) for b in worst]))
    for B in (1024, 2048, 4096, 8192):
        ms, _, _ = timed(sc, q0[:B], qd0[:B], 2, 100)
        print('B=%5d: %8.1f ms  %.3f M rollout-steps/s' % (B, ms, B * 100 / ms / 1e3), flush=True)
    for b in worst[:4]:
        ms, itb, stb = timed(sc, q0[b:b + 1], qd0[b:b + 1], 2, 100)
        print('rollout %d alone: %8.1f ms (reported newton %d, line search %d, status %d)' % (b, ms, itb[0, 0], itb[0, 1], stb[0]))
    ok = np.nonzero(st == 0)[0][:1]
    ms, itb, stb = timed(sc, q0[ok], qd0[ok], 2, 100)
    print('a converging rollout alone: %8.1f ms (newton %d)' % (ms, itb[0, 0]))
    # where does the time of the full batch go?  converging rollouts only / stalled rollouts only, with and without lockstep groups
    ms, it, st = timed(sc, q0, qd0, 2, 100)
    good, bad = np.nonzero(st == 0)[0], np.nonzero(st != 0)[0]
    print('B=8192: %d rollouts with a stalled step' % len(bad))
    for G in ('', '1'):
        if G:
            os.environ['RMX_GROUP'] = G
        else:
            os.environ.pop('RMX_GROUP', None)
        for name, sel in (('all', np.arange(8192)), ('converging only', good), ('stalled only', bad)):
            ms, _, _ = timed(sc, q0[sel], qd0[sel], 2, 100)
            print('RMX_GROUP=%-2s %-16s B=%5d: %8.1f ms  (%.1f slot-seconds over 740 slots)' % (G or '-', name, len(sel), ms, ms * 740 / 1e3), flush=True)
