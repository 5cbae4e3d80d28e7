"""Developer probe (round 2): data behind the scheduling and step-size decisions.
  * C3 (chain32 + ground friction, BDF2): distribution of per-rollout cost (Newton iterations, line-search evaluations) --
    how much of the makespan is one straggler's serial chain, how much is imbalance a scheduler could remove
  * C5 (chain64): fraction of rollouts whose Newton converges in every step, by step size
  * cudaHostRegister cost of the headline output buffers"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402
from redmax_b200 import _ffi  # noqa: E402


def run(sc, B, scheme, nsteps, **kw):
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
    return best, it.cpu().numpy(), st.cpu().numpy()


def cost_report(label, ms, it, st, nsteps, slots):
    B = len(st)
    # cost model: one Newton iteration (assembly + LU) ~ 4 residual evaluations
    cost = 4.0 * it[:, 0] + it[:, 1]
    order = np.sort(cost)[::-1]
    mean_load = cost.sum() / slots
    print('%s: %.1f ms, %.3e rollout-steps/s; status!=0 %.2f%%; newton/step mean %.2f max %.1f; ls/step mean %.2f max %.1f'
          % (label, ms, B * nsteps / (ms * 1e-3), 100.0 * (st != 0).mean(), it[:, 0].mean() / nsteps, it[:, 0].max() / nsteps,
             it[:, 1].mean() / nsteps, it[:, 1].max() / nsteps))
    print('   cost units (4 x newton + ls): total %.3e, per-slot mean %.3e (%d slots), largest single rollout %.3e '
          '(= %.2f x the per-slot mean), top-5 %s, median %.0f'
          % (cost.sum(), mean_load, slots, order[0], order[0] / mean_load, order[:5].astype(int).tolist(), np.median(cost)))
    print('   lower bound on the makespan by any scheduler = max(per-slot mean, largest rollout) = %.3e units -> at the measured '
          'rate of converging rollouts that is %.1fx the balanced time' % (max(mean_load, order[0]), max(mean_load, order[0]) / mean_load))
    sys.stdout.flush()


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0), flush=True)
    for h in (5e-4, 2e-4):
        sc = rb.chain_scene(32, ground=True, h=h, nsteps=100)
        sc.init()
        ms, it, st = run(sc, 4096, 2, 100)
        cost_report('C3 chain32+ground BDF2 h=%g' % h, ms, it, st, 100, 148 * 5)
    for h in (1e-3, 5e-4, 2e-4, 1e-4):
        sc = rb.chain_scene(64, h=h, nsteps=20)
        sc.init()
        # iterMaxFactor = 1 bounds what a stalled rollout costs in this probe; the status bits still tell who stalls
        ms, it, st = run(sc, 2048, 1, 20, iterMaxFactor=1)
        cost_report('C5 chain64 BDF1 h=%g (20 steps, iterMax = nr)' % h, ms, it, st, 20, 148 * 3)
        print('   status histogram', {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))}, flush=True)
    # page-locking cost of the headline output (2 x 105 MB)
    a = np.empty((4096, 100, 32))
    L = _ffi.lib()
    for r in range(3):
        t0 = time.perf_counter()
        _ffi.check(L.rmx_host_register(_ffi.ptr(a), a.nbytes), 'register')
        t1 = time.perf_counter()
        _ffi.check(L.rmx_host_unregister(_ffi.ptr(a)), 'unregister')
        t2 = time.perf_counter()
        print('cudaHostRegister %.0f MB: %.2f ms, unregister %.2f ms' % (a.nbytes / 1e6, 1e3 * (t1 - t0), 1e3 * (t2 - t1)), flush=True)
