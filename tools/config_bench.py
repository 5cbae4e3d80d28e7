"""Developer probe: device-resident timing of the BASELINE.json configurations other than the headline (they are parity-test
cases, not bench lines): C2 (10-link chain BDF1, B=1024), C3 (32-link chain + ground friction, SDIRK2+BDF2, B=4096), C4 (hand
tree, adjoint objective + gradient, B=2048 = one GPU's share of 8192 on 4), C5 shape (64-link chain BDF1, one GPU's share
B=8192 of 65536 on 8).  Prints rollout-steps/s, Newton iterations per step and status bits."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402


def fwd(label, sc, B, scheme, nsteps, reps=3):
    q0, qd0 = rb.synthetic_inputs(sc, B, seed=20260003)
    dq0, dqd0 = torch.from_numpy(q0).cuda(), torch.from_numpy(qd0).cuda()
    qo = torch.empty((B, nsteps, sc.nr), dtype=torch.float64, device='cuda')
    qdo = torch.empty_like(qo)
    st = torch.empty(B, dtype=torch.int32, device='cuda')
    it = torch.empty((B, 2), dtype=torch.int32, device='cuda')
    stream = torch.cuda.current_stream()
    best = 1e30
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=scheme, nsteps=nsteps, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    itc, stc = it.cpu().numpy(), st.cpu().numpy()
    print('%-34s B=%5d nsteps=%3d: %8.2f ms  %.3e rollout-steps/s  newton/step %.2f  ls/step %.2f  status!=0 %.1f%%  finite=%s'
          % (label, B, nsteps, best, B * nsteps / (best * 1e-3), itc[:, 0].mean() / nsteps, itc[:, 1].mean() / nsteps,
             100.0 * (stc != 0).mean(), bool(torch.isfinite(qo).all())), flush=True)


def adjoint(label, sc, B, reps=3):
    rng = np.random.Generator(np.random.PCG64(20260004))
    p = 0.01 * rng.uniform(-1, 1, (B, sc.nr))
    best = 1e30
    for r in range(reps):
        t0 = time.perf_counter()
        res = sc.rollout_adjoint(p)
        best = min(best, time.perf_counter() - t0)
    print('%-34s B=%5d nsteps=%3d: %8.2f ms  %.3e rollout-steps/s (host-pointer call: forward + tape + backward)  status!=0 %.1f%%'
          % (label, B, sc.nsteps, best * 1e3, B * sc.nsteps / best, 100.0 * (res['status'] != 0).mean()), flush=True)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    sc = rb.chain_scene(10, h=1e-3, nsteps=100)
    sc.init()
    fwd('C2 chain10 BDF1 h=1e-3', sc, 1024, 1, 100)
    sc = rb.chain_scene(32, ground=True, h=5e-4, nsteps=100)
    sc.init()
    fwd('C3 chain32+ground BDF2 h=5e-4', sc, 4096, 2, 100)
    sc = rb.hand_scene()
    sc.init()
    adjoint('C4 hand adjoint BDF1 h=1e-2', sc, 2048)
    sc = rb.chain_scene(64, h=1e-4, nsteps=20)
    sc.init()
    fwd('C5 chain64 BDF1 h=1e-4', sc, 8192, 1, 20, reps=2)
