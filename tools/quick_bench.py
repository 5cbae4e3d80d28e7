"""Developer timing probe (not the contract bench): device-resident rollouts, CUDA-event timing."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402


def run(n, B, nsteps, scheme, ground=False, h=1e-2, reps=3):
    sc = rb.chain_scene(n, ground=ground, h=h, nsteps=nsteps)
    sc.init()
    q0, qd0 = rb.synthetic_inputs(sc, B, seed=20260003)
    dq0 = torch.from_numpy(q0).cuda()
    dqd0 = torch.from_numpy(qd0).cuda()
    qo = torch.empty((B, nsteps, sc.nr), dtype=torch.float64, device='cuda')
    qdo = torch.empty_like(qo)
    st = torch.empty(B, dtype=torch.int32, device='cuda')
    it = torch.empty((B, 2), dtype=torch.int32, device='cuda')
    stream = torch.cuda.current_stream()
    best = 1e30
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=scheme, stream=stream)
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    itc = it.cpu().numpy()
    stc = st.cpu().numpy()
    print('n=%d B=%d nsteps=%d scheme=%d ground=%d: %.2f ms  %.3e rollout-steps/s  newton/step %.2f  ls/step %.2f  status|=%d  finite=%s'
          % (n, B, nsteps, scheme, ground, best, B * nsteps / (best * 1e-3), itc[:, 0].mean() / nsteps,
             itc[:, 1].mean() / nsteps, np.bitwise_or.reduce(stc), bool(torch.isfinite(qo).all())), flush=True)


if __name__ == '__main__':
    print(torch.cuda.get_device_name(0))
    run(10, 1024, 100, 1, h=1e-3)
    run(32, 4096, 100, 1, h=1e-3)
    run(32, 4096, 100, 2, h=1e-3)
    if os.environ.get("RMX_QUICK_BIG"):
        run(64, 8192, 50, 1, h=2e-4)
