"""Developer probe: time per rollout-step of the forward kernel as a function of how many blocks are resident per SM
(B = 148 * r rollouts -> r one-warp blocks per SM, one wave).  Tells whether the kernel is latency-bound (time per block
independent of r) or throughput-bound (time per block grows with r)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402

if __name__ == '__main__':
    n, nsteps = 32, 100
    sc = rb.chain_scene(n, h=1e-3, nsteps=nsteps)
    sc.init()
    stream = torch.cuda.current_stream()
    for r in (1, 2, 4, 6, 8):
        B = 148 * r
        q0, qd0 = rb.synthetic_inputs(sc, B, seed=20260003)
        dq0, dqd0 = torch.from_numpy(q0).cuda(), torch.from_numpy(qd0).cuda()
        qo = torch.empty((B, nsteps, sc.nr), dtype=torch.float64, device='cuda')
        qdo = torch.empty_like(qo)
        st = torch.empty(B, dtype=torch.int32, device='cuda')
        best = 1e30
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            sc.rollout_dev(dq0, dqd0, qo, qdo, st, None, scheme=1, stream=stream)
            e1.record(stream)
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print('blocks/SM %d  B=%5d : %.2f ms per wave, %.3e rollout-steps/s, %.1f us per rollout-step per block'
              % (r, B, best, B * nsteps / (best * 1e-3), best * 1e3 / nsteps), flush=True)
