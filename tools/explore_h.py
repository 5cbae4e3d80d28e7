"""Developer probe: Newton behaviour of the reference algorithm (as run on the GPU) versus time step."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402


def run(n, B, nsteps, scheme, h, ground=False, damping=0.0, gz=-40.0):
    sc = rb.chain_scene(n, ground=ground, h=h, nsteps=nsteps, ground_z=gz)
    for j in sc.joints:
        j.setDamping(damping)
    sc.init()
    q0, qd0 = rb.synthetic_inputs(sc, B, seed=20260003)
    dq0, dqd0 = torch.from_numpy(q0).cuda(), torch.from_numpy(qd0).cuda()
    qo = torch.empty((B, nsteps, sc.nr), dtype=torch.float64, device='cuda')
    qdo = torch.empty_like(qo)
    st = torch.empty(B, dtype=torch.int32, device='cuda')
    it = torch.empty((B, 2), dtype=torch.int32, device='cuda')
    best = 1e30
    for r in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=scheme, stream=torch.cuda.current_stream())
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    itc, stc = it.cpu().numpy(), st.cpu().numpy()
    print('n=%d B=%d ns=%d scheme=%d h=%g ground=%d damp=%g: %.2f ms %.3e steps/s newton/step %.2f (max %.1f) ls/step %.2f  frac status!=0 %.4f  |qd|max %.1f'
          % (n, B, nsteps, scheme, h, ground, damping, best, B * nsteps / (best * 1e-3), itc[:, 0].mean() / nsteps,
             itc[:, 0].max() / nsteps, itc[:, 1].mean() / nsteps, (stc != 0).mean(), float(qdo.abs().max())), flush=True)


if __name__ == '__main__':
    for gz in (-50.0, -52.0, -55.0):
        for h in (1e-3, 5e-4):
            run(32, 1024, 200, 2, h, ground=True, gz=gz)
    run(32, 1024, 200, 1, 1e-3, ground=True, gz=-52.0)
    run(64, 1024, 100, 1, 5e-4)
    run(64, 1024, 100, 1, 2e-4)
    run(6, 1024, 100, 2, 5e-4, ground=True, gz=-49.0)
