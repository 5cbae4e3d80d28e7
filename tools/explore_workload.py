"""Developer probe: which (h, damping) make the 32-link BDF1 workload a regime where the reference's Newton converges."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import redmax_b200 as rb  # noqa: E402


def run(n, B, nsteps, scheme, h, damping=0.0, stiffness=0.0, ground=False, gz=-40.0, amp=None):
    sc = rb.chain_scene(n, ground=ground, h=h, nsteps=nsteps, ground_z=gz)
    for j in sc.joints:
        j.setDamping(damping)
        j.setStiffness(stiffness)
    sc.init()
    q0, qd0 = rb.synthetic_inputs(sc, B, seed=20260003)
    dq0, dqd0 = torch.from_numpy(q0).cuda(), torch.from_numpy(qd0).cuda()
    qo = torch.empty((B, nsteps, sc.nr), dtype=torch.float64, device='cuda')
    qdo = torch.empty_like(qo)
    st = torch.empty(B, dtype=torch.int32, device='cuda')
    it = torch.empty((B, 2), dtype=torch.int32, device='cuda')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sc.rollout_dev(dq0, dqd0, qo, qdo, st, it, scheme=scheme, stream=torch.cuda.current_stream())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    itc, stc = it.cpu().numpy(), st.cpu().numpy()
    print('n=%d B=%d ns=%d scheme=%d h=%g damp=%g stiff=%g ground=%d: %.1f ms %.3e steps/s newton/step %.2f (max %.1f) ls/step %.2f  div %.4f maxit %.4f lsfail %.4f  |qd|max %.1f'
          % (n, B, nsteps, scheme, h, damping, stiffness, ground, ms, B * nsteps / (ms * 1e-3), itc[:, 0].mean() / nsteps,
             itc[:, 0].max() / nsteps, itc[:, 1].mean() / nsteps, ((stc & 1) != 0).mean(), ((stc & 2) != 0).mean(),
             ((stc & 4) != 0).mean(), float(qdo.abs().max())), flush=True)


if __name__ == '__main__':
    B = 1024
    for h in (1e-2, 5e-3, 2e-3, 1e-3):
        run(32, B, 100, 1, h)
    for d in (1e3, 1e4, 1e5, 1e6):
        run(32, B, 100, 1, 1e-2, damping=d)
    run(32, B, 100, 1, 1e-2, damping=1e4, stiffness=1e6)
    run(32, B, 100, 2, 1e-2, damping=1e5)
    run(32, B, 100, 2, 1e-3)
    for gz in (-40.0, -20.0):
        run(32, B, 100, 2, 5e-4, ground=True, gz=gz)
        run(32, B, 100, 2, 1e-3, ground=True, gz=gz, damping=1e4)
    run(10, B, 100, 1, 1e-2)
    run(64, B, 100, 1, 1e-2, damping=1e6)
    run(64, B, 100, 1, 1e-3)
